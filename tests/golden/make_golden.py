"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists):

  fire_rows.npz   src/FirePoints.xlsx converted row by row (68 rows of [x, y, area, weight, covered])
  targets.npz     Quadrotor_Targets.xlsx (root: 40 steps, src/: 120 steps): recorded MADS inputs of
                  UAV 1 as (x, y, z = R / tan(FOV/2))
  kat.json        known-answer vectors KAT-1..KAT-5 of SURVEY.md section 8c.  The expected numbers
                  in EXPECT below are the survey session's values (a separate restatement); this
                  script recomputes them with oracle/coverage_oracle.py and refuses to write the
                  file if any differs.

The reference is Julia and cannot run here, so these are not outputs of the reference itself:
parity stays "unpinned" (oracle/README.md).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import coverage_oracle as npo  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location(
    "fire_io", os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200", "fire_io.py"))
fire_io = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fire_io)

T = npo.TAN_HALF_FOV_DEFAULT

EXPECT = {
    "kat2": [((250.0, 250.0, 36.0), 164), ((250.0, 250.0, 35.75260777), 164), ((2.5, 2.5, 5.0), 1),
             ((0.0, 0.0, 10.0), 3), ((5.5, 6.5, 5.0), 3)],
    "kat3": {"count": 74, "area": 1850.0, "violation": 119.17535925942099, "objective": 11915685.925942099,
             "r_max": 35.7526077778263},
    "kat4": {"count": 312, "area": 7800.0, "violation": 1.2369611108685064, "objective": 115896.11108685064},
    "kat5": {"rows": 68, "entries": 6276, "unique": 3061, "entries_10": 455, "unique_10": 254,
             "mult_hist": {"1": 799, "2": 1309, "3": 953}},
}


def main():
    rows = fire_io.load_fire_rows_xlsx(os.path.join(REF, "src", "FirePoints.xlsx"))
    fire_io.save_fire_rows_npz(os.path.join(HERE, "fire_rows.npz"), rows)
    allp = np.concatenate(rows)
    uniq, cnt = np.unique(allp[:, :2], axis=0, return_counts=True)
    first10 = np.concatenate(rows[:10])
    k5 = {"rows": len(rows), "entries": int(len(allp)), "unique": int(len(uniq)),
          "entries_10": int(len(first10)), "unique_10": int(len(np.unique(first10[:, :2], axis=0))),
          "mult_hist": {str(k): int(v) for k, v in enumerate(np.bincount(cnt)) if v}}
    assert k5 == EXPECT["kat5"], k5
    # a disc containing the whole fire counts list entries, not unique cells
    k5["all_covering_disc"] = [250.0, 180.0, 400.0]
    k5["cumulative_entries"] = [int(sum(len(r) for r in rows[:k])) for k in (10, 20, 30, 40, 50, 68)]
    assert k5["cumulative_entries"] == [455, 1069, 1887, 2832, 3979, 6276]

    tg = {}
    for name, path in (("root", "Quadrotor_Targets.xlsx"), ("src", "src/Quadrotor_Targets.xlsx")):
        r = fire_io.load_fire_rows_xlsx.__globals__  # reuse the regexes
        import zipfile
        with zipfile.ZipFile(os.path.join(REF, path)) as z:
            xml = z.read("xl/worksheets/sheet1.xml").decode()
        vals = []
        for m in r["_ROW"].finditer(xml):
            row = []
            for c in r["_CELL"].finditer(m.group(1)):
                if c.group(4) is None or 't="s"' in c.group(3):
                    continue
                v = r["_VAL"].search(c.group(4))
                if v:
                    row.append(float(v.group(1)))
            if len(row) >= 3:
                vals.append(row[:3])
        tg[name] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **tg)

    pts = npo.createPOI(5.0, 5.0, 100.0, 100.0)
    kat = {"tan_half_fov": T, "kat1": {"P": int(len(pts)), "first": pts[0].tolist(), "second": pts[1].tolist(),
                                       "last": pts[-1].tolist()}}
    assert kat["kat1"] == {"P": 10000, "first": [2.5, 2.5, 25.0, 25.0, 0.0], "second": [2.5, 7.5, 25.0, 25.0, 0.0],
                           "last": [497.5, 497.5, 25.0, 25.0, 0.0]}
    k2 = []
    for disc, want in EXPECT["kat2"]:
        area, count = npo.calculateArea(np.array(disc), pts)
        assert count == want and area == 25.0 * want, (disc, count, area)
        k2.append({"disc": list(disc), "count": count, "area": area})
    kat["kat2"] = k2
    x3 = npo.allocate_even_circles(15.0, 5, 10 * T, 250.0, 250.0)
    r_max = np.full(5, 30.0 * T)
    obj, count = npo.objective(x3, pts, 5, r_max)
    e = EXPECT["kat3"]
    assert count == e["count"] and obj == e["objective"] and abs(r_max[0] - e["r_max"]) < 1e-12, (obj, count)
    kat["kat3"] = {"x": x3.tolist(), "r_max": r_max.tolist(), "count": count, "area": 25.0 * count, "objective": obj}
    x4 = np.array([265, 255, 238, 238, 255, 250, 264, 259, 241, 236, 36, 36, 36, 36, 36], dtype=np.float64)
    obj, count = npo.objective(x4, pts, 5, r_max)
    e = EXPECT["kat4"]
    assert count == e["count"] and obj == e["objective"], (obj, count)
    kat["kat4"] = {"x": x4.tolist(), "r_max": r_max.tolist(), "count": count, "area": 25.0 * count, "objective": obj}
    kat["kat5"] = k5
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden fixtures written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
