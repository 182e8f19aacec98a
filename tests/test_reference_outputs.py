"""Parity against the REFERENCE ITSELF, when its outputs are available.

tests/golden/reference_dump.jl runs the unmodified Julia reference on tests/golden/reference_inputs.txt and writes
tests/golden/reference_outputs.txt.  Julia exists neither in the build container nor on the GPU boxes, so that
file is normally ABSENT and the two pinning tests skip (parity then rests on the restatements and the known-answer
vectors: "unpinned by the reference").  A machine with Julia turns them on with one command (see the script).

The rail itself is tested here without Julia: the same comparison code runs against a dump EMULATED by the C
oracle, which checks the file formats, the parser and the comparison logic end to end.
"""
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
INPUTS = os.path.join(GOLDEN, "reference_inputs.txt")
OUTPUTS = os.path.join(GOLDEN, "reference_outputs.txt")


def _f(tokens):
    return np.array([int(t, 16) for t in tokens], dtype=np.uint64).view(np.float64)


def read_inputs(fire_rows):
    cases, cur = [], None
    with open(INPUTS) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "case":
                cur = {"name": tok[1], "N": int(tok[3]), "B": int(tok[5]), "P": int(tok[7]), "points": [], "X": []}
                cases.append(cur)
            elif tok[0] in ("r_max", "pre", "d_lim"):
                cur[tok[0]] = _f(tok[1:])
            elif tok[0] == "point":
                cur["points"].append(_f(tok[1:]))
            elif tok[0] == "x":
                cur["X"].append(_f(tok[1:]))
    for c in cases:
        c["X"] = np.array(c["X"])
        c["points"] = np.array(c["points"]) if c["points"] else None
        assert c["X"].shape == (c["B"], 3 * c["N"])
    return cases


def read_outputs(path):
    out, cur, tan = [], None, None
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "tan_half_fov":
                tan = float(_f(tok[1:2])[0])
            elif tok[0] == "case":
                cur = {"name": tok[1], "P": int(tok[3]), "y": [], "rmv": []}
                out.append(cur)
            elif tok[0] == "y":
                cur["y"].append(tok[1:])
            elif tok[0] == "rmv":
                cur["rmv"].append(tuple(int(t) for t in tok[1:]))
    for c in out:
        y = c.pop("y")
        c["obj"] = _f([t[0] for t in y])
        c["area"] = _f([t[1] for t in y])
        c["cons3"] = np.array([int(t[2]) for t in y], dtype=np.uint8)
        c["cons7"] = np.array([int(t[3]) for t in y], dtype=np.uint8)
        c["cons8"] = np.array([int(t[4]) for t in y], dtype=np.uint8)
        c["prog1"], c["prog2"], c["prog3"] = (_f([t[k] for t in y]) for k in (5, 6, 7))
    return tan, out


def case_points(c, orc):
    return c["points"] if c["points"] is not None else orc.createPOI(5.0, 5.0, 100.0, 100.0)


def emulate_dump(path, cases, orc, tan):
    """What reference_dump.jl writes, produced by the C oracle instead of the reference (format self-test)."""
    def hx(v):
        return f"{int(np.float64(v).view(np.uint64)):016x}"
    with open(path, "w") as f:
        f.write("# EMULATED by the C oracle (tests/test_reference_outputs.py), not the reference\n")
        f.write(f"tan_half_fov {hx(tan)}\n")
        for c in cases:
            pts = case_points(c, orc)
            f.write(f"case {c['name']} P {len(pts)}\n")
            for x in c["X"]:
                obj, _ = orc.objective(x, c["r_max"], pts)
                area, _, _ = orc.calculateArea(x, pts)
                f.write(" ".join(["y", hx(obj), hx(area), str(int(orc.cons3(x, c["pre"], tan, c["d_lim"]))),
                                  str(int(orc.cons7(x, tan))), str(int(orc.cons8(x, 15.0))),
                                  hx(orc.cons1_progressive(x, c["r_max"])), hx(orc.consK_progressive(x, c["r_max"], 2)),
                                  hx(orc.consK_progressive(x, c["r_max"], 3))]) + "\n")
            keys = [tuple(p[:2]) for p in pts]
            for b in range(0, len(c["X"]), 50):
                left = orc.rmvCoveredPOI(c["X"][b], pts)
                k, removed = 0, 0
                for idx, key in enumerate(keys, start=1):
                    if k < len(left) and tuple(left[k][:2]) == key:
                        k += 1
                    else:
                        removed += idx
                f.write(f"rmv {b} {len(left)} {removed}\n")


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


def compare_oracle(cases, tan, ref, orc, npo):
    """The C restatement (every candidate) and the NumPy restatement (a sample) against the dump, bit for bit."""
    assert [c["name"] for c in cases] == [r["name"] for r in ref]
    for c, r in zip(cases, ref):
        pts = case_points(c, orc)
        assert len(pts) == r["P"]
        N = c["N"]
        got = orc.eval_batch(c["X"], N, c["r_max"], pts, pre=c["pre"], d_lim=c["d_lim"], tan_half_fov=tan, want_prog=True)
        assert same(got["obj"], r["obj"]), np.flatnonzero(got["obj"] != r["obj"])[:10]
        assert same(got["progressive"], r["prog1"])
        area = np.array([orc.calculateArea(x, pts)[0] for x in c["X"]])
        assert same(area, r["area"])
        assert [int(orc.cons3(x, c["pre"], tan, c["d_lim"])) for x in c["X"]] == r["cons3"].tolist()
        assert [int(orc.cons7(x, tan)) for x in c["X"]] == r["cons7"].tolist()
        assert [int(orc.cons8(x, 15.0)) for x in c["X"]] == r["cons8"].tolist()
        assert same([orc.consK_progressive(x, c["r_max"], 2) for x in c["X"]], r["prog2"])
        assert same([orc.consK_progressive(x, c["r_max"], 3) for x in c["X"]], r["prog3"])
        for b, n_left, removed in r["rmv"]:
            left = orc.rmvCoveredPOI(c["X"][b], pts)
            assert len(left) == n_left
        for b in range(0, len(c["X"]), 97):
            x = c["X"][b]
            assert same([npo.objective(x, pts, N, c["r_max"])[0]], [r["obj"][b]])
            assert npo.cons3(x, c["pre"], tan, c["d_lim"]) == bool(r["cons3"][b])
        # the dump must exercise both verdicts of every constraint and non-zero progressive terms
        for k in ("cons3", "cons7", "cons8"):
            assert 0 < r[k].sum() < len(r[k]), k
        assert (r["prog2"] > 0).any() and (r["prog3"] > 0).any()


def compare_gpu(cases, tan, ref, cov, orc):
    """libcoverage_cuda through its C ABI against the dump: objective, area, constraint verdicts, progressive terms."""
    for c, r in zip(cases, ref):
        pts = case_points(c, orc)
        N = c["N"]
        with cov.CoverageEngine(0) as eng:
            eng.set_points(pts, 100, 100, 5.0, 5.0)
            assert eng.grid_info()["area_exact"] == 1
            for kernel in (cov.KERNEL_AUTO, cov.KERNEL_SPAN, cov.KERNEL_SPAN_GENERAL, cov.KERNEL_BRUTE):
                eng.set_option(cov.OPT_KERNEL, kernel)
                eng.set_params(N, c["r_max"], 1e5, prev_xyR=c["pre"], d_lim=c["d_lim"], tan_half_fov=tan)
                got = eng.eval_batch(c["X"], want_progressive=True)
                assert same(got["obj"], r["obj"]), (c["name"], kernel)
                assert got["feasible"].tolist() == r["cons3"].tolist()
                assert same(got["progressive"], r["prog1"])
                assert same(got["count"] * 25.0, r["area"])  # every entry of both lists weighs 25.0
                for which, key in ((2, "prog2"), (3, "prog3")):
                    eng.set_option(cov.OPT_PROGRESSIVE_INDEX, which)
                    assert same(eng.eval_batch(c["X"], want_progressive=True)["progressive"], r[key])
                eng.set_option(cov.OPT_PROGRESSIVE_INDEX, 0)
                eng.set_params(N, c["r_max"], 1e5, tan_half_fov=tan, use_cons7=True)
                assert eng.eval_batch(c["X"])["feasible"].tolist() == r["cons7"].tolist()
                eng.set_params(N, c["r_max"], 1e5, sep_min=15.0)
                assert eng.eval_batch(c["X"])["feasible"].tolist() == r["cons8"].tolist()
            for b, n_left, removed in r["rmv"][:6]:
                eng.set_points(pts, 100, 100, 5.0, 5.0)
                assert eng.remove_covered(c["X"][b]) == len(pts) - n_left


# ---------------------------------------------------------------- the rail, without Julia
def test_inputs_fixture_is_reproducible(fire_rows, orc):
    cases = read_inputs(fire_rows)
    assert [(c["name"], c["N"], c["B"]) for c in cases] == [("static", 5, 1000), ("fire", 5, 1000)]
    assert cases[0]["points"] is None and np.array_equal(cases[1]["points"], np.concatenate(fire_rows[:10]))
    assert len(cases[1]["points"]) == 455  # KAT-5: rows 1..10 of FirePoints.xlsx, duplicates kept


def test_rail_on_emulated_dump(fire_rows, orc, npo, tmp_path):
    cases = read_inputs(fire_rows)
    tan = math.tan((100 / 180 * math.pi) / 2)
    path = os.path.join(tmp_path, "reference_outputs.txt")
    emulate_dump(path, cases, orc, tan)
    tan2, ref = read_outputs(path)
    assert tan2 == tan and [len(r["obj"]) for r in ref] == [1000, 1000]
    compare_oracle(cases, tan2, ref, orc, npo)


@pytest.mark.gpu
def test_rail_on_emulated_dump_gpu(fire_rows, orc, cov, tmp_path):
    cases = read_inputs(fire_rows)
    tan = math.tan((100 / 180 * math.pi) / 2)
    path = os.path.join(tmp_path, "reference_outputs.txt")
    emulate_dump(path, cases, orc, tan)
    tan2, ref = read_outputs(path)
    compare_gpu(cases, tan2, ref, cov, orc)


# ---------------------------------------------------------------- the pin, when the reference has been run
needs_dump = pytest.mark.skipif(not os.path.exists(OUTPUTS), reason="tests/golden/reference_outputs.txt absent: run "
                                "`julia tests/golden/reference_dump.jl <reference checkout>` where Julia exists")


@needs_dump
def test_oracle_matches_the_reference(fire_rows, orc, npo):
    tan, ref = read_outputs(OUTPUTS)
    compare_oracle(read_inputs(fire_rows), tan, ref, orc, npo)


@needs_dump
@pytest.mark.gpu
def test_cuda_path_matches_the_reference(fire_rows, orc, cov):
    tan, ref = read_outputs(OUTPUTS)
    compare_gpu(read_inputs(fire_rows), tan, ref, cov, orc)
