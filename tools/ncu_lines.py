#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of an ncu report to CUDA source lines.

ncu's CSV source page lists SASS only; nvdisasm -g on the cubin of the same build gives the line of
every SASS instruction in the same order, so the two are joined by instruction index.

usage: python tools/ncu_lines.py <report.ncu-rep> <cubin name fragment> <mangled kernel fragment> [top N]
e.g.   python tools/ncu_lines.py gpurun_out/prof.ncu-rep cov_span_small span_small_kernelILb0ELi32E 40
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200", "libcoverage_cuda.so")
CSRC = os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200", "csrc")


def main():
    rep, cubin_frag, kern_frag = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if cubin_frag in f and f.count("-") == 0][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern_frag in l)
    ins, cur = [], None
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.lstrip().startswith(".section"):
            if ins:
                break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            ins.append(cur)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if r and r[0].startswith("0x")]
    ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if len(data) != len(ins):
        print(f"warning: {len(data)} profiled vs {len(ins)} disassembled instructions (first kernel of the report is used)")
        data = data[:len(ins)]
    agg = defaultdict(lambda: [0, 0, 0])
    for k, d in enumerate(data):
        a = agg[ins[k]]
        a[0] += int(d[ie])
        a[1] += int(d[sm])
        a[2] += 1
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[1] for v in agg.values())
    print(f"total warp instructions {tot}, samples {tots}, SASS instructions {len(ins)}")
    src = {}
    for key, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        f, l = key if key else ("?", 0)
        if f not in src and os.path.exists(os.path.join(CSRC, f)):
            src[f] = open(os.path.join(CSRC, f)).read().split("\n")
        text = src[f][l - 1].strip()[:90] if f in src and 0 < l <= len(src[f]) else ""
        print(f"{v[0] / 1e6:9.1f}M {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / max(tots, 1):5.1f}%  sass {v[2]:4d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
