"""CPU-only: the C-ABI library loads, exports every symbol include/coverage_cuda.h declares, and
fails loudly (no fallback) when there is no CUDA device.  No compute calls here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "coverage_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cov_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cov):
    names = declared_symbols()
    assert len(names) >= 40
    raw = ctypes.CDLL(cov._lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/coverage_cuda.h but not exported"
    assert set(names) == set(cov._lib.SIGNATURES), set(names) ^ set(cov._lib.SIGNATURES)
    assert cov._lib.lib.cov_abi_version() == 3


def test_no_oracle_in_product():
    """The product must never import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_limits_and_threshold_closed_form(cov, orc):
    lim = cov.limits()
    assert lim["max_uavs"] >= 200 and lim["max_nx"] >= 4096
    rng = np.random.default_rng(5)
    Rs = [5.0, 36.0, 15.0, 1.0, 2.0, 4.0, 0.5, 0.25, 3.0, 1e-3, 1e6, 35.7526077778263, 11.9175359259421,
          math.nextafter(1.0, 2.0), math.nextafter(2.0, 1.0), 1e-150, 1e150, 5e-324, 1e-310, 1.7e308]
    Rs += list(rng.random(300) * 40) + list(np.exp(rng.uniform(-40, 40, 200)))
    for R in Rs:
        assert cov.threshold(R) == orc.threshold_by_search(R), R
    assert cov.threshold(0.0) == 0.0 and cov.threshold(-3.0) == 0.0 and cov.threshold(float("nan")) == 0.0
    assert cov.threshold(math.inf) == math.inf


def test_create_fails_loudly_without_gpu(cov):
    if cov.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cov.CoverageError) as e:
        cov.CoverageEngine(0)
    assert e.value.code == cov._lib.COV_ERR_CUDA and "no CPU fallback" in e.value.message


def test_missing_library_raises(cov, tmp_path):
    with pytest.raises(ImportError):
        cov._lib.load(str(tmp_path / "libcoverage_cuda.so"))


def _build_c_demo(tmp_path):
    import subprocess
    exe = str(tmp_path / "poll_demo")
    pkg = os.path.join(ROOT, "maximumareacoverageoptimization.jl_b200")
    cmd = ["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c99", "-O2", "-Wall", "-Werror",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "poll_demo.c"), "-o", exe,
           "-L" + pkg, "-lcoverage_cuda", "-lm", "-Wl,-rpath," + pkg]
    subprocess.check_call(cmd)
    return exe


def test_header_is_plain_c_and_links(cov, tmp_path):
    """include/coverage_cuda.h compiles as C99 with -Wall -Werror and the example links against the .so."""
    import subprocess
    exe = _build_c_demo(tmp_path)
    if cov.device_count() == 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_demo_runs(cov, tmp_path):
    """The plain-C caller reproduces KAT-3 through cov_eval_one and evaluates a poll set."""
    import subprocess
    r = subprocess.run([_build_c_demo(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "11915685.925942099" in r.stdout and "poll winner" in r.stdout
