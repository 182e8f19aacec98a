for L in "" build/variants/lib_fs2.so build/variants/lib_fs2c.so build/variants/lib_fs2cp.so; do COVERAGE_CUDA_LIB=$L timeout 300 python tools/c2_quick.py; done
echo "--- main lib"; timeout 400 python tools/plane_mode_exp.py 2,3
echo "--- fs2cp"; COVERAGE_CUDA_LIB=build/variants/lib_fs2cp.so timeout 400 python tools/plane_mode_exp.py 2,3
COVERAGE_CUDA_LIB=build/variants/lib_fs2cp.so timeout 900 python -m pytest tests -m gpu -q -x -k "not bench_line and not reference_arm" 2>&1 | tail -15
