for L in build/variants/lib_v7.so; do echo "--- $L"; COVERAGE_CUDA_LIB=$L timeout 400 python tools/plane_mode_exp.py 2,3 | cut -c1-230; done
COVERAGE_CUDA_LIB=build/variants/lib_v7.so timeout 900 python -m pytest tests -m gpu -q -x -k "not bench_line and not reference_arm" 2>&1 | tail -5
