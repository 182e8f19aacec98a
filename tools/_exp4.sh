for L in "" build/variants/lib_v4.so; do COVERAGE_CUDA_LIB=$L timeout 300 python tools/c2_quick.py; done
for L in build/variants/lib_v4.so; do echo "--- $L"; COVERAGE_CUDA_LIB=$L timeout 400 python tools/plane_mode_exp.py 3 | cut -c1-140; done
