for L in "" build/variants/lib_c4.so; do echo "--- $L"; COVERAGE_CUDA_LIB=$L timeout 400 python tools/plane_mode_exp.py 3 | cut -c1-150; done
for L in build/variants/lib_l2.so build/variants/lib_l2c4.so; do echo "--- $L"; COVERAGE_CUDA_LIB=$L timeout 400 python tools/plane_mode_exp.py 3,4 | cut -c1-230; done
