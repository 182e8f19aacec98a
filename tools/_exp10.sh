for L in build/variants/lib_m4.so; do echo "--- $L"; COVERAGE_CUDA_LIB=$L timeout 600 python tools/plane_mode_exp.py -1,1,2,4 | cut -c1-330; done
echo "--- poll latency main"; timeout 200 python tools/poll_latency.py 2>&1 | grep zerocopy=1
echo "--- poll latency m4"; COVERAGE_CUDA_LIB=build/variants/lib_m4.so timeout 200 python tools/poll_latency.py 2>&1 | grep zerocopy=1
COVERAGE_CUDA_LIB=build/variants/lib_m4.so timeout 900 python -m pytest tests -m gpu -q -x -k "not bench_line and not reference_arm" 2>&1 | tail -5
