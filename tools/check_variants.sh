# the GPU suite under the two checking builds (make strict-atomics debug-bounds)
for V in strict_atomics debug_bounds; do
  echo "--- $V"
  COVERAGE_CUDA_LIB=$PWD/build/variants/lib_$V.so timeout 1200 python -m pytest tests -m gpu -q -k "not bench_line and not reference_arm" 2>&1 | tail -4
done
