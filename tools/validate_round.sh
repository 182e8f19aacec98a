set -x
python -m pytest tests -m gpu -q -k "not bench_line" 2>&1 | tail -5
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/issue_profile.py run > gpurun_out/issue_plain.log 2>&1 && ncu --metrics $M --clock-control none -k "regex:span_small_kernel|span_cta_kernel" --csv --log-file gpurun_out/issue.csv python tools/issue_profile.py run > gpurun_out/issue_ncu.log 2>&1
echo issue rc=$?
COV_BENCH_ALLOW_MISSING_PROFILE=1 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; echo bench rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo ref rc=$?
