set -x
python -m pytest tests -m gpu -q -k "not bench_line" 2>&1 | tail -8
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/issue_profile.py run > gpurun_out/issue_plain.log 2>&1 && ncu --metrics $M --clock-control none -k "regex:span_small_kernel|span_cta_kernel" --csv --log-file gpurun_out/issue.csv python tools/issue_profile.py run > gpurun_out/issue_ncu.log 2>&1
echo issue rc=$?
ncu --set full --clock-control none --import-source on -k regex:span_cta_kernel -s 1 -c 1 -f -o gpurun_out/r2_cta_c3_sweep python tools/issue_profile.py run > gpurun_out/ncu_c4.log 2>&1; echo rc=$?
echo skip
COV_BENCH_ALLOW_MISSING_PROFILE=1 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo bench rc=$?
