# round-end validation on one B200: the GPU suite, the issue profile of the bench kernels (stamped with the hash of
# csrc/), the bench line, the reference arm, and the launch list of a short bench command
set -x
python -m pytest tests -m gpu -q -k "not bench_line" 2>&1 | tail -5
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/issue_profile.py run > gpurun_out/issue_plain.log 2>&1 && ncu --metrics $M --clock-control none -k "regex:span_small_kernel|span_cta_kernel" --csv --log-file gpurun_out/issue.csv python tools/issue_profile.py run > gpurun_out/issue_ncu.log 2>&1
echo issue rc=$?
python tools/issue_profile.py parse gpurun_out/issue.csv gpurun_out/issue_run.json > gpurun_out/issue_parse.log 2>&1; echo parse rc=$?
cp profiles/r2_issue.json gpurun_out/r2_issue_box.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; echo bench rc=$?
python -m pytest tests -m gpu -q -k "bench_line" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo ref rc=$?
CMD="python bench.py --steps 2 --warmup 3 --launches 4 --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_packed_launches.csv $CMD > gpurun_out/bench_ncu.log 2>&1
echo launches rc=$?
