# round-end ncu evidence: the launch list of the bench command, and full captures of the three bench kernels
set -x
CMD="python bench.py --steps 2 --warmup 3 --launches 4 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv $CMD > gpurun_out/bench_ncu.log 2>&1
echo launches rc=$?
python tools/issue_profile.py run > gpurun_out/issue_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:span_small_kernel -s 1 -c 1 -f -o gpurun_out/r2_final_small_c2 python tools/issue_profile.py run > gpurun_out/ncu_a.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:span_cta_kernel -s 1 -c 1 -f -o gpurun_out/r2_final_cta_c3 python tools/issue_profile.py run > gpurun_out/ncu_b.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:span_cta_kernel -s 4 -c 1 -f -o gpurun_out/r2_final_cta_c4 python tools/issue_profile.py run > gpurun_out/ncu_c.log 2>&1; echo rc=$?
