N=8
export COV_BENCH_ALLOW_MISSING_PROFILE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py > gpurun_out/h2d_n$N.json 2> gpurun_out/h2d_n$N.err; echo h2d rc=$?; tail -2 gpurun_out/h2d_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?; tail -3 gpurun_out/bench_n$N.err; tail -c 600 gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo bench4 rc=$?; tail -c 400 gpurun_out/bench_n4.json
