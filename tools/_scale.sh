# usage: bash tools/_scale.sh N   -- bench.py and the H2D ceiling at N GPUs of one box
N=$1
export COV_BENCH_ALLOW_MISSING_PROFILE=1
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py > gpurun_out/h2d_n$N.json 2> gpurun_out/h2d_n$N.err; echo h2d rc=$?; tail -2 gpurun_out/h2d_n$N.json
python tools/multi_eval_rate.py > gpurun_out/multi_eval_n$N.json 2> gpurun_out/multi_eval_n$N.err; echo multi rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?; tail -3 gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json
