bash tools/_scale.sh 8
timeout 600 python tools/c4_sharded.py --out gpurun_out/c4_sharded_r2.json 2>&1 | tail -6
