#!/bin/bash
# strong scaling of config 5 (256x256x128 mesh) on $1 GPUs
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --scaling strong --steps 50 --warmup 5 > gpurun_out/r02_bench_strong_n$N.json 2> gpurun_out/r02_strong_n$N.err
grep -E "halo parity|Error|error|Traceback" gpurun_out/r02_strong_n$N.err | head -5 | cut -c1-400
cut -c1-900 gpurun_out/r02_bench_strong_n$N.json
