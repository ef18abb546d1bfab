#!/bin/bash
# round 2, call g (2 GPUs): NVLink halo parity + weak-scaling N=2 line, gloo host-logic is covered on CPU
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
grep -E "halo parity|Error|error|Traceback" gpurun_out/r02g_bench_n2.err | head
cat gpurun_out/r02g_bench_n2.json
