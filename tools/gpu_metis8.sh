#!/bin/bash
# N=8 on the reference's METIS partition of the 512x256x64 mesh (the weak-scaling N=8 mesh), next to the slab run
mkdir -p gpurun_out
( while true; do free -g | awk '/Mem:/ {print "used_gb", $3}'; sleep 10; done ) > gpurun_out/r02_metis8_mem.log 2>&1 &
MON=$!
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 50 --warmup 5 --part-file tools/parts/part_512x256x64_8.npy > gpurun_out/r02_bench_n8_metis.json 2> gpurun_out/r02_metis8.err
kill $MON
grep -E "halo parity|Error|error|Traceback" gpurun_out/r02_metis8.err | head -5 | cut -c1-400
sort -k2 -n gpurun_out/r02_metis8_mem.log | tail -1
cut -c1-1200 gpurun_out/r02_bench_n8_metis.json
