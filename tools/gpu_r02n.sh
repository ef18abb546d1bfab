#!/bin/bash
# round 2, call n (2 GPUs): bench on the reference's METIS partition (small), halo parity included
mkdir -p gpurun_out
python tools/make_metis_partition.py --nex 128x64x32 --nparts 2 --out gpurun_out/part_128x64x32_2.npy | tee gpurun_out/r02n_metis_part.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --nex 64x64x32 --part-file gpurun_out/part_128x64x32_2.npy > gpurun_out/r02n_bench_n2_metis_small.json 2> gpurun_out/r02n.err
grep -E "halo parity|Error|error|Traceback" gpurun_out/r02n.err | head -5
cut -c1-1500 gpurun_out/r02n_bench_n2_metis_small.json
