#!/bin/bash
# multi-GPU check: bench under torchrun at N=$1 (p2p halo), then the host-staged halo for comparison
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
cat gpurun_out/bench_n$N.json; grep -v Warning gpurun_out/bench_n$N.err | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 50 --warmup 5 --halo host > gpurun_out/bench_n${N}_host.json 2> gpurun_out/bench_n${N}_host.err; echo "bench host N=$N rc=$?"
cat gpurun_out/bench_n${N}_host.json; grep -v Warning gpurun_out/bench_n${N}_host.err | tail -5
