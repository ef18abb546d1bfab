#!/bin/bash
# round 2, call o: strong scaling N=1: config 5's 8.4M-element mesh on ONE B200 (host memory checked first)
mkdir -p gpurun_out
free -g | head -2 | tee gpurun_out/r02o_host.txt; nproc | tee -a gpurun_out/r02o_host.txt
avail=$(free -g | awk '/Mem:/ {print $7}')
if [ "$avail" -lt 130 ]; then echo "host has only ${avail} GB available: skipping the 8.4M single-GPU run"; exit 0; fi
( while true; do free -g | awk '/Mem:/ {print "used_gb", $3}'; sleep 20; done ) > gpurun_out/r02o_mem.log 2>&1 &
MON=$!
timeout 1500 python bench.py --scaling strong --steps 30 --warmup 5 > gpurun_out/r02o_bench_strong_n1.json 2> gpurun_out/r02o.err
kill $MON
tail -3 gpurun_out/r02o.err; sort -k2 -n gpurun_out/r02o_mem.log | tail -1
cut -c1-1200 gpurun_out/r02o_bench_strong_n1.json
