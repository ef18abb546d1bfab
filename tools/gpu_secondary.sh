#!/bin/bash
# per-launch durations of the secondary kernels (attenuation, acoustic, C-PML) at sizes >> L2: ncu launch lists
mkdir -p gpurun_out
for w in visco acoustic cpml; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$w.csv \
     python tools/run_secondary_kernels.py $w > gpurun_out/secondary_$w.log 2>&1; echo "$w rc=$?"
  grep "^{" gpurun_out/secondary_$w.log
  python tools/summarize_launches.py gpurun_out/launches_$w.csv | head -12
done
