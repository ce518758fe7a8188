#!/bin/bash
# Per-launch device times of one eager encoder pass (ncu, cold-cache + serialised: compare shares).
# usage: tools/gpu_launches.sh [out csv name] [launch count]
mkdir -p gpurun_out
export DLIMG_B200_GRAPHS=0
OUT=${1:-launches.csv}
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c ${2:-330} --csv \
    --log-file gpurun_out/$OUT python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"
