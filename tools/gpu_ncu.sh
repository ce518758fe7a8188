#!/bin/bash
# usage: tools/gpu_ncu.sh <kernel regex> <skip> <count> <out name>   -- ncu --set full on an eager encoder pass
mkdir -p gpurun_out
export DLIMG_B200_GRAPHS=0
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 \
    python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_$4.log 2>&1
echo "rc=$?"
