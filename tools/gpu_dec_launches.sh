#!/bin/bash
mkdir -p gpurun_out
python tools/dec_probe.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/dec_launches.csv python tools/dec_probe.py > gpurun_out/ncu_dec.log 2>&1
echo rc=$?
