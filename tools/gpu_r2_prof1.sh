#!/bin/bash
# launch lists + ncu --set full for the decoder and the pre/post kernels (each ncu run after a plain run that exited 0)
mkdir -p gpurun_out
export DLIMG_B200_GRAPHS=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/dec_probe.py > gpurun_out/plain_dec.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/dec_launches.csv python tools/dec_probe.py > gpurun_out/ncu_dec.log 2>&1
python tools/prepost_probe.py > gpurun_out/plain_pp.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/pp_launches.csv python tools/prepost_probe.py > gpurun_out/ncu_pp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'resize_tile|mask_post' -f -o gpurun_out/full_pp python tools/prepost_probe.py > gpurun_out/ncu_full_pp.log 2>&1
CALLS=1 ncu --set full --clock-control none --import-source on -k regex:'i2t_attention|t2i_flash|layernorm256_img|layernorm64|mask_dot|linear_small' -c 14 -f -o gpurun_out/full_dec python tools/dec_probe.py > gpurun_out/ncu_full_dec.log 2>&1
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 520 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_l.log 2>&1
echo done
