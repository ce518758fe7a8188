#!/bin/bash
# Round-end evidence: bench line, per-launch list of one eager pass, ncu --set full of the three top kernel families.
# Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
export DLIMG_B200_GRAPHS=0
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 520 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_l.log 2>&1
for spec in "gemm_tc:0:3:full_gemm" "mlp_fused:0:1:full_mlp" "window_attention:1:2:full_attn" "patch_embed:0:1:full_patch" "mbconv_tail:0:1:full_mbconv" "local_conv:0:3:full_lc"; do
  IFS=: read k s c o <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/$o \
      python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_$o.log 2>&1
done
echo done
