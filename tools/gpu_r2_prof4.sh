#!/bin/bash
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_decoder.py tests/test_gpu_edge_cases.py tests/test_gpu_hardening.py tests/test_gpu_api.py
timeout 600 python bench.py --only decoder > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench.json 2>&1 | cut -c1-1600 | sed -n 9,22p
export DLIMG_B200_GRAPHS=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/dec_probe.py > gpurun_out/plain_dec.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/dec_launches4.csv python tools/dec_probe.py > gpurun_out/ncu_dec.log 2>&1
CALLS=1 ncu --set full --clock-control none --import-source on -k regex:'t2i_mma|i2t_mma|token_attn_block|token_post|gemm_tc' -c 24 -f -o gpurun_out/full_dec4 python tools/dec_probe.py > gpurun_out/ncu_full_dec4.log 2>&1
echo done
