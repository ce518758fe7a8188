#!/bin/bash
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_decoder.py tests/test_gpu_edge_cases.py tests/test_gpu_hardening.py
timeout 600 python bench.py --only decoder > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench.json 2>&1 | cut -c1-1200 | sed -n 6,16p
export DLIMG_B200_GRAPHS=0
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/dec_probe.py > gpurun_out/plain_dec.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/dec_launches2.csv python tools/dec_probe.py > gpurun_out/ncu_dec.log 2>&1
python tools/prepost_probe.py > gpurun_out/plain_pp.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'mask_post_identity' -c 2 -f -o gpurun_out/full_id python tools/prepost_probe.py > gpurun_out/ncu_full_id.log 2>&1
echo done
