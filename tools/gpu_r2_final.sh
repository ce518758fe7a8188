#!/bin/bash
# Round-2 evidence run (one GPU): every -m gpu suite, the default bench line, the reference arm, the ncu launch list of the
# bench command and of one eager 64-prompt decoder pass.  Outputs under gpurun_out/.
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_prepost.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_api.py tests/test_gpu_edge_cases.py tests/test_gpu_dropin.py tests/test_gpu_hardening.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.json 2>/dev/null | cut -c1-1800
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --only none --quick > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
DLIMG_B200_GRAPHS=0 ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_decoder.csv python tools/dec_probe.py > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec rc=$?"
ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_prepost.csv python tools/prepost_probe.py > gpurun_out/ncu_pp.log 2>&1; echo "ncu pp rc=$?"
echo done
