#!/bin/bash
# all -m gpu suites file by file + the full bench line
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_prepost.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_api.py tests/test_gpu_edge_cases.py tests/test_gpu_dropin.py tests/test_gpu_hardening.py
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.json 2>/dev/null | head -80
