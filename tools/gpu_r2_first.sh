#!/bin/bash
# first GPU call of round 2: all -m gpu suites file by file + a bench line
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_gemm.py tests/test_gpu_prepost.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_api.py tests/test_gpu_edge_cases.py tests/test_gpu_dropin.py
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
