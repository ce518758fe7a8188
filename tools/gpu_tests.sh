#!/bin/bash
# Runs the -m gpu suites one file at a time so a hang or crash in one does not hide the others.
# usage: tools/gpu_tests.sh [files...]   (logs under gpurun_out/)
mkdir -p gpurun_out
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_gpu_gemm.py tests/test_gpu_prepost.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_api.py); fi
: > gpurun_out/summary.log
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv >> gpurun_out/summary.log 2>&1
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q -x -s > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc : $(tail -n 1 gpurun_out/$name.log)" >> gpurun_out/summary.log
  if [ "$name" = "test_gpu_gemm" ] && [ $rc -ne 0 ]; then echo "gemm failed; continuing with prepost only" >> gpurun_out/summary.log; timeout 900 python -m pytest tests/test_gpu_prepost.py -m gpu -q -s > gpurun_out/test_gpu_prepost.log 2>&1; echo "prepost rc=$?" >> gpurun_out/summary.log; break; fi
done
cat gpurun_out/summary.log
