#!/bin/bash
# Session-3 check (dev build): encoder suite, attention softmax modes A/B (0 = eager maxima, 1 = lazy, 2 = lazy + FMA-pipe exp2).
mkdir -p gpurun_out
tools/gpu_tests.sh tests/test_gpu_encoder.py
for r in 1 2; do
for m in 1 0 2; do
  out=$(DLIMG_B200_WA_MODE=$m python bench.py --steps 10 --warmup 3 --only none --cpu-sample 0 2>/dev/null | tail -n 1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode[$m]', round(d['value'],1), {k: round(v['ms_per_step'],3) for k,v in d.get('kernels',{}).items()})"
done
done
DLIMG_B200_WA_MODE=2 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x 2>&1 | tail -n 2
