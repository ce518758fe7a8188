#!/bin/bash
# Session-3 check (dev build): 7x7 attention with two heads per CTA (two CTAs per SM) against the release grouping.
mkdir -p gpurun_out
for r in 1 2; do
for m in 2 1; do
  out=$(DLIMG_B200_WA_HG=$m python bench.py --steps 10 --warmup 3 --only none --cpu-sample 0 2>/dev/null | tail -n 1)
  echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hg[$m]', round(d['value'],1), {k: round(v['ms_per_step'],3) for k,v in d.get('kernels',{}).items()})"
done
done
DLIMG_B200_WA_HG=1 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x 2>&1 | tail -n 2
