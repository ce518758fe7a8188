#!/bin/bash
# Quick GPU check used during development (run through gpurun): encoder/decoder parity tests + one bench line.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encoder.py tests/test_gpu_decoder.py -x -q -s -m gpu > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/quick_tests.log
tail -n 30 gpurun_out/quick_tests.log
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
print('images/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
for k, v in d.get('kernels', {}).items():
    print(f"  {k:24s} {v['ms_per_step']:.3f} ms  x{v['launches_per_step']}  {v['share']*100:.1f}%")
print('decoder masks/s', round(d['decoder']['value'], 1), 'e2e', round(d['decoder']['e2e']['value'], 1))
for k, v in d['decoder'].get('kernels', {}).items():
    print(f"  {k:24s} {v['ms_per_step']:.3f} ms  {v['share']*100:.1f}%")
print('roofline', d['roofline']['achieved'], d['roofline']['frac'])
PY
