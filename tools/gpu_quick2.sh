#!/bin/bash
# usage: tools/gpu_quick2.sh "<test files>" "<bench args>"   -- selected -m gpu suites + one bench run
mkdir -p gpurun_out
[ -n "$1" ] && tools/gpu_tests.sh $1
timeout 900 python bench.py $2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.json 2>&1 | cut -c1-1500 | head -60
