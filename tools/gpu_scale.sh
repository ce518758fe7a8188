#!/bin/bash
# usage (on the GPU box, N GPUs): tools/gpu_scale.sh N [extra bench args]  -- one torchrun bench at N ranks
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
    > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "rc=$?"; tail -n 3 gpurun_out/scale_$N.err
python tools/show_bench.py gpurun_out/scale_$N.json 2>&1 | cut -c1-900 | grep -E "^value|^decoder|^scaleout|^pcie|^host_binding|^sustained"
