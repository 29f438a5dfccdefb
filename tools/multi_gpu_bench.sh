#!/bin/bash
# N-GPU bench line the way the driver launches it:  gpurun --gpus N -- 'bash tools/multi_gpu_bench.sh N'
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-400
