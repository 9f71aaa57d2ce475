#!/bin/bash
# N-GPU bench lines of one box (weak: one field per GPU; strong: one field split by the sampler contract), each with the
# one-process replay check of the reduced flux.  usage (under gpurun --gpus N): bash tools/bench_multi_gpu.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "weak rc=$?"; tail -3 gpurun_out/r2_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --scaling strong > gpurun_out/r2_bench_${N}gpu_strong.json 2> gpurun_out/r2_bench_${N}gpu_strong.err; echo "strong rc=$?"; tail -3 gpurun_out/r2_bench_${N}gpu_strong.err
