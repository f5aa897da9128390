#!/bin/bash
# One 8-GPU box: (1) the PPO-BiHyb GED train step (BASELINE configs[4]) at 1/2/4/8 GPUs with the device-resident state,
# (2) strong scaling of the candidate solve at fixed totals.  Output: gpurun_out/r02_train_scale.jsonl, r02_strong_scale.jsonl
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" 2>>gpurun_out/r02_scale.err | grep '^{'; }
: > gpurun_out/r02_train_scale.jsonl; : > gpurun_out/r02_strong_scale.jsonl
for n in 1 2 4 8; do run $n scripts/bench_train.py --config aids-20-30 --pairs 64 --episodes 3 >> gpurun_out/r02_train_scale.jsonl; done
for n in 1 8; do run $n scripts/bench_train.py --config aids-50+ --pairs 32 --episodes 2 >> gpurun_out/r02_train_scale.jsonl; done
for n in 8 4 2 1; do run $n scripts/bench_strong.py --totals 131072,32768,4096 >> gpurun_out/r02_strong_scale.jsonl; done
cat gpurun_out/r02_train_scale.jsonl gpurun_out/r02_strong_scale.jsonl; tail -5 gpurun_out/r02_scale.err
