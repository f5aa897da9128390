#!/bin/bash
# Multi-GPU measurements of one box (gpurun --gpus 8):
#  (1) the PPO-BiHyb GED train step (BASELINE configs[4]) with the device-resident state, weak scaling (pairs per GPU fixed);
#  (2) strong scaling of the candidate solve at fixed totals (scripts/bench_strong.py).
# The 1-GPU points run on a 1-GPU box (NS="1"): a multi-GPU box is charged per GPU.
# Output: gpurun_out/${TAG}_train_scale.jsonl, gpurun_out/${TAG}_strong_scale.jsonl
NS=${NS:-"8 4 2"}
TAG=${TAG:-r02}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@" 2>>gpurun_out/${TAG}_scale.err | grep '^{'; }
for n in $NS; do run $n scripts/bench_train.py --config aids-20-30 --pairs 64 --episodes 3 >> gpurun_out/${TAG}_train_scale.jsonl; done
for n in $NS; do if [ $n = 8 ] || [ $n = 1 ]; then run $n scripts/bench_train.py --config aids-50+ --pairs 32 --episodes 2 >> gpurun_out/${TAG}_train_scale.jsonl; fi; done
for n in $NS; do run $n scripts/bench_strong.py --totals 131072,32768,4096 >> gpurun_out/${TAG}_strong_scale.jsonl; done
cat gpurun_out/${TAG}_train_scale.jsonl gpurun_out/${TAG}_strong_scale.jsonl; tail -5 gpurun_out/${TAG}_scale.err
