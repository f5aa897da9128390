#!/bin/bash
# End-of-round validation on one B200: GPU tests, smoke, the per-class instruction counts of the bench mix (ncu, limited metric
# set) written back into profiles/roofline_r02.json BEFORE bench.py reads it, the issue peaks, both bench arms, the launch list of
# the bench command (each ncu pass only after its own command exited 0 without ncu), then the callers either side of the path.
# Outputs under gpurun_out/ (copy what is to be judged into profiles/).
TAG=${TAG:-r02n}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
LIBVER=$(python -c "import ppo_bihyb_b200._lib as l; print(l.load().ged_version().decode())")
python scripts/prof_bench_step.py 2 > gpurun_out/${TAG}_prof_step.log 2>&1 && \
ncu --csv --clock-control none --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum -k regex:ged_solve_kernel -s 3 -c 3 --log-file gpurun_out/${TAG}_ncu_bench_mix.csv python scripts/prof_bench_step.py 2 > gpurun_out/${TAG}_ncu_mix.log 2>&1
scripts/ubench/issue_peak > gpurun_out/${TAG}_issue_peak.json
python scripts/ncu_roofline.py gpurun_out/${TAG}_ncu_bench_mix.csv "aids-50+|64|256|100" \
    "profiles/${TAG}_ncu_bench_mix.csv (ncu --metrics ... -k regex:ged_solve_kernel -s 3 -c 3 python scripts/prof_bench_step.py 2)" \
    --peaks=gpurun_out/${TAG}_issue_peak.json "--library=${LIBVER}" > gpurun_out/${TAG}_roofline_entry.json && cp profiles/roofline_r02.json gpurun_out/${TAG}_roofline_r02.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python scripts/bench_beam.py > gpurun_out/${TAG}_beam.log 2>&1
for c in aids-20-30:64 aids-50+:32; do python scripts/bench_train.py --config ${c%%:*} --pairs ${c##*:} --episodes 3 >> gpurun_out/${TAG}_train.log 2>&1; done
cut -c1-300 gpurun_out/${TAG}_bench.json; echo; cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json; echo; tail -2 gpurun_out/${TAG}_prof_step.log; cat gpurun_out/${TAG}_issue_peak.json; tail -4 gpurun_out/${TAG}_beam.log; tail -4 gpurun_out/${TAG}_train.log
