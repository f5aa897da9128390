#!/bin/bash
# Validation of the rebuilt tree on one B200 (tag r02k): GPU tests, smoke, both bench arms, then one `ncu --set full` capture of
# the two largest size-class launches of the bench mix (each only after its own command exited 0 without ncu).
TAG=${TAG:-r02k}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16) | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python scripts/prof_bench_step.py 2 > gpurun_out/${TAG}_prof_step.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ged_solve_kernel -s 3 -c 3 -o gpurun_out/${TAG}_full_bench_mix \
    python scripts/prof_bench_step.py 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
cut -c1-400 gpurun_out/${TAG}_bench.json; echo; cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json; echo; tail -2 gpurun_out/${TAG}_prof_step.log; tail -3 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out
