"""CPU: host-side rules added in round 2 that need no GPU -- the launch-shape rule (and the header / binding agreement on the
new launch hint) and the bench workload's weak / strong partitions."""
import argparse
import os
import re

import numpy as np

import bench
from ppo_bihyb_b200 import _lib, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_hint_rule():
    """<= TEAM_MAX_CANDIDATES: one CTA per candidate (2); up to SPREAD_MAX_CANDIDATES: one candidate per CTA (1); else packed (0)."""
    assert solver.launch_hint_for(1) == 2 and solver.launch_hint_for(270) == 2
    assert solver.launch_hint_for(solver.TEAM_MAX_CANDIDATES) == 2
    big = max(solver.TEAM_MAX_CANDIDATES, solver.SPREAD_MAX_CANDIDATES) + 1
    assert solver.launch_hint_for(big) == 0 and solver.launch_hint_for(16384) == 0
    hdr = open(os.path.join(ROOT, "include", "ged_b200.h")).read()
    assert "2: team" in hdr and re.search(r"#define GED_ABI_VERSION (\d+)", hdr).group(1) == str(_lib.ABI_VERSION)


def _args(**kw):
    d = dict(config="aids-20-30", pairs=8, edits=16, max_iter=100, scaling="weak", total=1024)
    d.update(kw)
    return argparse.Namespace(**d)


def test_bench_workload_partitions():
    """weak: every rank owns pairs x edits candidates of one job-wide list; strong: the job's total is fixed and split; the
    shards are disjoint, cover the list and keep every base pair on every rank (interleaved)."""
    for scaling, world, want_total in (("weak", 1, 128), ("weak", 4, 512), ("strong", 1, 1024), ("strong", 4, 1024)):
        args = _args(scaling=scaling)
        assert args.pairs * bench.per_pair_edits(args, world) == want_total
        shards = [bench.workload(args, r, world) for r in range(world)]
        sizes = [len(s[1]) for s in shards]
        assert sum(sizes) == want_total and max(sizes) - min(sizes) <= 1
        keys = np.concatenate([np.stack([b, u, v], 1) for _, b, u, v in shards])
        _, fb, fu, fv = bench.workload(args, 0, 1) if world == 1 else bench.workload(_args(scaling=scaling, edits=args.edits * (world if scaling == "weak" else 1)), 0, 1)
        full = np.stack([fb, fu, fv], 1)
        assert sorted(map(tuple, keys)) == sorted(map(tuple, full))          # the shards are a partition of the job's list
        for _, b, _, _ in shards:
            assert set(np.unique(b)) == set(range(args.pairs))
        assert ("in total" if scaling == "strong" else "per GPU") in bench.workload_name(args, world)
