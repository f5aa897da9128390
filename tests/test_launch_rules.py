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


def test_two_ended_queue_protocol_takes_every_candidate_once():
    """Model of the two-ended work queue of the tensor + global hybrid launches (csrc/ged_kernels.cu, ged_solve_kernel): one
    int32 holds both cursors (front: low 16 bits, back: high 16 bits); a taker adds 1 or 1 << 16, reads the OLD value, and takes
    index front (fast warps) or count - 1 - back (slow warps) while front + back < count.  Whatever the interleaving of fast and
    slow takers, every candidate is taken exactly once, and every warp's final (failing) attempt leaves the halves intact."""
    rng = np.random.default_rng(7)
    for count in (1, 2, 7, 270, 5120):
        for p_fast in (0.0, 0.3, 0.5, 0.9, 1.0):
            counter, taken, attempts = 0, [], 0
            warps = 2368
            done = np.zeros(warps, bool)
            fast = np.arange(warps) % 8 < 4                      # warps 0..3 of a CTA are the tensor-memory ones
            while not done.all():
                w = int(rng.choice(np.nonzero(~done)[0]))
                if p_fast in (0.0, 1.0) and fast[w] != (p_fast == 1.0) and (~done & (fast == (p_fast == 1.0))).any():
                    continue                                     # one kind of warp runs alone until it is exhausted
                old = counter
                counter += 1 if fast[w] else 1 << 16
                attempts += 1
                f, b = old & 0xffff, old >> 16
                if f + b < count:
                    taken.append(f if fast[w] else count - 1 - b)
                else:
                    done[w] = True
            assert sorted(taken) == list(range(count)), (count, p_fast)
            assert attempts == count + warps and (counter & 0xffff) + (counter >> 16) == attempts
            assert (counter & 0xffff) < 1 << 16 and (counter >> 16) < 1 << 16      # no carry between the halves (count + warps < 60000)
