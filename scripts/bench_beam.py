"""Row f-1 measurement: batched beam-search evaluation (ged_ppo_bihyb_eval.py `evaluate`: 10 test pairs, max_steps 10,
search_size 3) on synthetic AIDS-shaped pairs with a random-initialised policy of the pretrained architecture
(34 -> 64 -> 64 -> 64 GCN; the reference's checkpoints are not on the GPU box).  Prints wall time per evaluation, the number
of lower-level solves it issued, and the time the reference's own per-solve cost (tests/golden/solve_*.npz, measured with
the unmodified reference on one CPU thread) would imply for the same number of solves."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.beam_search import evaluate
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data
from ppo_bihyb_b200.policy import ActorCritic

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="aids-20-30,aids-30-50,aids-50+")
ap.add_argument("--pairs", type=int, default=10)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--beam", type=int, default=3)
ap.add_argument("--host-state", action="store_true", help="search over lists of Data graphs (GEDenv.step_batch) instead of device-resident states")
ap.add_argument("--hostloop", action="store_true", help="round 1's device-resident search with host bookkeeping per step (A/B against the sync-free search)")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
policy = ActorCritic(34, 64, False, 0, 3).to(dev).eval()
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class CountingEnv(GEDenv):
    solves = 0

    def step_batch(self, g1s, *a, **k):
        CountingEnv.solves += len(g1s)
        return super().step_batch(g1s, *a, **k)


from ppo_bihyb_b200.device_env import DeviceGEDState
_solve = DeviceGEDState.solve


def counting_solve(self, *a, **k):
    CountingEnv.solves += len(self)
    return _solve(self, *a, **k)


DeviceGEDState.solve = counting_solve
if args.hostloop:
    from ppo_bihyb_b200 import beam_search as _bs
    _bs.beam_search_batch_device = _bs.beam_search_batch_device_hostloop
print("state:", "host lists" if args.host_state else ("device resident, host loop" if args.hostloop else "device resident, no host synchronisation per step"))
print("| config | pairs | evaluate() wall s | lower-level solves | solves/s end to end | reference s/solve (1 CPU thread) | reference time for the same solves |")
print("|---|---|---|---|---|---|---|")
for cfg in args.configs.split(","):
    env = CountingEnv("ipfp", "AIDS-20-30")
    tuples = []
    for g1, g2 in synthetic.make_pairs(cfg, args.pairs, seed_offset=501):
        d1, d2 = Data.from_host(g1, dev), Data.from_host(g2, dev)
        ged0, k0, _ = env.solve_feasible_ged(d1, d2, "ipfp")
        tuples.append((d1, d2, k0, float(ged0), {"ipfp": float(ged0)}, None))
    evaluate(policy, env, tuples, max_steps=2, search_size=args.beam, host_state=args.host_state)          # warm-up
    CountingEnv.solves = 0
    torch.cuda.synchronize()
    t0 = time.time()
    res = evaluate(policy, env, tuples, max_steps=args.steps, search_size=args.beam, host_state=args.host_state)
    torch.cuda.synchronize()
    dt = time.time() - t0
    z = np.load(os.path.join(gold, "solve_" + cfg.replace("+", "plus") + ".npz"))
    ref_s = float(z["ref_seconds"].mean())
    print(f"| {cfg} | {args.pairs} | {dt:.2f} | {CountingEnv.solves} | {CountingEnv.solves / dt:,.0f} | {ref_s:.2f} | "
          f"{CountingEnv.solves * ref_s / 60:.1f} min | mean ratio {res['ratio']['mean']:.3f}", flush=True)
