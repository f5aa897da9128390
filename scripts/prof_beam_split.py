"""Development aid: device-time split of a sync-free beam search (solver calls vs policy forward + bookkeeping) and the host
enqueue time.   python scripts/prof_beam_split.py [config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.beam_search import beam_search_batch_device
from ppo_bihyb_b200.device_env import DeviceGEDState
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data
from ppo_bihyb_b200.policy import ActorCritic

dev = torch.device("cuda:0")
torch.manual_seed(0)
policy = ActorCritic(34, 64, False, 0, 3).to(dev).eval()
ev = []
_solve = DeviceGEDState.solve


def timed_solve(self, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = _solve(self, *a, **k)
    e1.record()
    ev.append((e0, e1))
    return out


DeviceGEDState.solve = timed_solve
for cfg in (sys.argv[1:] or ["aids-20-30", "aids-30-50", "aids-50+"]):
    env = GEDenv("ipfp", "AIDS-20-30")
    tuples = []
    for g1, g2 in synthetic.make_pairs(cfg, 10, seed_offset=501):
        d1, d2 = Data.from_host(g1, dev), Data.from_host(g2, dev)
        ged0, k0, _ = env.solve_feasible_ged(d1, d2, "ipfp")
        tuples.append((d1, d2, k0, float(ged0)))
    beam_search_batch_device(policy, env, tuples, 2, 3)
    torch.cuda.synchronize()
    ev.clear()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s0.record()
    beam_search_batch_device(policy, env, tuples, 10, 3)
    s1.record()
    t_host = time.time() - t0
    torch.cuda.synchronize()
    t_all = time.time() - t0
    solve_ms = sum(a.elapsed_time(b) for a, b in ev)
    print(f"{cfg}: host enqueue {t_host * 1e3:.1f} ms; wall {t_all * 1e3:.1f} ms; device span {s0.elapsed_time(s1):.1f} ms; solver calls {len(ev)} x {solve_ms / max(len(ev), 1):.2f} ms = "
          f"{solve_ms:.1f} ms; everything else on the device {s0.elapsed_time(s1) - solve_ms:.1f} ms")
