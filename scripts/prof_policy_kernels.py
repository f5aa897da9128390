"""Development aid: kernel list of one beam-search step outside the solver (policy propose + candidate bookkeeping)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.beam_search import beam_search_batch_device
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data
from ppo_bihyb_b200.policy import ActorCritic

dev = torch.device("cuda:0")
torch.manual_seed(0)
policy = ActorCritic(34, 64, False, 0, 3).to(dev).eval()
cfg = sys.argv[1] if len(sys.argv) > 1 else "aids-50+"
env = GEDenv("ipfp", "AIDS-20-30")
tuples = []
for g1, g2 in synthetic.make_pairs(cfg, 10, seed_offset=501):
    d1, d2 = Data.from_host(g1, dev), Data.from_host(g2, dev)
    ged0, k0, _ = env.solve_feasible_ged(d1, d2, "ipfp")
    tuples.append((d1, d2, k0, float(ged0)))
beam_search_batch_device(policy, env, tuples, 2, 3)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    beam_search_batch_device(policy, env, tuples, 4, 3)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
