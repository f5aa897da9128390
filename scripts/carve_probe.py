"""Development aid: solve latency with other kernels in between, under different shared-memory carve-outs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.device_env import DeviceGEDState
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data

dev = torch.device("cuda:0")
env = GEDenv("ipfp", "AIDS-20-30")
pairs = synthetic.make_pairs("aids-50+", 32, seed_offset=700)
g1 = [Data.from_host(a, dev) for a, _ in pairs]
g2 = [Data.from_host(b, dev) for _, b in pairs]
state = DeviceGEDState(env, g1, g2)
a = torch.randn(512, 512, device=dev)
acts = torch.stack([torch.zeros(32, dtype=torch.long), torch.ones(32, dtype=torch.long)]).to(dev)
prev = state.solve()
for between in ("nothing", "mm", "batch1"):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        for t in range(6):
            if between == "mm":
                for _ in range(20):
                    b = a @ a
            elif between == "batch1":
                b = state.batch1()
            r, prev = state.step(acts, prev)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 6 * 1e3
    print(f"carve={os.environ.get('GED_B200_CARVEOUT', 'default(100)')} between={between}: {dt:.1f} ms per step", flush=True)
