"""TEST INFRASTRUCTURE ONLY -- times the REFERENCE's own CPU implementation of the path (``GEDenv.step``,
``utils/ged_env.py:110-124`` -> ``solve_feasible_ged`` :140-158 -> ``construct_k`` -> ``ipfp_ged`` -> ``comp_ged``), unmodified,
imported through ``oracle/ref_shim.py`` from ``/root/reference`` or ``oracle/_ref`` (``oracle/build_ref.py``).

Used by ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm only.  Parallelism is the reference's own
(``ged_ppo_bihyb_train.py:259-260,294-295``): a pool of worker PROCESSES, one ``ged_env.step`` per task, one torch thread
each.  The workers are plain ``multiprocessing.Process`` objects with a pipe each, so that a candidate can be prepared
(graphs, ``ori_k`` of its base pair -- built once per tuple by ``generate_tuples`` :83 in the reference, not per step)
outside the timed region by the worker that later runs it.
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")          # the only place this module takes the reference from


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "utils", "ged_env.py"))


def _worker(conn, max_iter):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    os.environ["PPO_BIHYB_REFERENCE"] = REF_DIR
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import functools
    import torch
    torch.set_num_threads(1)
    from oracle import ref_shim
    from ppo_bihyb_b200 import synthetic
    ref = ref_shim.load_reference()
    env = ref_shim.make_reference_env("ipfp", "AIDS-50-100", synthetic.NUM_LABELS)
    orig_ipfp = ref.ipfp_ged

    def set_max_iter(m):         # warm-up steps only: the reference's ipfp_ged has max_iter as a keyword (:256)
        ref.ipfp_ged = orig_ipfp if m == 100 else functools.partial(orig_ipfp, max_iter=m)
    set_max_iter(max_iter)
    prepared = None
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg[0] == "stop":
            return
        if msg[0] == "max_iter":
            set_max_iter(int(msg[1]))
            conn.send("ok")
        elif msg[0] == "prepare":
            _, adj1, lab1, adj2, lab2, eu, ev = msg
            g1, g2 = synthetic.HostGraph(adj1, lab1), synthetic.HostGraph(adj2, lab2)
            d1 = ref_shim.make_data(torch.from_numpy(g1.features()), torch.from_numpy(g1.edge_index()))
            d2 = ref_shim.make_data(torch.from_numpy(g2.features()), torch.from_numpy(g2.edge_index()))
            ori_k = env.construct_k(d1, d2).squeeze(0)                    # generate_tuples :83
            prepared = (d1, d2, ori_k, torch.tensor([int(eu), int(ev)]))
            conn.send("prepared")
        elif msg[0] == "run":
            d1, d2, ori_k, act = prepared
            t0 = time.perf_counter()
            reward, _, new_solution = env.step(d1, d2, ori_k, act, 0.0)   # the hot path, as ged_ppo_bihyb_train.py:295 calls it
            conn.send((float(new_solution), time.perf_counter() - t0))


class ReferencePool:
    """``procs`` reference workers.  ``step(jobs)`` runs ``len(jobs) <= procs`` candidates, one per worker, and returns
    (objectives, wall seconds of the timed part, per-candidate seconds)."""

    def __init__(self, procs: int, max_iter: int = 100):
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for _ in range(procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, max_iter), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    def set_max_iter(self, m: int):
        for c in self.conns:
            c.send(("max_iter", m))
        for c in self.conns:
            assert c.recv() == "ok"

    def step(self, jobs):
        """jobs: list of (HostGraph g1, HostGraph g2, eu, ev)."""
        assert len(jobs) <= len(self.conns)
        used = self.conns[:len(jobs)]
        for c, (g1, g2, eu, ev) in zip(used, jobs):
            c.send(("prepare", g1.adj, g1.labels, g2.adj, g2.labels, eu, ev))
        for c in used:
            assert c.recv() == "prepared"
        t0 = time.perf_counter()
        for c in used:
            c.send(("run",))
        out = [c.recv() for c in used]
        wall = time.perf_counter() - t0
        return [o[0] for o in out], wall, [o[1] for o in out]

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except (BrokenPipeError, OSError):
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
