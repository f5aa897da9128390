#!/usr/bin/env python
"""Throughput of the GED lower-level solver hot path: IPFP pair-solves/sec at AIDS-50+ shape (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong --total 131072]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one pass of the hot path over one candidate batch: every candidate (base pair, edge edit) of the batch is
solved by IPFP (Hungarian start, <=100 Frank-Wolfe iterations with a SciPy-order LSAPE projection each) and scored on
its base pair -- what GEDenv.step does per candidate (reference utils/ged_env.py:110-158), for a whole beam-search /
rollout fan-out at once.  Workload = BASELINE.json configs[2]: synthetic AIDS-50+-shaped pairs (n = 50..100, mean ~60),
`--pairs` base pairs x `--edits` uniform random edge toggles per GPU (weak scaling: every rank owns its own shard of
the candidate list; no data-path collective, the B x 4-byte objectives are all-gathered at the end of each step).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the host-buffer
entry point (pinned host -> device copies, solve, objectives back to the host) inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ipfp_ged_pair_solves_per_sec"
UNIT = "pair-solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="aids-50+")
    ap.add_argument("--pairs", type=int, default=64, help="base pairs per GPU")
    ap.add_argument("--edits", type=int, default=256, help="candidate edge edits per base pair")
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): --pairs x --edits candidates PER GPU; strong: --total candidates over all GPUs")
    ap.add_argument("--total", type=int, default=131072, help="--scaling strong: candidates of the whole job")
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args, rank, world=1):
    """The job's candidate list: `pairs` base pairs x `edits * world` edge toggles (weak scaling: the candidate fan-out
    per base pair grows with the number of GPUs, the mix of pair sizes stays the same), sharded over the ranks with the
    interleaved partition of ppo_bihyb_b200.sharding (no data-path collective: every candidate is independent).
    Returns the base pairs and this rank's (base_idx, edit_u, edit_v)."""
    import numpy as np
    from ppo_bihyb_b200 import sharding, synthetic
    pairs = synthetic.make_pairs(args.config, args.pairs, seed_offset=100)
    base, eu, ev = synthetic.make_edits(pairs, per_pair_edits(args, world), seed_offset=100)
    if world > 1:      # the list is grouped by base pair: interleaving gives every rank the same mix of pairs
        idx = sharding.shard_indices(len(base), rank, world, interleave=True)
        base, eu, ev = base[idx], eu[idx], ev[idx]
    return pairs, base, eu, ev


def per_pair_edits(args, world):
    """edge-edit candidates per base pair of the whole job: fixed per GPU (weak) or fixed in total (strong)"""
    return args.edits * world if args.scaling == "weak" else max(1, args.total // args.pairs)


def workload_name(args, world):
    if args.scaling == "strong":
        return (f"{args.config} synthetic pairs, {args.pairs} base pairs x {per_pair_edits(args, world)} edge-edit candidates in total "
                f"({args.pairs * per_pair_edits(args, world)} candidate solves per step, interleaved over {world} GPU(s)), max_iter {args.max_iter}")
    return (f"{args.config} synthetic pairs, {args.pairs} base pairs x {args.edits} edge-edit candidates per GPU "
            f"({args.pairs * args.edits * world} candidate solves per step, interleaved over {world} GPU(s)), max_iter {args.max_iter}")


# ---- CPU arms: the reference itself (oracle/_ref, oracle/ref_runner.py) and the C port (oracle/ged_oracle.c) -------------
def host_cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def sample_indices(n, count, phase=0):
    """`count` candidate indices spread evenly over a batch of n; `phase` shifts the comb so that successive steps sample
    different candidates of the same workload."""
    return [int((n * (2 * t + 1) // (2 * count) + phase * 7919) % n) for t in range(count)]


def cpu_port(args, sample, threads, phase=0):
    """oracle/ged_oracle.c (the C restatement of the reference's path; `kind: port`) on a bounded sample of the workload,
    `threads` solves in flight (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ged_oracle as go
    from ppo_bihyb_b200.graphs import label_node_cost
    pairs, base, eu, ev = workload(args, 0)
    go.lib()
    idx = sample_indices(len(base), sample, phase)

    def one(b):
        g1, g2 = pairs[base[b]]
        return go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), edit=(int(eu[b]), int(ev[b])),
                        max_iter=args.max_iter)["ged"]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        geds = list(ex.map(one, idx))
    dt = time.perf_counter() - t0
    return len(idx) / dt, dt, idx, geds


class ReferenceArm:
    """The UNMODIFIED reference `GEDenv.step` (utils/ged_env.py:110-124) in a pool of `cores` worker processes, one torch
    thread each -- the reference's own parallelism (ged_ppo_bihyb_train.py:259-260,294-295).  A step = `cores` candidates
    of the workload, one per worker; ori_k of the base pair is built outside the timed part (generate_tuples does that
    once per tuple, :83)."""

    def __init__(self, args, cores):
        from oracle import ref_runner
        self.args, self.cores = args, cores
        self.pairs, self.base, self.eu, self.ev = workload(args, 0)
        self.pool = ref_runner.ReferencePool(cores)

    def jobs(self, phase):
        idx = sample_indices(len(self.base), self.cores, phase)
        return idx, [(self.pairs[self.base[b]][0], self.pairs[self.base[b]][1], int(self.eu[b]), int(self.ev[b])) for b in idx]

    def warm(self, steps):
        """untimed: pool start-up, page-in, first-call costs -- the same code at max_iter 3"""
        if steps <= 0:
            return
        self.pool.set_max_iter(3)
        for w in range(steps):
            self.pool.step(self.jobs(1000 + w)[1])
        self.pool.set_max_iter(self.args.max_iter)

    def step(self, phase):
        idx, jobs = self.jobs(phase)
        geds, wall, per = self.pool.step(jobs)
        return idx, geds, wall, per

    def close(self):
        self.pool.close()


def reference_available():
    """oracle/_ref in place?  (placed by __graft_entry__.build() / oracle/build_ref.py in the build container; it travels
    to the GPU box with the snapshot -- bench.py itself never reads /root/reference)"""
    try:
        from oracle import ref_runner
        return ref_runner.available()
    except Exception:
        return False


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (oracle/_ref when it
    is in place, else the C port), on OUR arm's workload, metric and unit; every step a bounded sample of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": 0,
    }
    port_value, port_dt, _, _ = cpu_port(args, args.cpu_sample or max(128, 8 * cores), cores)
    port = {"value": port_value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{args.cpu_sample or max(128, 8 * cores)} candidates, {port_dt:.1f} s, oracle/ged_oracle.c on {cores} threads"}
    if reference_available():
        arm = ReferenceArm(args, cores)
        try:
            arm.warm(args.warmup)
            walls, per = [], []
            for k in range(max(args.steps, 1)):
                _, _, wall, secs = arm.step(k)
                walls.append(wall)
                per += secs
        finally:
            arm.close()
        value = cores * len(walls) / sum(walls)
        sample = (f"{cores} candidates per step (one per worker process, spread evenly over the batch, a different comb every "
                  f"step), {len(walls)} steps, {sum(walls):.1f} s; unmodified utils/ged_env.py GEDenv.step through oracle/ref_shim.py, "
                  f"{cores} processes x 1 torch thread; mean {statistics.mean(per):.2f} s per candidate per core")
        line.update({"value": value, "ms_per_step": 1e3 * sum(walls) / len(walls),
                     "config": {"workload": workload_name(args, 1), "sample": f"{cores} candidates per step",
                                "warmup_steps": f"{args.warmup} untimed steps of the same code at max_iter 3"},
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                                      "cpu_model": host_cpu_model(), "per_core_value_all_cores_busy": 1.0 / statistics.mean(per),
                                      "port": port}})
    else:
        steps = max(args.steps, 1)
        vals = [cpu_port(args, args.cpu_sample or max(128, 8 * cores), cores, phase=k) for k in range(steps)]
        total = sum(v[1] for v in vals)
        n = sum(len(v[2]) for v in vals)
        line.update({"value": n / total, "ms_per_step": 1e3 * total / steps,
                     "config": {"workload": workload_name(args, 1), "sample": f"{n // steps} candidates per step"},
                     "cpu_baseline": dict(port, value=n / total, cpu_model=host_cpu_model(),
                                          note="oracle/_ref (the reference's own files) is not in place on this box: C port timed")})
    line["e2e"] = {"value": line["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line), flush=True)


# ---- clocks ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "power_w_max": max(pw) if pw else None}


def issue_peaks():
    """Issue roofs of THIS GPU, measured now by scripts/ubench/issue_peak (a few hundred ms); the copy stored with the
    profiles is the fallback when the binary did not travel."""
    exe = os.path.join(ROOT, "scripts", "ubench", "issue_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        peaks = json.loads(out)
        if "fma_winst_per_s" in peaks:
            return peaks, "measured in this run (scripts/ubench/issue_peak.cu)"
    except (OSError, ValueError, IndexError, subprocess.SubprocessError):
        pass
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_r02.json")))["issue_peaks"], "profiles/roofline_r02.json (stored measurement)"
    except (OSError, KeyError, ValueError):
        return None, "unavailable"


def library_version():
    from ppo_bihyb_b200 import _lib
    return _lib.load().ged_version().decode()


def issue_roofline(args, world, kernel_ms, lap_steps_per_step):
    """The binding roof of ged_solve_kernel is instruction issue (its busiest unit: the ALU pipe), not HBM or tensor cores
    (DESIGN.md section 5).  achieved = warp instructions the step's solve launches EXECUTE (ncu smsp__inst_executed.sum of
    this very workload, profiles/roofline_r02.json) / the step time measured here; peak = the FFMA issue rate measured on
    this GPU (1 warp instruction per clock per SM sub-partition).  The ALU-pipe roof, the LAP scan-step rate and HBM are
    reported beside it."""
    peaks, src = issue_peaks()
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_r02.json")))
    except (OSError, ValueError):
        pass
    sig = f"{args.config}|{args.pairs}|{args.edits}|{args.max_iter}" if args.scaling == "weak" else f"strong|{args.config}|{args.pairs}|{args.total}|{world}"
    w = prof.get("workloads", {}).get(sig)
    roof = {"bound": "issue", "unit": "warp-instructions/s", "kernel": "ged_solve_kernel", "peak_source": src,
            "lap_scan_steps_per_s": lap_steps_per_step / (kernel_ms * 1e-3) if lap_steps_per_step else None,
            "lap_scan_steps_per_step": lap_steps_per_step}
    if peaks:
        roof["peak"] = peaks["fma_winst_per_s"]
        roof["issue_peaks"] = peaks
    if w and peaks:
        inst = w["warp_instructions_per_step"]
        roof.update({"achieved": inst / (kernel_ms * 1e-3), "frac": inst / (kernel_ms * 1e-3) / peaks["fma_winst_per_s"],
                     "traffic": w.get("dram_bytes_per_step"), "warp_instructions_per_step": inst,
                     "instructions_source": f"ncu smsp__inst_executed.sum over the step's launches, {w.get('source')}",
                     "instructions_counted_on": w.get("library"), "library": library_version(),
                     "alu_pipe": {"achieved": w["alu_warp_instructions_per_step"] / (kernel_ms * 1e-3),
                                  "peak": peaks["alu_winst_per_s"], "unit": "warp-instructions/s",
                                  "frac": w["alu_warp_instructions_per_step"] / (kernel_ms * 1e-3) / peaks["alu_winst_per_s"]},
                     "per_class": w.get("per_class")})
    else:
        roof.update({"achieved": None, "frac": None, "traffic": None,
                     "note": f"no ncu instruction count stored for workload {sig!r} (profiles/roofline_r02.json)"})
    return roof


def algorithmic_bytes(pairs, base):
    """Compulsory HBM traffic of a step: per candidate the two uint8 adjacencies, the labels,
    the 12-byte (base, u, v) record in; the assignment (4(n1+n2)) and 16 bytes of objective/iters/status out."""
    tot = 0
    for b in base:
        n1, n2 = pairs[b][0].n, pairs[b][1].n
        tot += n1 * n1 + n2 * n2 + 4 * (n1 + n2) + 12 + 4 * (n1 + n2) + 16
    return tot


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:       # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    import torch.distributed as dist
    from ppo_bihyb_b200 import _lib, sharding, solver
    from ppo_bihyb_b200.graphs import PairSet

    _lib.require_cuda()                    # fails loudly when the CUDA library or the GPU is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pairs, base, eu, ev = workload(args, rank, world)
    B = len(base)
    total_candidates = args.pairs * per_pair_edits(args, world)
    B_max = -(-total_candidates // world)                                # shards differ by at most one candidate
    ps = PairSet.from_host(pairs, dev)
    base_d = torch.from_numpy(base).to(dev)
    eu_d, ev_d = torch.from_numpy(eu).to(dev), torch.from_numpy(ev).to(dev)
    plan = solver.make_plan(ps, base)
    host_batch = solver.HostBatch(pairs, base, eu, ev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        res = solver.solve_batch(ps, base_idx=base_d, edit_u=eu_d, edit_v=ev_d, max_iter=args.max_iter, plan=plan)
        if world > 1:                      # reward gather: the only collective of the path (B x 4 bytes per rank)
            mine = torch.zeros(B_max, dtype=torch.float32, device=dev)
            mine[:B] = res.ged
            allged = torch.empty(world * B_max, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allged, mine)
        return res

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed before each; returns ms list."""
        out = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            e1.synchronize()
            out.append((e0.elapsed_time(e1), r))
        return out

    # ---- device-resident timing (`value`) -----------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    runs = timed(step_device, args.steps)
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_total = sum(ms for ms, _ in runs)
    res = runs[-1][1]
    launches = sum(r.launches for _, r in runs)
    # LAP scan steps of one step (deterministic for a workload): one extra, untimed pass that asks the kernel for its counters
    counted = solver.solve_batch(ps, base_idx=base_d, edit_u=eu_d, edit_v=ev_d, max_iter=args.max_iter, plan=plan, lap_steps=True)
    lap_steps_local = torch.tensor([float(counted.trace["lap_steps"].sum().item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lap_steps_local)
    lap_steps_per_step = float(lap_steps_local[0])
    assert torch.equal(counted.ged, res.ged)
    status_bad = int((res.status != 0).sum().item())
    mean_iters = float(res.iters.float().mean().item())
    mean_ged = float(res.ged.mean().item())

    # ---- end-to-end timing through the host-buffer entry point (`e2e`) ---------------------------------------------------
    for _ in range(2):
        solver.solve_batch_host(host_batch, dev, max_iter=args.max_iter)
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ged_host = solver.solve_batch_host(host_batch, dev, max_iter=args.max_iter)     # H2D + solve + D2H + sync
        e1.record()
        e1.synchronize()
        e2e_ms.append(e0.elapsed_time(e1))
    barrier()
    assert np.array_equal(ged_host.numpy(), res.ged.cpu().numpy()), "host-buffer path and device path disagree"

    # ---- the reference's own call sizes (latency-bound): one beam-search step (10 pairs x 27 candidates, ged_ppo_bihyb_eval.py:95-100)
    # and one rollout step (32 pairs x 1 edit, ged_ppo_bihyb_train.py:294-298) of the same workload, device-timed, with the
    # launch shape the host layer picks for that size (CTA-per-candidate "team" kernel) and with one warp per candidate ------------
    latency = None
    if rank == 0:
        from ppo_bihyb_b200 import synthetic
        latency = {}
        for name, npairs, nedits in (("beam_step_10x27", 10, 27), ("rollout_step_32x1", 32, 1)):
            lp = synthetic.make_pairs(args.config, npairs, seed_offset=100)
            lb, lu, lv = synthetic.make_edits(lp, nedits, seed_offset=100)
            lps = PairSet.from_host(lp, dev)
            lplan = solver.make_plan(lps, lb)
            row = {"candidates": int(len(lb))}
            for label, hint in (("ms", None), ("ms_one_warp_per_candidate", 1)):
                best = None
                for rep in range(3):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    r = solver.solve_batch(lps, base_idx=lb, edit_u=lu, edit_v=lv, max_iter=args.max_iter, plan=lplan, launch_hint=hint)
                    e1.record()
                    e1.synchronize()
                    if rep:
                        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
                row[label] = best
                row.setdefault("checksum", float(r.ged.double().sum()))
                assert row["checksum"] == float(r.ged.double().sum()), "launch shapes disagree"
            row["solves_per_s"] = row["candidates"] / (row["ms"] * 1e-3)
            latency[name] = row
    barrier()

    tmax = torch.tensor([ms_total, sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, e2e_total = float(tmax[0]), float(tmax[1])

    if rank == 0:
        value = total_candidates * args.steps / (ms_total * 1e-3)
        e2e_value = total_candidates * args.steps / (e2e_total * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak, which = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        alg = algorithmic_bytes(pairs, base)
        kernel_ms = ms_total / args.steps             # the solve launches ARE the step (one kernel, one launch per size class)
        hbm_achieved = alg / (kernel_ms * 1e-3) / 1e9
        roof = issue_roofline(args, world, kernel_ms, lap_steps_per_step)
        roof["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                       "peak_source": which, "algorithmic_bytes_per_step": alg,
                       "note": "not the binding roof: ~1e3 executed instructions per compulsory byte"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "ms_steps_rank0": [round(ms, 2) for ms, _ in runs], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "l2": "flushed between steps (256 MiB write)",
                       "mean_ipfp_iterations": mean_iters, "mean_ged": mean_ged, "failed_solves": status_bad,
                       "parity": "bit-exact vs canonical-order oracle (tests/, smoke)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host_batch.h2d_bytes,
                    "d2h_bytes_per_step": host_batch.d2h_bytes, "ms_per_step": e2e_total / args.steps},
            "gpu_launches": launches,
            "latency_regime": latency,
            "clocks": clocks,
            "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:      # reported at N=1 only
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or max(128, 8 * cores)
            v, dt, idx, geds = cpu_port(args, sample, cores)
            ged_gpu = res.ged.cpu().numpy()
            same = bool(np.array_equal(np.asarray(geds, np.float32), ged_gpu[idx]))
            port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{sample} candidates of the same workload spread evenly over rank 0's batch, "
                              f"{dt:.1f} s, oracle/ged_oracle.c on {cores} threads",
                    "gpu_results_identical_on_sample": same}
            if reference_available():
                arm = ReferenceArm(args, cores)
                try:
                    arm.warm(1)
                    ridx, rgeds, wall, per = arm.step(0)
                finally:
                    arm.close()
                ours = ged_gpu[ridx]
                line["cpu_baseline"] = {
                    "value": cores / wall, "unit": UNIT, "cores": cores, "kind": "reference", "cpu_model": host_cpu_model(),
                    "sample": f"{cores} candidates of the same workload (one per worker process, spread evenly over rank 0's batch), "
                              f"{wall:.1f} s; unmodified utils/ged_env.py GEDenv.step through oracle/ref_shim.py, {cores} processes x 1 "
                              f"torch thread; {statistics.mean(per):.2f} s per candidate per core",
                    "per_core_value_all_cores_busy": 1.0 / statistics.mean(per),
                    "objective_on_sample": {"reference_mean": float(np.mean(rgeds)), "gpu_mean": float(np.mean(ours)),
                                            "identical": int(np.sum(np.asarray(rgeds, np.float32) == ours)), "of": len(ridx),
                                            "note": "the reference's fp32 trajectory is chaotic under summation order "
                                                    "(it disagrees with itself across MKL thread counts): DESIGN.md section 4"},
                    "port": port}
            else:
                line["cpu_baseline"] = dict(port, cpu_model=host_cpu_model())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
