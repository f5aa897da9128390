"""Batched entry points over the C-ABI (``include/ged_b200.h``).  Device tensors in, device tensors out, launches on
torch's current stream, no synchronisation unless ``check=True`` asks for the per-solve status."""
import ctypes
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .graphs import PairSet

SOLVERS = {"ipfp": _lib.GED_SOLVER_IPFP, "hungarian": _lib.GED_SOLVER_HUNGARIAN}
# Matrices per scratch row (ged_solve_desc.workspace_mats): X, X*A2^T and a third one that lets throughput launches of pairs
# above 64 nodes run extra warps whose LAP cost matrix lives in global memory (16 instead of 12 solves per SM at 65-85 nodes).
WORKSPACE_MATS = 3


def _workspace_mats() -> int:
    return 2 if os.environ.get("GED_B200_WORKSPACE_MATS") == "2" else WORKSPACE_MATS       # 2: tests / A-B runs


# calls of at most this many candidates in total are latency-bound (a rollout / beam-search step): one candidate per CTA
# (only reached when TEAM_MAX_CANDIDATES is set below it)
SPREAD_MAX_CANDIDATES = 512
# ... and of at most this many are solved by one CTA each (all warps of the CTA share the candidate: launch_hint 2)
TEAM_MAX_CANDIDATES = int(os.environ.get("GED_B200_TEAM_MAX", "1024"))


def launch_hint_for(total_candidates: int) -> int:
    """``ged_solve_desc.launch_hint`` for a call of this many candidates in total (include/ged_b200.h)."""
    if total_candidates <= TEAM_MAX_CANDIDATES:
        return 2
    return 1 if total_candidates <= SPREAD_MAX_CANDIDATES else 0


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@dataclass
class SolveResult:
    ged: torch.Tensor                 # [B] fp32, objective on the base (unedited) pair
    ged_edited: torch.Tensor          # [B] fp32, objective on the edited pair
    rowmatch: torch.Tensor            # [B, max_n1] int32 (column of row i; n2 = deleted)
    colmatch: torch.Tensor            # [B, max_n2] int32 (row of column a; n1 = inserted)
    iters: torch.Tensor               # [B] int32
    status: torch.Tensor              # [B] int32
    x: Optional[torch.Tensor] = None  # [B, mat_stride] dense 0/1 solutions (R x C with the pair's own C)
    trace: dict = field(default_factory=dict)
    launches: int = 0                 # ged_solve_kernel launches this call made

    def raise_on_error(self):
        st = self.status
        if bool((st != 0).any().item()):
            bad = int(torch.nonzero(st)[0].item())
            code = int(st[bad].item())
            if code & _lib.GED_ST_BAD_GRAPH:
                raise ValueError(f"solve {bad}: adjacency must be symmetric 0/1 off the diagonal (no multi-edges)")
            if code & _lib.GED_ST_INVALID_COST:
                raise ValueError(f"solve {bad}: matrix contains invalid numeric entries")   # scipy's message
            raise ValueError(f"solve {bad}: cost matrix is infeasible")
        return self


def _as_i32(t, device):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, np.int32))
    return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


SIZE_CLASS_STEP = 8          # node-count granularity of the size classes (one launch per class)


def plan_size_classes(n1: np.ndarray, n2: np.ndarray, step: int = SIZE_CLASS_STEP):
    """Group candidates into size classes, one launch each, so that the per-solve resources of a launch follow ITS pairs
    and not the batch maximum.  What a solve occupies is decided by the kernel's launch shape (csrc/ged_kernels.cu,
    plan_cta): S = ceil(max(n1, n2) / 32) register slots per lane and, with the cost matrix in tensor memory,
    next_pow2(n1 * S) TMEM columns per CTA -- so pairs that share (S, columns) share a class.  Pairs whose matrix does not
    fit tensor memory (> 512 columns) keep it in shared memory and are classed in `step`-node bins of max(n1, n2).
    Returns ``(order int32 [B], classes)`` with ``classes = [(offset, count, max_n1, max_n2), ...]``, largest class first;
    inside a class the heaviest solves come first (longest-processing-time order: the tail of a launch is light solves)."""
    n1 = np.asarray(n1, np.int64)
    n2 = np.asarray(n2, np.int64)
    m = np.maximum(n1, n2)
    S = np.maximum((m + 31) // 32, 1)
    cols = np.maximum(n1 * S, 32)
    cols = 1 << np.ceil(np.log2(cols)).astype(np.int64)
    cls = np.where(cols <= 512, S * 4096 + cols, S * 4096 + 2048 + (m + step - 1) // step)
    work = (n1 + n2) ** 3
    order = np.lexsort((-work, -cls)).astype(np.int32)
    classes, o = [], 0
    sc = cls[order]
    while o < len(order):
        e = o
        while e < len(order) and sc[e] == sc[o]:
            e += 1
        sel = order[o:e]
        classes.append((o, e - o, max(int(n1[sel].max()), 1), max(int(n2[sel].max()), 1)))
        o = e
    return order, classes


_side_streams = {}


def _streams(dev, n):
    pool = _side_streams.setdefault(str(dev), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def solve_batch(pairs: PairSet, base_idx=None, edit_u=None, edit_v=None, solver: str = "ipfp", max_iter: int = 100,
                x_init: Optional[torch.Tensor] = None, dense_x: bool = False, trace: bool = False,
                check: bool = False, score_adj1: Optional[torch.Tensor] = None,
                score_adj1_off: Optional[torch.Tensor] = None, score_idx=None, size_classes: bool = True,
                plan=None, score_adj1_ld: int = 0, lap_steps: bool = False, launch_hint: Optional[int] = None) -> SolveResult:
    """All candidate solves of one batch (``ged_solve_batch``): one launch per size class, concurrently on side
    streams, results in candidate order.

    Candidate b = base pair ``base_idx[b]`` with edge ``(edit_u[b], edit_v[b])`` of graph 1 toggled
    (reference ``GEDenv.step``, ``utils/ged_env.py:110-124``).  ``plan`` = a cached ``(order_device, classes)`` from
    :func:`make_plan` avoids re-planning when the same candidate list is solved repeatedly.  ``lap_steps`` also returns the
    number of LAP scan steps every solve executed (``result.trace["lap_steps"]``; the work figure of the roofline).
    ``launch_hint`` overrides the launch shape chosen from the call's size (0 packed, 1 spread, 2 team; tests / A-B runs)."""
    lib = _lib.require_cuda()
    if solver not in SOLVERS:
        raise NotImplementedError(f"{solver} is not implemented.")      # reference ged_env.py:156
    dev = pairs.device
    if plan is None and size_classes and not trace and x_init is None:
        plan = make_plan(pairs, base_idx)
    base_idx = _as_i32(base_idx, dev)
    edit_u, edit_v = _as_i32(edit_u, dev), _as_i32(edit_v, dev)
    B = pairs.num_pairs if base_idx is None else int(base_idx.numel())
    score_idx = _as_i32(score_idx, dev)
    ms, m1, m2 = pairs.mat_stride, pairs.max_n1, pairs.max_n2
    i32 = dict(device=dev, dtype=torch.int32)
    f32 = dict(device=dev, dtype=torch.float32)
    res = SolveResult(ged=torch.empty(B, **f32), ged_edited=torch.empty(B, **f32),
                      rowmatch=torch.full((B, m1), -1, **i32), colmatch=torch.full((B, m2), -1, **i32),   # -1 past n1 / n2
                      iters=torch.empty(B, **i32), status=torch.empty(B, **i32))
    # Scratch rows for the solves in flight (the iterate X and X*A2^T): one region per launch, sized by the number of
    # solves the device can hold at once, claimed row by row inside the kernel (ged_solve_desc.workspace_*).
    launches = [(0, B, m1, m2)] if (plan is None or len(plan[1]) <= 1) else list(plan[1])
    regions, workspace, locks, legacy, mats = [], None, None, False, _workspace_mats()
    if solver == "ipfp" or x_init is not None:
        with torch.cuda.device(dev):
            floats = rows = 0
            legacy = os.environ.get("GED_B200_WS_PER_CANDIDATE") == "1"      # experiments: one scratch row per candidate
            for (_, count, c1, c2) in launches:
                slots = B if legacy else max(1, min(count, int(lib.ged_solve_max_resident(c1, c2))))
                stride = (c1 + 1) * (c2 + 1)
                regions.append((floats, rows, slots, stride))
                floats += mats * slots * stride
                rows += slots
        workspace = torch.empty(floats, **f32)
        locks = torch.zeros(rows, **i32)
    # one work counter per launch: persistent warps pull candidates from it (ged_solve_desc.work_counter)
    counters = torch.zeros(len(launches), **i32) if os.environ.get("GED_B200_WORK_QUEUE", "1") != "0" else None
    if x_init is not None:
        x_init = x_init.to(**f32).contiguous()
        assert x_init.numel() == B * ms, "x_init must be [B, mat_stride]"
    if dense_x:
        res.x = torch.zeros(B, ms, **f32)
    T = max_iter
    tr = res.trace
    if trace:
        tr.update(cost=torch.zeros(B, T, ms, **f32), x_after=torch.zeros(B, T, ms, **f32),
                  scalars=torch.zeros(B, T, 6, device=dev, dtype=torch.float64),
                  rowmatch=torch.zeros(B, T, m1, **i32), colmatch=torch.zeros(B, T, m2, **i32),
                  init_cost=torch.zeros(B, ms, **f32), init_rowmatch=torch.zeros(B, m1, **i32),
                  init_colmatch=torch.zeros(B, m2, **i32), lap_steps=torch.zeros(B, device=dev, dtype=torch.int64))
    if lap_steps and "lap_steps" not in tr:
        tr["lap_steps"] = torch.zeros(B, device=dev, dtype=torch.int64)
    p = _lib.ptr
    desc = _lib.GedSolveDesc(pairs.c_struct(), B, p(base_idx, _lib.c_i32p), p(edit_u, _lib.c_i32p), p(edit_v, _lib.c_i32p),
                             p(score_adj1, _lib.c_u8p), p(score_adj1_off, _lib.c_i64p), p(score_idx, _lib.c_i32p),
                             int(score_adj1_ld), int(max_iter), SOLVERS[solver], p(x_init, _lib.c_f32p), ms, p(workspace, _lib.c_f32p),
                             p(locks, _lib.c_i32p), 0, 0, p(None, _lib.c_i32p), 0, p(None, _lib.c_i32p),
                             launch_hint_for(B) if launch_hint is None else int(launch_hint), mats)

    def set_region(k):
        if counters is not None:
            desc.work_counter = ctypes.cast(ctypes.c_void_p(counters.data_ptr() + 4 * k), _lib.c_i32p)
        if regions:
            f_off, r_off, slots, stride = regions[k]
            desc.workspace = ctypes.cast(ctypes.c_void_p(workspace.data_ptr() + 4 * f_off), _lib.c_f32p)
            desc.workspace_locks = ctypes.cast(ctypes.c_void_p(locks.data_ptr() + 4 * r_off), _lib.c_i32p)
            desc.workspace_slots, desc.workspace_stride = (0 if legacy else slots), stride

    out = _lib.GedSolveOut(p(res.ged, _lib.c_f32p), p(res.ged_edited, _lib.c_f32p), p(res.rowmatch, _lib.c_i32p),
                           p(res.colmatch, _lib.c_i32p), m1, m2, p(res.iters, _lib.c_i32p), p(res.status, _lib.c_i32p),
                           p(res.x, _lib.c_f32p),
                           p(tr.get("cost"), _lib.c_f32p), p(tr.get("x_after"), _lib.c_f32p), p(tr.get("scalars"), _lib.c_f64p),
                           p(tr.get("rowmatch"), _lib.c_i32p), p(tr.get("colmatch"), _lib.c_i32p),
                           p(tr.get("init_cost"), _lib.c_f32p), p(tr.get("init_rowmatch"), _lib.c_i32p),
                           p(tr.get("init_colmatch"), _lib.c_i32p), p(tr.get("lap_steps"), _lib.c_u64p))
    with torch.cuda.device(dev):
        if plan is None or len(plan[1]) <= 1:
            set_region(0)
            _lib.check(lib.ged_solve_batch(desc, out, _stream(dev)), "ged_solve_batch")
            res.launches = 1
        else:
            order_dev, classes = plan
            cur = torch.cuda.current_stream(dev)
            streams = _streams(dev, len(classes))
            ready = torch.cuda.Event()
            ready.record(cur)
            for k, ((off, count, c1, c2), st) in enumerate(zip(classes, streams)):
                st.wait_event(ready)
                set_region(k)
                desc.pairs.max_n1, desc.pairs.max_n2 = c1, c2
                desc.order = ctypes.cast(ctypes.c_void_p(order_dev.data_ptr() + 4 * off), _lib.c_i32p)
                desc.order_count = count
                _lib.check(lib.ged_solve_batch(desc, out, st.cuda_stream), "ged_solve_batch")
                done = torch.cuda.Event()
                done.record(st)
                cur.wait_event(done)
            res.launches = len(classes)
    if workspace is not None:
        workspace.record_stream(torch.cuda.current_stream(dev))
        locks.record_stream(torch.cuda.current_stream(dev))
    if counters is not None:
        counters.record_stream(torch.cuda.current_stream(dev))
    if check:
        res.raise_on_error()
    return res


def make_plan(pairs: PairSet, base_idx=None):
    """Size-class plan of a candidate list: ``(order on the device, classes)`` (see :func:`plan_size_classes`)."""
    if base_idx is None:
        hn1, hn2 = pairs.host_n1, pairs.host_n2
    else:
        hb = base_idx.detach().cpu().numpy() if torch.is_tensor(base_idx) else np.asarray(base_idx)
        hn1, hn2 = pairs.host_n1[hb], pairs.host_n2[hb]
    if len(hn1) == 0:
        return None
    order, classes = plan_size_classes(hn1, hn2)
    return torch.from_numpy(order).to(pairs.device, non_blocking=True), classes


def lsape_batch(cost: torch.Tensor, n1, n2, check: bool = False):
    """``hungarian_lap`` (reference ``utils/ged_env.py:329-351``) for B cost matrices.

    ``cost`` is [B, mat_stride] with matrix b stored R_b x C_b row-major at the start of its row, or [B, R, C] when all
    pairs share one size.  Returns (rowmatch [B,max_n1], colmatch [B,max_n2], lb [B], status [B])."""
    lib = _lib.require_cuda()
    dev = cost.device
    n1h = np.atleast_1d(np.asarray(n1, np.int32))
    n2h = np.atleast_1d(np.asarray(n2, np.int32))
    B = cost.shape[0]
    if n1h.size == 1 and B > 1:
        n1h, n2h = np.repeat(n1h, B), np.repeat(n2h, B)
    m1, m2 = max(int(n1h.max()), 1), max(int(n2h.max()), 1)
    cost = cost.to(torch.float32).reshape(B, -1).contiguous()
    need = (m1 + 1) * (m2 + 1)              # the C-ABI sizes its per-warp slices from (max_n1, max_n2): a ragged batch whose
    if cost.shape[1] < need:                # largest n1 and largest n2 sit in different matrices needs the padded stride
        cost = torch.nn.functional.pad(cost, (0, need - cost.shape[1]))
    ms = cost.shape[1]
    n1d, n2d = _as_i32(n1h, dev), _as_i32(n2h, dev)
    rm = torch.empty(B, m1, device=dev, dtype=torch.int32)
    cm = torch.empty(B, m2, device=dev, dtype=torch.int32)
    lb = torch.empty(B, device=dev, dtype=torch.float32)
    st = torch.empty(B, device=dev, dtype=torch.int32)
    p = _lib.ptr
    with torch.cuda.device(dev):
        _lib.check(lib.ged_lsape_batch(B, p(n1d, _lib.c_i32p), p(n2d, _lib.c_i32p), m1, m2, p(cost, _lib.c_f32p), ms,
                                       p(rm, _lib.c_i32p), m1, p(cm, _lib.c_i32p), m2, p(lb, _lib.c_f32p),
                                       p(st, _lib.c_i32p), _stream(dev)), "ged_lsape_batch")
    if check and bool((st != 0).any().item()):
        code = int(st[torch.nonzero(st)[0]].item())
        raise ValueError("matrix contains invalid numeric entries" if code & _lib.GED_ST_INVALID_COST
                         else "cost matrix is infeasible")
    return rm, cm, lb, st


def kv_batch(pairs: PairSet, x: torch.Tensor) -> torch.Tensor:
    """cost_b = K_b vec(X_b) without K (replaces ``torch.mm(k, v)``, reference ``utils/ged_env.py:269``).
    ``x`` and the result are [P, mat_stride]."""
    lib = _lib.require_cuda()
    dev = pairs.device
    x = x.to(device=dev, dtype=torch.float32).reshape(pairs.num_pairs, -1).contiguous()
    ms = x.shape[1]
    out = torch.zeros_like(x)
    scratch = torch.empty_like(x)
    p = _lib.ptr
    cs = pairs.c_struct()
    with torch.cuda.device(dev):
        _lib.check(lib.ged_kv_batch(cs, p(x, _lib.c_f32p), p(out, _lib.c_f32p), p(scratch, _lib.c_f32p), ms, _stream(dev)),
                   "ged_kv_batch")
    scratch.record_stream(torch.cuda.current_stream(dev))
    return out


def sinkhorn_batch(s: torch.Tensor, rows: torch.Tensor, cols: torch.Tensor, tau: torch.Tensor, iters: int) -> torch.Tensor:
    """``Sinkhorn(max_iter=iters, tau)(s, nrows, ncols)`` (reference ``utils/sinkhorn.py:143-239``, log domain) for a padded
    batch ``s`` [B, Rm, Cm] with per-matrix sizes ``rows`` / ``cols`` [B] and temperatures ``tau`` [B]: one launch, all passes
    in shared memory.  Returns exp(log_s) inside every block, 0 in the padding."""
    lib = _lib.require_cuda()
    dev = s.device
    s = s.to(torch.float32).contiguous()
    B, Rm, Cm = s.shape
    rows, cols = _as_i32(rows, dev), _as_i32(cols, dev)
    tau = tau.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    assert rows.numel() == cols.numel() == tau.numel() == B
    out = torch.empty_like(s)
    p = _lib.ptr
    with torch.cuda.device(dev):
        _lib.check(lib.ged_sinkhorn_batch(B, p(rows, _lib.c_i32p), p(cols, _lib.c_i32p), Rm, Cm, p(s, _lib.c_f32p), p(tau, _lib.c_f32p),
                                          int(iters), p(out, _lib.c_f32p), _stream(dev)), "ged_sinkhorn_batch")
    return out


def score_batch(pairs: PairSet, rowmatch: torch.Tensor, colmatch: torch.Tensor, base_idx=None) -> torch.Tensor:
    """``comp_ged`` of binary solutions given as assignments (reference ``utils/ged_env.py:245-253``)."""
    lib = _lib.require_cuda()
    dev = pairs.device
    rowmatch = rowmatch.to(device=dev, dtype=torch.int32).contiguous()
    colmatch = colmatch.to(device=dev, dtype=torch.int32).contiguous()
    base_idx = _as_i32(base_idx, dev)
    B = rowmatch.shape[0]
    out = torch.empty(B, device=dev, dtype=torch.float32)
    p = _lib.ptr
    cs = pairs.c_struct()
    with torch.cuda.device(dev):
        _lib.check(lib.ged_score_batch(cs, B, p(base_idx, _lib.c_i32p), p(rowmatch, _lib.c_i32p), rowmatch.shape[1],
                                       p(colmatch, _lib.c_i32p), colmatch.shape[1], p(out, _lib.c_f32p), _stream(dev)),
                   "ged_score_batch")
    return out


def quadform_dense(k: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """x^T K x against a dense K (``comp_ged`` on a materialised K).  k: [nk,nk] or [B,nk,nk]; x: [B,nk] or [nk]."""
    lib = _lib.require_cuda()
    dev = k.device
    k = k.to(torch.float32).contiguous()
    shared = k.dim() == 2
    nk = k.shape[-1]
    x = x.to(device=dev, dtype=torch.float32).reshape(-1, nk).contiguous()
    B = x.shape[0]
    if not shared:
        assert k.shape[0] == B
    out = torch.empty(B, device=dev, dtype=torch.float32)
    ws = torch.empty(max(1, lib.ged_quadform_ws_elems(B, nk)), device=dev, dtype=torch.float32)
    p = _lib.ptr
    with torch.cuda.device(dev):
        _lib.check(lib.ged_quadform_dense_batch(B, nk, p(k, _lib.c_f32p), 0 if shared else nk * nk, p(x, _lib.c_f32p),
                                                p(out, _lib.c_f32p), p(ws, _lib.c_f32p), _stream(dev)),
                   "ged_quadform_dense_batch")
    ws.record_stream(torch.cuda.current_stream(dev))
    return out


def match_to_dense(rowmatch: torch.Tensor, colmatch: torch.Tensor, n1: int, n2: int) -> torch.Tensor:
    """(rowmatch[:n1], colmatch[:n2]) -> the (n1+1) x (n2+1) 0/1 matrix ``hungarian_lap`` returns (``ged_env.py:344-347``)."""
    x = torch.zeros(n1 + 1, n2 + 1, device=rowmatch.device, dtype=torch.float32)
    if n1:
        x[torch.arange(n1, device=x.device), rowmatch[:n1].long()] = 1
    if n2:
        x[n1, :n2] = (colmatch[:n2] == n1).to(torch.float32)
    return x


class HostBatch:
    """A candidate batch in PINNED HOST memory, in the C-ABI's factor layout: what a caller of the reference-facing API
    holds before the solve (graphs as adjacency/label pools + (base pair, u, v) per candidate)."""

    def __init__(self, pairs, base_idx: np.ndarray, edit_u: np.ndarray, edit_v: np.ndarray):
        self.tensors, self.meta = PairSet.pack_host(pairs)
        pin = torch.cuda.is_available()
        mk = lambda a: (torch.from_numpy(np.ascontiguousarray(a, np.int32)).pin_memory() if pin
                        else torch.from_numpy(np.ascontiguousarray(a, np.int32)))
        self.base_idx, self.edit_u, self.edit_v = mk(base_idx), mk(edit_u), mk(edit_v)
        self.B = int(len(base_idx))
        order, self.classes = plan_size_classes(self.meta["host_n1"][base_idx], self.meta["host_n2"][base_idx])
        self.order = mk(order)
        self.ged_host = torch.empty(self.B, dtype=torch.float32).pin_memory() if pin else torch.empty(self.B)

    @property
    def h2d_bytes(self) -> int:
        ts = list(self.tensors.values()) + [self.base_idx, self.edit_u, self.edit_v, self.order]
        return int(sum(t.numel() * t.element_size() for t in ts))

    @property
    def d2h_bytes(self) -> int:
        return int(self.ged_host.numel() * self.ged_host.element_size())


def solve_batch_host(batch: HostBatch, device, solver: str = "ipfp", max_iter: int = 100) -> torch.Tensor:
    """End-to-end call with HOST buffers: pinned host -> device copies of the whole batch, the solve, and the
    device -> host read of the objectives (synchronises).  Returns the pinned host tensor of B objectives."""
    ps = PairSet.from_packed(batch.tensors, batch.meta, device)
    base = batch.base_idx.to(device, non_blocking=True)
    eu = batch.edit_u.to(device, non_blocking=True)
    ev = batch.edit_v.to(device, non_blocking=True)
    order = batch.order.to(device, non_blocking=True)
    res = solve_batch(ps, base_idx=base, edit_u=eu, edit_v=ev, solver=solver, max_iter=max_iter,
                      plan=(order, batch.classes))
    batch.ged_host.copy_(res.ged, non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    batch.last_result = res
    return batch.ged_host
