"""Row f-4: the reference's other lower-level GED solvers on the same kernels -- RRWM and graduated assignment (GA).

Reference (file:line under ``/root/reference/utils``):
  ``rrwm_ged``  ``ged_env.py:452-477``  reweighted random walk: K row-normalised, Hungarian start, v <- K v, L1-normalise,
                                        Sinkhorn(100 it, tau = max(v)/beta) jump mixed in with weight alpha, stop at |dv|_2 < 1e-3,
                                        project with ``hungarian_lap(-v)``
  ``ga_ged``    ``ged_env.py:424-449``  graduated assignment: uniform start, tau annealed 1 -> 0.1 (x0.95), per level up to 100 x
                                        (v <- Sinkhorn(10 it, tau)(-K v)), then hard iterations v <- hungarian_lap(-K v)
  ``Sinkhorn.forward_log``  ``sinkhorn.py:143-239`` (non-batched branch: alternate row / column log-normalisation of the
                                        whole (n1+1) x (n2+1) matrix, dummy row and column included)

B200 shape: all B pairs advance together.  ``K v`` is the factored-K kernel (``ged_kv_batch`` -- K is never built, the reference
multiplies by the dense ((n1+1)(n2+1))^2 matrix), both projections are the SciPy-order LSAPE kernel (``ged_lsape_batch``), the
start is the Hungarian kernel; the small elementwise / log-sum-exp glue between them is batched torch on padded [B, R, C]
tensors with per-pair convergence masks, so a pair stops changing exactly where the reference's per-pair loop breaks.
``ops`` (tests only) swaps the three kernel-backed operations of :class:`_Ragged` for the CPU oracle's, so the glue is checked
against the reference's goldens without a GPU; the product path constructs :class:`_Ragged`, which requires the CUDA library.
These are floating-point baselines (exp / log): parity with the reference is within tolerance on the iterate and exact on
the projected assignment wherever the iterate's margins exceed that tolerance (tests/test_gpu_baselines.py).
"""
from typing import Tuple

import numpy as np
import torch

from . import _lib, solver
from .graphs import PairSet

NEG_INF = float('-inf')


class _Ragged:
    """Index maps between the kernels' ragged layout ([P, mat_stride], matrix b = R_b x C_b row-major at the start of row b)
    and a padded [P, Rmax, Cmax] batch."""

    def __init__(self, ps: PairSet):
        _lib.require_cuda()                 # no CPU fallback: the three kernels below are the solver
        self.ps = ps
        dev = ps.device
        n1, n2 = np.asarray(ps.host_n1, np.int64), np.asarray(ps.host_n2, np.int64)
        self.R, self.C = n1 + 1, n2 + 1
        self.P, self.Rm, self.Cm = len(n1), int(self.R.max()), int(self.C.max())
        i = np.arange(self.Rm)[None, :, None]
        j = np.arange(self.Cm)[None, None, :]
        valid = (i < self.R[:, None, None]) & (j < self.C[:, None, None])
        flat = np.where(valid, i * self.C[:, None, None] + j, 0)
        self.valid = torch.from_numpy(valid).to(dev)
        self.flat = torch.from_numpy(flat.reshape(self.P, -1)).to(dev)
        self.ms = int(ps.mat_stride)

    def pad(self, ragged: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
        out = ragged.gather(1, self.flat).reshape(self.P, self.Rm, self.Cm)
        return torch.where(self.valid, out, torch.full_like(out, fill))

    def unpad(self, padded: torch.Tensor) -> torch.Tensor:
        out = torch.zeros(self.P, self.ms, device=padded.device, dtype=padded.dtype)
        src = torch.where(self.valid, padded, torch.zeros_like(padded)).reshape(self.P, -1)
        # padding entries all map to flat index 0 with value 0; scatter_add keeps entry (0,0) intact
        return out.scatter_add_(1, self.flat, src)

    def kv(self, padded: torch.Tensor) -> torch.Tensor:
        """K_b vec(V_b) for every pair, padded in and out."""
        return self.pad(solver.kv_batch(self.ps, self.unpad(padded)))

    def project(self, padded_cost: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``hungarian_lap`` of every pair's cost -> (0/1 padded matrix, rowmatch, colmatch)."""
        rm, cm, _, _ = solver.lsape_batch(self.unpad(padded_cost), self.ps.host_n1, self.ps.host_n2, check=True)
        return self.dense(rm, cm), rm, cm

    def dense(self, rm: torch.Tensor, cm: torch.Tensor) -> torch.Tensor:
        """(rowmatch, colmatch) -> padded 0/1 matrices (C-ABI convention: rowmatch[i] = n2 is a deletion, colmatch[j] = n1 an
        insertion; entries past n1_b / n2_b are padding)."""
        dev = rm.device
        x = torch.zeros(self.P, self.Rm, self.Cm, device=dev)
        n1 = torch.from_numpy(self.R - 1).to(dev)
        n2 = torch.from_numpy(self.C - 1).to(dev)
        b = torch.arange(self.P, device=dev)[:, None]
        rows = torch.arange(rm.shape[1], device=dev)[None, :].expand(self.P, -1)
        ok = rows < n1[:, None]
        x[b.expand_as(rows)[ok], rows[ok], rm.long()[ok]] = 1.0
        cols = torch.arange(cm.shape[1], device=dev)[None, :].expand(self.P, -1)
        ins = (cols < n2[:, None]) & (cm == n1[:, None])
        x[b.expand_as(cols)[ins], n1[:, None].expand_as(cols)[ins], cols[ins]] = 1.0
        return x

    def sinkhorn(self, s: torch.Tensor, tau: torch.Tensor, iters: int) -> torch.Tensor:
        """Log-domain Sinkhorn of every pair's matrix: the fused kernel (all passes in shared memory, one launch)."""
        if not hasattr(self, "_sizes"):
            self._sizes = (torch.from_numpy(self.R.astype(np.int32)).to(s.device), torch.from_numpy(self.C.astype(np.int32)).to(s.device))
        return solver.sinkhorn_batch(s, self._sizes[0], self._sizes[1], tau.reshape(-1).expand(self.P), iters)

    def start(self) -> torch.Tensor:
        """``hungarian_ged`` of every pair (ged_env.py:303-305), padded 0/1."""
        res = solver.solve_batch(self.ps, solver="hungarian", check=True)
        return self.dense(res.rowmatch, res.colmatch)


def _sinkhorn_log(s: torch.Tensor, valid: torch.Tensor, tau: torch.Tensor, iters: int) -> torch.Tensor:
    """``Sinkhorn(max_iter=iters, tau)(s, nrows, ncols)`` of sinkhorn.py:143-239 for a padded batch: log_s = s / tau, then
    alternately subtract the row (even i) / column (odd i) log-sum-exp; padding stays at -inf and is returned as 0."""
    log_s = torch.where(valid, s / tau, torch.full_like(s, NEG_INF))
    for i in range(iters):
        lse = torch.logsumexp(log_s, 2 if i % 2 == 0 else 1, keepdim=True)
        log_s = log_s - torch.where(torch.isfinite(lse), lse, torch.zeros_like(lse))
    return torch.exp(log_s)


def _l2(d: torch.Tensor) -> torch.Tensor:
    return d.reshape(d.shape[0], -1).norm(dim=1)


def rrwm_batch(ps: PairSet, max_iter: int = 100, sk_iter: int = 100, alpha: float = 0.2, beta: float = 100, trace=None, ops=None):
    """``rrwm_ged`` (ged_env.py:452-477) for every pair of ``ps`` -> (rowmatch, colmatch, iterations[P])."""
    rg = ops if ops is not None else _Ragged(ps)
    valid = rg.valid
    # row sums of K = K 1; the reference divides K by (max row sum + 1e-5 * min row sum) (:454-456)
    d = rg.kv(valid.to(torch.float32))
    dmax = torch.where(valid, d, torch.full_like(d, NEG_INF)).amax((1, 2), keepdim=True)
    dmin = torch.where(valid, d, torch.full_like(d, float('inf'))).amin((1, 2), keepdim=True)
    scale = 1.0 / (dmax + dmin * 1e-5)
    v = rg.start()                                                       # hungarian_ged (:459); scaling K does not move it
    active = torch.ones(rg.P, dtype=torch.bool, device=valid.device)
    iters = torch.zeros(rg.P, dtype=torch.int32, device=valid.device)
    for _ in range(max_iter):
        last_v = v
        nv = rg.kv(v) * scale
        nv = nv / nv.abs().sum((1, 2), keepdim=True)
        tau = nv.amax((1, 2), keepdim=True) / beta
        nv = alpha * rg.sinkhorn(-nv, tau, sk_iter) + (1 - alpha) * last_v
        nv = nv * (1.0 / nv.abs().sum((1, 2), keepdim=True))
        v = torch.where(active[:, None, None], nv, last_v)
        iters += active.to(torch.int32)
        active = active & ~(_l2(v - last_v) < 1e-3)
        if not bool(active.any()):
            break
    if trace is not None:
        trace["v"] = v
    _, rm, cm = rg.project(-v)
    return rm, cm, iters


def ga_batch(ps: PairSet, max_iter: int = 100, sk_iter: int = 10, tau0: float = 1, tau_min: float = 0.1, gamma: float = 0.95,
             trace=None, ops=None):
    """``ga_ged`` (ged_env.py:424-449) for every pair of ``ps`` -> (rowmatch, colmatch)."""
    rg = ops if ops is not None else _Ragged(ps)
    valid = rg.valid
    dev = valid.device
    size = torch.from_numpy((rg.R * rg.C).astype(np.float32)).to(dev)[:, None, None]
    v = valid.to(torch.float32) / size                                     # uniform start (:426)
    tau = tau0
    while tau >= tau_min:
        active = torch.ones(rg.P, dtype=torch.bool, device=dev)
        t = torch.full((rg.P, 1, 1), float(tau), device=dev)
        for _ in range(max_iter):
            last_v = v
            nv = rg.sinkhorn(-rg.kv(v), t, sk_iter)
            v = torch.where(active[:, None, None], nv, last_v)
            active = active & ~(_l2(v - last_v) < 1e-4)
            if not bool(active.any()):
                break
        tau = tau * gamma
    if trace is not None:
        trace["v_soft"] = v
    active = torch.ones(rg.P, dtype=torch.bool, device=dev)
    rm = cm = None
    for _ in range(max_iter):
        last_v = v
        kvv = rg.kv(v)
        if trace is not None and "kv_soft" not in trace:
            trace["kv_soft"] = kvv
        x, nrm, ncm = rg.project(-kvv)
        v = torch.where(active[:, None, None], x, last_v)
        rm = nrm if rm is None else torch.where(active[:, None], nrm, rm)
        cm = ncm if cm is None else torch.where(active[:, None], ncm, cm)
        active = active & ~(_l2(v - last_v) < 1e-4)
        if not bool(active.any()):
            break
    return rm, cm
