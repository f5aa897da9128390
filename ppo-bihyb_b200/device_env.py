"""Device-resident environment state for the callers of the hot path.

``GEDenv.step_batch`` keeps the reference's currency -- lists of ``Data`` graphs -- so every step packs graphs on the host,
copies them over and rebuilds edge lists.  :class:`DeviceGEDState` keeps B (graph_1, graph_2) pairs ON THE DEVICE across
steps in the layout both consumers read directly:

  * ``adj1`` / ``adj2``   uint8 [B, N, N] padded adjacency -- the solver reads it through ``ged_pairs.adj1_ld`` (row stride N,
                          ``adj1_off[b] = b * N * N``), the policy network reads ``adj1.float()`` as its dense batch;
  * ``x1`` / ``x2``       fp32 [B, N, F] node features (fixed: the reference never refreshes the degree one-hot after an edit,
                          ``utils/ged_env.py:110-121``);
  * ``score1``            the ORIGINAL graph 1 of every pair: rewards are scored on ``ori_k`` (``ged_env.py:122``).

``step(actions, prev)`` is the reference's ``GEDenv.step`` for the whole batch with no host round trip: one
``ged_solve_batch`` call (candidate b = pair b with the edge ``actions[:, b]`` toggled, scored on the original pair) and three
index operations that apply the toggles to ``adj1`` in place.  Nothing synchronises; an episode is enqueued ahead of the GPU.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, solver
from .graphs import PairSet
from .policy import DenseGraphBatch


class DeviceGEDState:
    def __init__(self, env, graphs_1: Sequence, graphs_2: Sequence, score_graphs_1: Optional[Sequence] = None, device=None):
        _lib.require_cuda()
        if env.solver_type not in solver.SOLVERS:
            raise NotImplementedError(f"device-resident state runs the {sorted(solver.SOLVERS)} solvers; "
                                      f"{env.solver_type!r} goes through GEDenv.step_batch")
        self.env = env
        dev = torch.device(device) if device is not None else env._dev(*(g.x for g in graphs_1))
        B = len(graphs_1)
        items = [env._host_pair(g1, g2) for g1, g2 in zip(graphs_1, graphs_2)]
        if any(c is not None for _, _, c in items):
            raise NotImplementedError("device-resident state needs one-hot node labels (explicit node-cost matrices are not kept)")
        n1 = np.asarray([h1.n for h1, _, _ in items], np.int32)
        n2 = np.asarray([h2.n for _, h2, _ in items], np.int32)
        N1, N2 = int(n1.max()), int(n2.max())

        def padded(adjs, N):
            out = np.zeros((B, N, N), np.uint8)
            for b, a in enumerate(adjs):
                out[b, :a.shape[0], :a.shape[1]] = np.minimum(a, 1)
            return torch.from_numpy(out).to(dev)

        def features(graphs, N):
            F = int(graphs[0].x.shape[1])
            out = torch.zeros(B, N, F, device=dev)
            for b, g in enumerate(graphs):
                out[b, :g.x.shape[0]] = g.x.to(device=dev, dtype=torch.float32)
            return out

        self.adj1 = padded([h1.adj for h1, _, _ in items], N1)
        self.adj2 = padded([h2.adj for _, h2, _ in items], N2)
        if score_graphs_1 is None:
            self.score1 = self.adj1.clone()
        else:
            self.score1 = padded([env._host_pair(g, g2)[0].adj for g, g2 in zip(score_graphs_1, graphs_2)], N1)
        self.x1, self.x2 = features(graphs_1, N1), features(graphs_2, N2)
        self.n1 = torch.from_numpy(n1).to(dev)
        self.n2 = torch.from_numpy(n2).to(dev)
        self.mask1 = torch.arange(N1, device=dev).unsqueeze(0) < self.n1.unsqueeze(1)
        self.mask2 = torch.arange(N2, device=dev).unsqueeze(0) < self.n2.unsqueeze(1)
        l1 = np.concatenate([h1.labels for h1, _, _ in items]).astype(np.int32)
        l2 = np.concatenate([h2.labels for _, h2, _ in items]).astype(np.int32)
        self._lab = (torch.from_numpy(l1).to(dev), torch.from_numpy(np.concatenate(([0], np.cumsum(n1[:-1]))).astype(np.int64)).to(dev),
                     torch.from_numpy(l2).to(dev), torch.from_numpy(np.concatenate(([0], np.cumsum(n2[:-1]))).astype(np.int64)).to(dev))
        self._host_n = (n1, n2)
        self._finish(dev)

    # ---- shared construction tail (also used by clone / select) -------------------------------------------------------
    def _finish(self, dev):
        B, N1, N2 = self.adj1.shape[0], self.adj1.shape[1], self.adj2.shape[1]
        self.device = dev
        self.rows = torch.arange(B, device=dev)
        self.rows_i32 = self.rows.to(torch.int32)
        self.off1 = self.rows * (N1 * N1)
        self.off2 = self.rows * (N2 * N2)
        n1, n2 = self._host_n
        self.ps = PairSet(self.n1, self.n2, self.adj1.view(-1), self.off1, self.adj2.view(-1), self.off2,
                          self._lab[0], self._lab[1], self._lab[2], self._lab[3], None, None, int(n1.max()), int(n2.max()), n1, n2)
        self.ps.adj1_ld, self.ps.adj2_ld = N1, N2
        self.plan = solver.make_plan(self.ps)
        self.status = torch.zeros(B, device=dev, dtype=torch.int32)      # OR of the GED_ST_* bits of every step so far

    def __len__(self) -> int:
        return self.adj1.shape[0]

    def clone(self) -> "DeviceGEDState":
        """A state that shares everything immutable and owns a copy of the edited adjacency (an episode's start)."""
        out = object.__new__(DeviceGEDState)
        out.__dict__.update(self.__dict__)
        out.adj1 = self.adj1.clone()
        out._finish(self.device)
        return out

    # ---- the two consumers ---------------------------------------------------------------------------------------------
    def batch1(self) -> DenseGraphBatch:
        """Graph 1 as the policy's dense batch (a fresh fp32 copy of the adjacency: safe to keep as a PPO memory entry)."""
        return DenseGraphBatch(self.x1, self.adj1.to(torch.float32), self.mask1, self.n1)

    def batch2(self) -> DenseGraphBatch:
        if not hasattr(self, "_batch2"):
            self._batch2 = DenseGraphBatch(self.x2, self.adj2.to(torch.float32), self.mask2, self.n2)
        return self._batch2

    def solve(self, solver_type: Optional[str] = None) -> torch.Tensor:
        """GED of the current pairs scored on the original pairs (``solve_feasible_ged`` with ``ori_k``), [B] on the device."""
        return self._solve(None, None, solver_type)

    def _solve(self, u, v, solver_type=None):
        env = self.env
        res = solver.solve_batch(self.ps, base_idx=self.rows_i32, edit_u=u, edit_v=v, solver=solver_type or env.solver_type,
                                 max_iter=env.max_iter, check=False, score_adj1=self.score1.view(-1), score_adj1_off=self.off1,
                                 score_idx=self.rows_i32, score_adj1_ld=self.adj1.shape[1], plan=self.plan)
        self.status |= res.status
        return res.ged

    def step(self, actions: torch.Tensor, prev_solutions: torch.Tensor):
        """``GEDenv.step`` (ged_env.py:110-124) for every pair: toggle edge ``actions[:, b]`` of graph 1, re-solve, score on the
        original pair.  ``actions`` [2, B] and ``prev_solutions`` [B] live on the device.  Returns (rewards [B], new solutions [B])
        and leaves the toggles applied to this state."""
        u, v = actions[0].to(self.device), actions[1].to(self.device)
        ged = self._solve(u.to(torch.int32), v.to(torch.int32))
        self.apply_edits(u, v)
        return prev_solutions.to(self.device) - ged, ged

    def apply_edits(self, u: torch.Tensor, v: torch.Tensor) -> None:
        """The edit of ``step`` (:111-121) on the device: the edge disappears if either direction is present, else it appears.
        ``u == v`` leaves the graph as the solver and the policy see it unchanged (a self loop enters neither K nor the GCN)."""
        u, v = u.long(), v.long()
        present = (self.adj1[self.rows, u, v] | self.adj1[self.rows, v, u]) != 0
        new = torch.where(u == v, self.adj1[self.rows, u, v], (~present).to(torch.uint8))
        self.adj1[self.rows, u, v] = new
        self.adj1[self.rows, v, u] = new

    def select(self, parents: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> "DeviceGEDState":
        """Successor states of a search step: state s = pair/state ``parents[s]`` with edge (u[s], v[s]) toggled."""
        parents = parents.to(self.device).long()
        out = object.__new__(DeviceGEDState)
        out.env = self.env
        out.adj1, out.adj2, out.score1 = self.adj1[parents].contiguous(), self.adj2[parents].contiguous(), self.score1[parents].contiguous()
        out.x1, out.x2 = self.x1[parents], self.x2[parents]
        out.n1, out.n2, out.mask1, out.mask2 = self.n1[parents], self.n2[parents], self.mask1[parents], self.mask2[parents]
        hp = parents.cpu().numpy()
        n1, n2 = self._host_n[0][hp], self._host_n[1][hp]
        out._host_n = (n1, n2)
        # labels: gather the parents' label segments into new compact pools
        l1, o1, l2, o2 = self._lab

        def regather(lab, off, n):
            starts = off[parents]
            length = torch.from_numpy(n.astype(np.int64)).to(self.device)
            new_off = torch.cumsum(length, 0) - length
            idx = torch.repeat_interleave(starts - new_off, length) + torch.arange(int(n.sum()), device=self.device)
            return lab[idx], new_off
        nl1, no1 = regather(l1, o1, n1)
        nl2, no2 = regather(l2, o2, n2)
        out._lab = (nl1, no1, nl2, no2)
        out._finish(self.device)
        out.apply_edits(u.to(self.device), v.to(self.device))
        return out

    def expand(self, repeat: int) -> "DeviceGEDState":
        """Every state ``repeat`` times, consecutively (state s -> entries s*repeat .. s*repeat + repeat - 1).  Sizes, labels,
        features, graph 2 and the scoring graph never change under edge toggles, so a search builds its state / candidate
        layouts ONCE with this and afterwards only moves ``adj1`` between them (:meth:`load_adj1`) -- no host work per step."""
        idx = torch.arange(len(self), device=self.device).repeat_interleave(repeat)
        out = object.__new__(DeviceGEDState)
        out.env = self.env
        out.adj1, out.adj2, out.score1 = self.adj1[idx].contiguous(), self.adj2[idx].contiguous(), self.score1[idx].contiguous()
        out.x1, out.x2 = self.x1[idx], self.x2[idx]
        out.n1, out.n2, out.mask1, out.mask2 = self.n1[idx], self.n2[idx], self.mask1[idx], self.mask2[idx]
        n1, n2 = np.repeat(self._host_n[0], repeat), np.repeat(self._host_n[1], repeat)
        out._host_n = (n1, n2)
        l1, o1, l2, o2 = self._lab
        out._lab = (l1, o1[idx], l2, o2[idx])          # label pools are shared: an offset may be used by many entries
        out._finish(self.device)
        return out

    def load_adj1(self, src: "DeviceGEDState", index: torch.Tensor) -> None:
        """``adj1[b] = src.adj1[index[b]]`` in place (states of the same pairs: sizes and labels already agree)."""
        torch.index_select(src.adj1, 0, index, out=self.adj1)

    def subset(self, idx: torch.Tensor) -> "DeviceGEDState":
        """States ``idx`` of this one, unchanged (the beams that survive a search step)."""
        zero = torch.zeros(len(idx), dtype=torch.long)
        out = self.select(idx, zero, zero)                  # u == v: no edit
        out.status |= self.status[idx.to(self.device).long()]
        return out

    # ---- leaving the device --------------------------------------------------------------------------------------------
    def raise_on_error(self) -> None:
        """One synchronising check of everything solved so far (the per-call check of ``step_batch``, deferred)."""
        st = self.status
        if bool((st != 0).any().item()):
            code = int(st[torch.nonzero(st)[0]].item())
            if code & _lib.GED_ST_BAD_GRAPH:
                raise ValueError("adjacency must be symmetric with 0/1 entries")
            raise ValueError("matrix contains invalid numeric entries" if code & _lib.GED_ST_INVALID_COST else "cost matrix is infeasible")

    def edge_lists(self) -> List[torch.Tensor]:
        """Current graph 1 of every pair as a sorted ``edge_index`` [2, 2E] (host), for comparison with the list-of-Data path."""
        adj = self.adj1.cpu().numpy()
        out = []
        for b, n in enumerate(self._host_n[0]):
            r, c = np.nonzero(adj[b, :n, :n])
            out.append(torch.from_numpy(np.stack([r, c]).astype(np.int64)))
        return out
