"""Multi-GPU partitioning of the reduced-order sweep path: one process per GPU, ``torch.distributed`` plumbing
(NCCL over NVLink on the B200 box; the same code runs over ``gloo`` on CPU tensors in the world_size-2 tests).

What shards and what is exchanged:

  stage 1  rows of the snapshot block are block-sharded; each rank forms an r x r partial Gram matrix, the
           partials are summed with ONE all-reduce per Cholesky-QR pass; the r x r factorisations are replicated
           (deterministic kernels on identical input -> identical factors on every rank, no broadcast needed);
  stage 2  operator rows are sharded conformally with Q; the SpMM needs the Q rows its column indices reach
           outside the local block -- a halo window exchanged point-to-point with the owning ranks (banded FEM
           ordering keeps it a few hundred rows); r x r partial projections are all-reduced;
  stage 3/4 sweep points are split into contiguous blocks, no communication; the F x M x M S-parameters are
           collected with an all-gather.

This module holds the partition arithmetic and the communication steps (device agnostic); the compute between
them is morfem_b200.device (CUDA only, no fallback) -- see ``ShardedHotPath``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------- partition arithmetic
def even_split(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``total`` items for ``rank``; the first ``total % world`` ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("even_split: need 0 <= rank < world")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_ranges(total: int, world: int) -> List[Tuple[int, int]]:
    return [even_split(total, world, r) for r in range(world)]


def column_window(indptr: np.ndarray, indices: np.ndarray, lo: int, hi: int, ncols: int) -> Tuple[int, int]:
    """Smallest contiguous range of column indices touched by CSR rows ``lo..hi-1`` (the halo window of the SpMM),
    always containing ``[lo, hi)`` itself clipped to ``ncols`` so that the local block sits inside the window."""
    s, e = int(indptr[lo]), int(indptr[hi])
    w0, w1 = min(lo, ncols), min(hi, ncols)
    if e > s:
        cols = indices[s:e]
        w0, w1 = min(w0, int(cols.min())), max(w1, int(cols.max()) + 1)
    return w0, w1


@dataclass
class HaloPlan:
    """Which Q rows this rank receives from / sends to every peer for one window ``[win0, win1)``."""
    rank: int
    world: int
    row0: int
    row1: int
    win0: int
    win1: int
    recv: List[Tuple[int, int, int]] = field(default_factory=list)   # (peer, global_lo, global_hi) rows received from peer
    send: List[Tuple[int, int, int]] = field(default_factory=list)   # (peer, global_lo, global_hi) rows sent to peer

    @property
    def halo_rows(self) -> int:
        return sum(hi - lo for _, lo, hi in self.recv)


def build_halo_plan(rank: int, world: int, n_rows: int, windows: List[Tuple[int, int]]) -> HaloPlan:
    """``windows[p]`` is rank p's column window; ownership is ``even_split(n_rows, world, p)``."""
    owned = owner_ranges(n_rows, world)
    row0, row1 = owned[rank]
    win0, win1 = windows[rank]
    plan = HaloPlan(rank, world, row0, row1, win0, win1)
    for p in range(world):
        if p == rank:
            continue
        lo, hi = max(win0, owned[p][0]), min(win1, owned[p][1])
        if hi > lo:
            plan.recv.append((p, lo, hi))
        lo, hi = max(windows[p][0], row0), min(windows[p][1], row1)
        if hi > lo:
            plan.send.append((p, lo, hi))
    return plan


def gather_windows(window: Tuple[int, int], group=None) -> List[Tuple[int, int]]:
    """All ranks' windows (plan time, host integers)."""
    world = dist.get_world_size(group)
    out: List[Optional[Tuple[int, int]]] = [None] * world
    dist.all_gather_object(out, (int(window[0]), int(window[1])), group=group)
    return [tuple(w) for w in out]


# ------------------------------------------------------------------------------------------ communication
def _as_real(t: torch.Tensor) -> torch.Tensor:
    return torch.view_as_real(t) if t.is_complex() else t


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of a (complex) tensor: the r x r Gram / projection partials."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(_as_real(t), op=dist.ReduceOp.SUM, group=group)
    return t


def exchange_halo(q_local: torch.Tensor, plan: HaloPlan, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """Assemble rows ``[win0, win1)`` of the global Q on this rank: the local block is copied, the rest is
    received straight into its slot of the window buffer from the owning ranks (point to point, one message per
    peer and direction)."""
    r = q_local.shape[1]
    nwin = plan.win1 - plan.win0
    if out is None or out.shape[0] < nwin or out.shape[1] != r or out.dtype != q_local.dtype or out.device != q_local.device:
        out = torch.empty((nwin, r), dtype=q_local.dtype, device=q_local.device)
    win = out[:nwin]
    win[plan.row0 - plan.win0:plan.row1 - plan.win0].copy_(q_local)
    ops = []
    for peer, lo, hi in plan.send:
        ops.append(dist.P2POp(dist.isend, _as_real(q_local[lo - plan.row0:hi - plan.row0]), peer, group=group))
    for peer, lo, hi in plan.recv:
        ops.append(dist.P2POp(dist.irecv, _as_real(win[lo - plan.win0:hi - plan.win0]), peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return win


def slab_halo_rows(rank_windows: List[Tuple[int, int]], owned: List[Tuple[int, int]]) -> Optional[int]:
    """Slab height H for the all-gather halo exchange, or None when point-to-point messages are needed.

    Every rank publishes its top H and bottom H local rows; the exchange works when every row a rank needs from a peer
    lies inside one of that peer's two slabs (banded operators: the window reaches a few hundred rows into the
    neighbouring blocks) and every rank owns at least H rows."""
    world = len(owned)
    h = 0
    for p in range(world):
        (w0, w1), (r0, r1) = rank_windows[p], owned[p]
        h = max(h, r0 - w0, w1 - r1)
    if h == 0:
        return 0
    if any(hi - lo < h for lo, hi in owned):
        return None
    for p in range(world):
        w0, w1 = rank_windows[p]
        for q in range(world):
            if q == p:
                continue
            lo, hi = max(w0, owned[q][0]), min(w1, owned[q][1])
            if hi <= lo:
                continue
            in_top = hi <= owned[q][0] + h
            in_bottom = lo >= owned[q][1] - h
            if not (in_top or in_bottom):
                return None
    return h


def exchange_halo_slabs(q_local: torch.Tensor, plan: HaloPlan, h: int, owned: List[Tuple[int, int]],
                        out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """Same result as ``exchange_halo`` with ONE fixed-size collective: every rank contributes its top and bottom ``h`` rows
    to an all-gather, the window is assembled from the peers' slabs with device copies.  Unlike the point-to-point version
    this is a plain NCCL collective on static shapes, so a CUDA graph can capture the whole multi-rank step."""
    r = q_local.shape[1]
    nwin = plan.win1 - plan.win0
    if out is None or out.shape[0] < nwin or out.shape[1] != r or out.dtype != q_local.dtype or out.device != q_local.device:
        out = torch.empty((nwin, r), dtype=q_local.dtype, device=q_local.device)
    win = out[:nwin]
    win[plan.row0 - plan.win0:plan.row1 - plan.win0].copy_(q_local)
    if h == 0:                                             # no rank needs a remote row (all ranks agree on h)
        return win
    slab = torch.empty((2 * h, r), dtype=q_local.dtype, device=q_local.device)
    slab[:h].copy_(q_local[:h])
    slab[h:].copy_(q_local[q_local.shape[0] - h:])
    gathered = torch.empty((plan.world * 2 * h, r), dtype=q_local.dtype, device=q_local.device)   # rank p's slabs at rows 2 h p ..
    dist.all_gather_into_tensor(_as_real(gathered), _as_real(slab), group=group)
    for peer, lo, hi in plan.recv:
        p0, p1 = owned[peer]
        base = 2 * h * peer
        if hi <= p0 + h:                                   # inside the peer's top slab (rows p0 .. p0 + h)
            src = gathered[base + lo - p0:base + hi - p0]
        else:                                              # inside its bottom slab (rows p1 - h .. p1)
            src = gathered[base + h + lo - (p1 - h):base + h + hi - (p1 - h)]
        win[lo - plan.win0:hi - plan.win0].copy_(src)
    return win


def gather_points(local: torch.Tensor, f_total: int, group=None) -> torch.Tensor:
    """All-gather of per-point results split with ``even_split``: returns the (f_total, ...) tensor on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [hi - lo for lo, hi in owner_ranges(f_total, world)]
    cmax = max(counts)
    tail = tuple(local.shape[1:])
    padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]].copy_(local)
    gathered = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(_as_real(gathered), _as_real(padded), group=group)
    if all(c == cmax for c in counts):
        return gathered
    return torch.cat([gathered[p * cmax:p * cmax + counts[p]] for p in range(world)], dim=0)


# -------------------------------------------------------------------------------------------- sharded path
class ShardedHotPath:
    """Stages 1-4 for one rank of a row-/point-sharded job.  Inputs are this rank's slices, already on its GPU:

      operators   list of three scipy csc arrays (GLOBAL; only the local CSR rows of ``a^T`` are uploaded) or None
      b           scipy csc port matrix (global, tiny)
      coefficient arrays c0, c1, c2, cb, zscale for the LOCAL sweep points (host ndarrays)
    """

    def __init__(self, operators, b, n_rows: int, f_total: int, coeffs_global, group=None):
        from . import device as dv
        self.dv = dv
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dev = dv.require_cuda()
        self.n_rows = n_rows
        self.row0, self.row1 = even_split(n_rows, self.world, self.rank)
        self.f_total = f_total
        self.f0, self.f1 = even_split(f_total, self.world, self.rank)
        self.ops = []
        self.plans = []
        self._windows = {}
        for a in operators:
            if a is None or a.nnz == 0:
                self.ops.append(None)
                self.plans.append(None)
                continue
            window = column_window(np.asarray(a.indptr), np.asarray(a.indices), self.row0, self.row1, a.shape[0])
            windows = gather_windows(window, group) if self.world > 1 else [window]
            plan = build_halo_plan(self.rank, self.world, n_rows, windows)
            self._windows[id(plan)] = windows
            csr = dv.csr_of_transpose(a, self.dev, row_range=(self.row0, self.row1), col_offset=plan.win0)   # window-relative columns
            self.ops.append(csr)
            self.plans.append(plan)
        # halo exchange: one slab all-gather (capturable by a CUDA graph) when every halo row sits within H rows of a block
        # boundary -- the banded FEM case; point-to-point messages otherwise
        self.owned = owner_ranges(n_rows, self.world)
        self.halo_h = 0
        self.halo_mode = "none"
        if self.world > 1:
            hs = []
            for a, plan in zip(operators, self.plans):
                if plan is None:
                    continue
                hs.append(slab_halo_rows(self._windows[id(plan)], self.owned))
            if any(h is None for h in hs):
                self.halo_mode = "p2p"
            else:
                self.halo_h = max(hs) if hs else 0
                self.halo_mode = "allgather"
        # operators with one sparsity pattern (Ct and Tt of a FEM model) are multiplied in one pass (device.spmm2)
        live = [i for i, c in enumerate(self.ops) if c is not None]
        self.pair = None
        if len(live) == 2 and self.plans[live[0]].win0 == self.plans[live[1]].win0 and self.plans[live[0]].win1 == self.plans[live[1]].win1:
            i0, i1 = live
            if dv.mark_same_pattern(self.ops[i0], self.ops[i1], operators[i0], operators[i1], row_range=(self.row0, self.row1)):
                self.pair = (i0, i1)
        self.b = dv.csc_to_device(b, self.dev)
        self.coeffs = [dv.upload(np.ascontiguousarray(c[self.f0:self.f1], dtype=np.float64), self.dev) for c in coeffs_global]
        self._win = None
        self.stage_events = None     # bench.py: set to a list to collect per-stage CUDA events
        self._graph = None           # (CUDAGraph, static input, static outputs, pinned flag buffer) of step_graph
        self.launches_per_graph = 0
        self._pending = []           # device flag blocks of step_deferred calls awaiting verify_deferred


    @property
    def graph_capable(self) -> bool:
        """True when ``step_graph`` can capture this rank's step (single rank, or a multi-rank step whose exchanges are
        all capturable NCCL collectives)."""
        return self.world == 1 or self.halo_mode == "allgather"

    def drop_graph(self) -> None:
        """Forget the captured step (its static outputs stay valid for whoever still references them)."""
        self._graph = None

    def _exchange(self, x: torch.Tensor, plan: HaloPlan, buf: Optional[torch.Tensor], group) -> torch.Tensor:
        if self.halo_mode == "allgather":
            return exchange_halo_slabs(x, plan, self.halo_h, self.owned, buf, group)
        return exchange_halo(x, plan, buf, group)

    def _mark(self, ev):
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append(e)

    def step_graph(self, s_local: torch.Tensor, want_x: bool = False, skip_sweep: bool = False):
        """``step`` replayed from a CUDA graph (single rank): the whole launch sequence -- optimistic CholeskyQR2 (two
        passes, success flags kept on the device), SpMMs, contractions, SVD rotation on its side stream, sweep -- is
        captured once per snapshot-block shape and replayed with one launch, which removes the host launch gaps between
        the ~45 short kernels of a step.  Call ``verify()`` before trusting the results: if a Cholesky broke down or the
        block needed a third pass it re-runs the adaptive ``step``.  Outputs are static tensors overwritten by the next
        replay.  ``skip_sweep`` captures stages 1 + 2 only (bench.py times the basis + projection stage with it)."""
        if not self.graph_capable:
            raise RuntimeError("step_graph: this rank's halo exchange needs point-to-point messages, which a CUDA graph cannot "
                               "capture; use step() / step_deferred()")
        gather = self.world > 1 and not skip_sweep
        key = (tuple(s_local.shape), s_local.dtype, bool(want_x), bool(skip_sweep))
        if self._graph is None or self._graph["key"] != key:
            static_s = s_local.clone()
            for _ in range(2):                                   # warm-up: workspaces, function attributes, side stream, NCCL
                self.step(static_s, want_x=want_x, gather=gather, optimistic=True, skip_sweep=skip_sweep)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)
                torch.cuda.synchronize()
            from . import _ffi
            graph = torch.cuda.CUDAGraph()
            launches0 = _ffi.launch_count()
            # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                out = self.step(static_s, want_x=want_x, gather=gather, optimistic=True, skip_sweep=skip_sweep)
            self.launches_per_graph = _ffi.launch_count() - launches0     # kernels of this library inside one replay
            self.dv.workspaces.pin_all()           # the replay writes into the workspace buffers captured here
            flags_host = torch.empty((2, 32), dtype=torch.uint8).pin_memory()
            self._graph = {"key": key, "graph": graph, "s": static_s, "out": out, "flags_host": flags_host, "want_x": want_x}
        gr = self._graph
        if s_local.data_ptr() != gr["s"].data_ptr():
            gr["s"].copy_(s_local)
        gr["graph"].replay()
        gr["flags_host"].copy_(gr["out"][4].flags, non_blocking=True)
        return gr["out"][:4]

    def step_deferred(self, s_local: torch.Tensor, want_x: bool = False, gather: bool = True):
        """Eager ``step`` with the optimistic CholeskyQR2 (no device->host read inside the step, so the launch queue never
        drains): the success flags of every call are kept and checked by ``verify_deferred()``.  Works under torchrun."""
        out = self.step(s_local, want_x=want_x, gather=gather, optimistic=True)
        self._pending.append(out[4].flags)
        return out[:4]

    def verify_deferred(self) -> bool:
        """Synchronise and check the flags of all ``step_deferred`` calls since the last check (all ranks agree: the
        Gram matrices are all-reduced, every rank factors the same matrix).  False = redo those steps with ``step``."""
        ok = all(self.dv.flags_ok(f.cpu()) for f in self._pending)
        self._pending = []
        return ok

    def verify(self):
        """Synchronise and check the deferred flags of the last ``step_graph``; on failure return the results of the
        adaptive path instead (None when the optimistic results stand)."""
        gr = self._graph
        if gr is None:
            return None
        torch.cuda.current_stream().synchronize()
        if self.dv.flags_ok(gr["flags_host"]):
            return None
        return self.step(gr["s"], want_x=gr["want_x"], gather=self.world > 1)

    def step(self, s_local: torch.Tensor, want_x: bool = False, gather: bool = True, optimistic: bool = False, skip_sweep: bool = False):
        """One pass of the hot path.  Returns (gsm_all or gsm_local, q_local, (a0_r, a1_r, a2_r, b_r), sweep result)
        (plus the BasisInfo when ``optimistic``: its ``flags`` still have to be verified, see ``step_graph``)."""
        dv = self.dv
        group = self.group if self.world > 1 else None
        ev = [] if self.stage_events is not None else None
        self._mark(ev)
        def project_block(x):
            """x^T (A_i x) and x^T b for the un-rotated Cholesky-QR block.  All row partials land in ONE flat buffer that is
            all-reduced once; operators with the same halo plan (same sparsity pattern) share one halo exchange."""
            r = x.shape[1]
            m = self.b.ncols
            live = [i for i, csr in enumerate(self.ops) if csr is not None]
            nlive = len(live)
            flat = torch.empty(nlive * r * r + r * m, dtype=x.dtype, device=x.device)
            g_list = [None] * len(self.ops)
            windows = {}
            if self.pair is not None and r <= dv.SPMM2_MAX_R:    # one pass over the shared pattern for both operators
                i0, i1 = self.pair
                plan = self.plans[i0]
                xin = x
                if self.world > 1:
                    self._win = self._exchange(x, plan, self._win, group)
                    xin = self._win[:plan.win1 - plan.win0]
                y0, y1 = dv.spmm2(self.ops[i0], self.ops[i1], xin)
                for k, (i, y) in enumerate(((i0, y0), (i1, y1))):
                    g_list[i] = dv.gemm_tn(y, x, conj=False, out=flat[k * r * r:(k + 1) * r * r].view(r, r))
                live = []
            for k, i in enumerate(live):
                csr, plan = self.ops[i], self.plans[i]
                dv.group_rows(csr, r, real=x.dtype == torch.float64)   # row-grouped operand, built once per operator (and element type)
                if self.world > 1:
                    key = (plan.win0, plan.win1, tuple(plan.send), tuple(plan.recv))
                    if key not in windows:
                        windows[key] = self._exchange(x, plan, self._win if not windows else None, group)
                        if len(windows) == 1:
                            self._win = windows[key]
                    y = dv.spmm(csr, windows[key][:plan.win1 - plan.win0])
                else:
                    y = dv.spmm(csr, x)
                g_list[i] = dv.gemm_tn(y, x, conj=False, out=flat[k * r * r:(k + 1) * r * r].view(r, r))
            bt = flat[nlive * r * r:].view(r, m)
            bt.copy_(dv.project_rhs(self.b, x, self.row0, conj=False))
            allreduce_sum_(flat, group)
            return g_list, bt

        # the sweep only needs the reduced model: it is launched before the stream is joined with the tall product q = x w
        q, reduced, b_r, info = dv.basis_and_projection(s_local, project_block, group=group, optimistic=optimistic, defer_q=True)
        sym = [None if o is None else dv.symmetrize(o) for o in reduced]
        c0, c1, c2, cb, zs = self.coeffs

        def join_q():
            if info.q_ready is not None:
                torch.cuda.current_stream().wait_event(info.q_ready)
                info.q_ready = None
            info.keepalive = None        # the side stream has been joined: the un-rotated block may be reused now

        if skip_sweep:                       # stages 1 + 2 only (timing of the basis + projection stage)
            join_q()
            self._mark(ev)
            out = (None, q, (reduced[0], reduced[1], reduced[2], b_r), None)
            return out + (info,) if optimistic else out
        self._mark(ev)
        res = dv.sweep(sym[0], sym[1], sym[2], b_r, c0, c1, c2, cb, zs, want_x=want_x, want_gsm=True)
        join_q()
        self._mark(ev)
        gsm = gather_points(res.gsm, self.f_total, group) if (gather and self.world > 1) else res.gsm
        self._mark(ev)
        if ev is not None:
            self.stage_events.append(ev)
        if optimistic:
            return gsm, q, (reduced[0], reduced[1], reduced[2], b_r), res, info
        return gsm, q, (reduced[0], reduced[1], reduced[2], b_r), res
