"""Multi-GPU path: column-block-cyclic assembly, LU and solves, one process per GPU.

Layout.  The (N+M) x (N+M) collocation matrix is split into column blocks of width ``nb``; block j
lives on rank ``j % world`` (a 1 x Q process grid of the 2-D block-cyclic family: every rank holds
all rows of its column blocks, row-major, padded leading dimension).  Choosing Q-only sharding is
deliberate for NVSwitch: partial pivoting then searches whole columns on ONE GPU (the cooperative
register-resident panel kernel, no per-column cross-GPU reduction -- 250 000 of them would be
latency-bound), and the only exchange per panel is one broadcast of the factored panel, which at
~770 GB/s per direction is far below the update's compute time.  Per panel k:

  owner(k+1) first applies panel k to the columns of block k+1 only, factors that block, posts its
  broadcast, and only then applies panel k to its remaining columns (look-ahead: the other ranks
  receive panel k+1 while they are still busy with update k);
  every rank: row interchanges on its local columns, U12 = L11^-1 A12 (recursive TRSM against the
  received panel), Schur update A22 -= L21 U12 (DMMA GEMM with the A operand TMA-loaded from the
  panel buffer).

Assembly needs no communication: each rank fills exactly the column blocks it owns.

The numerical kernels are reached through a small backend interface; the product backend
(``CudaBackend``) calls the C-ABI of include/updes_b200.h.  tests/ supplies a numpy backend so the
host logic (ownership maps, look-ahead order, pivot plumbing, solves) runs under gloo on CPUs.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .assembly import padded_ld
from .rbf import RBF_CODES


class ColumnBlockCyclic:
    """Index maps of the 1 x Q block-cyclic column distribution."""

    def __init__(self, n: int, nb: int, world: int):
        assert nb % 32 == 0, "block width must be a multiple of 32 (panel / TRSM / GEMM tiles)"
        self.n, self.nb, self.world = n, nb, world
        self.nblocks = (n + nb - 1) // nb

    def owner(self, j: int) -> int:
        return j % self.world

    def width(self, j: int) -> int:
        return min(self.nb, self.n - j * self.nb)

    def local_offset(self, j: int) -> int:
        """Column offset of global block j inside its owner's local matrix."""
        return (j // self.world) * self.nb

    def local_blocks(self, rank: int):
        return list(range(rank, self.nblocks, self.world))

    def local_cols(self, rank: int) -> int:
        return sum(self.width(j) for j in self.local_blocks(rank))

    def first_local_block_after(self, rank: int, j: int):
        """Smallest global block index > j owned by rank (or None)."""
        k = j + 1 + ((rank - (j + 1)) % self.world)
        return k if k < self.nblocks else None

    def local_offset_after(self, rank: int, j: int) -> int:
        """Local column offset where the blocks with global index > j start on `rank`."""
        k = self.first_local_block_after(rank, j)
        return self.local_cols(rank) if k is None else self.local_offset(k)


class CudaBackend:
    """The product backend: local matrix + two panel buffers in HBM, kernels through the C-ABI."""

    def __init__(self, layout: ColumnBlockCyclic, rank: int, gemm_sms_reserved: int = 0):
        self.torch = torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.layout, self.rank = layout, rank
        n, nb = layout.n, layout.nb
        self.n, self.nb = n, nb
        self.cols = layout.local_cols(rank)
        self.ld = max(padded_ld(self.cols), 16)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.local = torch.zeros((n, self.ld), dtype=torch.float64, device=dev)
        # panel buffer: [n][nb] panel rows followed by nb pivot slots (as float64), one flat tensor
        self.bufs = [torch.zeros(n * nb + nb, dtype=torch.float64, device=dev) for _ in range(2)]
        self.ipiv = torch.zeros(n, dtype=torch.int32, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        h = ctypes.c_void_p()
        _lib.check(self.lib.updes_lu_create(ctypes.byref(h), n, self.ld), "updes_lu_create")
        self.h = h
        _lib.check(self.lib.updes_lu_bind(h, 0, self.local.data_ptr(), n, self.ld), "updes_lu_bind")
        for s, b in enumerate(self.bufs):
            _lib.check(self.lib.updes_lu_bind(h, 1 + s, b.data_ptr(), n, nb), "updes_lu_bind")
        if gemm_sms_reserved > 0:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            _lib.check(self.lib.updes_lu_set_gemm_ctas(h, max(sms - gemm_sms_reserved, 1)), "updes_lu_set_gemm_ctas")

    # ---- assembly ---------------------------------------------------------------------------------
    def assemble(self, rows, kind, param, M):
        """Fill the owned column blocks (updes_assemble_block); no communication."""
        N = rows.N
        Ni = rows.table.Ni
        code = RBF_CODES[kind]
        st = _lib.stream_ptr()
        ranges = [(0, Ni, rows.mask_internal), (Ni, N - Ni, rows.mask_boundary), (N, M, 7)]
        for j in self.layout.local_blocks(self.rank):
            c0, w, lc = j * self.nb, self.layout.width(j), self.layout.local_offset(j)
            for r0, nr, mask in ranges:
                if nr <= 0:
                    continue
                out = self.local.data_ptr() + 8 * (r0 * self.ld + lc)
                rc = self.lib.updes_assemble_block(code, float(param), N, M, rows.centres.data_ptr(),
                                                   ctypes.byref(rows.struct), r0, nr, c0, w, mask, out, self.ld, st)
                _lib.check(rc, "updes_assemble_block")

    # ---- factorisation building blocks ------------------------------------------------------------------
    def panel_factor(self, r0, lc, w):
        rc = self.lib.updes_lu_panel_factor(self.h, 0, r0, lc, w, self.ipiv.data_ptr(), self.info.data_ptr(),
                                            _lib.stream_ptr())
        _lib.check(rc, "updes_lu_panel_factor")

    def pack(self, slot, r0, lc, w):
        buf = self.bufs[slot]
        panel = buf[: self.n * self.nb].view(self.n, self.nb)
        panel[r0:, :w].copy_(self.local[r0:, lc:lc + w])
        buf[self.n * self.nb: self.n * self.nb + w].copy_(self.ipiv[r0:r0 + w])

    def unpack_pivots(self, slot, r0, w):
        buf = self.bufs[slot]
        self.ipiv[r0:r0 + w].copy_(buf[self.n * self.nb: self.n * self.nb + w])

    def message(self, slot, r0):
        """Contiguous slice that travels: panel rows r0.. and the pivot tail."""
        return self.bufs[slot][r0 * self.nb:]

    def apply_swaps(self, r0, w, c_lo, c_hi):
        if c_hi <= c_lo:
            return
        rc = self.lib.updes_lu_apply_swaps(self.h, 0, c_lo, c_hi, r0, w, self.ipiv.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_lu_apply_swaps")

    def apply_panel(self, slot, r0, w, c_lo, c_hi):
        """Panel (rows r0.., width w, in buffer `slot`) applied to local columns [c_lo, c_hi)."""
        if c_hi <= c_lo:
            return
        st = _lib.stream_ptr()
        self.apply_swaps(r0, w, c_lo, c_hi)
        rc = self.lib.updes_lu_trsm(self.h, 1 + slot, r0, 0, w, 0, r0, c_lo, c_hi - c_lo, st)
        _lib.check(rc, "updes_lu_trsm")
        m = self.n - (r0 + w)
        if m > 0:
            rc = self.lib.updes_lu_gemm(self.h, 1 + slot, r0 + w, 0, 0, r0, c_lo, 0, r0 + w, c_lo, m, c_hi - c_lo, w, st)
            _lib.check(rc, "updes_lu_gemm")

    # ---- solve building blocks ----------------------------------------------------------------------------
    def set_pivots(self):
        _lib.check(self.lib.updes_lu_set_pivots(self.h, self.ipiv.data_ptr(), _lib.stream_ptr()), "updes_lu_set_pivots")

    def permute_rhs(self, b):
        x = self.torch.empty_like(b)
        rc = self.lib.updes_lu_permute_rhs(self.h, b.data_ptr(), b.shape[-1], 1, x.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_lu_permute_rhs")
        return x

    def block_sweep(self, upper, r0, lc, w, x):
        rc = self.lib.updes_tri_block_sweep(self.h, 0, 1 if upper else 0, r0, lc, w, x.data_ptr(), 1, _lib.stream_ptr())
        _lib.check(rc, "updes_tri_block_sweep")

    def block_gemv(self, r0, w, c_lo, c_hi, xl, out):
        """out[:w] = local[r0:r0+w, c_lo:c_hi] @ xl[c_lo:c_hi] (xl indexed by local column); zeros for an empty range."""
        if c_hi <= c_lo:
            out.zero_()
            return
        rc = self.lib.updes_block_gemv(self.h, 0, r0, w, c_lo, c_hi, xl.data_ptr(), out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_block_gemv")

    def diag_solve(self, upper, r0, lc, w, x):
        """Triangular solve with the diagonal block of global block (rows r0.., local columns lc..) in place on x[r0:r0+w]."""
        rc = self.lib.updes_tri_diag_solve(self.h, 0, 1 if upper else 0, r0, lc, w, x.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_tri_diag_solve")

    def zeros(self, k):
        return self.torch.zeros(k, dtype=self.torch.float64, device=self.device)

    def vector(self, host_array):
        return self.torch.as_tensor(np.ascontiguousarray(host_array), dtype=self.torch.float64).to(self.device)

    def zero_pivot(self):
        return int(self.info.item())

    def nbytes(self):
        return (self.local.numel() + sum(b.numel() for b in self.bufs)) * 8

    def check_sweeps(self):
        """Raise if a triangular sweep of this rank timed out waiting for a solved block."""
        flags = ctypes.c_int32(0)
        _lib.check(self.lib.updes_lu_status(self.h, ctypes.byref(flags), _lib.stream_ptr()), "updes_lu_status")
        if flags.value:
            raise RuntimeError("updes_b200: a triangular sweep timed out waiting for a solved block (status %d)" % flags.value)

    # ---- row equilibration (rows are spread over all ranks: the maxima are combined by the caller) ---------
    def row_absmax(self):
        out = self.torch.empty(self.n, dtype=self.torch.float64, device=self.device)
        rc = self.lib.updes_row_absmax(self.local.data_ptr(), self.n, self.cols, self.ld, out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_row_absmax")
        return out

    def apply_row_scale(self, absmax):
        """Scale the local columns by 2^-floor(log2 absmax) per row and make the solves scale right-hand sides."""
        self.scale = self.torch.empty_like(absmax)
        st = _lib.stream_ptr()
        _lib.check(self.lib.updes_scale_from_absmax(absmax.data_ptr(), self.n, self.scale.data_ptr(), st), "updes_scale_from_absmax")
        _lib.check(self.lib.updes_row_scale(self.local.data_ptr(), self.n, self.cols, self.ld, self.scale.data_ptr(), st), "updes_row_scale")
        _lib.check(self.lib.updes_lu_set_row_scale(self.h, self.scale.data_ptr()), "updes_lu_set_row_scale")

    def event(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def close(self):
        if getattr(self, "h", None):
            self.lib.updes_lu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistributedLU:
    """Column-block-cyclic LU with partial pivoting and look-ahead of one panel."""

    def __init__(self, layout: ColumnBlockCyclic, rank: int, backend, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.layout, self.rank, self.be, self.group = layout, rank, backend, group
        self.factored = False
        self.timeline = None          # set to [] to record per-panel events (bench.py's critical-path breakdown)
        self.solve_variant = "left"   # "left": left-looking substitution (default); "right": first-generation column sweep

    def _src(self, r):
        """`layout` ranks are ranks of `group`; torch.distributed.broadcast wants the global rank."""
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _bcast(self, slot, r0, src, async_op):
        return self.dist.broadcast(self.be.message(slot, r0), src=self._src(src), group=self.group, async_op=async_op)

    def equilibrate(self):
        """Row equilibration before pivoting: per-row max |entry| over ALL ranks' columns (all-reduce MAX of n
        doubles), then every rank scales its columns by the same exact power of two."""
        m = self.be.row_absmax()
        self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX, group=self.group)
        self.be.apply_row_scale(m)
        return self

    def factor(self):
        L, be, me = self.layout, self.be, self.rank
        nb = L.nb
        ncols_local = L.local_cols(me)
        tl = self.timeline
        mark = (lambda: be.event()) if tl is not None else (lambda: None)
        # panel 0
        if L.owner(0) == me:
            be.panel_factor(0, 0, L.width(0))
            be.pack(0, 0, 0, L.width(0))
        pending = self._bcast(0, 0, L.owner(0), async_op=True)
        for k in range(L.nblocks):
            slot, r0, w = k % 2, k * nb, L.width(k)
            t_begin = mark()
            pending.wait()                       # panel k (and its pivots) is in bufs[slot]
            t_have = mark()
            t_look = t_fact = None
            if L.owner(k) != me:
                be.unpack_pivots(slot, r0, w)
            # interchanges on the already-factored columns to the left (L stored in final row order)
            left_end = L.local_offset(k) if L.owner(k) == me else L.local_offset_after(me, k)
            be.apply_swaps(r0, w, 0, left_end)
            right0 = L.local_offset_after(me, k)  # local columns of the blocks with global index > k
            nxt = k + 1
            if nxt < L.nblocks:
                r1, w1 = nxt * nb, L.width(nxt)
                if L.owner(nxt) == me:
                    lc1 = L.local_offset(nxt)
                    be.apply_panel(slot, r0, w, lc1, lc1 + w1)          # look-ahead: next panel's columns first
                    t_look = mark()
                    be.panel_factor(r1, lc1, w1)
                    be.pack(1 - slot, r1, lc1, w1)
                    t_fact = mark()
                    pending = self._bcast(1 - slot, r1, me, async_op=True)
                    be.apply_panel(slot, r0, w, lc1 + w1, ncols_local)  # the rest of update k
                else:
                    pending = self._bcast(1 - slot, r1, L.owner(nxt), async_op=True)
                    be.apply_panel(slot, r0, w, right0, ncols_local)
            # last panel: nothing to the right
            if tl is not None:
                tl.append((k, t_begin, t_have, t_look, t_fact, mark()))
        be.set_pivots()
        self.factored = True
        return self

    def timeline_ms(self):
        """Per-panel durations on this rank (ms): wait for the panel broadcast, look-ahead update of the next panel's
        columns, its factorisation + packing (owner only), the rest of the trailing update."""
        out = []
        for k, t0, t1, t2, t3, t4 in self.timeline or []:
            rec = {"k": k, "wait": t0.elapsed_time(t1)}
            if t2 is not None:
                rec.update(lookahead=t1.elapsed_time(t2), panel=t2.elapsed_time(t3), update=t3.elapsed_time(t4))
            else:
                rec.update(update=t1.elapsed_time(t4))
            out.append(rec)
        return out

    def solve(self, b_host):
        """Solve K x = b (b: length-n host array or device vector, identical on all ranks) -> x replicated on every rank.

        Left-looking block substitution.  Row block j of L (or U) is spread over all ranks by columns, and each rank
        only ever multiplies it with the solution entries of the blocks it owns and solved itself:
            every rank:  p_r = L[rows j, my columns of blocks < j] . y[those]     (contiguous row segments, HBM rate,
                                                                                   all ranks in parallel)
            reduce p_r -> owner(j)   (width doubles);   owner:  y_j = L_jj^-1 (b_j - sum_r p_r)
        so the chain per block is one small reduction + one diagonal-block solve, and the O(n^2) reads are shared by
        all GPUs.  (The first version swept column blocks on the owner alone and broadcast the whole running
        right-hand side after every block: serial in the matrix reads, 140 ms of a 2.4 s step at 8 GPUs.)"""
        assert self.factored
        if self.solve_variant == "right":
            return self.solve_right_looking(b_host)
        L, be, me, dist = self.layout, self.be, self.rank, self.dist
        nb, n = L.nb, L.n
        ncols = L.local_cols(me)
        x = be.permute_rhs(b_host if hasattr(b_host, "data_ptr") else be.vector(b_host))
        xl = be.zeros(max(ncols, 2))            # solved entries of MY blocks, indexed by local column
        part = be.zeros(nb)
        SUM = dist.ReduceOp.SUM
        for j in range(L.nblocks):                                   # forward, unit lower
            r0, w, o = j * nb, L.width(j), L.owner(j)
            lc = L.local_offset(j) if o == me else None
            hi = lc if o == me else L.local_offset_after(me, j)      # my columns of the blocks with global index < j
            p = part[:w]
            be.block_gemv(r0, w, 0, hi, xl, p)
            dist.reduce(p, dst=self._src(o), op=SUM, group=self.group)
            if o == me:
                x[r0:r0 + w].sub_(p)
                be.diag_solve(False, r0, lc, w, x)
                xl[lc:lc + w].copy_(x[r0:r0 + w])
        for j in range(L.nblocks - 1, -1, -1):                       # backward, upper
            r0, w, o = j * nb, L.width(j), L.owner(j)
            lc = L.local_offset(j) if o == me else None
            lo = lc + w if o == me else L.local_offset_after(me, j)  # my columns of the blocks with global index > j
            p = part[:w]
            be.block_gemv(r0, w, lo, ncols, xl, p)
            dist.reduce(p, dst=self._src(o), op=SUM, group=self.group)
            if o == me:
                x[r0:r0 + w].copy_(xl[lc:lc + w])
                x[r0:r0 + w].sub_(p)
                be.diag_solve(True, r0, lc, w, x)
                xl[lc:lc + w].copy_(x[r0:r0 + w])
        out = be.zeros(n)                                            # every rank contributes the blocks it solved
        for j in L.local_blocks(me):
            r0, w, lc = j * nb, L.width(j), L.local_offset(j)
            out[r0:r0 + w].copy_(xl[lc:lc + w])
        dist.all_reduce(out, op=SUM, group=self.group)
        return out

    def solve_right_looking(self, b_host):
        """First-generation solve, kept for comparison (bench.py --dist-solve right) and as a second opinion in the
        tests: column-sweep substitution; the owner of block j finishes x_j, subtracts its block's contribution from
        the remaining rows and broadcasts them."""
        assert self.factored
        L, be, me, dist = self.layout, self.be, self.rank, self.dist
        nb, n = L.nb, L.n
        x = be.permute_rhs(b_host if hasattr(b_host, "data_ptr") else be.vector(b_host))
        for j in range(L.nblocks):                                   # forward, unit lower
            r0, w = j * nb, L.width(j)
            if L.owner(j) == me:
                be.block_sweep(False, r0, L.local_offset(j), w, x)
            dist.broadcast(x[r0:], src=self._src(L.owner(j)), group=self.group)
        for j in range(L.nblocks - 1, -1, -1):                       # backward, upper
            r0, w = j * nb, L.width(j)
            if L.owner(j) == me:
                be.block_sweep(True, r0, L.local_offset(j), w, x)
            dist.broadcast(x[:r0 + w], src=self._src(L.owner(j)), group=self.group)
        return x
