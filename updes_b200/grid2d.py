"""P x Q (2-D) block-cyclic variant of the multi-GPU LU ("2D block-cyclic P x Q (1x2, 2x2, 2x4)").

STATUS: (i) host logic -- index maps, panel gather / scatter, row exchange, broadcasts, solves -- verified under gloo on
CPUs (tests/test_grid2d_cpu.py: factors, pivots and solutions equal LAPACK's on 2x2, 2x3, 3x2, 1xQ and Px1 grids);
(ii) every CUDA wrapper of ``CudaKernels2D`` verified on a B200 on the 1 x 1 grid (tools/grid2d_gpu_check.py,
profiles/r02_grid2d_gpu_check_1x1.log: tiles bit-identical to the single-GPU assembly, pivots identical to the
single-GPU LU, backward error 4e-18); (iii) NOT yet run on several GPUs and NOT timed (the round's GPU budget ended).
It is opt-in (``enable_distributed(grid=(P, Q))``); the default and measured layout stays 1 x Q, for the reason
DESIGN.md 3.4 gives (on NVSwitch the panel broadcast volume that P x Q saves is not what binds).

Layout.  nb x nb blocks; block (I, J) lives on the process (I % P, J % Q), rank = p * Q + q.  Every rank stores
its blocks as one row-major local matrix with whole blocks (the ragged last block is padded with zeros): local row
of global row r = (r // nb // P) * nb + r % nb, same for columns.

Per panel k (block column k on process column qk = k % Q, diagonal block on process row pk = k % P):

  1. the P ranks of process column qk send their pieces of the panel to the diagonal owner d = (pk, qk), which
     factors the whole (n - r0) x nb panel with the single-GPU register-resident panel kernels -- pivot search stays
     on ONE GPU (no per-column cross-GPU arg-max: 90 000 - 250 000 latency-bound reductions otherwise);
  2. d returns to every rank of its process column the rows that rank owns (+ the pivots); each of them broadcasts
     its rows ALONG ITS PROCESS ROW.  A rank therefore receives (n / P) x nb panel entries instead of n x nb;
  3. row interchanges: the pivot list is composed into one permutation of the <= 2 nb involved rows; inside every
     process column the involved rows are exchanged with one all-reduce of a (rows x local columns) staging buffer
     (each row is contributed by exactly one rank);
  4. process row pk computes U12 = L11^-1 A12 for its columns and broadcasts it DOWN ITS PROCESS COLUMN;
  5. every rank updates its trailing blocks with one DMMA GEMM:  local -= L_rows(mine) . U12_cols(mine).

Solve: left-looking block substitution.  Row block j lives on process row j % P spread over Q ranks by columns:
each forms its partial product with the solution entries of its own columns, one reduction along the process row
to the diagonal owner, which solves the diagonal block and broadcasts x_j down its process column.

The numerical kernels are reached through a small interface (``CudaKernels2D``: the C-ABI of
include/updes_b200.h on torch CUDA tensors; tests/numpy_kernels2d.py: numpy on CPU tensors for gloo).  Everything
else -- index maps, staging, communication -- is the same code on both.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .rbf import RBF_CODES


class BlockCyclic2D:
    """Index maps of the P x Q block-cyclic distribution (rank = p * Q + q)."""

    def __init__(self, n: int, nb: int, P: int, Q: int):
        if nb % 32:
            raise ValueError("block size must be a multiple of 32 (panel / TRSM / GEMM tiles)")
        if P < 1 or Q < 1:
            raise ValueError("process grid must be at least 1 x 1")
        self.n, self.nb, self.P, self.Q = n, nb, P, Q
        self.world = P * Q
        self.nblocks = (n + nb - 1) // nb
        if self.nblocks < max(P, Q):
            raise ValueError("matrix of %d blocks is too small for a %d x %d process grid" % (self.nblocks, P, Q))

    def coords(self, rank: int):
        return divmod(rank, self.Q)

    def rank_of(self, p: int, q: int) -> int:
        return p * self.Q + q

    def width(self, j: int) -> int:
        return min(self.nb, self.n - j * self.nb)

    def row_blocks(self, p: int):
        return list(range(p, self.nblocks, self.P))

    def col_blocks(self, q: int):
        return list(range(q, self.nblocks, self.Q))

    def local_rows(self, p: int) -> int:
        """Rows of the local matrix of process row p (whole blocks, padding included)."""
        return len(range(p, self.nblocks, self.P)) * self.nb

    def local_cols(self, q: int) -> int:
        return len(range(q, self.nblocks, self.Q)) * self.nb

    def _pad(self) -> int:
        return self.nb - self.width(self.nblocks - 1)

    def valid_rows(self, p: int) -> int:
        """Local rows that hold matrix rows (the padding of the ragged last block excluded)."""
        return self.local_rows(p) - (self._pad() if (self.nblocks - 1) % self.P == p else 0)

    def valid_cols(self, q: int) -> int:
        return self.local_cols(q) - (self._pad() if (self.nblocks - 1) % self.Q == q else 0)

    def lrow(self, I: int) -> int:
        """Local row offset of global row block I on its owner."""
        return (I // self.P) * self.nb

    def lcol(self, J: int) -> int:
        return (J // self.Q) * self.nb

    @staticmethod
    def _first_at_or_after(idx: int, stride: int, k: int) -> int:
        return k + ((idx - k) % stride)

    def first_row_block_from(self, p: int, k: int) -> int:
        """Smallest global block index >= k owned by process row p (may be >= nblocks: none)."""
        return self._first_at_or_after(p, self.P, k)

    def lrow_from(self, p: int, k: int) -> int:
        """Local row offset where the row blocks with global index >= k start on process row p."""
        return min((self.first_row_block_from(p, k) // self.P) * self.nb, self.local_rows(p))

    def lcol_from(self, q: int, k: int) -> int:
        return min((self._first_at_or_after(q, self.Q, k) // self.Q) * self.nb, self.local_cols(q))

    def row_owner(self, r: int) -> int:
        return (r // self.nb) % self.P

    def local_row_of(self, r: int) -> int:
        return (r // self.nb // self.P) * self.nb + r % self.nb


def compose_interchanges(r0: int, piv):
    """LAPACK-style interchanges (r0 + t <-> piv[t], t = 0, 1, ...), applied in order, as ONE permutation:
    returns {destination row: source row} for the rows whose content changes (source = the row, BEFORE any
    interchange, whose content ends up at the destination)."""
    content = {}
    for t, p in enumerate(piv):
        a, b = r0 + t, int(p)
        if a == b:
            continue
        ca, cb = content.get(a, a), content.get(b, b)
        content[a], content[b] = cb, ca
    return {d: s for d, s in content.items() if d != s}


def permutation_from_pivots(ipiv: np.ndarray) -> np.ndarray:
    """perm with (P b)[i] = b[perm[i]] for the interchange list ipiv (row k <-> ipiv[k], k ascending)."""
    perm = np.arange(len(ipiv))
    for k, p in enumerate(ipiv.tolist()):
        if p != k:
            perm[k], perm[p] = perm[p], perm[k]
    return perm


class CudaKernels2D:
    """The product kernels: C-ABI of include/updes_b200.h on four bound buffers (slot 0 local matrix, 1 gathered
    panel, 2 this rank's rows of the factored panel, 3 the received U12 block row)."""

    SLOT = {"local": 0, "G": 1, "L": 2, "U": 3}

    def __init__(self):
        self.torch = torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.h = None
        self.t = {}

    def bind(self, n_padded, local, local_rows, G, G_rows, L, U):
        """local: (mloc, ld) with `local_rows` real rows; G: (n_padded, nb) with G_rows = n real rows; L: (mloc, nb);
        U: (nb, ld)."""
        h = ctypes.c_void_p()
        _lib.check(self.lib.updes_lu_create(ctypes.byref(h), n_padded, local.shape[1]), "updes_lu_create")
        self.h = h
        self.t = {"local": local, "G": G, "L": L, "U": U}
        for name, t, rows in (("local", local, local_rows), ("G", G, G_rows), ("L", L, L.shape[0]), ("U", U, U.shape[0])):
            _lib.check(self.lib.updes_lu_bind(h, self.SLOT[name], t.data_ptr(), max(int(rows), 1), t.shape[1]), "updes_lu_bind")

    def panel_factor(self, r0, w, ipiv, info):
        rc = self.lib.updes_lu_panel_factor(self.h, 1, r0, 0, w, ipiv.data_ptr(), info.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_lu_panel_factor")

    def trsm(self, lr, w, rb, cb, ncols):
        rc = self.lib.updes_lu_trsm(self.h, 2, lr, 0, w, 0, rb, cb, ncols, _lib.stream_ptr())
        _lib.check(rc, "updes_lu_trsm")

    def gemm(self, ra, b_name, rb, cb, rc_, cc, m, n, k):
        rc = self.lib.updes_lu_gemm(self.h, 2, ra, 0, self.SLOT[b_name], rb, cb, 0, rc_, cc, m, n, k, _lib.stream_ptr())
        _lib.check(rc, "updes_lu_gemm")

    def gemv(self, r0, w, c_lo, c_hi, xl, out):
        if c_hi <= c_lo:
            out.zero_()
            return
        rc = self.lib.updes_block_gemv(self.h, 0, r0, w, c_lo, c_hi, xl.data_ptr(), out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_block_gemv")

    def diag_solve(self, upper, r0, c0, w, xt):
        rc = self.lib.updes_tri_diag_solve(self.h, 0, 1 if upper else 0, r0, c0, w, xt.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "updes_tri_diag_solve")

    def row_absmax(self, cols):
        A = self.t["local"]
        out = self.torch.zeros(A.shape[0], dtype=self.torch.float64, device=self.device)
        if cols > 0:
            _lib.check(self.lib.updes_row_absmax(A.data_ptr(), A.shape[0], cols, A.shape[1], out.data_ptr(), _lib.stream_ptr()),
                       "updes_row_absmax")
        return out

    def scale_from_absmax(self, absmax):
        scale = self.torch.empty_like(absmax)
        _lib.check(self.lib.updes_scale_from_absmax(absmax.data_ptr(), absmax.numel(), scale.data_ptr(), _lib.stream_ptr()),
                   "updes_scale_from_absmax")
        return scale

    def row_scale(self, scale, cols):
        A = self.t["local"]
        if cols > 0:
            _lib.check(self.lib.updes_row_scale(A.data_ptr(), A.shape[0], cols, A.shape[1], scale.data_ptr(), _lib.stream_ptr()),
                       "updes_row_scale")

    def assemble_tile(self, rows, code, param, N, M, r0, nr, c0, w, mask, lr, lc):
        """Entries [r0, r0+nr) x [c0, c0+w) of the collocation matrix into local[lr.., lc..] (updes_assemble_block)."""
        A = self.t["local"]
        out = A.data_ptr() + 8 * (lr * A.shape[1] + lc)
        rc = self.lib.updes_assemble_block(code, float(param), N, M, rows.centres.data_ptr(), ctypes.byref(rows.struct),
                                           r0, nr, c0, w, mask, out, A.shape[1], _lib.stream_ptr())
        _lib.check(rc, "updes_assemble_block")

    def check_sweeps(self):
        flags = ctypes.c_int32(0)
        _lib.check(self.lib.updes_lu_status(self.h, ctypes.byref(flags), _lib.stream_ptr()), "updes_lu_status")
        if flags.value:
            raise RuntimeError("updes_b200: a triangular sweep timed out waiting for a solved block (status %d)" % flags.value)

    def close(self):
        if self.h:
            self.lib.updes_lu_destroy(self.h)
            self.h = None
            self.t = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_SUBGROUPS = {}


def _grid_groups(dist, parent, P, Q):
    """Row and column communicators of the grid (created once per (parent group, P, Q): every rank of the parent
    group must reach this call)."""
    key = (id(parent) if parent is not None else None, P, Q)
    if key not in _SUBGROUPS:
        glob = (lambda r: r) if parent is None else (lambda r: dist.get_global_rank(parent, r))
        rows = [dist.new_group([glob(p * Q + q) for q in range(Q)]) for p in range(P)]
        cols = [dist.new_group([glob(p * Q + q) for p in range(P)]) for q in range(Q)]
        _SUBGROUPS[key] = (rows, cols)
    return _SUBGROUPS[key]


class DistributedLU2D:
    """P x Q block-cyclic LU with partial pivoting (centralised panel factorisation) and distributed solves."""

    def __init__(self, layout: BlockCyclic2D, rank: int, kernels, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.layout, self.rank, self.K, self.group = layout, rank, kernels, group
        self.p, self.q = layout.coords(rank)
        L, nb = layout, layout.nb
        self.row_groups, self.col_groups = _grid_groups(dist, group, L.P, L.Q)
        dev = kernels.device
        self.device = dev
        f64 = torch.float64
        self.mloc, self.ld = L.local_rows(self.p), L.local_cols(self.q)
        self.mvalid, self.cvalid = L.valid_rows(self.p), L.valid_cols(self.q)
        self.npad = L.nblocks * nb
        self.local = torch.zeros((self.mloc, self.ld), dtype=f64, device=dev)
        self.Gbuf = torch.zeros((self.npad, nb), dtype=f64, device=dev)             # gathered panel (diagonal owners)
        self.Lflat = torch.zeros(self.mloc * nb + nb, dtype=f64, device=dev)        # my rows of the panel + pivot tail
        self.Lbuf = self.Lflat[: self.mloc * nb].view(self.mloc, nb)
        self.Ubuf = torch.zeros((nb, self.ld), dtype=f64, device=dev)               # received U12 block row
        mmax = max(L.local_rows(p) for p in range(L.P))
        self.stage = torch.zeros(mmax * nb + nb, dtype=f64, device=dev)             # send / receive staging on diagonal owners
        self.X = torch.zeros((2 * nb, self.ld), dtype=f64, device=dev)              # row-exchange staging
        self.ipiv_dev = torch.zeros(self.npad, dtype=torch.int32, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ipiv = np.arange(L.n, dtype=np.int64)                                  # complete pivot list (host)
        self.scale_global = None
        self.perm_dev = None
        self.factored = False
        self.always_stage_u = False     # test hook: route U12 through the staging buffer even on the owning process row
        kernels.bind(self.npad, self.local, self.mvalid, self.Gbuf, L.n, self.Lbuf, self.Ubuf)

    # ---- helpers -------------------------------------------------------------------------------------------
    def _g(self, p, q):
        """global rank of grid position (p, q)"""
        r = self.layout.rank_of(p, q)
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def fill_from_global(self, Kfull):
        """Test helper: load this rank's blocks from a global host matrix."""
        L, nb = self.layout, self.layout.nb
        Kp = np.zeros((self.npad, self.npad))
        Kp[: L.n, : L.n] = Kfull
        blocks = Kp.reshape(L.nblocks, nb, L.nblocks, nb)[self.p::L.P, :, self.q::L.Q, :]
        self.local.copy_(self.torch.from_numpy(np.ascontiguousarray(blocks).reshape(self.mloc, self.ld)))
        self.factored = False

    def assemble(self, rows, kind, param, M):
        """Fill the owned tiles of the collocation matrix (updes_assemble_block per tile; no communication).  A row
        block is cut where the jet mask changes (internal rows | boundary rows | P^T rows)."""
        L, nb = self.layout, self.layout.nb
        N, Ni = rows.N, rows.table.Ni
        code = RBF_CODES[kind]
        cuts = [(0, Ni, rows.mask_internal), (Ni, N, rows.mask_boundary), (N, N + M, 7)]
        for I in L.row_blocks(self.p):
            g0, g1 = I * nb, I * nb + L.width(I)
            for a, b, mask in cuts:
                r0, r1 = max(g0, a), min(g1, b)
                if r1 <= r0:
                    continue
                for J in L.col_blocks(self.q):
                    self.K.assemble_tile(rows, code, param, N, M, r0, r1 - r0, J * nb, L.width(J), mask,
                                         L.lrow(I) + (r0 - g0), L.lcol(J))
        self.factored = False

    def equilibrate(self):
        """Row equilibration: per-row max |entry| over the process row (all-reduce MAX), every rank scales its
        columns by the same exact power of two; right-hand sides are scaled in solve()."""
        L, nb, dist, torch = self.layout, self.layout.nb, self.dist, self.torch
        m = self.K.row_absmax(self.cvalid)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.row_groups[self.p])
        scale = self.K.scale_from_absmax(m)
        self.K.row_scale(scale, self.cvalid)
        g = torch.zeros(self.npad, dtype=torch.float64, device=self.device)
        if self.q == 0:
            g.view(L.nblocks, nb)[self.p::L.P].copy_(scale.view(-1, nb))
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        self.scale_global = g[: L.n].clone()
        return self

    # ---- factorisation ---------------------------------------------------------------------------------------
    def _exchange_rows(self, moves):
        """Apply {destination row: source row} to ALL local columns.  Inside a process column every involved row is
        owned by exactly one rank: the owners deposit their rows in a staging buffer, one all-reduce(SUM) hands every
        rank all of them, and every rank overwrites the destination rows it owns."""
        L, torch, dist = self.layout, self.torch, self.dist
        rows = sorted(set(moves) | set(moves.values()))
        pos = {r: i for i, r in enumerate(rows)}
        X = self.X[: len(rows)]
        mine = [i for i, r in enumerate(rows) if L.row_owner(r) == self.p]
        if L.P > 1:
            X.zero_()
        if mine:
            src_local = torch.as_tensor([L.local_row_of(rows[i]) for i in mine], dtype=torch.int64, device=self.device)
            X.index_copy_(0, torch.as_tensor(mine, dtype=torch.int64, device=self.device), self.local.index_select(0, src_local))
        if L.P > 1:
            dist.all_reduce(X, op=dist.ReduceOp.SUM, group=self.col_groups[self.q])
        dest = [d for d in moves if L.row_owner(d) == self.p]
        if dest:
            dl = torch.as_tensor([L.local_row_of(d) for d in dest], dtype=torch.int64, device=self.device)
            sp = torch.as_tensor([pos[moves[d]] for d in dest], dtype=torch.int64, device=self.device)
            self.local.index_copy_(0, dl, X.index_select(0, sp))

    def factor(self):
        L, nb, dist, torch, K = self.layout, self.layout.nb, self.dist, self.torch, self.K
        P, Q, p, q = L.P, L.Q, self.p, self.q
        mloc = self.mloc
        Gb = self.Gbuf.view(L.nblocks, nb, nb)
        tail = self.Lflat[mloc * nb:]
        self.info.zero_()
        for k in range(L.nblocks):
            r0, w, pk, qk = k * nb, L.width(k), k % P, k % Q
            d = L.rank_of(pk, qk)
            lrs = L.lrow_from(p, k)                       # my local rows of the blocks >= k start here
            lc = L.lcol(k)
            # 1. gather the panel on the diagonal owner and factor it there
            if q == qk:
                if self.rank == d:
                    Gb[k::P].copy_(self.local[lrs:, lc:lc + nb].reshape(-1, nb, nb))
                    for p2 in range(P):
                        if p2 == pk:
                            continue
                        cnt = L.local_rows(p2) - L.lrow_from(p2, k)
                        if cnt > 0:
                            buf = self.stage[: cnt * nb]
                            dist.recv(buf, src=self._g(p2, qk), group=self.group)
                            Gb[L.first_row_block_from(p2, k)::P].copy_(buf.view(-1, nb, nb))
                    K.panel_factor(r0, w, self.ipiv_dev, self.info)
                    # 2a. hand every rank of the process column its rows of the factored panel + the pivots
                    pivots = self.ipiv_dev[r0:r0 + w].to(torch.float64)
                    for p2 in range(P):
                        if p2 == pk:
                            continue
                        cnt = L.local_rows(p2) - L.lrow_from(p2, k)
                        msg = self.stage[self.stage.numel() - (cnt * nb + nb):]
                        if cnt > 0:
                            msg[: cnt * nb].view(-1, nb, nb).copy_(Gb[L.first_row_block_from(p2, k)::P])
                        msg[cnt * nb: cnt * nb + w].copy_(pivots)
                        dist.send(msg, dst=self._g(p2, qk), group=self.group)
                    self.Lbuf[lrs:].view(-1, nb, nb).copy_(Gb[k::P])
                    tail[:w].copy_(pivots)
                else:
                    if lrs < mloc:
                        self.Lbuf[lrs:].copy_(self.local[lrs:, lc:lc + nb])
                        dist.send(self.Lflat[lrs * nb: mloc * nb], dst=self._g(pk, qk), group=self.group)
                    dist.recv(self.Lflat[lrs * nb:], src=self._g(pk, qk), group=self.group)
            # 2b. my rows of the panel (and the pivots) along the process row
            if Q > 1:
                dist.broadcast(self.Lflat[lrs * nb:], src=self._g(p, qk), group=self.row_groups[p])
            piv = tail[:w].cpu().numpy().astype(np.int64)
            self.ipiv[r0:r0 + w] = piv
            # 3. row interchanges on every local column; the panel's own columns then take the factored rows
            moves = compose_interchanges(r0, piv)
            if moves:
                self._exchange_rows(moves)
            if q == qk and lrs < mloc:
                self.local[lrs:, lc:lc + nb].copy_(self.Lbuf[lrs:])
            # 4. U12 = L11^-1 A12 on process row pk, then down the process columns
            cr = L.lcol_from(q, k + 1)
            ncr = self.cvalid - cr
            if ncr > 0:
                lr0 = L.lrow(k)
                if p == pk:
                    K.trsm(lr0, w, lr0, cr, ncr)
                    if P > 1:
                        dist.broadcast(self.local[lr0:lr0 + w], src=self._g(pk, q), group=self.col_groups[q])
                elif P > 1:
                    dist.broadcast(self.Ubuf[:w], src=self._g(pk, q), group=self.col_groups[q])
                # 5. trailing update of my blocks below / right of the panel
                lrb = L.lrow_from(p, k + 1)
                m = self.mvalid - lrb
                if m > 0:
                    if p == pk and self.always_stage_u:
                        self.Ubuf[:w].copy_(self.local[lr0:lr0 + w])
                        K.gemm(lrb, "U", 0, cr, lrb, cr, m, ncr, w)
                    elif p == pk:
                        K.gemm(lrb, "local", lr0, cr, lrb, cr, m, ncr, w)
                    else:
                        K.gemm(lrb, "U", 0, cr, lrb, cr, m, ncr, w)
        self.perm_dev = torch.as_tensor(permutation_from_pivots(self.ipiv), dtype=torch.int64, device=self.device)
        self.factored = True
        return self

    def zero_pivot(self):
        """LAPACK-style status combined over all ranks: < 0 internal failure on some rank, > 0 a zero pivot."""
        dist = self.dist
        hi, lo = self.info.clone(), self.info.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        return int(lo.item()) if int(lo.item()) < 0 else int(hi.item())

    # ---- solve ---------------------------------------------------------------------------------------------------
    def solve(self, b):
        """Solve K x = b (b: length-n host array or device vector, identical on all ranks) -> x on every rank."""
        assert self.factored
        L, nb, dist, torch, K = self.layout, self.layout.nb, self.dist, self.torch, self.K
        P, Q, p, q = L.P, L.Q, self.p, self.q
        SUM = dist.ReduceOp.SUM
        b = torch.as_tensor(np.ascontiguousarray(b) if isinstance(b, np.ndarray) else b, dtype=torch.float64).to(self.device)
        if self.scale_global is not None:
            b = b * self.scale_global
        x = b.index_select(0, self.perm_dev)
        xl = torch.zeros(max(self.ld, 2), dtype=torch.float64, device=self.device)      # solved entries by local column
        xt = torch.zeros(max(self.mloc, 2), dtype=torch.float64, device=self.device)    # block right-hand sides by local row
        part = torch.zeros(nb, dtype=torch.float64, device=self.device)
        for upper in (False, True):
            order = range(L.nblocks - 1, -1, -1) if upper else range(L.nblocks)
            for j in order:
                r0, w, pj, qj = j * nb, L.width(j), j % P, j % Q
                lc = L.lcol(j)
                if p == pj:
                    lr = L.lrow(j)
                    out = part[:w]
                    if upper:
                        K.gemv(lr, w, L.lcol_from(q, j + 1), self.cvalid, xl, out)     # my columns of the blocks > j
                    else:
                        K.gemv(lr, w, 0, L.lcol_from(q, j), xl, out)                   # my columns of the blocks < j
                    if Q > 1:
                        dist.reduce(out, dst=self._g(pj, qj), op=SUM, group=self.row_groups[p])
                    if q == qj:
                        rhs = xl[lc:lc + w] if upper else x[r0:r0 + w]
                        xt[lr:lr + w].copy_(rhs - out)
                        K.diag_solve(upper, lr, lc, w, xt)
                        xl[lc:lc + w].copy_(xt[lr:lr + w])
                if q == qj and P > 1:
                    dist.broadcast(xl[lc:lc + w], src=self._g(pj, qj), group=self.col_groups[q])
        out = torch.zeros(L.n, dtype=torch.float64, device=self.device)
        for j in range(L.nblocks):
            if j % P == p and j % Q == q:
                out[j * nb: j * nb + L.width(j)].copy_(xl[L.lcol(j): L.lcol(j) + L.width(j)])
        dist.all_reduce(out, op=SUM, group=self.group)
        return out

    def nbytes(self):
        return 8 * (self.local.numel() + self.Gbuf.numel() + self.Lflat.numel() + self.Ubuf.numel() + self.stage.numel()
                    + self.X.numel())

    def close(self):
        if hasattr(self.K, "close"):
            self.K.close()
