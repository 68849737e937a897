"""CPU stand-in for updes_b200.distributed.CudaBackend, used ONLY by tests to exercise the host logic
of the multi-GPU driver (ownership maps, look-ahead order, pivot plumbing, solves) under gloo.
Same method names and semantics; numpy arithmetic; panel buffers are torch CPU tensors so that
torch.distributed can broadcast them."""
import numpy as np
import torch


class NumpyBackend:
    def __init__(self, layout, rank):
        self.layout, self.rank = layout, rank
        self.n, self.nb = layout.n, layout.nb
        self.cols = layout.local_cols(rank)
        self.ld = max((self.cols + 15) // 16 * 16, 16)
        self.local = np.zeros((self.n, self.ld))
        self.bufs = [torch.zeros(self.n * self.nb + self.nb, dtype=torch.float64) for _ in range(2)]
        self.ipiv = np.zeros(self.n, dtype=np.int32)
        self.info = 0
        self.calls = []
        self.scale = None

    def fill_from_global(self, K):
        for j in self.layout.local_blocks(self.rank):
            w, lc = self.layout.width(j), self.layout.local_offset(j)
            self.local[:, lc:lc + w] = K[:, j * self.nb:j * self.nb + w]

    def _panel(self, slot):
        return self.bufs[slot][: self.n * self.nb].view(self.n, self.nb).numpy()

    def panel_factor(self, r0, lc, w):
        A = self.local
        self.calls.append(("panel", r0, lc, w))
        for j in range(w):
            col = np.abs(A[r0 + j:, lc + j])
            p = r0 + j + int(np.argmax(col))
            self.ipiv[r0 + j] = p
            if A[p, lc + j] == 0.0 and self.info == 0:
                self.info = r0 + j + 1
            if p != r0 + j:
                A[[r0 + j, p], lc:lc + w] = A[[p, r0 + j], lc:lc + w]
            if A[r0 + j, lc + j] != 0.0:
                A[r0 + j + 1:, lc + j] /= A[r0 + j, lc + j]
                A[r0 + j + 1:, lc + j + 1:lc + w] -= np.outer(A[r0 + j + 1:, lc + j], A[r0 + j, lc + j + 1:lc + w])

    def pack(self, slot, r0, lc, w):
        self._panel(slot)[r0:, :w] = self.local[r0:, lc:lc + w]
        self.bufs[slot][self.n * self.nb: self.n * self.nb + w] = torch.from_numpy(self.ipiv[r0:r0 + w].astype(np.float64))

    def unpack_pivots(self, slot, r0, w):
        self.ipiv[r0:r0 + w] = self.bufs[slot][self.n * self.nb: self.n * self.nb + w].numpy().astype(np.int32)

    def message(self, slot, r0):
        return self.bufs[slot][r0 * self.nb:]

    def apply_swaps(self, r0, w, c_lo, c_hi):
        if c_hi <= c_lo:
            return
        for t in range(w):
            p = int(self.ipiv[r0 + t])
            if p != r0 + t:
                self.local[[r0 + t, p], c_lo:c_hi] = self.local[[p, r0 + t], c_lo:c_hi]

    def apply_panel(self, slot, r0, w, c_lo, c_hi):
        if c_hi <= c_lo:
            return
        self.calls.append(("apply", slot, r0, w, c_lo, c_hi))
        self.apply_swaps(r0, w, c_lo, c_hi)
        P = self._panel(slot)
        L11 = np.tril(P[r0:r0 + w, :w], -1) + np.eye(w)
        self.local[r0:r0 + w, c_lo:c_hi] = np.linalg.solve(L11, self.local[r0:r0 + w, c_lo:c_hi])
        self.local[r0 + w:, c_lo:c_hi] -= P[r0 + w:, :w] @ self.local[r0:r0 + w, c_lo:c_hi]

    def set_pivots(self):
        perm = np.arange(self.n)
        for k in range(self.n):
            p = int(self.ipiv[k])
            perm[[k, p]] = perm[[p, k]]
        self.perm = perm

    def vector(self, host_array):
        return torch.from_numpy(np.array(host_array, dtype=np.float64))

    def permute_rhs(self, b):
        v = b.numpy() if self.scale is None else b.numpy() * self.scale
        return torch.from_numpy(v[self.perm].copy())

    def row_absmax(self):
        return torch.from_numpy(np.abs(self.local[:, :self.cols]).max(axis=1) if self.cols else np.zeros(self.n))

    def apply_row_scale(self, absmax):
        m = absmax.numpy()
        sc = np.ones(self.n)
        ok = (m > 0) & np.isfinite(m)
        sc[ok] = np.ldexp(1.0, 1 - np.frexp(m[ok])[1])
        self.scale = sc
        self.local *= sc[:, None]

    def event(self):
        import time

        class _E:
            def __init__(self): self.t = time.perf_counter()
            def elapsed_time(self, other): return (other.t - self.t) * 1e3
        return _E()

    def block_sweep(self, upper, r0, lc, w, x):
        xv = x.numpy()
        T = self.local[r0:r0 + w, lc:lc + w]
        if not upper:
            xv[r0:r0 + w] = np.linalg.solve(np.tril(T, -1) + np.eye(w), xv[r0:r0 + w])
            xv[r0 + w:] -= self.local[r0 + w:, lc:lc + w] @ xv[r0:r0 + w]
        else:
            xv[r0:r0 + w] = np.linalg.solve(np.triu(T), xv[r0:r0 + w])
            xv[:r0] -= self.local[:r0, lc:lc + w] @ xv[r0:r0 + w]

    def block_gemv(self, r0, w, c_lo, c_hi, xl, out):
        self.calls.append(("gemv", r0, w, c_lo, c_hi))
        if c_hi <= c_lo:
            out.zero_()
        else:
            out.numpy()[:] = self.local[r0:r0 + w, c_lo:c_hi] @ xl.numpy()[c_lo:c_hi]

    def diag_solve(self, upper, r0, lc, w, x):
        xv = x.numpy()
        T = self.local[r0:r0 + w, lc:lc + w]
        xv[r0:r0 + w] = np.linalg.solve(np.triu(T) if upper else np.tril(T, -1) + np.eye(w), xv[r0:r0 + w])

    def zeros(self, k):
        return torch.zeros(k, dtype=torch.float64)

    def zero_pivot(self):
        return self.info
