"""CPU stand-in for updes_b200.grid2d.CudaKernels2D, used ONLY by tests: the P x Q block-cyclic driver
(index maps, staging, panel gather / scatter, row exchange, broadcasts, left-looking solves) then runs under
gloo with numpy arithmetic on CPU torch tensors.  Same method names, argument meaning and in-place semantics."""
import numpy as np
import torch


class NumpyKernels2D:
    def __init__(self):
        self.device = torch.device("cpu")
        self.calls = []

    def bind(self, n_padded, local, local_rows, G, G_rows, L, U):
        self.local, self.G, self.L, self.U = local.numpy(), G.numpy(), L.numpy(), U.numpy()
        self.local_rows, self.G_rows = int(local_rows), int(G_rows)

    def panel_factor(self, r0, w, ipiv, info):
        """LU with partial pivoting of rows [r0, G_rows) x columns [0, w) of the gathered panel; pivots are rows of G."""
        A, n = self.G, self.G_rows
        self.calls.append(("panel", r0, w))
        piv = ipiv.numpy()
        for j in range(w):
            col = np.abs(A[r0 + j:n, j])
            p = r0 + j + int(np.argmax(col))
            piv[r0 + j] = p
            if A[p, j] == 0.0 and int(info[0]) == 0:
                info[0] = r0 + j + 1
            if p != r0 + j:
                A[[r0 + j, p], :w] = A[[p, r0 + j], :w]
            if A[r0 + j, j] != 0.0:
                A[r0 + j + 1:n, j] /= A[r0 + j, j]
                A[r0 + j + 1:n, j + 1:w] -= np.outer(A[r0 + j + 1:n, j], A[r0 + j, j + 1:w])

    def trsm(self, lr, w, rb, cb, ncols):
        assert w % 32 == 0 or w == 16, "the CUDA TRSM takes widths of 16 or multiples of 32"
        L11 = np.tril(self.L[lr:lr + w, :w], -1) + np.eye(w)
        self.local[rb:rb + w, cb:cb + ncols] = np.linalg.solve(L11, self.local[rb:rb + w, cb:cb + ncols])

    def gemm(self, ra, b_name, rb, cb, rc, cc, m, n, k):
        assert k % 16 == 0 and cc % 2 == 0, "the DMMA GEMM needs k % 16 == 0 and an even C column"
        B = self.local if b_name == "local" else self.U
        assert rc + m <= self.local_rows
        self.calls.append(("gemm", b_name, m, n, k))
        self.local[rc:rc + m, cc:cc + n] -= self.L[ra:ra + m, :k] @ B[rb:rb + k, cb:cb + n]

    def gemv(self, r0, w, c_lo, c_hi, xl, out):
        assert c_lo % 2 == 0
        if c_hi <= c_lo:
            out.zero_()
        else:
            out.numpy()[:] = self.local[r0:r0 + w, c_lo:c_hi] @ xl.numpy()[c_lo:c_hi]

    def diag_solve(self, upper, r0, c0, w, xt):
        assert r0 + w <= self.local_rows and c0 % 2 == 0
        T = self.local[r0:r0 + w, c0:c0 + w]
        v = xt.numpy()
        v[r0:r0 + w] = np.linalg.solve(np.triu(T) if upper else np.tril(T, -1) + np.eye(w), v[r0:r0 + w])

    def row_absmax(self, cols):
        return torch.from_numpy(np.abs(self.local[:, :cols]).max(axis=1) if cols else np.zeros(self.local.shape[0]))

    def scale_from_absmax(self, absmax):
        m = absmax.numpy()
        sc = np.ones_like(m)
        ok = (m > 0) & np.isfinite(m)
        sc[ok] = np.ldexp(1.0, 1 - np.frexp(m[ok])[1])
        return torch.from_numpy(sc)

    def row_scale(self, scale, cols):
        self.local[:, :cols] *= scale.numpy()[:, None]
