"""GPU parity of the dense LU building blocks and the factor/solve path (through the C-ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _matrix(n, seed=0, ld=None, dominant=False):
    import torch
    from updes_b200.assembly import padded_ld
    ld = ld or padded_ld(n)
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn((n, n), generator=g, dtype=torch.float64)
    if dominant:
        A += n * torch.eye(n, dtype=torch.float64)
    K = torch.zeros((n, ld), dtype=torch.float64)
    K[:, :n] = A
    return A, K.cuda()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 11, 15, 8])
@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (256, 128, 64), (300, 200, 32), (1000, 96, 128), (77, 500, 48),
                                   (2048, 2048, 512), (129, 33, 16), (64, 32, 32), (4000, 64, 256), (5000, 3001, 96), (6000, 2500, 512)])
def test_dgemm_sub(m, n, k, variant):
    """C -= A @ B on sub-blocks of one matrix vs torch (DMMA + TMA kernel, all tile shapes / edges)."""
    import torch
    from updes_b200.linalg import LUFactorization
    N = max(m, k) + k + 64
    N = (N + 31) // 32 * 32
    cols = (n + k + 64 + 31) // 32 * 32
    size = max(N, cols)
    _, K = _matrix(size, seed=m + n + k)
    lu = LUFactorization(K, size)
    lu.set_gemm_variant(variant)          # bit 0: ping-pong (two 128x64 CTAs per SM); bit 1: 32-deep stages; bit 3: read-modify-write epilogue
    # A at (ra=k+32.., ca=0), B at (rb=0, cb=k+32), C at (k+32, k+32): an LU-like arrangement
    ra, ca, rb, cb = 32 + k, 0, 0, 32 + k
    rc, cc = 32 + k, 32 + k
    m = min(m, size - ra); n = min(n, size - cb)
    ref = K.clone()
    ref[rc:rc + m, cc:cc + n] -= ref[ra:ra + m, ca:ca + k] @ ref[rb:rb + k, cb:cb + n]
    lu.gemm_sub(rc, cc, ra, ca, rb, cb, m, n, k)
    torch.cuda.synchronize()
    err = (K - ref).abs().max().item()
    assert err <= 1e-12 * k * ref.abs().max().item(), err
    # nothing outside the C block may change
    mask = torch.ones_like(K, dtype=torch.bool); mask[rc:rc + m, cc:cc + n] = False
    assert torch.equal(K[mask], ref[mask])


@pytest.mark.parametrize("n", [32, 64, 100, 257, 1000, 2051])
def test_lu_factor_reconstructs(n):
    """P K = L U: reconstruct from the factors; pivots must equal LAPACK-style partial pivoting."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=n)
    lu = LUFactorization(K, n).factor()
    torch.cuda.synchronize()
    assert lu.zero_pivot() == 0
    F = K[:, :n].cpu()
    L = torch.tril(F, -1) + torch.eye(n, dtype=torch.float64)
    U = torch.triu(F)
    PA = A.clone()
    piv = lu.ipiv.cpu().numpy()
    for kk in range(n):
        p = int(piv[kk])
        assert kk <= p < n
        if p != kk:
            PA[[kk, p]] = PA[[p, kk]]
    err = (L @ U - PA).abs().max().item() / A.abs().max().item()
    assert err <= 1e-13 * n, err
    assert L.abs().max().item() <= 1.0 + 1e-12            # partial pivoting bounds the multipliers
    # same pivots as LAPACK (torch.linalg.lu_factor: 1-based)
    _, piv_ref = torch.linalg.lu_factor(A)
    assert np.array_equal(piv, piv_ref.numpy() - 1)


@pytest.mark.parametrize("cap", [600, 300])
def test_lu_tall_panel_paths(cap):
    """Panels taller than the 32-wide register-resident kernel fall back to 16- and 8-wide base panels
    (250k-row panels of the multi-GPU path); force that with a small capacity and compare with LAPACK."""
    import torch
    from updes_b200.linalg import LUFactorization
    n = 1000                                          # cap=600 -> 16-wide at the top; cap=300 -> 8-wide, then 16, then 32
    A, K = _matrix(n, seed=77)
    lu = LUFactorization(K, n)
    lu.set_panel_capacity(cap)
    lu.factor()
    torch.cuda.synchronize()
    assert lu.zero_pivot() == 0
    lu_ref, piv_ref = torch.linalg.lu_factor(A)
    assert np.array_equal(lu.ipiv.cpu().numpy(), piv_ref.numpy() - 1)
    assert (K[:, :n].cpu() - lu_ref).abs().max().item() <= 1e-10 * lu_ref.abs().max().item()


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("n,nrhs", [(64, 1), (300, 3), (1000, 1), (2051, 5), (128, 2), (129, 1), (20000, 2), (20000, 1), (5003, 4)])
def test_lu_solve(n, nrhs, variant):
    """Triangular solves: variant 2 = row-block streaming sweeps (default; solved blocks published through the
    data), variant 1 = step-synchronous persistent sweeps, variant 0 = one launch per 128-row block."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=10 + n, dominant=(n >= 20000))
    lu = LUFactorization(K, n).factor()
    lu.set_solve_variant(variant)
    g = torch.Generator().manual_seed(1)
    B = torch.randn((nrhs, n), generator=g, dtype=torch.float64)
    X = lu.solve(B.cuda().clone()).cpu()
    ref = torch.linalg.solve(A, B.T).T
    err = (X - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-9, err
    res = (A @ X.T - B.T).abs().max().item() / (A.abs().max().item() * X.abs().max().item() * n)
    assert res <= 1e-14, res


@pytest.mark.parametrize("n,nrhs", [(64, 1), (300, 3), (1000, 2), (2051, 5), (129, 1)])
def test_lu_solve_transposed(n, nrhs):
    """K^T X = B with the factors of K (adjoint solves)."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=30 + n)
    lu = LUFactorization(K, n).factor()
    g = torch.Generator().manual_seed(2)
    B = torch.randn((nrhs, n), generator=g, dtype=torch.float64)
    X = lu.solve(B.cuda().clone(), transpose=True).cpu()
    ref = torch.linalg.solve(A.T, B.T).T
    assert (X - ref).abs().max().item() / ref.abs().max().item() <= 1e-9
    res = (A.T @ X.T - B.T).abs().max().item() / (A.abs().max().item() * X.abs().max().item() * n)
    assert res <= 1e-14, res


@pytest.mark.parametrize("n", [100, 640, 2051, 9000, 12000])
def test_panel_variants_agree_with_lapack(n):
    """Second-generation panel kernels (implicit pivoting, REDUX arg-max, pushed cluster exchange) must give the
    same pivots as the first-generation ones and as LAPACK; n = 9000 / 12000 cross the 8 192-row boundary between
    the cluster kernel and the grid-wide kernel."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K0 = _matrix(n, seed=3 * n)
    _, piv_ref = torch.linalg.lu_factor(A.cuda())
    piv_ref = piv_ref.cpu().numpy() - 1
    facs = []
    for variant in (2, 1, 0):
        K = K0.clone()
        lu = LUFactorization(K, n)
        lu.set_panel_variant(variant)
        lu.factor()
        torch.cuda.synchronize()
        assert lu.check() == 0
        assert np.array_equal(lu.ipiv.cpu().numpy(), piv_ref), "panel variant %d: pivots differ from LAPACK" % variant
        facs.append(K[:, :n].clone())
    scale = facs[0].abs().max().item()
    assert (facs[0] - facs[1]).abs().max().item() <= 1e-11 * scale
    assert (facs[0] - facs[2]).abs().max().item() <= 1e-11 * scale


def test_panel_ties_take_the_first_row_like_lapack():
    """Equal magnitudes in a column: idamax takes the smallest row index, also after interchanges."""
    import torch
    from updes_b200.linalg import LUFactorization
    n = 700
    g = torch.Generator().manual_seed(9)
    A = torch.randint(-3, 4, (n, n), generator=g, dtype=torch.int64).to(torch.float64)   # many exact ties
    A += torch.diag(torch.full((n,), 0.5, dtype=torch.float64))
    from updes_b200.assembly import padded_ld
    for variant in (2, 1):
        K = torch.zeros((n, padded_ld(n)), dtype=torch.float64); K[:, :n] = A
        K = K.cuda()
        lu = LUFactorization(K, n)
        lu.set_panel_variant(variant)
        lu.factor()
        torch.cuda.synchronize()
        _, piv_ref = torch.linalg.lu_factor(A)
        # LAPACK blocks differently, so rounding may break later ties differently; the first 32-column panel is exact
        assert np.array_equal(lu.ipiv.cpu().numpy()[:8], piv_ref.numpy()[:8] - 1)
        F = K[:, :n].cpu()
        L = torch.tril(F, -1) + torch.eye(n, dtype=torch.float64)
        PA = A.clone()
        for kk, p in enumerate(lu.ipiv.cpu().tolist()):
            if p != kk:
                PA[[kk, p]] = PA[[p, kk]]
        assert (L @ torch.triu(F) - PA).abs().max().item() <= 1e-10 * n
        assert L.abs().max().item() <= 1.0 + 1e-12


@pytest.mark.parametrize("n", [300, 2051])
def test_row_equilibrated_factorisation(n):
    """updes_lu_factor_scaled: rows scaled by exact powers of two before pivoting; solves (plain and transposed)
    still solve the ORIGINAL system, and the backward error no longer depends on the row scaling."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=n + 5)
    g = torch.Generator().manual_seed(3)
    rs = 10.0 ** torch.randint(-8, 9, (n,), generator=g).to(torch.float64)       # rows of wildly different scale
    A = A * rs[:, None]
    K[:, :n] = A.cuda()
    lu = LUFactorization(K, n).factor(equilibrate=True)
    torch.cuda.synchronize()
    assert lu.check() == 0
    sc = lu.scale.cpu()
    m = A.abs().max(dim=1).values * sc
    assert bool(((m >= 1.0) & (m < 2.0)).all()) and bool((torch.frexp(sc)[0] == 0.5).all())
    B = torch.randn((2, n), generator=g, dtype=torch.float64)
    X = lu.solve((B * rs[None]).cuda().clone()).cpu()            # right-hand side scaled like the rows
    ref = torch.linalg.solve(A / rs[:, None], B.T).T
    assert (X - ref).abs().max().item() / ref.abs().max().item() <= 1e-9
    Xt = lu.solve(B.cuda().clone(), transpose=True).cpu()
    res = (A.T @ Xt.T - B.T).abs().max().item() / ((A.abs().sum(0).max() * Xt.abs().max()).item() + B.abs().max().item())
    assert res <= 1e-13, res
    # the pivots are those LAPACK finds on the equilibrated matrix
    _, piv_ref = torch.linalg.lu_factor(A * sc[:, None])
    assert np.array_equal(lu.ipiv.cpu().numpy(), piv_ref.numpy() - 1)


def test_internal_failure_raises():
    """A timed-out grid barrier is reported as info = -1; check() must raise instead of returning garbage."""
    import torch
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(64, seed=1)
    lu = LUFactorization(K, 64).factor()
    torch.cuda.synchronize()
    assert lu.check() == 0
    lu.info.fill_(-1)
    with pytest.raises(RuntimeError, match="grid barrier"):
        lu.check()


def test_lu_singular_reports_info():
    import torch
    from updes_b200.linalg import LUFactorization
    n = 96
    A, K = _matrix(n, seed=4)
    K[:, 40] = 0.0                                    # an exactly zero column -> zero pivot at column 41
    lu = LUFactorization(K, n).factor()
    assert lu.zero_pivot() == 41


@pytest.mark.parametrize("n,r0,w,c_lo,c_hi", [(1000, 128, 128, 0, 128), (1000, 256, 100, 0, 991), (2051, 1024, 512, 512, 2051),
                                            (2051, 2048, 3, 64, 2048), (700, 0, 64, 0, 0), (5000, 1024, 1024, 0, 4096)])
def test_block_gemv_partial_products(n, r0, w, c_lo, c_hi):
    """updes_block_gemv: out = A[r0:r0+w, c_lo:c_hi] @ x[c_lo:c_hi] (the left-looking distributed substitution's partials)."""
    import ctypes
    import torch
    from updes_b200 import _lib
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=n + w)
    lu = LUFactorization(K, n)
    lib = _lib.load()
    _lib.check(lib.updes_lu_bind(lu._handle, 0, K.data_ptr(), n, K.shape[1]), "bind")
    x = torch.randn(K.shape[1], dtype=torch.float64, device="cuda")
    out = torch.full((w,), 7.0, dtype=torch.float64, device="cuda")
    _lib.check(lib.updes_block_gemv(lu._handle, 0, r0, w, c_lo, c_hi, x.data_ptr(), out.data_ptr(), _lib.stream_ptr()), "gemv")
    ref = K[r0:r0 + w, c_lo:c_hi] @ x[c_lo:c_hi]
    scale = float((K[r0:r0 + w, c_lo:c_hi].abs() @ x[c_lo:c_hi].abs()).max()) if c_hi > c_lo else 1.0
    assert float((out - ref).abs().max()) <= 1e-14 * max(scale, 1.0)
    # argument validation through the ABI: odd first column, range past the leading dimension
    assert lib.updes_block_gemv(lu._handle, 0, r0, w, 1, 8, x.data_ptr(), out.data_ptr(), _lib.stream_ptr()) == -5
    assert lib.updes_block_gemv(lu._handle, 0, n, 1, 0, 8, x.data_ptr(), out.data_ptr(), _lib.stream_ptr()) == -3


@pytest.mark.parametrize("n,r0,w", [(1024, 0, 1024), (2051, 1024, 1027), (1000, 256, 512), (1000, 64, 96), (3000, 2944, 56), (600, 128, 32)])
@pytest.mark.parametrize("upper", [0, 1])
def test_tri_diag_solve_touches_only_its_block(n, r0, w, upper):
    """updes_tri_diag_solve: the w x w diagonal block at (r0, c0) solved in place on x[r0:r0+w]; nothing else moves.
    The block sits at local column c0 != r0, as on a rank that owns only some of the column blocks."""
    import torch
    from updes_b200 import _lib
    from updes_b200.linalg import LUFactorization
    A, K = _matrix(n, seed=3 * n + w, dominant=True)
    lu = LUFactorization(K, n)
    lib = _lib.load()
    _lib.check(lib.updes_lu_bind(lu._handle, 0, K.data_ptr(), n, K.shape[1]), "bind")
    c0 = min((r0 // 2 + 64) & ~1, (K.shape[1] - w) & ~1)
    idx = torch.arange(w, device="cuda")
    K[r0 + idx, c0 + idx] += float(n)                  # a random triangular block would be exponentially ill-conditioned
    T = K[r0:r0 + w, c0:c0 + w]
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    x0 = x.clone()
    _lib.check(lib.updes_tri_diag_solve(lu._handle, 0, upper, r0, c0, w, x.data_ptr(), _lib.stream_ptr()), "diag_solve")
    Tm = torch.triu(T) if upper else torch.tril(T, -1) + torch.eye(w, dtype=torch.float64, device="cuda")
    ref = torch.linalg.solve_triangular(Tm, x0[r0:r0 + w, None], upper=bool(upper))[:, 0]
    assert float((x[r0:r0 + w] - ref).abs().max() / ref.abs().max()) <= 1e-11
    keep = torch.ones(n, dtype=torch.bool, device="cuda"); keep[r0:r0 + w] = False
    assert torch.equal(x[keep], x0[keep])
    assert lu.check() == 0


@pytest.mark.parametrize("n1,ncols,base", [(128, 200, 128), (128, 9600, 128), (96, 19100, 128), (128, 40000, 128), (64, 777, 128),
                                          (512, 3000, 128), (512, 3000, 32), (1024, 1500, 128)])
def test_trsm_unit_lower_blocks(n1, ncols, base):
    """updes_lu_trsm: B <- L^-1 B for a unit-lower n1 x n1 block against ncols columns of the same buffer (the 128-row
    substitution kernel in all three CTA sizes, the 32-row kernel, and the recursion above them) vs torch."""
    import torch
    from updes_b200 import _lib
    from updes_b200.assembly import padded_ld
    g = torch.Generator(device="cpu").manual_seed(n1 + ncols)
    rows, ld = n1 + 40, padded_ld(n1 + 64 + ncols)
    K = torch.randn((rows, ld), generator=g, dtype=torch.float64).cuda()
    K[:, :n1 + 64] *= 0.05                                   # |multipliers| < 1 as after partial pivoting; keeps the solve tame
    rl, cl, rb, cb = 8, 32, 8, n1 + 64                       # L at (8, 32), B at (8, n1 + 64): same rows, as inside the LU
    ref = K.clone()
    L = torch.tril(ref[rl:rl + n1, cl:cl + n1], -1) + torch.eye(n1, dtype=torch.float64, device="cuda")
    ref[rb:rb + n1, cb:cb + ncols] = torch.linalg.solve_triangular(L, ref[rb:rb + n1, cb:cb + ncols], upper=False, unitriangular=True)
    lib = _lib.load()
    import ctypes
    h = ctypes.c_void_p()
    _lib.check(lib.updes_lu_create(ctypes.byref(h), rows, ld), "create")
    _lib.check(lib.updes_lu_bind(h, 0, K.data_ptr(), rows, ld), "bind")
    _lib.check(lib.updes_lu_set_trsm_base(h, base), "base")
    _lib.check(lib.updes_lu_trsm(h, 0, rl, cl, n1, 0, rb, cb, ncols, _lib.stream_ptr()), "trsm")
    torch.cuda.synchronize()
    err = float((K - ref).abs().max() / ref[rb:rb + n1, cb:cb + ncols].abs().max())
    assert err <= 1e-12, err
    keep = torch.ones_like(K, dtype=torch.bool); keep[rb:rb + n1, cb:cb + ncols] = False
    assert torch.equal(K[keep], ref[keep])
    lib.updes_lu_destroy(h)


@pytest.mark.parametrize("rows,ncols,k0,npiv", [(400, 300, 0, 32), (3000, 1000, 64, 512), (5000, 777, 100, 70), (2048, 4096, 0, 2048), (900, 64, 10, 1)])
def test_apply_swaps_many_pivots_in_one_launch(rows, ncols, k0, npiv):
    """updes_lu_apply_swaps: interchanges (k0+t <-> ipiv[k0+t]), t = 0..npiv-1, in order, on a column range; every
    batch of 32 inside one launch.  Pivot lists with repeats, self-swaps and targets inside later batches."""
    import ctypes
    import torch
    from updes_b200 import _lib
    from updes_b200.assembly import padded_ld
    g = torch.Generator(device="cpu").manual_seed(rows + npiv)
    ld = padded_ld(ncols + 40)
    K = torch.randn((rows, ld), generator=g, dtype=torch.float64).cuda()
    piv = torch.zeros(rows, dtype=torch.int32)
    for t in range(npiv):
        r = int(torch.randint(0, 10, (1,), generator=g))
        if r < 2:
            p = k0 + t                                           # self
        elif r < 5:
            p = min(rows - 1, k0 + t + int(torch.randint(0, 40, (1,), generator=g)))   # a nearby row (often a later diagonal row)
        else:
            p = int(torch.randint(k0 + t, rows, (1,), generator=g))
        piv[k0 + t] = p
    ref = K.clone()
    c_lo, c_hi = 16, 16 + ncols
    for t in range(npiv):
        p = int(piv[k0 + t])
        if p != k0 + t:
            tmp = ref[k0 + t, c_lo:c_hi].clone()
            ref[k0 + t, c_lo:c_hi] = ref[p, c_lo:c_hi]
            ref[p, c_lo:c_hi] = tmp
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.updes_lu_create(ctypes.byref(h), rows, ld), "create")
    _lib.check(lib.updes_lu_bind(h, 0, K.data_ptr(), rows, ld), "bind")
    pd = piv.cuda()
    _lib.check(lib.updes_lu_apply_swaps(h, 0, c_lo, c_hi, k0, npiv, pd.data_ptr(), _lib.stream_ptr()), "swaps")
    torch.cuda.synchronize()
    assert torch.equal(K, ref)
    lib.updes_lu_destroy(h)
