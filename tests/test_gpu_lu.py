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


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
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
    lu.set_gemm_variant(variant)          # 1 = ping-pong schedule (two 128x64 CTAs per SM) for wide updates
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


@pytest.mark.parametrize("variant", [1, 0])
@pytest.mark.parametrize("n,nrhs", [(64, 1), (300, 3), (1000, 1), (2051, 5), (128, 2), (129, 1), (20000, 2)])
def test_lu_solve(n, nrhs, variant):
    """Triangular solves: variant 1 = persistent pipelined sweeps (one cooperative launch per direction),
    variant 0 = one launch per 128-row block."""
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


def test_lu_singular_reports_info():
    import torch
    from updes_b200.linalg import LUFactorization
    n = 96
    A, K = _matrix(n, seed=4)
    K[:, 40] = 0.0                                    # an exactly zero column -> zero pivot at column 41
    lu = LUFactorization(K, n).factor()
    assert lu.zero_pivot() == 41
