"""Factor + solve once at a moderate size (for ncu captures of the panel and sweep kernels)."""
import sys
import torch
sys.path.insert(0, ".")
from updes_b200.assembly import padded_ld
from updes_b200.linalg import LUFactorization
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
K = torch.randn((n, padded_ld(n)), dtype=torch.float64, device="cuda")
K[:, :n] += n ** 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
lu = LUFactorization(K, n).factor()
x = torch.randn(n, dtype=torch.float64, device="cuda")
lu.solve(x)
torch.cuda.synchronize()
print("ok", lu.zero_pivot())
