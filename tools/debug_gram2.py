import numpy as np, torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine
n = 200
parent = np.full(n, n - 1, np.int32); parent[-1] = -1
ctx = engine.TreeContext(parent, np.ones(n))
N = 300
rng = np.random.default_rng(0)
np.set_printoptions(linewidth=250)
for case in ("rank1_k0", "rank1_k77", "dense_mixed_scales"):
    P = np.zeros((n, N), dtype=np.float32)
    if case == "rank1_k0": P[0] = np.arange(1, N + 1)
    elif case == "rank1_k77": P[77] = np.arange(1, N + 1) * 0.37
    else:
        P = (rng.random((n, N)) * (rng.random((n, N)) < 0.4)).astype(np.float32) * (10.0 ** rng.integers(-6, 1, (n, 1))).astype(np.float32)
    PT = torch.zeros((n, engine.padded_samples(N)), dtype=torch.float32, device="cuda"); PT[:, :N] = torch.from_numpy(P)
    D, q, err = ctx.pairwise_gram(PT, N)
    torch.cuda.synchronize()
    Dn = D.cpu().numpy()
    P64 = P.astype(np.float64)
    exp = ((P64[:, :, None] - P64[:, None, :]) ** 2).sum(0)
    scale = (P64 ** 2).sum(0); sc = scale[:, None] + scale[None, :]
    iu = np.triu_indices(N, 1)
    print(f"{case}: max |D - exact| / (qi+qj) = {(np.abs(Dn - exp) / sc)[iu].max():.3e}; max rel = {(np.abs(Dn - exp)[iu] / exp[iu]).max():.3e}; sym {np.array_equal(Dn, Dn.T)} diag0 {not np.diag(Dn).any()}; q rel err {np.abs(q.cpu().numpy()[:N] - scale).max() / scale.max():.2e}; err stat max {err.max().item():.2e}")
