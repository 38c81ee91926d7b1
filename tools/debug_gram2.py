import numpy as np, torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine
n = 64
parent = np.full(n, n - 1, np.int32); parent[-1] = -1
ctx = engine.TreeContext(parent, np.ones(n))
N = 256
np.set_printoptions(linewidth=250)
for k0 in (0, 1, 9, 33):
    PT = torch.zeros((n, N), dtype=torch.float32, device="cuda")
    PT[k0] = torch.arange(1, N + 1, dtype=torch.float32, device="cuda")
    D, q, _ = ctx.pairwise_gram(PT, N)
    torch.cuda.synchronize()
    Dn = D.cpu().numpy()
    idx = np.arange(N, dtype=float)
    exp = (idx[:, None] - idx[None, :]) ** 2
    print(f"k0={k0}: D exact: {np.array_equal(Dn, exp)}  max abs diff {np.abs(Dn - exp).max()}")
