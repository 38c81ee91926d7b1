import numpy as np, torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine
n = 200
parent = np.full(n, n - 1, np.int32); parent[-1] = -1
ctx = engine.TreeContext(parent, np.ones(n))
N = 300
P = np.zeros((n, N), dtype=np.float32); P[0] = np.arange(1, N + 1)
PT = torch.zeros((n, engine.padded_samples(N)), dtype=torch.float32, device="cuda"); PT[:, :N] = torch.from_numpy(P)
D, q, err = ctx.pairwise_gram(PT, N)
Dn = D.cpu().numpy(); qs = q.cpu().numpy()[:N]
G = (qs[:, None] + qs[None, :] - Dn) / 2
a = P[0].astype(np.float64); S = 1024.0
t = a / S * 127; f0 = np.rint(t); t = (t - f0) * 254; f1 = np.rint(t); t = (t - f1) * 254; f2 = np.rint(t)
w0 = S * S / 127 ** 2
c0 = np.outer(f0, f0); c01 = np.outer(f0, f1); c10 = np.outer(f1, f0); c11 = np.outer(f1, f1); c02 = np.outer(f0, f2); c20 = np.outer(f2, f0)
cands = {
    "full": w0 * (c0 + (c01 + c10) / 254 + (c11 + c02 + c20) / 254 ** 2),
    "c0 only": w0 * c0,
    "c0 + c1": w0 * (c0 + (c01 + c10) / 254),
    "c0 + c01 only": w0 * (c0 + c01 / 254),
    "c0 + c10 only": w0 * (c0 + c10 / 254),
    "c0 + 2*c01": w0 * (c0 + 2 * c01 / 254),
    "c0 + c2/254 (class swap)": w0 * (c0 + (c11 + c02 + c20) / 254 + (c01 + c10) / 254 ** 2),
}
iu = np.triu_indices(N, 1)
for k, v in cands.items():
    print(f"{k:28s} max |G_gpu - cand| / max|G| = {np.abs(G - v)[iu].max() / np.abs(G).max():.3e}")
print("G_gpu[0,1:6]", G[0, 1:6], "\nfull   ", cands["full"][0, 1:6], "\nc0 only", cands["c0 only"][0, 1:6])
print("digits of a=1..5:", f0[:5], f1[:5], f2[:5])
