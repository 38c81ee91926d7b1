import numpy as np, torch, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine, synth
from fununifrac_b200.factory import make_emd_input, make_tree
with tempfile.TemporaryDirectory() as tmp:
    path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
ctx = engine.TreeContext(inp.parent, inp.edge_length)
leaves = synth.leaves_of(inp.parent)
N = 260
X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=7)
rng = np.random.default_rng(7)
for base, copies, jitter in ((0, 6, 1e-3), (10, 5, 1e-5), (130, 4, 1e-7), (200, 3, 0.0)):
    for c in range(1, copies):
        X[base + c] = X[base] * (1 + jitter * rng.random(ctx.n))
        X[base + c] /= X[base + c].sum()
for mode in (0, 1):
    job = ctx.push_up_job(torch.from_numpy(X).cuda(), mode)
    raw = ctx.pairwise(job.PT, N, mode).cpu().numpy()
    Dn = ctx.all_pairs_host(X, mode, refine=False)
    d = np.abs(Dn - raw)
    print("mode", mode, "noref host vs raw: ndiff", (d > 0).sum(), "max", d.max(), "where", np.argwhere(d > 0)[:10].tolist())
    Dr = ctx.all_pairs_host(X, mode, refine=True)
    D = ctx.distances(job, mode).cpu().numpy()
    print("refined host vs dev equal:", np.array_equal(Dr, D), "refined count", ctx.refine_count())
    Dn = ctx.all_pairs_host(X, mode, refine=False)
    d = np.abs(Dn - raw)
    print("again noref host vs raw: ndiff", (d > 0).sum(), "max", d.max(), "where", np.argwhere(d > 0)[:10].tolist())
    P64T = ctx.push_up_f64(torch.from_numpy(X).cuda(), mode)
    a, b = job.P64T[:, :N].cpu().numpy(), P64T[:, :N].cpu().numpy()
    d = np.abs(a - b)
    print("P64T max abs", d.max(), "max rel", (d / np.maximum(np.abs(b), 1e-300)).max())
