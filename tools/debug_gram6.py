import numpy as np, torch, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine, synth
from fununifrac_b200.factory import make_emd_input, make_tree
from oracle import oracle_c
with tempfile.TemporaryDirectory() as tmp:
    path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
ctx = engine.TreeContext(inp.parent, inp.edge_length)
leaves = synth.leaves_of(inp.parent)
N = 768
iu = np.triu_indices(N, 1)
X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=11)
rng = np.random.default_rng(3)
# graded similarity: sample 2i+1 = mix of sample 2i and an independent one
for i in range(0, 200, 2):
    lam = 10 ** rng.uniform(-3, 0)
    X[i + 1] = (1 - lam) * X[i] + lam * X[i + 1]
for tag, Xs in (("weighted", X), ("unweighted", np.where(X > 0, 1.0, 0.0))):
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, Xs, True), True)
    job = ctx.push_up_job(torch.from_numpy(Xs).cuda(), engine.MODE_L2)
    D, q, err = ctx.pairwise_gram(job.PT, N, err_stat=job.err_stat)
    hs = q.cpu().numpy()[:N]
    ratio = (Dref / (hs[:, None] + hs[None, :]))[iu]
    abs_err = np.abs(D.cpu().numpy()[iu] - Dref[iu])
    rel = abs_err / np.maximum(Dref[iu], 1e-300)
    ok = Dref[iu] > 0
    es = err.cpu().numpy()[:N]; rr = (np.sqrt(es)[:, None] + np.sqrt(es)[None, :])[iu]
    round_bound = 2 * np.sqrt(Dref[iu]) * rr + rr ** 2
    excess = np.maximum(abs_err - round_bound, 0) / np.sqrt(np.outer(hs, hs))[iu]
    print(f"{tag}: truncation: max (|err| - rounding bound)+ / (h_i h_j) = {excess.max():.3e} (rigorous bound 1.44e-9); D/(h2i+h2j) quantiles {np.quantile(ratio[ok], [0, 0.001, 0.01, 0.5])}")
    for lo, hi in ((0, 0.005), (0.005, 0.02), (0.02, 0.05), (0.05, 0.2), (0.2, 10)):
        m = ok & (ratio >= lo) & (ratio < hi)
        if m.any(): print(f"   ratio in [{lo},{hi}): {m.sum()} pairs, max rel {rel[m].max():.3e}")
    Df = ctx.distances(job, engine.MODE_L2, gram=True).cpu().numpy()
    relf = np.abs(Df[iu] - Dref[iu]) / np.maximum(Dref[iu], 1e-300)
    print(f"   after refinement: max rel {relf[ok].max():.3e}, refined {ctx.gram_refined}")
