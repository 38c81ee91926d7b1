import numpy as np, torch, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine, synth
from fununifrac_b200.factory import make_emd_input, make_tree
from oracle import oracle_c
with tempfile.TemporaryDirectory() as tmp:
    path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
ctx = engine.TreeContext(inp.parent, inp.edge_length)
leaves = synth.leaves_of(inp.parent)
for N in (130, 300):
    X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=7)
    X[5] = X[4] * (1 + 1e-4 * np.random.default_rng(0).random(ctx.n)); X[5] /= X[5].sum()
    X[N - 1] = X[7]
    Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, True), True)
    job = ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L2)
    D, q, _ = ctx.pairwise_gram(job.PT, N)
    torch.cuda.synchronize()
    G = D.cpu().numpy()
    iu = np.triu_indices(N, 1)
    rel = np.abs(G[iu] - Dref[iu]) / np.maximum(Dref[iu], 1e-300)
    big = Dref[iu] > 1e-6
    print(f"N={N} raw gram: max rel err ordinary pairs {rel[big].max():.3e}, median {np.median(rel[big]):.3e}; symmetric {np.array_equal(G, G.T)}; diag zero {not np.diag(G).any()}")
    qs = q.cpu().numpy()[:N]; ratio = Dref / (qs[:, None] + qs[None, :])
    print("   D/(qi+qj) quantiles", np.quantile(ratio[iu], [0, 0.01, 0.5, 1]))
    D2 = ctx.distances_gram(job).cpu().numpy()
    rel2 = np.abs(D2[iu] - Dref[iu]) / np.maximum(Dref[iu], 1e-300)
    print(f"   refined: max rel err {rel2.max():.3e} gram-refined {ctx.gram_refined} D[7,N-1]={D2[7, N-1]}")
# timing at larger N (random device data)
import time
for N in (2048, 8192):
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.zeros((N, ctx.n), dtype=torch.float64, device="cuda")
    lv = torch.from_numpy(leaves).cuda()
    vals = torch.rand((N, len(leaves)), generator=g, device="cuda", dtype=torch.float64) * (torch.rand((N, len(leaves)), generator=g, device="cuda") < 0.2)
    X[:, lv] = vals; X /= X.sum(1, keepdim=True); del vals
    job = ctx.push_up_job(X, engine.MODE_L2)
    D = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    ctx.pairwise_gram(job.PT, N, D=D); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.pairwise_gram(job.PT, N, D=D); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); T = (N + 127) // 128; tiles = T * (T + 1) // 2
    ex = 3 * 2 * 128 * 128 * ctx.n * tiles
    print(f"gram N={N}: {ms:.2f} ms total (prep+tile+mirror); executed {ex/ms/1e9:.1f} TFLOP/s TF32, {N*(N-1)/2/ms/1e3:.3e} pairs/s")
    D2 = torch.zeros_like(D)
    e0.record(); ctx.pairwise(job.PT, N, engine.MODE_L2, D=D2); e1.record(); torch.cuda.synchronize()
    print(f"   direct FFMA2 kernel: {e0.elapsed_time(e1):.2f} ms; max rel diff gram vs direct {( (D-D2).abs() / D2.clamp_min(1e-300)).max().item():.3e}")
    del X, job, D, D2
