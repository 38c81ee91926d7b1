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
N = 512
X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=11)
Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, X, True), True)
job = ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L2)
iu = np.triu_indices(N, 1)
g = torch.Generator(device="cuda").manual_seed(0)
Nb = 4096
Xb = torch.zeros((Nb, ctx.n), dtype=torch.float64, device="cuda"); lv = torch.from_numpy(leaves).cuda()
vals = torch.rand((Nb, len(leaves)), generator=g, device="cuda", dtype=torch.float64) * (torch.rand((Nb, len(leaves)), generator=g, device="cuda") < 0.2)
Xb[:, lv] = vals; Xb /= Xb.sum(1, keepdim=True); del vals
jobb = ctx.push_up_job(Xb, engine.MODE_L2); Db = torch.zeros((Nb, Nb), dtype=torch.float64, device="cuda")
D, h2, _ = ctx.pairwise_gram(job.PT, N)
G = D.cpu().numpy()
rel = np.abs(G[iu] - Dref[iu]) / Dref[iu]
ctx.pairwise_gram(jobb.PT, Nb, D=Db); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.pairwise_gram(jobb.PT, Nb, D=Db); e1.record(); torch.cuda.synchronize()
ms, fl = ctx.gram_last_kernel()
print(f"int8 gram: N={N} max rel {rel.max():.3e} median {np.median(rel):.3e}; N={Nb} total {e0.elapsed_time(e1):.2f} ms, tile kernel {ms:.2f} ms = {fl/ms/1e9:.0f} TOP/s int8")
Dd = torch.zeros_like(Db)
e0.record(); ctx.pairwise(jobb.PT, Nb, engine.MODE_L2, D=Dd); e1.record(); torch.cuda.synchronize()
print(f"direct fp32 kernel N={Nb}: {e0.elapsed_time(e1):.2f} ms; max rel diff gram vs direct {((Db-Dd).abs()/Dd.clamp_min(1e-300)).max().item():.3e}")
