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
sets = {}
X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=11); sets["weighted"] = X
Xu = X.copy(); Xu[Xu > 0] = 1; sets["unweighted"] = Xu
g = torch.Generator(device="cuda").manual_seed(0)
Nb = 8192
Xb = torch.zeros((Nb, ctx.n), dtype=torch.float64, device="cuda"); lv = torch.from_numpy(leaves).cuda()
vals = torch.rand((Nb, len(leaves)), generator=g, device="cuda", dtype=torch.float64) * (torch.rand((Nb, len(leaves)), generator=g, device="cuda") < 0.2)
Xb[:, lv] = vals; Xb /= Xb.sum(1, keepdim=True); del vals
jobb = ctx.push_up_job(Xb, engine.MODE_L2); Db = torch.zeros((Nb, Nb), dtype=torch.float64, device="cuda")
for nd in ("64", "128", "256", "512", "1024"):
    os.environ["FUF_GRAM_DIRECT"] = nd
    msg = f"direct rows {nd}:"
    for tag, Xs in sets.items():
        Dref = oracle_c.pairwise(oracle_c.push_up(inp.parent, inp.edge_length, Xs, True), True)
        job = ctx.push_up_job(torch.from_numpy(Xs).cuda(), engine.MODE_L2)
        D, q, _ = ctx.pairwise_gram(job.PT, N, err_stat=job.err_stat)
        rel = np.abs(D.cpu().numpy()[iu] - Dref[iu]) / Dref[iu]
        msg += f" {tag} max {rel.max():.2e} med {np.median(rel):.2e};"
    ctx.pairwise_gram(jobb.PT, Nb, D=Db); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.pairwise_gram(jobb.PT, Nb, D=Db); e1.record(); torch.cuda.synchronize()
    print(msg, f"N={Nb} total {e0.elapsed_time(e1):.2f} ms (tile kernel {ctx.gram_last_kernel()[0]:.2f})")
