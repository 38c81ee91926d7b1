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
for fs in ("16", "4", "2", "1"):
    os.environ["FUF_GRAM_FLUSH"] = fs
    D, q, _ = ctx.pairwise_gram(job.PT, N)
    G = D.cpu().numpy()
    rel = np.abs(G[iu] - Dref[iu]) / Dref[iu]
    ctx.pairwise_gram(jobb.PT, Nb, D=Db); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.pairwise_gram(jobb.PT, Nb, D=Db); e1.record(); torch.cuda.synchronize()
    print(f"flush={fs} stages: N={N} max rel {rel.max():.3e} median {np.median(rel):.3e} p99.9 {np.quantile(rel, 0.999):.3e}; N={Nb} {e0.elapsed_time(e1):.2f} ms")
Dd = ctx.pairwise(job.PT, N, engine.MODE_L2).cpu().numpy()
rel = np.abs(Dd[iu] - Dref[iu]) / Dref[iu]
print(f"direct fp32 kernel vs oracle: max rel {rel.max():.3e} median {np.median(rel):.3e}")
P64 = job.P64T[:, :N].double()
c = job.center[:, None]
Dx = ctx.pairwise_f64((job.PT[:, :N].double()).contiguous(), N, engine.MODE_L2).cpu().numpy()
rel = np.abs(Dx[iu] - Dref[iu]) / Dref[iu]
print(f"float64 arithmetic on the fp32-rounded centred operands vs oracle (pure input rounding): max rel {rel.max():.3e} median {np.median(rel):.3e}")
os.environ["FUF_GRAM_FLUSH"] = "4"
D, q, _ = ctx.pairwise_gram(job.PT, N); G = D.cpu().numpy()
rel = np.abs(G[iu] - Dx[iu]) / Dx[iu]
print(f"gram vs float64 on the same fp32 operands (arithmetic error only): max rel {rel.max():.3e} median {np.median(rel):.3e}")
rel = np.abs(Dd[iu] - Dx[iu]) / Dx[iu]
print(f"direct vs float64 on the same fp32 operands: max rel {rel.max():.3e} median {np.median(rel):.3e}")
