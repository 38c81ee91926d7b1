import numpy as np, torch, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine, synth
from fununifrac_b200.factory import make_emd_input, make_tree
with tempfile.TemporaryDirectory() as tmp:
    path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
ctx = engine.TreeContext(inp.parent, inp.edge_length)
leaves = synth.leaves_of(inp.parent)
internal = np.setdiff1d(np.arange(ctx.n), leaves)
N = 512
X = synth.synth_profiles_dense(leaves, ctx.n, N, seed=11)
job = ctx.push_up_job(torch.from_numpy(X).cuda(), engine.MODE_L2)
iu = np.triu_indices(N, 1)
def run(PT, tag):
    ref = ctx.pairwise_f64(PT.double(), N, engine.MODE_L2).cpu().numpy()
    for fs in ("4", "1"):
        os.environ["FUF_GRAM_FLUSH"] = fs
        D, q, _ = ctx.pairwise_gram(PT, N)
        G = D.cpu().numpy()
        rel = np.abs(G[iu] - ref[iu]) / ref[iu]
        qs = q.cpu().numpy()[:N]
        ratio = ref[iu] / (qs[:, None] + qs[None, :])[iu]
        w = np.argmax(rel)
        print(f"{tag} flush={fs}: max rel {rel.max():.3e} median {np.median(rel):.3e}; ratio D/(qi+qj) at worst pair {ratio[w]:.3f}, min ratio {ratio.min():.3f}, corr(rel, 1/ratio)={np.corrcoef(rel, 1/ratio)[0,1]:.2f}")
run(job.PT, "all nodes     ")
PTl = job.PT.clone(); PTl[torch.from_numpy(internal).cuda()] = 0
run(PTl, "leaves only   ")
PTi = job.PT.clone(); PTi[torch.from_numpy(leaves).cuda()] = 0
run(PTi, "internal only ")
