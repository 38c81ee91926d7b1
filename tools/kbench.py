"""Kernel micro-timings on the synthetic full tree (device-generated random profiles; not the bench workload)."""
import os, sys, tempfile, argparse
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fununifrac_b200 import engine, synth
from fununifrac_b200.factory import make_emd_input, make_tree

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--what", default="push,pair")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
with tempfile.TemporaryDirectory() as tmp:
    path = synth.synth_ko_tree(os.path.join(tmp, "t.txt"), seed=0)
    inp = make_emd_input.tree_to_EMDU_input(make_tree.import_graph(path), "ko00001")
ctx = engine.TreeContext(inp.parent, inp.edge_length)
n, N = ctx.n, a.N
leaves = torch.from_numpy(synth.leaves_of(inp.parent)).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
X = engine.profile_matrix(N, n, torch.float64, "cuda") if not os.environ.get("KB_UNPADDED") else torch.zeros((N, n), dtype=torch.float64, device="cuda")
vals = torch.rand((N, len(leaves)), generator=g, device="cuda", dtype=torch.float64)
vals = vals * (torch.rand((N, len(leaves)), generator=g, device="cuda") < 0.2)
X[:, leaves] = vals
X /= X.sum(1, keepdim=True)
del vals
def timeit(f, reps=a.reps):
    f(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        e[i].record(); f()
    e[reps].record(); torch.cuda.synchronize()
    return min(e[i].elapsed_time(e[i + 1]) for i in range(reps))
job = engine.Pushed(ctx, N, True)
job.center = ctx.push_up_center(X, a.mode)
job.P64T = ctx.new_profile_buffer(N, dtype=torch.float64)
if "push" in a.what:
    ms = timeit(lambda: ctx.push_up(X, a.mode, job.PT, 0, job.center, job.P64T, job.err_stat))
    print(f"push_up f64->f32+f64+stat N={N}: {ms:.3f} ms  {N*n*20/ms/1e6:.0f} GB/s (20 B/entry)")
    ms = timeit(lambda: ctx.push_up(X, a.mode, job.PT, 0, job.center, None, job.err_stat))
    print(f"push_up f64->f32+stat     N={N}: {ms:.3f} ms  {N*n*12/ms/1e6:.0f} GB/s (12 B/entry)  [TMA kernel]")
    os.environ["FUF_PUSHUP_NO_TMA"] = "1"
    ms = timeit(lambda: ctx.push_up(X, a.mode, job.PT, 0, job.center, None, job.err_stat))
    print(f"push_up f64->f32+stat     N={N}: {ms:.3f} ms  {N*n*12/ms/1e6:.0f} GB/s (12 B/entry)  [chunk kernel]")
    del os.environ["FUF_PUSHUP_NO_TMA"]
    ms = timeit(lambda: ctx.push_up(X, a.mode, job.PT, 0, job.center, None, None))
    print(f"push_up f64->f32 no stat  N={N}: {ms:.3f} ms  {N*n*12/ms/1e6:.0f} GB/s (12 B/entry)  [TMA kernel]")
    a_ = torch.empty(N * n * 12 // 16, dtype=torch.float64, device="cuda"); b_ = torch.empty_like(a_)
    ms = timeit(lambda: b_.copy_(a_))
    print(f"torch copy of {a_.numel()*8/1e9:.2f} GB (read + write = {a_.numel()*16/1e9:.2f} GB): {ms:.3f} ms  {a_.numel()*16/ms/1e6:.0f} GB/s")
    del a_, b_
    X32 = engine.profile_matrix(N, n, torch.float32, "cuda"); X32.copy_(X)
    ms = timeit(lambda: ctx.push_up(X32, a.mode, job.PT, 0, job.center, None, None))
    print(f"push_up f32->f32          N={N}: {ms:.3f} ms  {N*n*8/ms/1e6:.0f} GB/s (8 B/entry)")
    del X32
ctx.push_up(X, a.mode, job.PT, 0, job.center, job.P64T, job.err_stat)
if "pair" in a.what:
    D = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    ms = timeit(lambda: ctx.pairwise(job.PT, N, a.mode, D=D), reps=3)
    pairs = N * (N - 1) / 2
    print(f"pairwise mode={a.mode} N={N}: {ms:.3f} ms  {pairs/ms/1e3:.4g} pairs/s  {2*pairs*n/ms/1e9/37.22496:.4f} of fp32 peak")
    ms = timeit(lambda: ctx.refine(job.P64T, N, a.mode, job.err_stat, D), reps=3)
    print(f"refine scan N={N}: {ms:.3f} ms, refined {ctx.refine_count()}")
if "host" in a.what:
    import time
    Xh = torch.empty((N, n), dtype=torch.float64, pin_memory=True); Xh.copy_(X); Xn = Xh.numpy()
    Dp = torch.zeros((N, N), dtype=torch.float64, pin_memory=True); Dn = Dp.numpy()
    for zc in ("",) if os.environ.get("KB_BANDS") else ("", "1"):
        for bands in (os.environ["KB_BANDS"].split(",") if os.environ.get("KB_BANDS") else ("1", "6", "10", "16")):
            os.environ["FUF_HOST_BANDS"] = bands
            if zc: os.environ["FUF_ZEROCOPY"] = "1"
            else: os.environ.pop("FUF_ZEROCOPY", None)
            ctx.all_pairs_host(Xn, a.mode, Dn)
            t = time.perf_counter()
            for _ in range(3): ctx.all_pairs_host(Xn, a.mode, Dn)
            dt = (time.perf_counter() - t) / 3
            print(f"host path N={N} bands={bands} zero_copy={'yes' if zc else 'no'}: {dt*1e3:.1f} ms  timing(h2d+push first band, rest, total)={ctx.last_timing()}")
