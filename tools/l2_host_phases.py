#!/usr/bin/env python
"""Phase timing of the --L2 host-buffer path on one GPU (fuf_all_pairs_host, Gram route): the CUDA-event phases the
library records (fuf_last_timing) and the wall clock of the call, for N samples (FUF_N, default 50,000).

    python tools/l2_host_phases.py            # one line per repetition
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fununifrac_b200 import engine  # noqa: E402


def main():
    N = int(os.environ.get("FUF_N", "50000"))
    reps = int(os.environ.get("FUF_REPS", "3"))
    torch.cuda.set_device(0)
    inp, leaves = bench.build_tree()
    n = len(inp.basis)
    ctx = engine.TreeContext(inp.parent, inp.edge_length, device=0)
    X = bench.device_profiles(leaves, n, N, seed=1)
    X_pin = torch.empty((N, n), dtype=torch.float64, pin_memory=True)
    X_pin.copy_(X[:, :n])
    del X
    torch.cuda.empty_cache()
    D_pin = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
    Xh, Dh = X_pin.numpy(), D_pin.numpy()
    for it in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.all_pairs_host(Xh, engine.MODE_L2, Dh)
        wall = (time.perf_counter() - t0) * 1e3
        h2d_push, pair, total = ctx.last_timing()
        gk = ctx.gram_last_kernel()[0]
        print(f"N={N} rep {it}: wall {wall:8.1f} ms | events: h2d+first push-up {h2d_push:7.1f}  rest {pair:7.1f}  "
              f"total {total:7.1f} | gram kernel {gk:7.1f} ms | H2D {N * n * 8 / 1e9:.1f} GB, D2H {N * N * 8 / 1e9:.1f} GB",
              flush=True)
    sym = float(np.abs(Dh[:2048, :2048] - Dh[:2048, :2048].T).max())
    print("symmetry of the first 2048 block:", sym, "diag max", float(np.abs(np.diagonal(Dh)).max()))


if __name__ == "__main__":
    main()
