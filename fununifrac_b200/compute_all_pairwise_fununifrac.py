#!/usr/bin/env python
"""Driver of the all-pairs FunUniFrac path: validate → tree → ingest → push-up → pairwise → save.

Same ``main(args)`` contract, log milestones and output files as the reference's
``src/compute_all_pairwise_fununifrac.py:20-108``; the compute section (:64-73) is one call into the
CUDA engine instead of N Python push-ups and N(N-1)/2 numpy pair evaluations.
"""
import datetime
import glob
import logging
import os
import pickle
from pathlib import Path

import numpy as np

import fununifrac_b200.factory.make_emd_input as make_emd_input
import fununifrac_b200.factory.make_tree as make_tree
import fununifrac_b200.utility.constant as constant
import fununifrac_b200.utility.data_paths as data
import fununifrac_b200.utility.kegg_db as kegg_db
from fununifrac_b200 import engine, sparse_npz
from fununifrac_b200.algorithms.emd_unifrac import EarthMoverDistanceUniFracSolver
from fununifrac_b200.objects.func_tree import FuncTreeEmduInput


def _ingest(files, input, abundance_key, unweighted, num_threads, python_ingest):
    """Un-pushed profile matrix [len(files), n] of gather CSVs, in file order, in PAGE-LOCKED host memory: the rows are
    parsed straight into the buffer the GPU then reads by DMA (a pageable matrix would cross PCIe through the driver's
    staging buffers, serialised with the compute)."""
    n = len(input.basis)
    if not files:
        return np.zeros((0, n))
    X = engine.profile_matrix(len(files), n, pin_memory=True).numpy()     # rows 128-byte aligned
    if python_ingest:
        # the reference's own route: pandas.read_csv + a process pool (src/factory/make_emd_input.py:58-122)
        results = make_emd_input.parallel_functional_profile_to_vector(num_threads, files, input, abundance_key,
                                                                       unweighted)
        got = [f for f, _ in results]
        for i, (_, P) in enumerate(results):
            X[i, :] = P
    else:
        # native ingest (csrc/ingest.cu): C++ thread pool, only the two columns the path uses are tokenised;
        # rows are bit-identical to the route above
        X[...] = 0
        got, X = make_emd_input.native_profiles_to_matrix(files, input, abundance_key, unweighted, num_threads, out=X)
    assert got == list(files)
    return X


def _main_multi_gpu(args, input, fun_files, is_L2, make_diffab, save_Ppushed, python_ingest):
    """The compute section under ``torchrun`` (one process per GPU of ONE node): every rank ingests and pushes up its own
    block of files, the pushed profiles are exchanged over NVLink, every rank computes its tile rows and stores them into
    one shared page-locked host matrix (``sharded.ShardedAllPairs.step_host``); rank 0 writes the output files from it.
    ``--diffab``: every rank streams its own contiguous range of pairs over its own PCIe link and compacts it; the
    entries are joined on rank 0.  ``--Ppushed``: every rank pushes its own files up in float64 (reference order)."""
    import torch
    import torch.distributed as dist

    from fununifrac_b200 import sharded

    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    if local_world != world:
        # the result matrix is one shared-memory mapping: that exists on one node only.  Fail before any rank waits.
        raise RuntimeError(f"the multi-GPU driver shards over the GPUs of ONE node (LOCAL_WORLD_SIZE={local_world}, "
                           f"WORLD_SIZE={world}); launch one job per node")
    torch.cuda.set_device(local_rank)
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        N = len(fun_files)
        mode = engine.MODE_L2 if is_L2 else engine.MODE_L1
        ctx = engine.context_for(input, local_rank)
        solver = EarthMoverDistanceUniFracSolver(device=local_rank)
        job = sharded.ShardedAllPairs(sharded.CudaOps(ctx), N, mode, world, rank, dist)
        X = _ingest(fun_files[job.lo:job.hi], input, args.abundance_key, args.unweighted, args.threads, python_ingest)
        shared = sharded.SharedHostMatrix(N, world, rank, dist)
        dists = job.step_host(X, shared)              # the shared mapping itself: nobody copies N x N doubles
        diffabs_sparse = Ps_pushed = None
        if save_Ppushed:
            rows = [None] * world
            dist.gather_object(solver.push_up_batch(X, input, is_L2), rows if rank == 0 else None, dst=0)
            if rank == 0:
                pushed = np.concatenate(rows, axis=0)
                Ps_pushed = {file: pushed[i] for i, file in enumerate(fun_files)}
        if make_diffab:
            # every rank needs every sample's float64 profile: all-gather the un-pushed shards (N x n doubles; --diffab
            # output is N^2 n / 2 entries, so N is small whenever this is feasible at all) and push them up locally
            dev = f"cuda:{local_rank}"
            Xs = torch.zeros((job.shard, ctx.n), dtype=torch.float64, device=dev)
            Xs[:job.hi - job.lo] = torch.from_numpy(X).to(dev)
            Xall = torch.empty((world * job.shard, ctx.n), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(Xall, Xs)
            keep = torch.cat([torch.arange(r * job.shard, r * job.shard + hi - lo, device=dev)
                              for r in range(world) for lo, hi in [sharded.sample_range(N, world, r, job.shard)]])
            pushed = ctx.push_up_job(Xall[keep].contiguous(), mode)
            p0, p1 = engine.diffab_pair_range(N, world, rank)
            part = solver.diffab_entries(ctx, pushed.P64T, N, mode, pair_begin=p0, pair_end=p1)
            parts = [None] * world
            dist.gather_object(part, parts if rank == 0 else None, dst=0)
            if rank == 0:
                diffabs_sparse = solver.coo_from_entries(parts, N, ctx.n)
        if rank == 0:
            _write_outputs(args, input, fun_files, dists, diffabs_sparse, Ps_pushed)
        dist.barrier()                                # the other ranks keep their mappings until the file is written
        shared.close()
    finally:
        if own_group:
            dist.destroy_process_group()


def _sample_files(args):
    """Validation of the inputs (reference :37-46: unknown BRITE id -> ValueError, missing edge list or no matching sample
    file -> Exception) and the sample order: sorted glob, which is the row / column order of the output matrix."""
    if args.brite not in kegg_db.instance.brites:
        raise ValueError(f"{args.brite} is not a valid BRITE ID. Choices are: {kegg_db.instance.brites}")
    data.check_data_abspath(args.edge_list)
    pattern = os.path.join(args.file_dir, args.file_pattern)
    data.check_data_abspaths(pattern)
    return sorted(glob.glob(pattern))


def _write_outputs(args, input: FuncTreeEmduInput, fun_files, dists, diffabs_sparse, Ps_pushed):
    """The files of reference :76-108, same names (utility/constant.py) and contents."""
    run_id = args.out_id if args.out_id else datetime.datetime.now().strftime('%Y%m%d-%H%M%S')
    target = Path(args.out_dir)
    target.mkdir(parents=True, exist_ok=True)

    matrix_path = target / constant.FUNUNIFRAC_OUT__MAIN_FILE_NAME.format(run_id)
    np.save(matrix_path, dists)
    (target / constant.FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME.format(run_id)).write_text(
        "".join(f"{sample}\n" for sample in fun_files))
    logging.info("distance matrix (%d x %d) written to %s", len(fun_files), len(fun_files), matrix_path)
    if diffabs_sparse is not None:
        diffab_path = str(target / constant.FUNUNIFRAC_OUT__DIFFAB_FILE_NAME.format(run_id))
        sparse_npz.save_npz(diffab_path, diffabs_sparse)
        (target / constant.FUNUNIFRAC_OUT__DIFFAB_NODES_FILE_NAME.format(run_id)).write_text(
            "".join(f"{input.idx_to_node[node]}\n" for node in input.basis))
        logging.info("difference vectors (%d non-zero entries) written to %s.npz", diffabs_sparse.nnz, diffab_path)
    if Ps_pushed is not None:
        pushed_path = target / constant.FUNUNIFRAC_OUT__PUSH_FILE_NAME.format(run_id)
        with open(pushed_path, 'wb') as f:
            pickle.dump(Ps_pushed, f)
        logging.info("pushed profiles written to %s", pushed_path)


def main(args):
    logging.basicConfig(level=args.loglevel, format='%(asctime)s %(levelname)s: %(message)s')
    make_diffab, is_L2, save_Ppushed = args.diffab, args.L2, args.Ppushed
    validate = bool(getattr(args, "validate", False))
    python_ingest = bool(getattr(args, "python_ingest", False))
    fun_files = _sample_files(args)

    logging.info("reading the tree %s", args.edge_list)
    solver = EarthMoverDistanceUniFracSolver(validate=validate)
    tree = make_tree.import_graph(args.edge_list, directed=True)
    input: FuncTreeEmduInput = make_emd_input.tree_to_EMDU_input(tree, args.brite)
    logging.info("%d tree nodes below %s, %d samples", len(input.basis), args.brite, len(fun_files))

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1 and not validate:
        # launched with torchrun: shard over the GPUs of the node; rank 0 writes the output files
        return _main_multi_gpu(args, input, fun_files, is_L2, make_diffab, save_Ppushed, python_ingest)
    if rank != 0:
        logging.info("--validate runs on one GPU: this rank has nothing to do")
        return

    X = _ingest(fun_files, input, args.abundance_key, args.unweighted, args.threads, python_ingest)
    logging.info("profiles read; computing all pairs on the GPU")
    mode = engine.MODE_L2 if is_L2 else engine.MODE_L1
    N = len(fun_files)
    diffabs_sparse = Ps_pushed = None
    if save_Ppushed:
        pushed = solver.push_up_batch(X, input, is_L2)  # float64, the reference's operation order
        Ps_pushed = {file: pushed[i] for i, file in enumerate(fun_files)}
    if not make_diffab:
        # page-locked result: finished bands of the matrix leave the device while later ones are still computed
        dists = solver.all_pairs(X, input, is_L2, engine.pinned_empty((N, N), engine._torch().float64)[1] if N else None)
    else:
        torch = engine._torch()
        ctx = engine.context_for(input)
        Xd = torch.from_numpy(X).to(f"cuda:{ctx.device}")
        if validate:
            PT = ctx.push_up_f64(Xd, mode)
            dists = ctx.pairwise_f64(PT, N, mode).cpu().numpy()
        else:
            job = ctx.push_up_job(Xd, mode)
            # the difference vectors come from the float64 profiles (the reference's P - Q is float64: fp32 operands
            # would lose near-equal samples' differences altogether); the distances from the fp32 tiles + refinement
            PT = job.P64T
            dists = ctx.distances(job, mode).cpu().numpy()
        diffabs_sparse = solver.diffab_coo(ctx, PT, N, mode)
    _write_outputs(args, input, fun_files, dists, diffabs_sparse, Ps_pushed)
