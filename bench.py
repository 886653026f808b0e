#!/usr/bin/env python
"""Benchmark of the all-pairs beta-diversity hot path (sample-pair distances per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload weighted|unweighted|braycurtis|jaccard] [--samples n] [--tips t]

Workload at N=1: BASELINE.json configs[1] - weighted normalized UniFrac, 10,000 samples x
100,000 tips (199,999 nodes), random-join binary tree, iid Poisson(0.02) counts.  A "step" is
one full pass of the hot path over that batch: embedding build (CSR scatter + tree
propagation) + all pair-tile kernel launches, inputs already resident in HBM.  With N > 1
ranks (torchrun, one process per GPU) the sample count grows by sqrt(N) so that the pairs per
GPU stay fixed (weak scaling); rank r owns the row stripes dealt to shard r and there is no
inter-GPU collective on the data path.

One JSON line is printed by rank 0.  `value` = pairs / device time of the step (CUDA events
recorded by the library on its own stream, max over ranks); `e2e` = the same metric through
the one-shot C-ABI call with HOST buffers (H2D of tree + CSR and D2H of the condensed result
inside the timed region); `roofline` is for the dominant kernel (k_pair_weighted, FP64-pipe
bound - see DESIGN.md); `cpu_baseline` times the oracle port of the reference's CPU path on a
bounded sample of the same workload.

--impl reference times the reference's own CPU algorithm (the NumPy port in oracle/, which
follows scikit-bio's per-pair NumPy code line by line and is driven by the same scipy pdist
call) on all host cores (row blocks over a process pool), on a bounded sample of the workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# DRAM bytes (read + write) of one launch at the default weighted configuration, from ncu captures
NCU_TRAFFIC_WEIGHTED = {"k_pair_tips": 21.980974e9 + 0.402909e9, "k_pair_weighted_mask": 17.406591e9 + 0.437857e9}
METRIC_NAME = "sample-pair distances/sec"
UNIT = "pairs/s"

WORKLOADS = {
    # name: (metric key, default samples, default tips/features, lambda, description)
    "weighted": ("weighted_unifrac_normalized", 10000, 100000, 0.02,
                 "weighted normalized UniFrac, 10,000 samples x 100,000 tips (BASELINE configs[1])"),
    "unweighted": ("unweighted_unifrac", 10000, 100000, 0.02,
                   "unweighted UniFrac, random-join tree"),
    "weighted_unnorm": ("weighted_unifrac", 10000, 100000, 0.02,
                        "weighted (unnormalized) UniFrac"),
    "braycurtis": ("braycurtis", 10000, 50000, 0.01, "Bray-Curtis, sparse Poisson counts"),
    "jaccard": ("jaccard", 10000, 50000, 0.01, "Jaccard, sparse Poisson counts"),
}


def make_inputs(workload, n, tips, tree="random"):
    from skbio_b200 import synthetic

    key, _, _, lam, _ = WORKLOADS[workload]
    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    if workload in ("weighted", "unweighted", "weighted_unnorm"):
        gen = synthetic.caterpillar_tree if tree == "caterpillar" else synthetic.random_join_tree
        parent, length, names, tip_node = gen(tips, seed=1)
        return dict(kind="tree", key=key, n=n, parent=parent, length=length, indptr=indptr,
                    ids=tip_node[cols].astype(np.int32), cols=cols, vals=vals, tip_node=tip_node,
                    m=len(parent))
    return dict(kind="features", key=key, n=n, indptr=indptr, ids=cols.astype(np.int32),
                cols=cols, vals=vals.astype(np.float64), m=tips)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smmax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_mhz_min": float(min(sm)) if sm else None,
                "sm_max_mhz": float(max(smmax)) if smmax else None,
                "power_w_median": float(np.median(power)) if power else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------
# CPU side (oracle port of the reference path) - bounded samples
# ----------------------------------------------------------------------------------------
_JOB = None  # (X, bl, tips, d, normalized, unweighted): inherited by forked workers, not pickled


def _np_port_rows(rows):
    """Rows [r0, r1) of the pair triangle with the NumPy port (one process)."""
    from functools import partial

    from oracle import oracle_np

    X, bl, tips, d, normalized, unweighted = _JOB
    r0, r1 = rows
    n = X.shape[0]
    out = []
    if unweighted:
        f = partial(oracle_np.unweighted_unifrac_pair, branch_lengths=bl)
    elif normalized:
        def f(u, v):
            ut = np.take(u, tips).sum()
            vt = np.take(v, tips).sum()
            return oracle_np.weighted_unifrac_normalized_pair(u, v, ut, vt, bl, d)
    else:
        def f(u, v):
            ut = np.take(u, tips).sum()
            vt = np.take(v, tips).sum()
            return oracle_np.weighted_unifrac_pair(u, v, ut, vt, bl)[0]
    for i in range(r0, r1):
        for j in range(i + 1, n):
            out.append(f(X[i], X[j]))
    return out


def cpu_reference_pairs_per_s(inp, workload, n_sub, procs):
    """Time the NumPy port (reference algorithm: dense int64 node matrix + one NumPy-ufunc
    pair function call per pair, /root/reference/skbio/diversity/beta/_unifrac.py:340-471) on
    the first n_sub samples of the workload, over `procs` processes."""
    from oracle import oracle_fast, oracle_np
    from skbio_b200 import synthetic

    n_sub = min(n_sub, inp["n"])
    ip = inp["indptr"][: n_sub + 1]
    cols = inp["cols"][: ip[-1]]
    vals = inp["vals"][: ip[-1]]
    P = n_sub * (n_sub - 1) // 2
    t0 = time.perf_counter()
    if inp["kind"] == "tree":
        tips_total = len(inp["tip_node"])
        dense = synthetic.csr_to_dense(ip, cols, vals.astype(np.int64), tips_total)
        unweighted = workload == "unweighted"
        if unweighted:
            dense = (dense > 0).astype(np.int64)
        # _nodes_by_counts / _traverse_reduce: compiled in the reference too (Cython)
        X = oracle_fast.counts_by_node(dense, inp["tip_node"], inp["parent"])
        bl = np.ascontiguousarray(inp["length"], dtype=np.float64)
        tipidx = oracle_np.tip_indices_from_parent(inp["parent"])
        d = oracle_np.tip_distances(bl, inp["parent"], tipidx)
        normalized = workload == "weighted"  # "weighted_unnorm" -> unnormalized
        global _JOB
        _JOB = (X, bl, tipidx, d, normalized, unweighted)
        jobs = _balanced_rows(n_sub, procs)
        if procs > 1:
            import multiprocessing as mp

            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_np_port_rows, jobs)
        else:
            res = [_np_port_rows(j) for j in jobs]
        vec = np.array([x for r in res for x in r])
    else:
        from scipy.spatial.distance import pdist

        dense = synthetic.csr_to_dense(ip, cols, vals, inp["m"])
        with np.errstate(all="ignore"):
            vec = pdist(dense > 0, "jaccard") if workload == "jaccard" else pdist(dense, "braycurtis")
        procs = 1
    dt = time.perf_counter() - t0
    return P / dt, dt, vec, procs


def _balanced_rows(n, parts):
    """Split rows 0..n-2 of the triangle into `parts` ranges of about equal pair count."""
    total = n * (n - 1) // 2
    bounds, start, acc, target = [], 0, 0, total / max(parts, 1)
    for i in range(n - 1):
        acc += n - 1 - i
        if acc >= target * (len(bounds) + 1) and len(bounds) < parts - 1:
            bounds.append((start, i + 1))
            start = i + 1
    bounds.append((start, n - 1))
    return [b for b in bounds if b[1] > b[0]]


def c_oracle_pairs_per_s(inp, workload, n_sub):
    """The C/OpenMP oracle on the same bounded sample (a stronger CPU implementation than
    the reference's Python pair loop; reported for context)."""
    from oracle import oracle_fast
    from skbio_b200 import synthetic

    n_sub = min(n_sub, inp["n"])
    ip = inp["indptr"][: n_sub + 1]
    cols = inp["cols"][: ip[-1]]
    vals = inp["vals"][: ip[-1]]
    P = n_sub * (n_sub - 1) // 2
    if inp["kind"] != "tree":
        return None
    dense = synthetic.csr_to_dense(ip, cols, vals.astype(np.int64), len(inp["tip_node"]))
    bl = np.ascontiguousarray(inp["length"], dtype=np.float64)
    t0 = time.perf_counter()
    if workload == "unweighted":
        X = oracle_fast.counts_by_node((dense > 0).astype(np.int64), inp["tip_node"], inp["parent"])
        oracle_fast.unweighted(X, bl)
    else:
        X = oracle_fast.counts_by_node(dense, inp["tip_node"], inp["parent"])
        oracle_fast.weighted(X, bl, inp["parent"], True)
    dt = time.perf_counter() - t0
    return {"value": P / dt, "unit": UNIT, "threads": oracle_fast.max_threads(),
            "sample": f"first {n_sub} samples, all nodes, {dt:.1f} s"}


# ----------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload
    key, n_def, tips_def, lam, desc = WORKLOADS[workload]
    n = args.samples or int(round(n_def * np.sqrt(args.gpus)))
    tips = args.tips or tips_def
    procs = min(os.cpu_count() or 1, 32)
    n_sub = args.cpu_samples or (12 * procs if workload in ("weighted", "unweighted", "weighted_unnorm")
                                 else 1500)
    # only the first n_sub samples are ever touched: generate just those
    inp = make_inputs(workload, n_sub, tips, args.tree)
    vals = []
    for step in range(args.warmup + args.steps):
        v, dt, _, used = cpu_reference_pairs_per_s(inp, workload, n_sub, procs)
        if step >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = (f"first {n_sub} samples of the workload ({n_sub * (n_sub - 1) // 2} pairs), all "
              f"{inp['m']} nodes/features; NumPy port of the reference pair functions over "
              f"{used} processes")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "samples": n, "tips_or_features": tips, "lambda": lam,
                   "metric_key": key},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch

    from skbio_b200 import _capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    emulate = max(int(args.emulate_shards), 1)  # one process computing shard 0 of `emulate`
    if not _capi.available():
        raise RuntimeError("libskbb200.so missing or no CUDA device: nothing to benchmark "
                           "(the engine has no CPU path)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    workload = args.workload
    key, n_def, tips_def, lam, desc = WORKLOADS[workload]
    n = args.samples or int(round(n_def * np.sqrt(world)))
    tips = args.tips or tips_def
    if args.lam:
        WORKLOADS[workload] = (key, n_def, tips_def, args.lam, desc)
        lam = args.lam
    inp = make_inputs(workload, n, tips, args.tree)
    m = inp["m"]
    P_total = n * (n - 1) // 2
    shard_rank, shard_world = (rank, world) if emulate == 1 else (0, emulate)

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(device)

    def make_problem():
        if inp["kind"] == "tree":
            return _capi.Problem.tree(inp["parent"], inp["length"], inp["indptr"], inp["ids"],
                                      inp["vals"], n, device=device, shard=shard_rank,
                                      n_shards=shard_world)
        return _capi.Problem.features(inp["indptr"], inp["ids"], inp["vals"], n, m, device=device,
                                      shard=shard_rank, n_shards=shard_world)

    # ---------------- device-resident steps ----------------
    prob = make_problem()
    P_shard = prob.result_len()
    for _ in range(args.warmup):
        prob.compute(key)
    sampler = ClockSampler(device)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, pair_ms, embed_ms, prep_ms, launches = [], [], [], [], 0
    for _ in range(args.steps):
        prob.compute(key)  # blocks until the shard's distances are complete in HBM
        t = prob.timings()
        dev_ms.append(t["compute_ms"])
        pair_ms.append(t["pair_ms"])
        embed_ms.append(t["embed_ms"])
        prep_ms.append(t["skip_prep_ms"] + t["tips_ms"])
        launches += t["kernels"]   # every kernel the library launched for this step
        pair_launches_per_step = t["pair_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    step_ms = float(np.mean(dev_ms))
    tt = torch.tensor([step_ms, wall_ms / args.steps], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms_max, wall_step_ms_max = float(tt[0]), float(tt[1])
    P_done = P_total if emulate == 1 else P_shard   # pairs actually computed by this job
    value = P_done / (step_ms_max * 1e-3)

    # ---------------- end to end through the one-shot C-ABI call, host buffers ----------------
    prob.destroy()
    h2d = (inp["indptr"].nbytes + inp["ids"].nbytes + (inp["vals"].nbytes if inp["vals"] is not None else 0)
           + (inp["parent"].nbytes + inp["length"].nbytes if inp["kind"] == "tree" else 0))
    d2h = P_shard * 8
    # one rank: the full condensed vector; several ranks: each rank's own slices, compact
    # (rank r's page-locked buffer holds only what rank r computed)
    compact = world > 1 or emulate > 1
    out = _capi.PinnedBuffer(P_shard if compact else P_total)
    # inputs in page-locked host memory as well: the H2D copies inside the timed region are DMA
    pinned_inputs = {k: _capi.PinnedBuffer.like(inp[k]) for k in ("parent", "length", "indptr", "ids", "vals")
                     if inp.get(k) is not None}
    inp_pageable = inp
    inp = dict(inp)
    for k, b in pinned_inputs.items():
        inp[k] = b.array
    e2e_times = []
    for it in range(5):
        barrier()
        t0 = time.perf_counter()
        with make_problem() as p2:
            p2.compute(key)
            if compact:
                p2.fetch_compact(out.array)
            else:
                p2.fetch(out.array)
        torch.cuda.synchronize(device)
        e2e_times.append(time.perf_counter() - t0)
    inp = inp_pageable
    e2e_s = min(e2e_times)
    if rank == 0:
        print("e2e repetitions (s):", [round(x, 4) for x in e2e_times], file=sys.stderr)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{device}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = P_done / float(te[0])
    checksum = float(np.nansum(out.array[: min(len(out.array), 1 << 20)])) if rank == 0 else 0.0

    line = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        # dominant kernel roofline
        fp64_peak = _capi.microbench(0, device)      # DFMA instructions / s measured now
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        avg_launch_ms = (float(np.mean(pair_ms)) - float(np.mean(prep_ms))) / max(pair_launches_per_step, 1)
        nodes_per_launch = m / max(pair_launches_per_step, 1)
        sm_count = 148
        if workload in ("weighted", "weighted_unnorm", "braycurtis"):
            kernel = "k_pair_weighted_mask" if t["visits_dense"] else "k_pair_weighted_v3"
            units_per_term = 2.0   # DADD (pu - pv) + DFMA (|.| * len + acc)
            bound, unit = "fp64", "Tinst/s (FP64 pipe instructions)"
            peak = fp64_peak
            peak_source = ("b2d_microbench(0): dependent DFMA chains on all SMs, measured in this "
                           "run; nominal 148 SM x 64 lanes x 1.965 GHz = 18.6")
        elif workload == "unweighted":
            kernel = "k_pair_unweighted_lut"
            units_per_term = 1.0 / 8.0   # one 64-bit LUT gather (+ one DADD) per 8 node-pair terms
            bound, unit = "smem", "Tgather/s (64-bit shared-memory gathers)"
            peak = _capi.microbench(2, device)
            peak_source = ("b2d_microbench(2): bank-conflict-free 64-bit shared-memory gathers on "
                           "all SMs, measured in this run; nominal 148 SM x 16/clk x 1.965 GHz = 4.65")
        else:
            kernel = "k_pair_popcount"
            units_per_term = 1.0 / 32.0  # one POPC per 32 feature-pair terms
            bound, unit = "alu", "Tinst/s (POPC)"
            peak = sm_count * 16 * 1.965e9
            peak_source = "nominal 148 SM x 16 POPC/clk x 1.965 GHz (not measured)"
        alg_units = P_shard * nodes_per_launch * units_per_term
        sparse_path = inp["kind"] == "features" and t["nodes_resident"] == 0
        if sparse_path:
            # integer feature table below 2.5 % density: intersection kernel.  Algorithmic unit =
            # one accumulator update per (sample pair, shared feature) = sum_f C(n_f, 2).
            nf = np.bincount(inp["cols"], minlength=m).astype(np.float64)
            matches = float((nf * (nf - 1.0) / 2.0).sum()) * (P_shard / max(P_total, 1))
            kernel = "k_pair_sparse"
            bound, unit = "smem-atomic", "Tatom/s (shared-memory int32 atomic adds)"
            peak = _capi.microbench(3, device)
            peak_source = ("b2d_microbench(3): shared-memory atomicAdd on random cells of a 128x128 "
                           "int32 tile, all SMs, measured in this run")
            alg_units = matches
            units_per_term = matches / max(P_shard * nodes_per_launch, 1)
            sparse_bytes = P_shard * 24.0 + 2.0 * len(inp["cols"]) * 8.0  # result write + finalize r/w, buckets
        # Zero-block skipping: the kernel executes only `visited` of the node-pair terms (the rest
        # has a closed form).  The roofline counts the instructions actually executed; the
        # dense-equivalent rate (all terms / time) is reported beside it.
        visited = (t["visits_executed"] / t["visits_dense"]) if t.get("visits_dense") else 1.0
        dense_equiv = alg_units / (avg_launch_ms * 1e-3) / 1e12
        if not sparse_path:
            alg_units *= visited
        achieved = alg_units / (avg_launch_ms * 1e-3) / 1e12
        elem_bytes = 8.0 if workload in ("weighted", "weighted_unnorm", "braycurtis") else 1.0 / 8.0
        # unique embedding bytes a launch must read at least once
        alg_bytes = n * nodes_per_launch * elem_bytes
        if sparse_path:
            alg_bytes = sparse_bytes
        # DRAM bytes of ONE launch of the dominant kernel from `ncu --set full` captures of this
        # very configuration (profiles/r1_ncu_*_one_launch.txt); null for any other shape
        # per-launch DRAM bytes from `ncu --set full` captures of this very configuration (default
        # shapes only; profiles/r1_ncu_*): null for any other shape
        default_shape = not (args.samples or args.tips or args.lam) and world == 1 and emulate == 1
        TRAFFIC = {"k_pair_unweighted_lut": 0.133509e9 + 0.346140e9}
        TRAFFIC.update(NCU_TRAFFIC_WEIGHTED)
        traffic = TRAFFIC.get(kernel) if default_shape else None
        roofline = {
            "kernel": kernel, "bound": bound, "achieved": achieved,
            "peak": peak / 1e12, "unit": unit,
            "frac": achieved / (peak / 1e12),
            "peak_source": peak_source,
            "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, "
                              "profiles/r1_ncu_*_one_launch.txt" if traffic else None,
            "avg_launch_ms": avg_launch_ms, "launches_per_step": pair_launches_per_step,
            "algorithmic_terms_per_launch": P_shard * nodes_per_launch,
            "units_per_term": units_per_term,
            "terms_visited_frac": visited,
            "dense_equivalent": dense_equiv,
            "skip_prep_and_tips_ms": float(np.mean(prep_ms)),
            "tips_ms": float(t["tips_ms"]),
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                    "achieved_gbs": alg_bytes / (avg_launch_ms * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": peak_src,
                    "note": "not the binding resource: a 128x128 pair tile does 16384 terms per "
                            "256 embedding elements read"},
        }
        # Sparse rows (tips and nodes with few tips below) go through the intersection kernel; when it
        # takes longer than the dense launches it is the dominant kernel and gets the roofline, with
        # the dense kernel's numbers kept beside it.  Its unit of work: one (sample pair, row) match
        # = two 32-bit shared-memory atomic adds.
        if t.get("tips_ms", 0.0) > avg_launch_ms * pair_launches_per_step and t.get("sparse_row_matches", 0.0) > 0:
            atom_peak = _capi.microbench(3, device)
            entry_bytes = 4.0 if workload == "unweighted" else 16.0   # bare keys / {key, fixed-point weight}
            atoms = 2.0 * t["sparse_row_matches"]
            a_rate = atoms / (t["tips_ms"] * 1e-3) / 1e12
            roofline = {
                "kernel": "k_pair_utips" if workload == "unweighted" else "k_pair_tips", "bound": "smem-atomic",
                "achieved": a_rate,
                "peak": atom_peak / 1e12, "unit": "Tatom/s (shared-memory int32 atomic adds)",
                "frac": a_rate / (atom_peak / 1e12),
                "peak_source": "b2d_microbench(3): shared-memory atomicAdd on random cells of a 128x128 int32 "
                               "tile, all SMs, measured in this run",
                "traffic": TRAFFIC.get("k_pair_tips") if default_shape and workload == "weighted" else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the launch, ncu --set full, "
                                  "profiles/r1_ncu_*_one_launch.txt" if default_shape and workload == "weighted" else None,
                "avg_launch_ms": float(t["tips_ms"]), "launches_per_step": 1,
                "matches_per_launch": t["sparse_row_matches"], "units_per_match": 2.0,
                "sparse_rows_max_tips": t["sparse_rows_max_tips"],
                "sparse_row_entries": t["sparse_row_entries"],
                # every bucket is read by about n/128 tiles (as row or as column block), 16 B per entry
                "algorithmic_bytes_per_launch": entry_bytes * t["sparse_row_entries"] * ((n + 127) // 128 + 1) / max(world * emulate, 1),
                "hbm": {"algorithmic_bytes_per_launch": entry_bytes * t["sparse_row_entries"] * ((n + 127) // 128 + 1) / max(world * emulate, 1),
                        "achieved_gbs": entry_bytes * t["sparse_row_entries"] * ((n + 127) // 128 + 1) / max(world * emulate, 1)
                                        / (t["tips_ms"] * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "peak_source": peak_src,
                        "note": "entries are re-read by every tile of their block row / column, mostly from L2; "
                                "not the binding resource"},
                "dense_kernel": roofline,
            }
            # the dense kernel's time is a difference of event times and overlaps with the tail of the
            # intersection kernel on the other stream: a fraction above 1 only says the split is unreliable
            if roofline["dense_kernel"]["frac"] > 1.0:
                roofline["dense_kernel"]["frac"] = None
                roofline["dense_kernel"]["note"] = ("time by difference (pair phase - preparation - intersection "
                                                    "kernel), the two pair kernels overlap on two streams")
        cpu = None
        if not args.no_cpu and world == 1 and emulate == 1:
            n_sub = args.cpu_samples or (64 if inp["kind"] == "tree" else 1000)
            v, dt, vec, used = cpu_reference_pairs_per_s(inp, workload, n_sub, 1)
            # parity spot check of the bench output against the CPU port on the same inputs
            got = np.empty(n_sub * (n_sub - 1) // 2)
            k = 0
            for i in range(n_sub - 1):
                a = i * n - (i + 2) * (i + 1) // 2 + (i + 1)
                cnt = n_sub - 1 - i
                got[k:k + cnt] = out.array[a:a + cnt]
                k += cnt
            owned = ~np.isnan(got) if world == 1 else None
            with np.errstate(all="ignore"):
                rel = np.abs(got - vec) / np.maximum(np.abs(vec), 1e-300)
            max_rel = float(np.nanmax(rel)) if world == 1 and len(rel) else None
            cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                   "sample": f"first {n_sub} samples ({len(vec)} pairs), all {m} nodes/features, "
                             f"{dt:.1f} s; NumPy port of the reference pair functions, 1 process",
                   "max_rel_err_vs_gpu": max_rel,
                   "c_openmp_oracle": c_oracle_pairs_per_s(inp, workload, 256)}
        line = {
            "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc if not (args.samples or args.tips) else f"{key} custom size",
                       "metric_key": key, "samples": n, "tips_or_features": tips, "nodes": m,
                       "lambda": lam, "pairs": P_total, "pairs_computed": P_done,
                       "sharding": f"row-stripe snake x{world}" if emulate == 1 else
                                   f"shard 0 of {emulate} emulated on one GPU",
                       "l2_policy": "inputs larger than L2 (embedding %.1f GB per pass)"
                                    % (n * m * elem_bytes / 1e9)},
            "wall_ms_per_step": wall_step_ms_max,
            "embed_ms": float(np.mean(embed_ms)), "pair_ms": float(np.mean(pair_ms)),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "seconds": float(te[0]),
                    "host_buffers": "page-locked (b2d_host_alloc) for inputs and result",
                    "timed": "create (H2D) + compute + fetch (D2H) + destroy, best of 5"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "checksum_first_1M": checksum,
        }
    out.free()
    for b in pinned_inputs.values():
        b.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="weighted", choices=list(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--tips", type=int, default=0)
    ap.add_argument("--cpu-samples", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lam", type=float, default=0.0)
    ap.add_argument("--tree", default="random", choices=["random", "caterpillar"])
    ap.add_argument("--emulate-shards", type=int, default=1,
                    help="compute only shard 0 of N on this GPU (sizing runs for multi-GPU configs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
