#!/usr/bin/env python
"""Benchmark of the all-pairs beta-diversity hot path (sample-pair distances per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload weighted|unweighted|braycurtis|jaccard] [--samples n] [--tips t]
                    [--configs auto|none|cfg3,cfg4,cfg5]

Headline workload at N=1: BASELINE.json configs[1] - weighted normalized UniFrac, 10,000
samples x 100,000 tips (199,999 nodes), random-join binary tree, iid Poisson(0.02) counts.  A
"step" is one full pass of the hot path over that batch: embedding build (CSR scatter + tree
propagation) + all pair-tile kernel launches, inputs already resident in HBM.  With N > 1 ranks
(torchrun, one process per GPU) the sample count grows by sqrt(N) so that the pairs per GPU stay
fixed (weak scaling); rank r owns the row stripes dealt to shard r and there is no inter-GPU
collective on the data path.

One JSON line is printed by rank 0.  `value` = pairs / device time of the step (CUDA events
recorded by the library on its own stream, max over ranks); `e2e` = the same metric through the
one-shot C-ABI path with HOST buffers (H2D of tree + CSR and D2H of the condensed result inside
the timed region); `api_e2e` = the public `beta_diversity()` call itself (Newick text + sparse
table + taxa names in, DistanceMatrix out); `roofline` is for the dominant kernel;
`cpu_baseline` times scikit-bio's own `beta_diversity` (the unmodified reference built by
tools/build_ref.sh into oracle/_ref) on a bounded sample of the same workload; `parity` is a
check of every rank's first owned row stripe against the CPU oracle (at every N).

The same line carries a `configs` block filled by the same invocation with the other BASELINE
configurations at FULL size: cfg3 (Bray-Curtis and Jaccard, 50,000 x 50,000 sparse counts,
strong scaling: the same problem at every N), and at N = 8 cfg4 (unweighted UniFrac, 100,000
samples x 500,000 tips) and cfg5 (weighted unnormalized UniFrac, 25,000 samples, 1M-tip
caterpillar tree) - each with pairs/s, e2e, the roofline of its dominant kernel and a parity
sample.

--impl reference times scikit-bio's own CPU `beta_diversity` (oracle/_ref; the NumPy port in
oracle/ only when that build is missing) on all host cores: every worker process runs the
reference's public API on its own disjoint block of samples of the workload.
"""

import argparse
import io
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

METRIC_NAME = "sample-pair distances/sec"
UNIT = "pairs/s"
TREE_KEYS = ("weighted_unifrac_normalized", "weighted_unifrac", "unweighted_unifrac")

WORKLOADS = {
    # name: (metric key, default samples, default tips/features, lambda, tree, description)
    "weighted": ("weighted_unifrac_normalized", 10000, 100000, 0.02, "random",
                 "weighted normalized UniFrac, 10,000 samples x 100,000 tips (BASELINE configs[1])"),
    "unweighted": ("unweighted_unifrac", 10000, 100000, 0.02, "random",
                   "unweighted UniFrac, random-join tree"),
    "weighted_unnorm": ("weighted_unifrac", 10000, 100000, 0.02, "random",
                        "weighted (unnormalized) UniFrac"),
    "braycurtis": ("braycurtis", 10000, 50000, 0.01, None, "Bray-Curtis, sparse Poisson counts"),
    "jaccard": ("jaccard", 10000, 50000, 0.01, None, "Jaccard, sparse Poisson counts"),
}
# the other BASELINE configurations, run at full size inside the same invocation
CONFIGS = {
    "cfg3_braycurtis": dict(key="braycurtis", n=50000, tips=50000, lam=0.01, tree=None, min_world=1,
                            desc="Bray-Curtis, 50,000 samples x 50,000 features, sparse Poisson(0.01) "
                                 "counts (BASELINE configs[2]), strong scaling"),
    "cfg3_jaccard": dict(key="jaccard", n=50000, tips=50000, lam=0.01, tree=None, min_world=1,
                         desc="Jaccard, 50,000 samples x 50,000 features, sparse Poisson(0.01) counts "
                              "(BASELINE configs[2]), strong scaling"),
    "cfg4_unweighted": dict(key="unweighted_unifrac", n=100000, tips=500000, lam=0.004, tree="random",
                            min_world=8,
                            desc="unweighted UniFrac, 100,000 samples x 500,000 tips (BASELINE configs[3]), "
                                 "full matrix over 8 GPUs"),
    "cfg5_caterpillar": dict(key="weighted_unifrac", n=25000, tips=1000000, lam=0.002, tree="caterpillar",
                             min_world=8,
                             desc="weighted unnormalized UniFrac, 25,000 samples, 1,000,000-tip caterpillar "
                                  "tree (BASELINE configs[4]), over 8 GPUs"),
}


# ----------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------
def make_inputs(key, n, tips, lam, tree="random", need_values=True):
    from skbio_b200 import synthetic

    indptr, cols, vals = synthetic.poisson_csr(n, tips, lam, seed=0)
    if key in TREE_KEYS:
        gen = synthetic.caterpillar_tree if tree == "caterpillar" else synthetic.random_join_tree
        parent, length, names, tip_node = gen(tips, seed=1)
        if key == "unweighted_unifrac" and not need_values:
            vals = None
        return dict(kind="tree", key=key, n=n, parent=parent, length=length, indptr=indptr,
                    ids=tip_node[cols].astype(np.int32), cols=cols, vals=vals, tip_node=tip_node,
                    names=names, m=len(parent), tips=tips, lam=lam)
    return dict(kind="features", key=key, n=n, indptr=indptr, ids=cols.astype(np.int32),
                cols=cols, vals=vals.astype(np.float64), m=tips, tips=tips, lam=lam)


def dense_rows(inp, rows):
    """Dense int64 counts (len(rows) x tips/features) of the given samples."""
    out = np.zeros((len(rows), inp["tips"]), dtype=np.int64)
    ip, cols, vals = inp["indptr"], inp["cols"], inp["vals"]
    for k, s in enumerate(rows):
        a, b = ip[s], ip[s + 1]
        out[k, cols[a:b]] = 1 if vals is None else vals[a:b].astype(np.int64)
    return out


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


def sm_clock_hz(device):
    """Current SM clock (the pair kernels have just run: the boost clock under load)."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout
        return float(out.strip().splitlines()[0]) * 1e6
    except Exception:
        return 1.965e9


def ncu_traffic():
    """Per-launch DRAM bytes of the pair kernels from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, regenerated by tools/ncu_summary.py), keyed by workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------
# CPU side: the reference itself (oracle/_ref) or, without that build, the NumPy port
# ----------------------------------------------------------------------------------------
_REF_JOB = None   # inherited by forked workers, never pickled


def _ref_tree(inp):
    """skbio.TreeNode of the synthetic tree, through the reference's own Newick reader."""
    from oracle import ref_skbio
    from skbio_b200 import synthetic

    skbio = ref_skbio.load()
    text = synthetic.to_newick(inp["parent"], inp["length"], inp["names"])
    return skbio.TreeNode.read(io.StringIO(text))


def _ref_call(block):
    """One call of the reference's public API on one block of samples (run in a worker)."""
    from oracle import ref_skbio

    skbio = ref_skbio.load()
    from skbio.diversity import beta_diversity

    inp, tree, taxa, blocks = _REF_JOB
    rows = blocks[block]
    key = inp["key"]
    t0 = time.perf_counter()
    dense = dense_rows(inp, rows)
    ids = [str(r) for r in rows]
    t1 = time.perf_counter()
    with np.errstate(all="ignore"):
        if key == "unweighted_unifrac":
            dm = beta_diversity("unweighted_unifrac", dense, ids=ids, tree=tree, taxa=taxa)
        elif key in ("weighted_unifrac", "weighted_unifrac_normalized"):
            dm = beta_diversity("weighted_unifrac", dense, ids=ids, tree=tree, taxa=taxa,
                                normalized=key.endswith("normalized"))
        else:
            dm = beta_diversity(key, dense, ids=ids)
    t2 = time.perf_counter()
    n = len(rows)
    data = np.asarray(dm.data)
    vec = data[np.triu_indices(n, k=1)]
    return {"pairs": n * (n - 1) // 2, "call_s": t2 - t1, "densify_s": t1 - t0, "vec": vec}


def _port_call(block):
    """Same job with the NumPy port of the reference's pair functions (no oracle/_ref build)."""
    from functools import partial

    from oracle import oracle_fast, oracle_np

    inp, _, _, blocks = _REF_JOB
    rows = blocks[block]
    key = inp["key"]
    t1 = time.perf_counter()
    dense = dense_rows(inp, rows)
    n = len(rows)
    with np.errstate(all="ignore"):
        if inp["kind"] == "tree":
            bl = np.ascontiguousarray(inp["length"], dtype=np.float64)
            if key == "unweighted_unifrac":
                X = oracle_fast.counts_by_node((dense > 0).astype(np.int64), inp["tip_node"], inp["parent"])
                f = partial(oracle_np.unweighted_unifrac_pair, branch_lengths=bl)
            else:
                X = oracle_fast.counts_by_node(dense, inp["tip_node"], inp["parent"])
                tipidx = oracle_np.tip_indices_from_parent(inp["parent"])
                d = oracle_np.tip_distances(bl, inp["parent"], tipidx)
                if key.endswith("normalized"):
                    def f(u, v):
                        return oracle_np.weighted_unifrac_normalized_pair(
                            u, v, np.take(u, tipidx).sum(), np.take(v, tipidx).sum(), bl, d)
                else:
                    def f(u, v):
                        return oracle_np.weighted_unifrac_pair(
                            u, v, np.take(u, tipidx).sum(), np.take(v, tipidx).sum(), bl)[0]
            vec = np.array([f(X[i], X[j]) for i in range(n) for j in range(i + 1, n)])
        else:
            from scipy.spatial.distance import pdist

            vec = pdist(dense > 0, "jaccard") if key == "jaccard" else pdist(dense.astype(np.float64), "braycurtis")
    t2 = time.perf_counter()
    return {"pairs": n * (n - 1) // 2, "call_s": t2 - t1, "densify_s": 0.0, "vec": vec}


class CpuReference:
    """Runs the reference CPU path on disjoint sample blocks of a workload, one block per
    worker process (the reference's beta_diversity is single-threaded)."""

    def __init__(self, inp, blocks, procs):
        global _REF_JOB
        from oracle import ref_skbio

        self.real = ref_skbio.available()
        self.inp = inp
        self.blocks = blocks
        self.procs = max(1, min(procs, len(blocks)))
        tree = taxa = None
        self.tree_build_s = 0.0
        if self.real and inp["kind"] == "tree":
            t0 = time.perf_counter()
            tree = _ref_tree(inp)
            taxa = [f"T{i}" for i in range(inp["tips"])]
            self.tree_build_s = time.perf_counter() - t0
        _REF_JOB = (inp, tree, taxa, blocks)
        self.pool = None
        if self.procs > 1:
            import multiprocessing as mp

            self.pool = mp.get_context("fork").Pool(self.procs)

    def step(self):
        fn = _ref_call if self.real else _port_call
        t0 = time.perf_counter()
        if self.pool is not None:
            res = self.pool.map(fn, range(len(self.blocks)), chunksize=1)
        else:
            res = [fn(b) for b in range(len(self.blocks))]
        wall = time.perf_counter() - t0
        return wall, res

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

    def describe(self, wall, res):
        pairs = sum(r["pairs"] for r in res)
        per = len(self.blocks[0])
        call = float(np.mean([r["call_s"] for r in res]))
        what = ("skbio.diversity.beta_diversity of the unmodified reference (oracle/_ref, built by "
                "tools/build_ref.sh)") if self.real else \
               "NumPy port of the reference pair functions (oracle/_ref not built)"
        return (f"{len(self.blocks)} disjoint blocks of {per} samples of the workload "
                f"({pairs} pairs in all), all {self.inp['m']} nodes/features, dense counts as the "
                f"reference requires; {what}; one block per process, {self.procs} processes; "
                f"{wall:.1f} s wall per step, {call:.1f} s mean per call (tree validation and "
                "vectorisation included)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload
    key, n_def, tips_def, lam, tree_kind, desc = WORKLOADS[workload]
    lam = args.lam or lam
    n = args.samples or int(round(n_def * np.sqrt(args.gpus)))
    tips = args.tips or tips_def
    procs = min(os.cpu_count() or 1, 32)
    per = args.cpu_samples or (24 if key in TREE_KEYS else 256)
    n_sub = per * procs
    # only the first n_sub samples are ever touched: generate just those
    inp = make_inputs(key, min(n_sub, n), tips, lam, args.tree or tree_kind or "random")
    n_sub = inp["n"]
    blocks = [list(range(b, min(b + per, n_sub))) for b in range(0, n_sub, per)]
    blocks = [b for b in blocks if len(b) >= 2]
    ref = CpuReference(inp, blocks, procs)
    vals = []
    for step in range(args.warmup + args.steps):
        wall, res = ref.step()
        if step >= args.warmup:
            vals.append((sum(r["pairs"] for r in res) / wall, wall))
    ref.close()
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "metric_key": key, "samples": n, "tips_or_features": tips,
                   "nodes": inp["m"], "lambda": lam,
                   "samples_computed_per_step": n_sub,
                   "pairs_computed_per_step": int(sum(len(b) * (len(b) - 1) // 2 for b in blocks))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.procs,
                         "kind": "reference" if ref.real else "port",
                         "build": "oracle/_ref (tools/build_ref.sh)" if ref.real else None,
                         "host_cores_available": os.cpu_count(),
                         "sample": ref.describe(wall, res)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------
# parity sample: a rank's first owned row stripe against the CPU oracle
# ----------------------------------------------------------------------------------------
def oracle_pairs(inp, rows):
    """Condensed distances among `rows` from the C oracle (oracle/oracle_c.c; scipy for the
    feature metrics) - the checker, never the thing measured."""
    from oracle import oracle_fast

    dense = dense_rows(inp, rows)
    key = inp["key"]
    if inp["kind"] == "tree":
        bl = np.ascontiguousarray(inp["length"], dtype=np.float64)
        if key == "unweighted_unifrac":
            X = oracle_fast.counts_by_node((dense > 0).astype(np.int64), inp["tip_node"], inp["parent"])
            return oracle_fast.unweighted(X, bl)
        X = oracle_fast.counts_by_node(dense, inp["tip_node"], inp["parent"])
        return oracle_fast.weighted(X, bl, inp["parent"], key.endswith("normalized"))
    from scipy.spatial.distance import pdist

    with np.errstate(all="ignore"):
        return pdist(dense > 0, "jaccard") if key == "jaccard" else pdist(dense.astype(np.float64), "braycurtis")


def parity_sample(inp, compact, layout, n, k_rows):
    """Compare the pairs among the first k_rows samples of this rank's first owned stripe."""
    if not len(layout):
        return None
    first = int(layout[0][0])
    # stripe start row i0 from its first condensed index: cond_row_start(i0) == first
    i0 = 0
    while i0 * n - (i0 * (i0 + 1)) // 2 < first:
        i0 += 128
    rows = list(range(i0, min(i0 + k_rows, n)))
    if len(rows) < 2:
        return None
    exp = oracle_pairs(inp, rows)
    got = np.empty(len(exp))
    k = 0
    for a, i in enumerate(rows):
        row_start = i * n - (i * (i + 1)) // 2 - first
        cnt = len(rows) - 1 - a
        got[k:k + cnt] = compact[row_start:row_start + cnt]
        k += cnt
    exact = inp["kind"] == "features"
    if exact:
        ok = bool(np.array_equal(got, exp, equal_nan=True))
        return {"rows": [rows[0], rows[-1]], "pairs": len(exp), "bit_exact": ok,
                "max_rel_err": 0.0 if ok else float("inf")}
    with np.errstate(all="ignore"):
        rel = np.abs(got - exp) / np.maximum(np.abs(exp), 1e-300)
        rel = np.where(exp == 0.0, np.where(got == 0.0, 0.0, np.inf), rel)
    return {"rows": [rows[0], rows[-1]], "pairs": len(exp), "max_rel_err": float(np.max(rel))}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def _capi_mod():
    from skbio_b200 import _capi
    return _capi


class Ctx:
    def __init__(self, args):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.emulate = max(int(args.emulate_shards), 1)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.device = self.local_rank
        self.coll = None
        if self.dist is not None and self.emulate == 1 and not getattr(args, "no_exchange", False):
            from skbio_b200.distributed import TorchCollectives

            self.coll = TorchCollectives(self.local_rank)
            # every rank uploads only the CSR rows of its own build range; the others arrive over NVLink
            _capi_mod().set_option("partial_upload", 1)
        if self.emulate > 1:
            self.shard, self.n_shards = 0, self.emulate
        else:
            self.shard, self.n_shards = self.rank, self.world

    def barrier(self):
        self.torch.cuda.synchronize(self.device)
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=f"cuda:{self.device}")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def min_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=f"cuda:{self.device}")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return [float(x) for x in t]


def roofline_block(ctx, inp, t, pair_ms, prep_ms, P_shard, P_total, default_shape, traffic_key):
    """Roofline of the dominant pair kernel from the library's event timings of the last steps."""
    from skbio_b200 import _capi

    key, n, m = inp["key"], inp["n"], inp["m"]
    device = ctx.device
    peaks, peak_src = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    launches = max(int(t["pair_launches"]), 1)
    avg_launch_ms = max(pair_ms - prep_ms - (t["tips_ms"] if t["tips_ms"] > 0 else 0.0), 1e-6) / launches
    nodes_per_launch = m / launches
    traffic_tab = ncu_traffic().get(traffic_key, {}) if default_shape else {}
    if key in ("weighted_unifrac_normalized", "weighted_unifrac", "braycurtis"):
        kernel = "k_pair_weighted_mask" if t["visits_dense"] else "k_pair_weighted_v3"
        units_per_term = 2.0   # DADD (pu - pv) + DFMA (|.| * len + acc)
        bound, unit = "fp64", "Tinst/s (FP64 pipe instructions)"
        peak = _capi.microbench(0, device)
        peak_source = ("b2d_microbench(0): dependent DFMA chains on all SMs, measured in this run; "
                       "nominal 148 SM x 64 lanes x 1.965 GHz = 18.6")
        elem_bytes = 8.0
    elif key == "unweighted_unifrac":
        kernel = "k_pair_unweighted_lut"
        units_per_term = 1.0 / 8.0   # one 64-bit LUT gather (+ one DADD) per 8 node-pair terms
        bound, unit = "smem", "Tgather/s (64-bit shared-memory gathers)"
        peak = _capi.microbench(2, device)
        peak_source = ("b2d_microbench(2): bank-conflict-free 64-bit shared-memory gathers on all SMs, "
                       "measured in this run; nominal 148 SM x 16/clk x 1.965 GHz = 4.65")
        elem_bytes = 1.0 / 8.0
    else:
        kernel = "k_pair_popcount"
        units_per_term = 1.0 / 32.0  # one POPC per 32 feature-pair terms
        bound, unit = "alu", "Tinst/s (POPC)"
        peak = 148 * 16 * 1.965e9
        peak_source = "nominal 148 SM x 16 POPC/clk x 1.965 GHz (not measured)"
        elem_bytes = 1.0 / 8.0
    alg_units = P_shard * nodes_per_launch * units_per_term
    alg_bytes = n * nodes_per_launch * elem_bytes
    sparse_path = inp["kind"] == "features" and t["nodes_resident"] == 0
    atom_peak = atom_peak_cf = None
    if sparse_path or t["tips_ms"] > 0:
        atom_peak = _capi.microbench(6, device)      # random cells, 32 warps / SM
        atom_peak_cf = _capi.microbench(5, device)   # conflict-free, 64 warps / SM
    atom_src = ("b2d_microbench(6): fire-and-forget shared-memory atomicAdd on pseudo-random cells "
                "(bank conflicts as they fall), 32 warps/SM, all SMs, measured in this run")
    if sparse_path:
        # integer feature table below 2.5 % density: intersection kernel.  Algorithmic unit = one
        # accumulator update per (sample pair, shared feature) = sum_f C(n_f, 2).
        nf = np.bincount(inp["cols"], minlength=m).astype(np.float64)
        matches = float((nf * (nf - 1.0) / 2.0).sum()) * (P_shard / max(P_total, 1))
        kernel = "k_pair_isect" if t.get("isect_generation", 2) == 2 else "k_pair_sparse"
        bound, unit = "smem-atomic", "Tatom/s (shared-memory int32 atomic adds)"
        peak, peak_source = atom_peak, atom_src
        alg_units = matches
        units_per_term = matches / max(P_shard * nodes_per_launch, 1)
        alg_bytes = P_shard * 24.0 + 2.0 * len(inp["cols"]) * 8.0  # result write + finalize r/w, buckets
    # Zero-block skipping: the kernel executes only `visited` of the node-pair terms (the rest has a
    # closed form).  The roofline counts the instructions actually executed; the dense-equivalent
    # rate (all terms / time) is reported beside it.
    visited = (t["visits_executed"] / t["visits_dense"]) if t.get("visits_dense") else 1.0
    dense_equiv = alg_units / (avg_launch_ms * 1e-3) / 1e12
    if not sparse_path:
        alg_units *= visited
    achieved = alg_units / (avg_launch_ms * 1e-3) / 1e12
    roofline = {
        "kernel": kernel, "bound": bound, "achieved": achieved, "peak": peak / 1e12, "unit": unit,
        "frac": achieved / (peak / 1e12), "peak_source": peak_source,
        "traffic": traffic_tab.get(kernel),
        "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, "
                           "profiles/ncu_traffic.json") if traffic_tab.get(kernel) else None,
        "avg_launch_ms": avg_launch_ms, "launches_per_step": launches,
        "algorithmic_terms_per_launch": P_shard * nodes_per_launch,
        "units_per_term": units_per_term, "terms_visited_frac": visited,
        "dense_equivalent": dense_equiv,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                "achieved_gbs": alg_bytes / (avg_launch_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "peak_source": peak_src,
                "note": "not the binding resource: a 128x128 pair tile does 16384 terms per 256 "
                        "embedding elements read"},
    }
    if sparse_path:
        roofline["frac_vs_conflict_free_peak"] = achieved / (atom_peak_cf / 1e12)
    # Sparse rows (tips and nodes with few tips below) go through the intersection kernel; when it
    # takes longer than the dense launches it is the dominant kernel and gets the roofline, with the
    # dense kernel's numbers kept beside it.  Unit of work: one (sample pair, row) match = two
    # 32-bit shared-memory atomic adds (a 64-bit sum kept as two halves).
    if t["tips_ms"] > 0 and t["sparse_row_matches"] > 0:
        # which kernel the library chose: the pipelined one (b2d_isect.cuh) or the first generation
        kname = "k_pair_isect" if t.get("isect_generation", 2) == 2 else (
            "k_pair_utips" if key == "unweighted_unifrac" else "k_pair_tips")
        entry_bytes = 4.0 if key == "unweighted_unifrac" else 16.0
        atoms = 2.0 * t["sparse_row_matches"]
        a_rate = atoms / (t["tips_ms"] * 1e-3) / 1e12
        nbk = (n + 127) // 128 + 1
        bucket_bytes = entry_bytes * t["sparse_row_entries"] * nbk / max(ctx.n_shards, 1)
        tips_roof = {
            "kernel": kname, "mode": "unweighted (keys + length table)" if key == "unweighted_unifrac"
            else "weighted (fixed-point entries)",
            "bound": "smem-atomic", "achieved": a_rate, "peak": atom_peak / 1e12,
            "unit": "Tatom/s (shared-memory int32 atomic adds)", "frac": a_rate / (atom_peak / 1e12),
            "peak_source": atom_src,
            "frac_vs_conflict_free_peak": a_rate / (atom_peak_cf / 1e12),
            "conflict_free_peak": atom_peak_cf / 1e12,
            "conflict_free_peak_source": "b2d_microbench(5): one bank per lane by construction, 64 warps/SM",
            "traffic": traffic_tab.get(kname),
            "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum of the launch, ncu --set full, "
                               "profiles/ncu_traffic.json") if traffic_tab.get(kname) else None,
            "avg_launch_ms": float(t["tips_ms"]), "launches_per_step": 1,
            "matches_per_launch": t["sparse_row_matches"], "units_per_match": 2.0,
            # what ncu shows to be the binding resource: every shared-memory load and atomic is a
            # wavefront of the SM's data pipe (one per clock).  Algorithmic wavefronts per match = the
            # column entry read by a full warp (entry bytes / 128) + two atomic adds of a full warp
            # without bank conflicts (2 / 32); the executed count is higher by the bank conflicts of the
            # scattered atomics and partly filled warps (profiles/: measured pipe utilisation)
            "shared_memory_pipe": (lambda wpm, pk: {
                "algorithmic_wavefronts_per_match": wpm,
                "achieved": wpm * t["sparse_row_matches"] / (t["tips_ms"] * 1e-3) / 1e9,
                "peak": pk / 1e9, "unit": "Gwavefront/s (l1tex data pipe, 1 per clock per SM)",
                "frac": wpm * t["sparse_row_matches"] / (t["tips_ms"] * 1e-3) / pk,
                "peak_source": "148 SMs x SM clock sampled in this run (nvidia-smi)",
                "measured_pipe_utilisation": traffic_tab.get(kname + ":lsu_pipe_pct"),
                "measured_source": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed of the "
                                    "launch, ncu --set full, profiles/ncu_traffic.json")
                if traffic_tab.get(kname + ":lsu_pipe_pct") else None,
            })(entry_bytes / 128.0 + 2.0 / 32.0, 148 * sm_clock_hz(device)),
            "sparse_rows_max_tips": t["sparse_rows_max_tips"],
            "sparse_row_entries": t["sparse_row_entries"],
            # every bucket is read by about n/128 tiles (as row or as column block)
            "algorithmic_bytes_per_launch": bucket_bytes,
            "hbm": {"algorithmic_bytes_per_launch": bucket_bytes,
                    "achieved_gbs": bucket_bytes / (t["tips_ms"] * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": peak_src,
                    "note": "entries are re-read by every tile of their block row / column, mostly "
                            "from L2; not the binding resource"},
        }
        # the dense kernel's time is a difference of event times and overlaps with the tail of the
        # intersection kernel on the other stream: a fraction above 1 only says the split is unreliable
        if roofline["frac"] > 1.0:
            roofline["frac"] = None
            roofline["note"] = ("time by difference (pair phase - preparation - intersection kernel), "
                                "the two pair kernels overlap on two streams")
        if t["tips_ms"] >= avg_launch_ms * launches:
            tips_roof["dense_kernel"] = roofline
            roofline = tips_roof
        else:
            roofline["sparse_row_kernel"] = tips_roof
    return roofline


def bench_problem(ctx, inp, steps, warmup, e2e_reps, parity_rows, default_shape=False, traffic_key=None,
                  sample_clocks=False):
    """Device-resident steps, the one-shot C-ABI path with host buffers, parity and roofline of
    one workload on this rank's shard.  Returns a dict (complete on rank 0)."""
    from skbio_b200 import _capi

    key, n, m = inp["key"], inp["n"], inp["m"]
    device = ctx.device
    P_total = n * (n - 1) // 2

    def make_problem(src):
        if src["kind"] == "tree":
            prob = _capi.Problem.tree(src["parent"], src["length"], src["indptr"], src["ids"],
                                      src["vals"], n, device=device, shard=ctx.shard,
                                      n_shards=ctx.n_shards)
            if ctx.coll is not None:   # distributed build: the ranks exchange the per-block products (NCCL)
                ctx.coll.install(prob)
            return prob
        return _capi.Problem.features(src["indptr"], src["ids"], src["vals"], n, m, device=device,
                                      shard=ctx.shard, n_shards=ctx.n_shards)

    # ---------------- device-resident steps ----------------
    prob = make_problem(inp)
    P_shard = prob.result_len()
    for _ in range(warmup):
        prob.compute(key)
    sampler = ClockSampler(device) if sample_clocks else None
    ctx.barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, pair_ms, embed_ms, prep_ms, tips_ms, launches = [], [], [], [], [], 0
    xchg_ms, xchg_bytes = [], []
    for _ in range(steps):
        prob.compute(key)  # blocks until the shard's distances are complete in HBM
        t = prob.timings()
        xchg_ms.append(t.get("exchange_ms", 0.0))
        xchg_bytes.append(t.get("exchange_bytes", 0.0))
        dev_ms.append(t["compute_ms"])
        pair_ms.append(t["pair_ms"])
        embed_ms.append(t["embed_ms"])
        prep_ms.append(t["skip_prep_ms"])
        tips_ms.append(t["tips_ms"])
        launches += t["kernels"]   # every kernel the library launched for this step
    ctx.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    step_ms = float(np.mean(dev_ms))
    step_ms_max, wall_step_ms_max, embed_max, prep_max, xchg_max, xchg_bytes_max = ctx.max_over_ranks(
        [step_ms, wall_ms / steps, float(np.mean(embed_ms)), float(np.mean(prep_ms)), float(np.mean(xchg_ms)),
         float(np.mean(xchg_bytes))])
    P_done = P_total if ctx.emulate == 1 else P_shard   # pairs actually computed by this job
    value = P_done / (step_ms_max * 1e-3)

    # ---------------- parity: this rank's first owned stripe against the oracle ----------------
    compact, layout = prob.fetch_compact()
    par = parity_sample(inp, compact, layout, n, parity_rows) if parity_rows else None
    worst = par["max_rel_err"] if par else 0.0
    worst_all, = ctx.max_over_ranks([worst if np.isfinite(worst) else 1e300])
    checksum = float(np.nansum(compact[: min(len(compact), 1 << 20)]))
    keep_rows = min(n, 512)   # rows 0..511 of stripe 0 (rank 0): for the CPU-baseline comparison
    stripe0_prefix = np.array(compact[: keep_rows * n - (keep_rows * (keep_rows + 1)) // 2]) \
        if ctx.rank == 0 and ctx.shard == 0 else None
    del compact
    prob.destroy()

    # ---------------- end to end through the C ABI, host buffers ----------------
    h2d = (inp["indptr"].nbytes + inp["ids"].nbytes + (inp["vals"].nbytes if inp["vals"] is not None else 0)
           + (inp["parent"].nbytes + inp["length"].nbytes if inp["kind"] == "tree" else 0))
    d2h = P_shard * 8
    e2e = None
    if e2e_reps:
        # rank r's page-locked buffer holds what rank r computed (compact: owned stripes back to
        # back); inputs are page-locked as well so that the uploads are plain DMA
        out = _capi.PinnedBuffer(P_shard)
        pinned = {k: _capi.PinnedBuffer.like(inp[k]) for k in ("parent", "length", "indptr", "ids", "vals")
                  if inp.get(k) is not None}
        src = dict(inp)
        for k, b in pinned.items():
            src[k] = b.array
        times = []
        csr_up = None
        for _ in range(e2e_reps):
            ctx.barrier()
            t0 = time.perf_counter()
            with make_problem(src) as p2:
                p2.compute(key)
                p2.fetch_compact(out.array)
                csr_up = p2.timings().get("csr_upload_bytes")
            ctx.torch.cuda.synchronize(device)
            times.append(time.perf_counter() - t0)
        if csr_up:   # what this rank really uploaded (partial upload: its own rows of the CSR)
            h2d = csr_up + (inp["parent"].nbytes + inp["length"].nbytes if inp["kind"] == "tree" else 0)
        h2d, = ctx.max_over_ranks([float(h2d)])
        med, best = float(np.median(times)), float(min(times))
        med_max, best_max = ctx.max_over_ranks([med, best])
        e2e = {"value": P_done / med_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "seconds": med_max, "seconds_best": best_max,
               "host_buffers": "page-locked (b2d_host_alloc) for inputs and result",
               "h2d_note": ("max over ranks; with several ranks a rank uploads the CSR rows of its own block range "
                            "only, the other rows arrive from the peers over NVLink inside the timed region"
                            if ctx.coll is not None and inp["kind"] == "tree" else "the whole CSR and the tree"),
               "timed": f"create (H2D) + compute + fetch (D2H) + destroy, median of {e2e_reps} "
                        "(max over ranks)"}
        if ctx.rank == 0:
            print(f"[{key} n={n}] e2e repetitions (s):", [round(x, 4) for x in times], file=sys.stderr)
        out.free()
        for b in pinned.values():
            b.free()

    res = None
    if ctx.rank == 0:
        roofline = roofline_block(ctx, inp, t, float(np.mean(pair_ms)), float(np.mean(prep_ms)), P_shard,
                                  P_total, default_shape, traffic_key)
        res = {
            "value": value, "unit": UNIT, "ms_per_step": step_ms_max, "steps": steps, "warmup": warmup,
            "wall_ms_per_step": wall_step_ms_max,
            "samples": n, "tips_or_features": inp["tips"], "nodes": m, "lambda": inp["lam"],
            "pairs": P_total, "pairs_computed": P_done,
            "sharding": (f"row-stripe snake x{ctx.world}" if ctx.emulate == 1 else
                         f"shard 0 of {ctx.emulate} emulated on one GPU"),
            "embed_ms": float(np.mean(embed_ms)), "pair_ms": float(np.mean(pair_ms)),
            "prologue_ms": {"embed": embed_max, "sparse_rows_and_skip_prep": prep_max,
                            "exchange": xchg_max, "exchange_bytes_received_per_rank": xchg_bytes_max,
                            "note": ("distributed build: every rank builds the per-block products of its own "
                                     "block range, the ranks all-gather them over NCCL (exchange is part of "
                                     "sparse_rows_and_skip_prep); max over ranks") if xchg_bytes_max > 0 else
                                    "replicated on every rank (all samples); max over ranks"},
            "passes": int(t["passes"]),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "parity": {"checker": "oracle/oracle_c.c (C restatement of the reference, pinned to the "
                                  "reference's fixtures) / scipy pdist for the feature metrics",
                       "rank0": par, "max_rel_err_all_ranks": worst_all if worst_all < 1e299 else float("inf"),
                       "tolerance": 0.0 if inp["kind"] == "features" else 1e-12,
                       "what": f"pairs among the first {parity_rows} samples of every rank's first owned "
                               "row stripe"},
            "tips_flagged": int(t["tips_flagged"]),
            "checksum_first_1M": checksum,
            "clocks": clocks,
            "_stripe0_prefix": stripe0_prefix,
        }
    return res


def cpu_baseline_block(inp, n_sub, gpu_vec_fn):
    """scikit-bio's own beta_diversity (oracle/_ref) on the first n_sub samples, 1 process,
    plus the parity of the GPU result on those samples against it."""
    rows = list(range(min(n_sub, inp["n"])))
    ref = CpuReference(inp, [rows], 1)
    wall, res = ref.step()
    ref.close()
    vec = res[0]["vec"]
    got = gpu_vec_fn(rows)
    with np.errstate(all="ignore"):
        rel = np.abs(got - vec) / np.maximum(np.abs(vec), 1e-300)
        rel = np.where(vec == 0.0, np.where(got == 0.0, 0.0, np.inf), rel)
    out = {"value": res[0]["pairs"] / res[0]["call_s"], "unit": UNIT, "cores": 1,
           "kind": "reference" if ref.real else "port",
           "build": "oracle/_ref (tools/build_ref.sh)" if ref.real else None,
           "host_cores_available": os.cpu_count(),
           "sample": ref.describe(wall, res),
           "max_rel_err_vs_gpu": float(np.nanmax(rel)) if len(rel) else None}
    return out


def api_e2e_block(ctx, inp, reps=5):
    """The public call a user makes: beta_diversity(metric, sparse table, ids, tree=<Newick text
    read into flat arrays>, taxa=<names>) -> DistanceMatrix, host objects in, host object out."""
    import scipy.sparse as sp

    import skbio_b200
    from skbio_b200 import synthetic

    key, n, tips = inp["key"], inp["n"], inp["tips"]
    text = synthetic.to_newick(inp["parent"], inp["length"], inp["names"])
    table = sp.csr_matrix((inp["vals"] if inp["vals"] is not None else np.ones(len(inp["cols"]), dtype=np.int64),
                           inp["cols"], inp["indptr"]), shape=(n, tips))
    taxa = [f"T{i}" for i in range(tips)]
    ids = [f"S{i}" for i in range(n)]
    metric = "weighted_unifrac" if key.startswith("weighted") else key
    kw = {"normalized": True} if key.endswith("normalized") else {}
    times, splits = [], []
    for _ in range(reps):
        ctx.torch.cuda.synchronize(ctx.device)
        t0 = time.perf_counter()
        tree = skbio_b200.read_newick_flat(text)
        t1 = time.perf_counter()
        dm = skbio_b200.beta_diversity(metric, table, ids=ids, tree=tree, taxa=taxa, condensed=True,
                                       devices=[ctx.device], **kw)
        t2 = time.perf_counter()
        times.append(t2 - t0)
        splits.append({"newick_to_flat_s": t1 - t0, "beta_diversity_s": t2 - t1,
                       **getattr(skbio_b200, "last_call_profile", lambda: {})()})
        del dm
    k = int(np.argsort(times)[len(times) // 2])
    P = n * (n - 1) // 2
    return {"value": P / times[k], "unit": UNIT, "seconds": times[k], "split": splits[k],
            "input": "Newick text (read_newick_flat) + scipy.sparse CSR table + taxa names + ids",
            "output": "DistanceMatrix(condensed=True)", "reps": reps}


def run_ours(args):
    from skbio_b200 import _capi

    ctx = Ctx(args)
    if not _capi.available():
        raise RuntimeError("libskbb200.so missing or no CUDA device: nothing to benchmark "
                           "(the engine has no CPU path)")
    # this process creates and destroys the same large problem many times: let the library keep
    # its freed device buffers (embedding included) between problems - opt-in, 4 GiB by default
    _capi.set_cache_limit(120 << 30)
    workload = args.workload
    key, n_def, tips_def, lam, tree_kind, desc = WORKLOADS[workload]
    lam = args.lam or lam
    n = args.samples or int(round(n_def * np.sqrt(ctx.world)))
    tips = args.tips or tips_def
    inp = make_inputs(key, n, tips, lam, args.tree or tree_kind or "random")
    default_shape = not (args.samples or args.tips or args.lam or args.tree) and ctx.world == 1 and ctx.emulate == 1
    head = bench_problem(ctx, inp, args.steps, args.warmup, 5, 48, default_shape=default_shape,
                         traffic_key=workload, sample_clocks=True)

    line = None
    if ctx.rank == 0:
        cpu = None
        if not args.no_cpu and ctx.world == 1 and ctx.emulate == 1:
            n_sub = args.cpu_samples or (32 if inp["kind"] == "tree" else 512)

            def gpu_vec(rows):
                # the benchmark's own output: rows 0..k-1 are a prefix of stripe 0
                compact = head["_stripe0_prefix"]
                k = len(rows)
                got = np.empty(k * (k - 1) // 2)
                pos = 0
                for i in range(k - 1):
                    a = i * n - (i * (i + 1)) // 2
                    got[pos:pos + k - 1 - i] = compact[a:a + k - 1 - i]
                    pos += k - 1 - i
                return got

            cpu = cpu_baseline_block(inp, n_sub, gpu_vec)
        api = None
        if not args.no_api and ctx.emulate == 1 and inp["kind"] == "tree":
            api = api_e2e_block(ctx, inp)
        elem_bytes = 8.0 if key != "unweighted_unifrac" and key != "jaccard" else 1.0 / 8.0
        line = {
            "metric": METRIC_NAME, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": (desc if not (args.samples or args.tips) else f"{key} custom size")
                                   + (f"; weak scaling: {n} samples at {ctx.world} GPUs (pairs per GPU fixed)"
                                      if ctx.world > 1 and not args.samples else ""),
                       "metric_key": key, "samples": n, "tips_or_features": tips, "nodes": inp["m"],
                       "lambda": lam, "pairs": head["pairs"], "pairs_computed": head["pairs_computed"],
                       "sharding": head["sharding"],
                       "l2_policy": "inputs larger than L2 (embedding %.1f GB per pass)"
                                    % (n * inp["m"] * elem_bytes / 1e9),
                       "device_buffer_cache": "opt-in: skbio_b200.set_cache_limit(120 GiB) - freed device "
                                              "buffers are reused by the next problem of this process "
                                              "(library default: 4 GiB parked at most)"},
            "wall_ms_per_step": head["wall_ms_per_step"],
            "embed_ms": head["embed_ms"], "pair_ms": head["pair_ms"], "prologue_ms": head["prologue_ms"],
            "e2e": head["e2e"], "api_e2e": api,
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
            "roofline": head["roofline"], "cpu_baseline": cpu, "parity": head["parity"],
            "tips_flagged": head["tips_flagged"], "checksum_first_1M": head["checksum_first_1M"],
        }
    del inp

    # ---------------- the other BASELINE configurations, full size, same invocation ----------------
    want = args.configs
    names = []
    if want == "auto":
        names = [k for k, c in CONFIGS.items() if ctx.world >= c["min_world"]] if ctx.emulate == 1 and \
            not (args.samples or args.tips or args.lam or args.tree or args.workload != "weighted") else []
    elif want != "none":
        names = [k for k in CONFIGS if any(k == w or k.startswith(w + "_") for w in want.split(","))]
    cfg_out = {}
    for name in names:
        c = CONFIGS[name]
        t0 = time.perf_counter()
        cin = make_inputs(c["key"], c["n"], c["tips"], c["lam"], c["tree"] or "random", need_values=False)
        gen_s = time.perf_counter() - t0
        csteps = max(1, min(args.steps, 3 if c["n"] >= 25000 and c["tree"] else 5))
        r = bench_problem(ctx, cin, csteps, 3, 3, 32)
        if ctx.rank == 0:
            r["workload"] = c["desc"]
            r["metric_key"] = c["key"]
            r["scaling"] = "strong" if c["min_world"] == 1 else "single configuration (8 GPUs)"
            r["input_generation_s"] = gen_s
            r.pop("clocks", None)
            r.pop("_stripe0_prefix", None)
            cfg_out[name] = r
        del cin
    if line is not None:
        line["configs"] = cfg_out
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
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
    ap.add_argument("--no-api", action="store_true")
    ap.add_argument("--lam", type=float, default=0.0)
    ap.add_argument("--tree", default="", choices=["", "random", "caterpillar"])
    ap.add_argument("--configs", default="auto",
                    help="auto: cfg3 at every N, cfg4 and cfg5 at N = 8; none; or a comma list")
    ap.add_argument("--no-exchange", action="store_true",
                    help="N > 1: every rank builds the per-block products of all samples itself (no collective)")
    ap.add_argument("--emulate-shards", type=int, default=1,
                    help="compute only shard 0 of N on this GPU (sizing runs for multi-GPU configs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
