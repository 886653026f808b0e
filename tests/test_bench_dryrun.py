"""bench.py's control flow and JSON line on CPU: the native engine is replaced by a stand-in that
answers with the oracle on tiny shapes (test infrastructure only - nothing here measures anything).
Catches breakage of the benchmark script itself, which otherwise only runs on the GPU box."""
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _FakeProblem:
    def __init__(self, kind, args, n, m, shard, n_shards):
        self.kind, self.args, self.n, self.m = kind, args, n, m
        self.shard, self.n_shards = shard, n_shards
        self.key = None

    @classmethod
    def tree(cls, parent, length, indptr, node_ids, values, n, device=0, shard=0, n_shards=1, node_of_col=None):
        return cls("tree", (parent, length, indptr, node_ids, values), n, len(parent), shard, n_shards)

    @classmethod
    def features(cls, indptr, ids, values, n, m, device=0, shard=0, n_shards=1):
        return cls("features", (indptr, ids, values), n, m, shard, n_shards)

    def _layout(self):
        from skbio_b200 import _capi

        return _capi.shard_layout(self.n, self.shard, self.n_shards)

    def result_len(self):
        return self._layout()[1]

    def compute(self, key):
        self.key = key

    def _full(self):
        from oracle import oracle_fast
        from scipy.spatial.distance import pdist

        n = self.n
        if self.kind == "tree":
            parent, length, indptr, ids, vals = self.args
            m = len(parent)
            X = np.zeros((n, m), dtype=np.int64)
            rows = np.repeat(np.arange(n), np.diff(indptr))
            X[rows, ids] = 1 if vals is None or self.key == "unweighted_unifrac" else vals
            for k in range(m - 1):
                X[:, parent[k]] += X[:, k]
            X = np.ascontiguousarray(X)
            bl = np.ascontiguousarray(length, dtype=np.float64)
            if self.key == "unweighted_unifrac":
                return oracle_fast.unweighted(X, bl)
            return oracle_fast.weighted(X, bl, parent, self.key.endswith("normalized"))
        indptr, ids, vals = self.args
        D = np.zeros((n, self.m))
        D[np.repeat(np.arange(n), np.diff(indptr)), ids] = vals
        with np.errstate(all="ignore"):
            return pdist(D > 0, "jaccard") if self.key == "jaccard" else pdist(D, "braycurtis")

    def fetch_compact(self, out=None):
        full = self._full()
        offs, total = self._layout()
        parts = [full[a:a + c] for a, c in offs]
        vec = np.concatenate(parts) if parts else np.zeros(0)
        if out is not None:
            out[:len(vec)] = vec
            vec = out
        return vec, offs

    def timings(self):
        return {"embed_ms": 1.0, "pair_ms": 3.0, "compute_ms": 4.0, "fetch_ms": 0.1, "upload_ms": 0.1,
                "pair_launches": 1, "nodes_resident": 0 if self.kind == "features" else self.m, "passes": 1,
                "visits_executed": 10.0, "visits_dense": 40.0, "skip_prep_ms": 0.5, "kernels": 12,
                "tips_ms": 1.5, "tips_flagged": 0, "sparse_rows_max_tips": 4, "sparse_row_entries": 1000,
                "sparse_row_matches": 5000.0, "isect_generation": 2}

    def destroy(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass


class _FakePinned:
    def __init__(self, n, dtype=np.float64):
        self.array = np.zeros(int(n), dtype=dtype)

    @classmethod
    def like(cls, a):
        b = cls(a.size, a.dtype)
        b.array[:] = np.ascontiguousarray(a).ravel()
        return b

    def free(self):
        pass


@pytest.fixture
def fake_engine(monkeypatch):
    import torch

    from skbio_b200 import _capi

    monkeypatch.setattr(_capi, "Problem", _FakeProblem)
    monkeypatch.setattr(_capi, "PinnedBuffer", _FakePinned)
    monkeypatch.setattr(_capi, "available", lambda: True)
    monkeypatch.setattr(_capi, "microbench", lambda kind, device=0: 1e12)
    monkeypatch.setattr(_capi, "set_cache_limit", lambda nbytes: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, dtype=None, device=None: real_tensor(data, dtype=dtype))
    import bench

    monkeypatch.setattr(bench, "api_e2e_block", lambda ctx, inp, reps=3: {"value": 1.0, "unit": bench.UNIT})
    return bench


@pytest.mark.parametrize("workload", ["weighted", "unweighted", "braycurtis"])
def test_bench_line_has_the_contract_keys(fake_engine, monkeypatch, capsys, workload):
    bench = fake_engine
    tiny = {"cfg3_braycurtis": dict(bench.CONFIGS["cfg3_braycurtis"], n=150, tips=400, lam=0.05),
            "cfg4_unweighted": dict(bench.CONFIGS["cfg4_unweighted"], n=140, tips=300, lam=0.05, min_world=1)}
    monkeypatch.setattr(bench, "CONFIGS", tiny)
    argv = ["bench.py", "--steps", "2", "--warmup", "1", "--workload", workload, "--samples", "200", "--tips", "300",
            "--cpu-samples", "6", "--configs", "cfg3,cfg4"]
    monkeypatch.setattr(sys, "argv", argv)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert bench.main() == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline",
                "cpu_baseline", "parity", "configs"):
        assert key in line, key
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["max_rel_err_vs_gpu"] <= 1e-12
    assert line["parity"]["max_rel_err_all_ranks"] <= 1e-12
    assert set(line["configs"]) == {"cfg3_braycurtis", "cfg4_unweighted"}
    for c in line["configs"].values():
        assert c["parity"]["max_rel_err_all_ranks"] <= 1e-12 and c["e2e"]["value"] > 0 and "roofline" in c


def test_reference_arm_prints_its_line(monkeypatch, capsys):
    import bench

    argv = ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--samples", "40", "--tips", "200",
            "--cpu-samples", "5"]
    monkeypatch.setattr(sys, "argv", argv)
    monkeypatch.setattr(os, "cpu_count", lambda: 2)
    assert bench.main() == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
