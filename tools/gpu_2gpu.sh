mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --samples 6000 --tips 30000 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 2500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
python bench.py --steps 2 --warmup 3 --samples 4243 --tips 30000 --no-cpu > gpurun_out/bench_1gpu_small.json 2> gpurun_out/bench_1gpu_small.err
tail -c 1200 gpurun_out/bench_1gpu_small.json
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import skbio_b200
from skbio_b200 import synthetic
n, tips = 700, 900
parent, length, names, tip_node = synthetic.random_join_tree(tips, seed=2)
indptr, cols, vals = synthetic.poisson_csr(n, tips, 0.05, seed=1)
dense = synthetic.csr_to_dense(indptr, cols, vals, tips)
tree = synthetic.to_nodes(parent, length, names)
taxa = [f"T{i}" for i in range(tips)]
a = skbio_b200.beta_diversity("unweighted_unifrac", dense, tree=tree, taxa=taxa).condensed_form()
b = skbio_b200.beta_diversity("unweighted_unifrac", dense, tree=tree, taxa=taxa, devices=[0, 1]).condensed_form()
c = skbio_b200.beta_diversity("braycurtis", dense, devices=[1, 0]).condensed_form()
d = skbio_b200.beta_diversity("braycurtis", dense).condensed_form()
print("two-device thread mode equal:", np.array_equal(a, b), np.array_equal(c, d, equal_nan=True))
PY
