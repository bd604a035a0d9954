import json, sys, itertools
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import theoremsearch_b200 as ts
from theoremsearch_b200 import synthetic
from bench_extra import timed
dev = torch.device("cuda", 0)
rows, nlist = 40_000_000, 16384
index = ts.TheoremIndex(1024, rows, dtype="bf16", device=dev)
centers = synthetic.fill_index_clustered(index, rows, nlist, 1.0, seed=0)
q_all = synthetic.make_clustered_queries(4096, centers, 1.0)
index.ivf_train(nlist, n_sample=1_000_000, iters=5, seed=0)
index.ivf_build("fp8")
for nq in (16, 64, 128, 256, 512, 1024, 4096):
    res = {"nq": nq}
    for name, mn in (("k4b", 0), ("k4d", 1)):
        ts.set_tunable("ivf.group_min_nq", mn)
        index._ws = {}
        q = q_all[:nq].contiguous()
        ms = timed(lambda: index.ivf_search(q, 10, nprobe=32, rescore_k=100), 2, 5)
        res[name + "_ms"] = ms
    print(json.dumps(res))
