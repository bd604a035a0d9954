import json, sys, torch
sys.path.insert(0, '.')
import theoremsearch_b200 as ts
from theoremsearch_b200 import synthetic
dev = torch.device('cuda', 0)
q = synthetic.make_queries(64, 1024, dev)
out = {}
for rnd in range(2):
    for no_vmm in (0, 1):
        ts.set_tunable('store.no_vmm', no_vmm)
        ix = ts.TheoremIndex(1024, 10_000_000, device=dev)
        ts.set_tunable('store.no_vmm', 0)
        synthetic.fill_index(ix, 0, 10_000_000, seed=0)
        for i in range(5): ix.search(q[i:i+1], 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100): ix.search(q[i % 64:i % 64 + 1], 10)
        e1.record(); torch.cuda.synchronize()
        out[f'round{rnd}_no_vmm{no_vmm}_ms'] = e0.elapsed_time(e1) / 100
        ix.close(); del ix; torch.cuda.empty_cache()
print(json.dumps(out))
