#!/usr/bin/env python3
"""Steady-state device time of permute_expression_pair and of the graph evaluator (several calls each)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectre_b200 import halo2
be = halo2.Backend([0])
dev = torch.device("cuda", 0)
for k in (16, 20):
    n = 1 << k
    usable = n - 6
    table = torch.zeros((n, 4), dtype=torch.int64, device=dev); table[:, 0] = torch.arange(n, device=dev) % (1 << min(k - 1, 19))
    lk_in = table[torch.randint(0, usable, (n,), device=dev)].contiguous()
    a = torch.empty_like(table); b = torch.empty_like(table)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        be.permute_expression_pair_dev(lk_in.data_ptr(), table.data_ptr(), usable, a.data_ptr(), b.data_ptr())
        torch.cuda.synchronize()
        print("permute k=%d iter %d: %.3f ms" % (k, it, (time.perf_counter() - t0) * 1e3), flush=True)
# graph evaluator: 95 calculations over 2^22 rows
E = 1 << 22
cols = [torch.randint(0, 2**60, (E, 4), dtype=torch.int64, device=dev) for _ in range(20)]
vals = torch.zeros((E, 4), dtype=torch.int64, device=dev)
prog = []
for a_i in range(19):
    t = 5 * a_i
    prog += [2, t, 3, a_i | (1 << 16), 3, a_i | (2 << 16), 0, t + 1, 3, a_i, 1, t, 1, t + 2, 1, t + 1, 3, a_i | (3 << 16), 2, t + 3, 2, 0, 1, t + 2,
             6 | (1 << 8), t + 4, (10 if a_i == 0 else 1), (0 if a_i == 0 else t - 1), 9, 0, 1, t + 3]
prog = np.array(prog, dtype=np.uint32)
zero = np.zeros((1, 4), dtype=np.uint64); y = np.array([[3, 5, 7, 11]], dtype=np.uint64)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    be.graph_evaluate_dev(prog, 95, 95, zero, np.array([0, 1, 2, 3], dtype=np.int32), [cols[19].data_ptr()], [c.data_ptr() for c in cols[:19]], [], zero, y, y, y, y, vals.data_ptr(), E, 4)
    torch.cuda.synchronize()
    print("graph 95 calcs x 2^22 rows iter %d: %.3f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
be.close()
