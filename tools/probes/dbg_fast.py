import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from opticalimageprocessor_b200 import ops
oracle.build()
ctx = ops.Context(0)
rng = np.random.default_rng(7)
src = rng.integers(0, 65536, (300, 1024), dtype=np.uint16)
want = oracle.prestitch_shift(src, 1.37, -2.61)
got = ops.prestitch_shift(ctx, torch.from_numpy(src).cuda(), 1.37, -2.61).cpu().numpy()
bad = np.argwhere(got != want)
print("mismatches", len(bad), bad[:10].tolist())
