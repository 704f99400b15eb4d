"""Times the pieces of one end-to-end fit step (reset, accumulate from host buffers, finalize)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from pypolymlp_b200.libmlpcpp import PotentialXtX
from pypolymlp_b200.params import make_params_dict
import cases  # bench puts tests/ on sys.path

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pd = make_params_dict(**cases.cfg2_model_kwargs(4))
acc = PotentialXtX(pd, device=0)
axis, pcs, tys, w, y = bench.make_batch(S, 0)
batch = acc.stage(axis, pcs, tys, [True] * S, w, y)
w_h, y_h = np.ascontiguousarray(w), np.ascontiguousarray(y)
for it in range(4):
    t0 = time.perf_counter(); acc.reset(); acc.context.synchronize()
    t1 = time.perf_counter(); acc.add_batch(batch, w_h, y_h); t2 = time.perf_counter(); acc.context.synchronize()
    t3 = time.perf_counter(); r = acc.finalize(); t4 = time.perf_counter()
    print(f"reset {1e3*(t1-t0):.2f} ms  add_batch call {1e3*(t2-t1):.2f} ms (+sync {1e3*(t3-t2):.2f})  finalize {1e3*(t4-t3):.2f} ms  total {1e3*(t4-t0):.2f}")
