"""Prints the per-layer parity report (CUDA path vs oracle) for a few batch sizes.  Debug aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ga3c_b200
from _parity import make_case, layer_report

batches = [int(b) for b in sys.argv[1:]] or [1, 3, 8, 33, 128]
for b in batches:
    params, x, y_r, a = make_case(b)
    net = ga3c_b200.Network("gpu:0", "dbg", 6, max_batch=max(b, 32))
    t = time.time()
    rep = layer_report(net, params, x, y_r, a)
    print(f"--- B={b} ({time.time()-t:.1f}s)")
    for k, (d, r) in rep.items():
        print(f"  {k:24s} abs {d:.3e} rel {r:.3e}")
    del net
