import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
os.environ["GA3C_MLP_TC"] = sys.argv[1] if len(sys.argv) > 1 else "1"
from ga3c_b200 import mlp_network as mlp
from oracle import oracle_mlp as om
kind, s, a, b = "fork_vp", 3, 1, int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(7)
params = om.init_params(rng, kind, s, a)
x = rng.uniform(-1, 1, size=(b, s)).astype(np.float32); y_r = rng.uniform(-1, 1, size=b).astype(np.float32)
act = rng.uniform(-1, 1, size=(b, a)).astype(np.float32)
net = mlp.NetworkVP("gpu:0", "dbg", a, s, max_batch=b)
net.set_variables(params); net.beta = 0.01
print("losses", net.losses(x, y_r, act))
ref_l, ref_g = om.loss_and_grads(params, x, y_r, act, kind, beta=0.01)
print("ref   ", ref_l)
# forward activations
h = x.astype(np.float64); acts = []
sig = lambda z: 1 / (1 + np.exp(-z))
for name, _, actn in om.FORK_VP_LAYERS:
    z = h @ params[name + "/w:0"].astype(np.float64) + params[name + "/b:0"]
    h = sig(z) if actn == "sigmoid" else z
    acts.append(h)
for l in range(5):
    got = net.workspace(0, l, b)
    d = np.abs(got - acts[l])
    print(f"act[{l}] max err {d.max():.3e} (max |ref| {np.abs(acts[l]).max():.3f}) worst idx {np.unravel_index(d.argmax(), d.shape)}")
    if d.max() > 1e-3:
        bad = (d > 1e-3)
        print("   bad rows:", np.unique(np.nonzero(bad)[0])[:20], " bad cols:", np.unique(np.nonzero(bad)[1])[:40], "count", bad.sum(), "of", bad.size)
        r, c = np.unravel_index(d.argmax(), d.shape)
        print("   got", got[r, c:c+4], "ref", acts[l][r, c:c+4])
g = net.get_gradients()
for k in ref_g:
    d = np.abs(g[k] - ref_g[k]).max(); print(f"grad {k}: max abs err {d:.3e} (max |ref| {np.abs(ref_g[k]).max():.3e})")
