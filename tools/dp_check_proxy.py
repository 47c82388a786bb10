"""What bench.py's dp_check oracle leg will see at world = 2, 4, 8 -- estimated on ONE GPU.

The data-parallel step equals a single trainer fed the concatenated batch, and its distance from the oracle is
set by bf16 activations that round the other way than in the fp64-accumulating oracle -- row by row, no matter which rank
holds the row.  So one Network trained on the concatenated `64 * world` rows of dp_check's own seeded data (seed 4242, the
same draws in the same order) lands within a few ulp of what the N-rank run reports as `max_abs_vs_oracle`; the script prints
that distance next to dp_check's tolerance for every world size, so a world size that was never run on real GPUs (N = 4) is
known to clear the check before the driver's scaling run meets it.  (It is how the first, rows-proportional tolerance of
dp_check was found to fail at N = 4 -- one ReLU-gate flip, 2.8e-5 against 2e-5 -- and replaced by bench.oracle_distance.)

    python tools/dp_check_proxy.py            # on a GPU box
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ga3c_b200                               # noqa: E402
from oracle import oracle_np as onp            # noqa: E402  (checker only)
from bench import oracle_distance              # noqa: E402  (the criterion dp_check applies)

NUM_ACTIONS = 6


def leg(world, rows=64):
    g = np.random.default_rng(4242)
    params = onp.init_params(g, NUM_ACTIONS)
    net = ga3c_b200.Network("gpu:0", f"proxy{world}", NUM_ACTIONS, max_batch=rows * world)
    net.set_variables(params)
    net.set_slots({k: np.ones_like(v) for k, v in params.items()}, {k: np.zeros_like(v) for k, v in params.items()})
    ms, mom = onp.rmsprop_init(params)
    ref = params
    for _ in range(2):
        x = onp.synth_frames(g, rows * world)
        y_r, a = onp.synth_targets(g, rows * world, NUM_ACTIONS)
        net.train(x, y_r, a, None, None, 0)
        _, _, ref, ms, mom = onp.train_step(ref, ms, mom, x, y_r, a, lr=net.learning_rate, beta=net.beta, quant="bf16")
        ref = {k: v.astype(np.float32) for k, v in ref.items()}
    got = net.get_variables()
    return dict(world=world, rows=rows * world, **oracle_distance(got, ref, params))


if __name__ == "__main__":
    worlds = [int(w) for w in sys.argv[1:]] or [2, 4, 8]
    for w in worlds:
        print(json.dumps(leg(w)), flush=True)
