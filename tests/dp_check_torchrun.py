"""Test infrastructure (imports the oracle as the checker; launched by tests/test_gpu_parity.py on boxes with >= 2 GPUs, or by hand).
Data-parallel parity check, one rank per GPU (launch with torchrun): every rank trains on its row
shard of one global batch; afterwards (1) all replicas hold bit-identical weights and (2) they match
the oracle's single-process step on the concatenated batch.  Prints 'DP_CHECK OK' on rank 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import ga3c_b200
from ga3c_b200.dataparallel import shard_rows
from oracle import oracle_np as onp        # checker only

B = 16 * world
rng = np.random.default_rng(12345)
params = onp.init_params(rng, 6)
net = ga3c_b200.Network(f"gpu:{local}", "dp", 6, max_batch=64)        # no seed: every rank draws its own initial weights ...


def replicas_identical():
    got = net.get_variables()
    ms_, _ = net.get_slots()
    flat = torch.from_numpy(np.concatenate([got[k].ravel() for k in sorted(got)] + [ms_[k].ravel() for k in sorted(ms_)] +
                                           [net.workspace(6)])).cuda()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    return all(torch.equal(gathered[0], g) for g in gathered)


assert replicas_identical(), "replicas differ after construction without a seed"      # ... and rank 0's are broadcast
net.set_variables(params)
assert net.dp_mode in ("fused", "nccl")
print(f"rank {rank}: dp_mode={net.dp_mode}", flush=True)
ms, mom = onp.rmsprop_init(params)
ref = params
for step in range(3):
    x = onp.synth_frames(rng, B)
    y_r, a = onp.synth_targets(rng, B)
    lo, hi = shard_rows(B, rank, world)
    if step == 1:                    # a lock-step round in which the last rank has no rows: the one before it takes both shards
        if rank == world - 1:
            lo = hi
        elif rank == world - 2:
            hi = shard_rows(B, world - 1, world)[1]
    net.train(x[lo:hi], y_r[lo:hi], a[lo:hi], None, None, 0)
    _, _, ref, ms, mom = onp.train_step(ref, ms, mom, x, y_r, a, lr=net.learning_rate, beta=net.beta, quant="bf16")
    ref = {k: v.astype(np.float32) for k, v in ref.items()}
got = net.get_variables()
identical = replicas_identical()
worst = max(float(np.abs(got[k] - ref[k]).max()) for k in got)
if rank == 0:
    print(f"replicas identical: {identical}; max |w - oracle| after 3 steps: {worst:.3e}")
    assert identical and worst <= 1e-5, (identical, worst)
    print("DP_CHECK OK")
dist.barrier()
dist.destroy_process_group()
