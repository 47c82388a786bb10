"""BASELINE configs[4] (SURVEY 8d run 5): the GA3C loop end to end with synthetic agents -- agent processes -> prediction
queue -> ThreadPredictor -> Network.predict_p_and_v -> replies; experiences -> training queue -> ThreadTrainer ->
Network.train -- reporting PPS (predictions/s) and TPS (training frames/s), per GPU and whole job.

    python tools/loop_bench.py --agents 256 --procs 16 --seconds 10 [--transport slab|queue] [--uint8]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/loop_bench.py ...      (one loop per GPU, data parallel train)

The environment is a stub (a fresh frame from a pool of random frames every step, `done` every 1000 steps); everything
else is the real path: ga3c_b200.ThreadPredictor / ThreadTrainer, the reference's experience handling (returns
R_t = gamma^(n-1-t) r_last as ProcessAgent._accumulate_rewards computes them, one-hot actions, np.random.choice sampling,
TIME_MAX = 5 as upstream) and ga3c_b200.Network.  `--transport queue` runs the same loop over multiprocessing.Queue, the
reference's transport, for comparison on the same box.  Each agent process hosts agents/procs agents (the box has far
fewer cores than 256) and keeps one request per agent in flight, which is the reference's per-agent contract.
"""
import argparse
import json
import multiprocessing as mp
import os
import queue
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
S, A, T_MAX, GAMMA = 84 * 84 * 4, 6, 5, 0.99


def agent_worker(wid, aids, pq, tq, wait_qs, stop, counters, uint8, seed):
    """Hosts len(aids) agents.  Per step and agent: env stub -> predict request -> (p, v) -> sample action -> experience;
    every T_MAX steps the experiences go to the trainer (ProcessAgent.py:117-176)."""
    rng = np.random.default_rng(seed)
    np.random.seed(seed % (2 ** 31))
    pool = rng.integers(0, 256, size=(32, S), dtype=np.uint8)
    if not uint8:
        pool = pool.astype(np.float32) / np.float32(128.0) - np.float32(1.0)      # Environment.py:60
    k = len(aids)
    xs = np.zeros((k, T_MAX, S), dtype=pool.dtype)
    acts = np.zeros((k, T_MAX), dtype=np.int64)
    actions = np.arange(A)
    eye = np.eye(A, dtype=np.float32)
    disc = GAMMA ** np.arange(T_MAX - 1, -1, -1, dtype=np.float64)
    slab = hasattr(pq, "post")
    t = 0
    n_pred = n_train = 0
    parent = os.getppid()
    while not stop.value and os.getppid() == parent:      # never outlive the server process
        frames = rng.integers(0, 32, size=k)
        for j, aid in enumerate(aids):
            st = pool[frames[j]]
            xs[j, t] = st
            if slab:
                pq.state_row(aid)[...] = st
                pq.post(aid)
            else:
                pq.put((aid, st))
        for j, aid in enumerate(aids):
            while True:
                try:
                    p, v = wait_qs[j].get(timeout=1.0)
                    break
                except queue.Empty:
                    if stop.value or os.getppid() != parent:
                        return
            acts[j, t] = np.random.choice(actions, p=p)                            # ProcessAgent.py:110-115
        n_pred += k
        t += 1
        if t == T_MAX:
            for j, aid in enumerate(aids):
                r_last = float(rng.uniform(-1, 1))
                item = (xs[j].copy(), disc * r_last, eye[acts[j]], xs[j], np.zeros(T_MAX, dtype=bool))
                (tq[j] if slab else tq).put(item)
            n_train += k * T_MAX
            t = 0
        counters[2 * wid] = n_pred
        counters[2 * wid + 1] = n_train


def agent_worker_vec(wid, aids, pq, tq, stop, counters, uint8, seed):
    """The same agents, stepped as a vector by one process (slab transport only): one gather of k frames, one post, one wait,
    inverse-CDF sampling for all k agents at once (what np.random.choice does per agent: searchsorted(cumsum(p), u))."""
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 256, size=(32, S), dtype=np.uint8)
    if not uint8:
        pool = pool.astype(np.float32) / np.float32(128.0) - np.float32(1.0)
    aids = np.asarray(aids)
    k = len(aids)
    xs = np.zeros((T_MAX, k, S), dtype=pool.dtype)
    acts = np.zeros((T_MAX, k), dtype=np.int64)
    p = np.zeros((k, A), np.float32)
    v = np.zeros(k, np.float32)
    eye = np.eye(A, dtype=np.float32)
    disc = GAMMA ** np.arange(T_MAX - 1, -1, -1, dtype=np.float64)
    done = np.zeros(T_MAX, dtype=bool)
    tqs = [tq.for_agent(int(a)) for a in aids]
    parent = os.getppid()
    t = 0
    n_pred = n_train = 0
    while not stop.value and os.getppid() == parent:
        np.take(pool, rng.integers(0, 32, size=k), axis=0, out=xs[t])
        pq.post_many(aids, xs[t])
        got = 0
        while got < k:
            got = pq.wait_many(aids, p, v, timeout=1.0, start=got)
            if got < k and (stop.value or os.getppid() != parent):
                return
        cdf = np.cumsum(p.astype(np.float64), axis=1)
        cdf /= cdf[:, -1:]
        acts[t] = np.minimum((cdf <= rng.random(k)[:, None]).sum(axis=1), A - 1)          # searchsorted(cdf, u, side='right')
        n_pred += k
        t += 1
        if t == T_MAX:
            r_last = rng.uniform(-1, 1, size=k)
            for j in range(k):
                tqs[j].put((xs[:, j], disc * r_last[j], eye[acts[:, j]], xs[:, j], done))
            n_train += k * T_MAX
            t = 0
        counters[2 * wid] = n_pred
        counters[2 * wid + 1] = n_train


class MiniServer:
    """The three attributes the thread classes use (Server.py:70-150)."""

    def __init__(self, model, training_q, agents):
        self.model, self.training_q, self.agents = model, training_q, agents
        self.frames = 0
        self.batches = 0

    def train_model(self, x, r, a, x2, done, tid):
        self.model.train(x, r, a, x2, done, tid)
        self.frames += x.shape[0]
        self.batches += 1


class _QueueAgent:
    def __init__(self, wait_q):
        self.wait_q = wait_q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=256)
    ap.add_argument("--procs", type=int, default=min(16, os.cpu_count() or 1))
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--transport", default="slab", choices=["slab", "queue"])
    ap.add_argument("--uint8", action="store_true", help="agents ship raw uint8 frames (SURVEY 8f F2)")
    ap.add_argument("--vector-agents", action="store_true",
                    help="each agent process steps its agents as one numpy vector (slab transport only)")
    ap.add_argument("--predictors", type=int, default=2)
    ap.add_argument("--native-predictor", action="store_true",
                    help="the predictor loop inside the C library (ga3c_b200.NativePredictor, slab transport only) instead of "
                         "--predictors Python threads")
    ap.add_argument("--trainers", type=int, default=2)
    ap.add_argument("--min-train-batch", type=int, default=512, help="Config.TRAINING_MIN_BATCH_SIZE")
    ap.add_argument("--independent-replicas", action="store_true",
                    help="N > 1: no gradient exchange, every GPU trains its own replica (round 1's mode); the default is the "
                         "data-parallel train step driven by ga3c_b200.LockstepTrainer")
    ap.add_argument("--tick", type=float, default=0.004, help="LockstepTrainer: seconds a rank waits for rows before it ticks")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    ctx = mp.get_context("fork")
    procs_n = max(1, min(args.procs // max(world, 1) if world > 1 else args.procs, args.agents))
    dtype = np.uint8 if args.uint8 else np.float32

    # transport objects and agent processes FIRST (fork before CUDA is initialised)
    from ga3c_b200.transport import SlabPredictionQueue, SlabTrainingQueue
    stop = ctx.Value("i", 0)
    counters = ctx.Array("q", 2 * procs_n, lock=False)
    groups = [list(range(w, args.agents, procs_n)) for w in range(procs_n)]
    if args.transport == "slab":
        pq = SlabPredictionQueue(args.agents, S, A, dtype=dtype, ctx=ctx)
        tq = SlabTrainingQueue(args.agents, T_MAX, S, A, blocks_per_agent=2, dtype=dtype, ctx=ctx)
        agents = []
        if args.vector_agents:
            workers = [ctx.Process(target=agent_worker_vec, daemon=True,
                                   args=(w, g, pq, tq, stop, counters, args.uint8, 1000 * rank + w)) for w, g in enumerate(groups)]
        else:
            workers = [ctx.Process(target=agent_worker, daemon=True,
                                   args=(w, g, pq, [tq.for_agent(a) for a in g], [pq.wait_q(a) for a in g], stop, counters,
                                         args.uint8, 1000 * rank + w)) for w, g in enumerate(groups)]
    else:
        pq, tq = ctx.Queue(maxsize=100), ctx.Queue(maxsize=100)                  # Config.MAX_QUEUE_SIZE
        wait = [ctx.Queue(maxsize=1) for _ in range(args.agents)]
        agents = [_QueueAgent(q) for q in wait]
        workers = [ctx.Process(target=agent_worker, daemon=True,
                               args=(w, g, pq, tq, [wait[a] for a in g], stop, counters, args.uint8, 1000 * rank + w))
                   for w, g in enumerate(groups)]
    for w in workers:
        w.start()

    import torch
    import ga3c_b200
    from ga3c_b200 import LockstepTrainer, NativePredictor, ThreadPredictor, ThreadTrainer
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    class Cfg(ga3c_b200.Config):
        TRAINING_MIN_BATCH_SIZE = args.min_train_batch
        PREDICTION_BATCH_SIZE = 128
    # N > 1: the data-parallel train step needs every rank to enter every step, the asynchronous loop trains whenever its own
    # queue yields a batch -- LockstepTrainer bridges the two (one tiny host allreduce per round; a rank whose round came up
    # empty enters the step with 0 rows).  Replicas start identical (no seed: rank 0's weights are broadcast) and must end so.
    dp = world > 1 and not args.independent_replicas
    model = ga3c_b200.Network(f"gpu:{local}", "loop", A, max_batch=4096, seed=None if dp else 12345, config=Cfg, data_parallel=dp)
    server = MiniServer(model, tq, agents)
    if args.native_predictor:
        preds = [NativePredictor(server, 0, pq, config=Cfg)]
    else:
        preds = [ThreadPredictor(server, i, S, pq, config=Cfg) for i in range(args.predictors)]
    trains = [LockstepTrainer(server, 0, config=Cfg, tick=args.tick)] if dp else [ThreadTrainer(server, i, config=Cfg) for i in range(args.trainers)]
    for th in preds + trains:
        th.start()
    time.sleep(2.0)                                                               # warm-up
    c0 = np.array(counters[:]).reshape(-1, 2).sum(axis=0)
    f0, b0, r0, pb0 = server.frames, server.batches, sum(p.rows for p in preds), sum(p.batches for p in preds)
    t0 = time.perf_counter()
    time.sleep(args.seconds)
    dt = time.perf_counter() - t0
    c1 = np.array(counters[:]).reshape(-1, 2).sum(axis=0)
    f1, b1, r1, pb1 = server.frames, server.batches, sum(p.rows for p in preds), sum(p.batches for p in preds)
    stop.value = 1
    dead = [w.exitcode for w in workers if not w.is_alive()]
    if dead:
        print(f"[loop_bench] rank {rank}: {len(dead)} agent processes died (exit codes {sorted(set(dead))})", file=sys.stderr, flush=True)
    res = {"pps": (r1 - r0) / dt, "tps_frames": (f1 - f0) / dt, "tps_batches": (b1 - b0) / dt,
           "mean_predict_batch": (r1 - r0) / max(pb1 - pb0, 1), "mean_train_batch": (f1 - f0) / max(b1 - b0, 1),
           "agent_side_pps": float(c1[0] - c0[0]) / dt}
    extra = {}
    if dp:
        # stop training on every rank after the same number of steps, then compare the replicas
        for th in trains:
            th.exit_flag = True
        for th in trains:
            th.join(timeout=60)
        w = model.get_variables()
        ms, _ = model.get_slots()
        flat = torch.from_numpy(np.concatenate([w[k].ravel() for k in sorted(w)] + [ms[k].ravel() for k in sorted(ms)] +
                                               [model.workspace(6)])).to(f"cuda:{local}")
        got = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(got, flat)
        steps = torch.tensor([trains[0].steps, trains[0].empty_steps], dtype=torch.int64, device=f"cuda:{local}")
        all_steps = [torch.empty_like(steps) for _ in range(world)]
        dist.all_gather(all_steps, steps)
        extra = {"data_parallel": "fused peer-memory exchange, LockstepTrainer", "replicas_identical_after_run": bool(all(torch.equal(got[0], g) for g in got)),
                 "exchange_steps_per_rank": [int(t[0]) for t in all_steps], "empty_steps_per_rank": [int(t[1]) for t in all_steps],
                 "global_step": model.get_global_step()}
        model.dp_check()
    if world > 1:
        t = torch.tensor([res["pps"], res["tps_frames"]], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t)
        res["job_pps"], res["job_tps_frames"] = float(t[0]), float(t[1])
    if rank == 0:
        out = {"config": f"{args.agents} synthetic agents per GPU in {procs_n} processes, transport={args.transport}, "
                         f"frames={'uint8' if args.uint8 else 'fp32'}, {'vectorised agent processes, ' if args.vector_agents else ''}T_MAX={T_MAX}, predictors={'native (C library)' if args.native_predictor else args.predictors}, "
                         f"trainers={args.trainers}, TRAINING_MIN_BATCH_SIZE={args.min_train_batch}, {os.cpu_count()} host cores",
               "n_gpus": world, "seconds": round(dt, 2), **{k: round(v, 1) for k, v in res.items()}, **extra}
        print(json.dumps(out), flush=True)
    for th in preds + trains:
        th.exit_flag = True
    time.sleep(0.2)
    for w in workers:
        w.terminate()
    if world > 1:
        dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
