"""Host-side batching threads with the reference's queue contracts.

`ThreadPredictor(server, id, state_dim, prediction_q)` -- ThreadPredictor.py:34-66
    drains `prediction_q` items `(agent_id, state)` into a batch of at most
    Config.PREDICTION_BATCH_SIZE rows (block for the first item, then take whatever is already
    queued, no timeout), calls `server.model.predict_p_and_v`, and answers every agent on its own
    `wait_q` with `(p_row, v_scalar)`.
`ThreadTrainer(server, id)` -- ThreadTrainer.py:33-62
    concatenates `(x, r, a, x2, done)` items from `server.training_q` until the row count exceeds
    Config.TRAINING_MIN_BATCH_SIZE, then calls `server.train_model(x, r, a, x2, done, id)`.

They only need a `server` exposing `.model`, `.agents`, `.training_q`, `.train_model` (and
`.network_tester_process` / `.replay_q` when those Config switches are on), as in Server.py:70-150.
"""
from __future__ import annotations

from threading import Thread

import numpy as np

from .config import Config as _DefaultConfig

NETWORK_TESTER_ID = 100      # ThreadPredictor.py:64-66


class ThreadPredictor(Thread):
    def __init__(self, server, id, state_dim, prediction_q, config=None):
        super().__init__(daemon=True)
        self.server = server
        self.id = id
        self.state_dim = state_dim
        self.prediction_q = prediction_q
        self.exit_flag = False
        self.config = config or getattr(server, "config", None) or _DefaultConfig
        self.batches = 0
        self.rows = 0

    def _run_slab(self, q, cap):
        """Batch calls of ga3c_b200.transport.SlabPredictionQueue: one gather into a pinned buffer (which Network copies
        to the device without another host pass) and one reply scatter per batch."""
        dtype = q._states.array.dtype
        try:
            import torch
            tdt = torch.uint8 if dtype == np.uint8 else torch.float32
            states = torch.empty((cap, self.state_dim), dtype=tdt, pin_memory=torch.cuda.is_available()).numpy()
        except Exception:
            states = np.zeros((cap, self.state_dim), dtype=dtype)
        while not self.exit_flag:
            ids = q.get_batch(cap, states, timeout=0.05)
            if ids is None:
                continue
            n = ids.size
            p, v = self.server.model.predict_p_and_v(states[:n])
            self.batches += 1
            self.rows += n
            q.reply_batch(ids, p, v)

    def run(self):
        cap = self.config.PREDICTION_BATCH_SIZE
        if hasattr(self.prediction_q, "get_batch"):
            return self._run_slab(self.prediction_q, cap)
        ids = np.zeros(cap, dtype=np.uint16)                      # uint16 as in ThreadPredictor.py:46
        try:        # pinned, so that Network copies the batch to the device without another host pass
            import torch
            states = torch.zeros((cap, self.state_dim), dtype=torch.float32, pin_memory=torch.cuda.is_available()).numpy()
        except Exception:
            states = np.zeros((cap, self.state_dim), dtype=np.float32)
        q = self.prediction_q
        while not self.exit_flag:
            ids[0], states[0] = q.get()
            n = 1
            while n < cap and not q.empty():
                ids[n], states[n] = q.get()
                n += 1
            p, v = self.server.model.predict_p_and_v(states[:n])
            self.batches += 1
            self.rows += n
            agents = self.server.agents
            for i in range(n):
                aid = int(ids[i])
                if aid < len(agents):
                    agents[aid].wait_q.put((p[i], v[i]))
                if aid == NETWORK_TESTER_ID and getattr(self.config, "USE_NETWORK_TESTER", False):
                    self.server.network_tester_process.wait_q.put((p[i], v[i]))


class ThreadTrainer(Thread):
    def __init__(self, server, id, config=None):
        super().__init__(daemon=True)
        self.server = server
        self.id = id
        self.exit_flag = False
        self.config = config or getattr(server, "config", None) or _DefaultConfig

    def _next_batch(self):
        """Accumulate agent batches until the row count EXCEEDS TRAINING_MIN_BATCH_SIZE (the
        reference loops `while batch_size <= MIN`, ThreadTrainer.py:49)."""
        parts, rows = [], 0
        while rows <= self.config.TRAINING_MIN_BATCH_SIZE:
            item = self.server.training_q.get()
            parts.append(item)
            rows += item[0].shape[0]
        if len(parts) == 1:
            return parts[0]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))

    def _run_slab(self, q):
        """Batch call of ga3c_b200.transport.SlabTrainingQueue: the agents' blocks are gathered straight into pinned arrays
        (which Network.train copies to the device without another host pass) until the row count exceeds
        TRAINING_MIN_BATCH_SIZE, exactly where the reference stops concatenating (ThreadTrainer.py:49)."""
        min_rows = self.config.TRAINING_MIN_BATCH_SIZE
        cap = min_rows + q.max_rows
        dtype = q._x.array.dtype
        try:
            import torch
            pin = torch.cuda.is_available()
            x = torch.empty((cap, q.state_dim), dtype=torch.uint8 if dtype == np.uint8 else torch.float32, pin_memory=pin).numpy()
        except Exception:
            x = np.zeros((cap, q.state_dim), dtype=dtype)
        r = np.zeros(cap, dtype=np.float64)
        a = np.zeros((cap, q.num_actions), dtype=np.float32)
        done = np.zeros(cap, dtype=np.bool_)
        x2 = np.zeros((cap, 0), dtype=dtype)          # the A3C networks ignore next_state (NetworkVP.py:254-257)
        while not self.exit_flag:
            n = q.get_batch(min_rows, x, r, a, done, timeout=0.05, stop=lambda: self.exit_flag)
            if n is None or self.exit_flag:
                continue
            if self.config.TRAIN_MODELS:
                self.server.train_model(x[:n], r[:n], a[:n], x2[:n], done[:n], self.id)

    def run(self):
        q = self.server.training_q
        if hasattr(q, "get_batch") and not getattr(self.config, "USE_REPLAY_MEMORY", False) and not q.ship_next_state:
            return self._run_slab(q)
        while not self.exit_flag:
            if getattr(self.config, "USE_REPLAY_MEMORY", False):
                x, a, r, done, x2 = self.server.replay_q.get()     # ThreadTrainer.py:45-46 (DDPG only)
            else:
                x, r, a, x2, done = self._next_batch()
            if self.config.TRAIN_MODELS:
                self.server.train_model(x, r, a, x2, done, self.id)


class LockstepTrainer(ThreadTrainer):
    """The trainer thread of ONE rank of a data-parallel job (every GPU owns a slice of the agents and its own
    queues; gradients are summed across the ranks inside `model.train`).

    The reference's trainer is fed asynchronously (ThreadTrainer.py:42-62): it trains whenever its own queue yields a batch.
    The gradient exchange needs every rank to enter every step, so this thread ticks: each round it gathers what its
    `training_q` yields within `tick` seconds under the reference's stop rule (rows > TRAINING_MIN_BATCH_SIZE ends the
    round early), the ranks agree with one tiny host allreduce on (rows in total, anybody leaving?), and then ALL of them
    call `server.train_model` -- a rank whose round came up empty with a 0-row batch, which contributes a zero gradient and
    applies the same update as everybody else (`Network.train` -> ga3c_train_step with batch = 0).  Rounds in which no rank
    has rows are skipped by all; when any rank raises its exit flag all ranks leave after the same number of steps.
    One LockstepTrainer per rank (Config.TRAINERS = 1 per GPU): two would interleave their steps differently on different
    ranks.  `group`: a torch.distributed group whose backend takes CPU tensors (gloo); default: a new gloo group over all
    ranks (a collective call -- construct the trainer on every rank)."""

    def __init__(self, server, id, config=None, group=None, tick=0.002):
        super().__init__(server, id, config)
        import torch.distributed as dist
        self.tick = float(tick)
        self.group = group if group is not None else (dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else None)
        self.steps = 0              # exchange steps taken (identical on every rank)
        self.empty_steps = 0        # ... of which this rank contributed no rows

    def _collect(self, q):
        import queue as _queue
        import time
        parts, rows = [], 0
        deadline = time.monotonic() + self.tick
        while rows <= self.config.TRAINING_MIN_BATCH_SIZE:
            left = deadline - time.monotonic()
            try:
                item = q.get(True, left) if left > 0 else q.get(False)
            except _queue.Empty:
                break
            parts.append(item)
            rows += item[0].shape[0]
        return parts, rows

    def run(self):
        import torch
        import torch.distributed as dist
        q = self.server.training_q
        model = self.server.model
        sdim, na = int(model.state_dim), int(model.num_actions)
        while True:
            parts, rows = self._collect(q)
            flags = torch.tensor([rows, 1 if self.exit_flag else 0], dtype=torch.int64)
            dist.all_reduce(flags, op=dist.ReduceOp.SUM, group=self.group)
            if int(flags[1]) > 0:
                break
            if int(flags[0]) == 0:
                continue
            if not parts:
                item = (np.zeros((0, sdim), np.float32), np.zeros(0, np.float64), np.zeros((0, na), np.float32),
                        np.zeros((0, 0), np.float32), np.zeros(0, np.bool_))
                self.empty_steps += 1
            elif len(parts) == 1:
                item = parts[0]
            else:
                item = tuple(np.concatenate([p[k] for p in parts]) for k in range(5))
            self.steps += 1
            if self.config.TRAIN_MODELS:
                self.server.train_model(*item, self.id)
