"""`NativePredictor` -- ThreadPredictor (ThreadPredictor.py:34-66) with its loop inside the C library.

Same place in the system, same contracts: one object per predictor, `start()`, `exit_flag`, `batches` / `rows` counters; it
serves a `ga3c_b200.transport.SlabPredictionQueue` (agents `put((id, state))` / `post`, answers arrive on the agents' own
`wait_q`).  What changes is where the loop runs: `ga3c_batcher_*` (include/ga3c_b200.h) scans the pending bytes, copies the
pending rows from the page-locked state slab into the device batch with one kernel, runs `ga3c_predict`, writes the replies
and posts the agents' semaphores -- without entering the interpreter, so agents are answered while the trainer thread holds
the GIL, and the batch is never gathered on the host.  Model calls are serialised with `Network.train` by the handle's mutex.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


def _sem_handle(sem) -> int:
    """sem_t* of a multiprocessing.Semaphore (CPython's SemLock keeps it as an integer handle)."""
    return int(sem._semlock.handle)


class NativePredictor:
    def __init__(self, server, id, prediction_q, config=None):
        model = server.model
        q = prediction_q
        if not hasattr(q, "_states"):
            raise TypeError("NativePredictor serves a ga3c_b200.transport.SlabPredictionQueue")
        self.server, self.id, self.prediction_q = server, id, q
        cfg = config or getattr(server, "config", None) or model.config
        states = q._states.array
        if states.dtype not in (np.uint8, np.float32) or states.shape[1] != model.state_dim:
            raise ValueError(f"state slab must be uint8 or float32 [agents, {model.state_dim}], got {states.dtype} {states.shape}")
        self._lib = _capi.load()
        self._wake = (C.c_void_p * q.num_agents)(*[_sem_handle(s) for s in q._wake])
        c = _capi.ga3c_batcher_config(
            device=model._ordinal, num_agents=q.num_agents, state_bytes=int(states.shape[1] * states.itemsize),
            x_u8=int(states.dtype == np.uint8), max_batch=int(cfg.PREDICTION_BATCH_SIZE), num_actions=model.num_actions,
            states=states.ctypes.data, pending=q._pending.array.ctypes.data, reply_p=q._reply_p.array.ctypes.data,
            reply_v=q._reply_v.array.ctypes.data, work_sem=_sem_handle(q._work), wake_sems=self._wake)
        h = C.c_void_p()
        _capi.check(self._lib.ga3c_batcher_create(model._h, C.byref(c), C.byref(h)), "ga3c_batcher_create")
        self._h = h
        self._model = model          # keeps the network (and its handle) alive for the thread
        self._running = False

    def start(self):
        _capi.check(self._lib.ga3c_batcher_start(self._h), "ga3c_batcher_start")
        self._running = True

    def _stats(self):
        b, r, e = C.c_int64(), C.c_int64(), C.c_int32()
        rc = self._lib.ga3c_batcher_stats(self._h, C.byref(b), C.byref(r), C.byref(e))
        if rc != 0:
            _capi.check(rc, "native predictor batcher")
        return b.value, r.value

    @property
    def batches(self):
        return self._stats()[0]

    @property
    def rows(self):
        return self._stats()[1]

    @property
    def exit_flag(self):
        return not self._running

    @exit_flag.setter
    def exit_flag(self, value):
        if value and self._running:
            self._lib.ga3c_batcher_stop(self._h)
            self._running = False

    def join(self, timeout=None):
        self.exit_flag = True

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ga3c_batcher_destroy(self._h)
            self._h = None
            self._running = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
