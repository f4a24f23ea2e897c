"""Batching front end: many single-threaded sessions, one GPU pass.

The reference serves every user from its own single-threaded ``WebSession`` Ray actor
(seesaw/web/web_session_actor.py:13-16, 70-82) and each of them scans the read-only index on its own.
With the database in the HBM of one GPU, concurrent ``query`` calls are better served together: the
batched kernel reads the database ONCE for up to 64 queries (SSW_MAX_BATCH).  ``ScanBatcher`` collects
the stage-1 requests of concurrent callers for at most ``max_wait_s`` (or until ``max_batch`` are
waiting), issues one ``scan_topk`` and hands each caller its own rows of the result.

``scanner`` is anything with ``scan_topk(queries[nq, d], k, exclude=[ids_0, ...]) -> dict(dbidx, score,
row, count)`` — a :class:`seesaw_b200.engine.PatchDatabase` in production, a stub in the CPU tests."""
from __future__ import annotations

import threading
import time
from concurrent.futures import Future

import numpy as np


class ScanBatcher:
    def __init__(self, scanner, *, max_batch=64, max_wait_s=0.0005):
        self.scanner, self.max_batch, self.max_wait_s = scanner, int(max_batch), float(max_wait_s)
        self._cv = threading.Condition()
        self._pending = []            # (vector, k, exclude, future)
        self._closed = False
        self.batches_issued = 0
        self.queries_served = 0
        self._thread = threading.Thread(target=self._run, name="ssw-scan-batcher", daemon=True)
        self._thread.start()

    # ---- caller side --------------------------------------------------------------------
    def submit(self, vector, k, exclude=None) -> Future:
        """Enqueue one query; the Future resolves to dict(dbidx, score, row) trimmed to the count."""
        fut = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("ScanBatcher is closed")
            self._pending.append((np.asarray(vector, dtype=np.float32).reshape(-1), int(k), exclude, fut))
            self._cv.notify_all()
        return fut

    def scan_topk_one(self, vector, k, exclude=None, timeout=None):
        """Blocking form, what an index's ``_query_prelim`` calls."""
        return self.submit(vector, k, exclude).result(timeout)

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join()

    # ---- worker ---------------------------------------------------------------------------
    def _take_batch(self):
        with self._cv:
            while not self._pending and not self._closed:
                self._cv.wait()
            if not self._pending:
                return None
            deadline = time.monotonic() + self.max_wait_s
            while len(self._pending) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
            return batch

    def _run(self):
        while True:
            batch = self._take_batch()
            if batch is None:
                return
            try:
                k = max(b[1] for b in batch)                    # one pass at the largest k; callers get their own head
                q = np.stack([b[0] for b in batch])
                ex = [b[2] for b in batch]
                r = self.scanner.scan_topk(q, k, exclude=ex if any(e is not None for e in ex) else None)
                self.batches_issued += 1
                self.queries_served += len(batch)
                for i, (_, ki, _, fut) in enumerate(batch):
                    n = min(int(r["count"][i]), ki)
                    fut.set_result(dict(dbidx=r["dbidx"][i, :n].copy(), score=r["score"][i, :n].copy(),
                                        row=r["row"][i, :n].copy()))
            except BaseException as e:  # noqa: BLE001 - every waiting caller must learn about the failure
                for *_, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)
