"""Batching front end: many single-threaded sessions, one GPU pass.

The reference serves every user from its own single-threaded ``WebSession`` Ray actor PROCESS
(seesaw/web/web_session_actor.py:13-16, 70-82) and each of them scans the read-only index on its own.
With the database in the HBM of one GPU, concurrent ``query`` calls are better served together: the
batched kernel reads the database ONCE for up to 64 queries (SSW_MAX_BATCH).

Two layers:

* :class:`ScanBatcher` (in one process) collects the stage-1 requests of concurrent callers for at most
  ``max_wait_s`` (or until ``max_batch`` are waiting), issues one ``scan_topk`` and hands each caller its own
  rows of the result.
* :class:`ScanServer` / :class:`ScanClient` put a process boundary in front of it: ONE process owns the GPU
  and the database and listens on a Unix-domain socket; every session process connects a ``ScanClient``, which
  offers the methods the index classes call on a database (``scan_topk``, ``rescore``, ``score_all``,
  ``topk_from_scores``) — so ``B200MultiscaleIndex.from_database(ScanClient(path), vector_meta)`` in a session
  process is the whole integration.  A connection is served by one thread, request after request: the serial
  per-session contract of the reference's actors.  Stage-1 scans of all connections meet in the server's
  ScanBatcher; the other calls go straight to the handle (the library serialises them on its mutex).
  The client side imports neither CUDA nor torch.

``scanner`` is anything with ``scan_topk(queries[nq, d], k, exclude=[ids_0, ...]) -> dict(dbidx, score,
row, count)`` — a :class:`seesaw_b200.engine.PatchDatabase` in production, a stub in the CPU tests."""
from __future__ import annotations

import os
import threading
import time
from concurrent.futures import Future
from multiprocessing.connection import Client, Listener

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


# ---------------------------------------------------------------------------------------------------
# process boundary: one GPU-owning server, one client per session process
# ---------------------------------------------------------------------------------------------------
class ScanServer:
    """Owns ``db`` (a PatchDatabase or anything with its host-buffer methods) and serves it on a Unix-domain
    socket.  ``serve_forever()`` blocks; ``start()`` runs the accept loop on a daemon thread."""

    def __init__(self, db, address, *, max_batch=64, max_wait_s=0.0005, authkey=b"seesaw_b200"):
        self.db, self.address = db, address
        if os.path.exists(address):
            os.unlink(address)
        self._authkey = authkey
        self._listener = Listener(address, family="AF_UNIX", authkey=authkey)
        self.batcher = ScanBatcher(db, max_batch=max_batch, max_wait_s=max_wait_s)
        self._stop = threading.Event()
        self._threads = []
        self.connections_served = 0

    def info(self):
        db = self.db
        return dict(n_rows=int(db.n_rows), n_images=int(db.n_images), dim=int(db.dim), dtype=np.dtype(db.dtype).str,
                    device=int(getattr(db, "device", 0)))

    def _serve_connection(self, conn):
        try:
            while True:
                try:
                    op, args = conn.recv()
                except (EOFError, OSError):
                    return
                try:
                    if op == "scan1":                      # one query of one session: through the batcher
                        out = self.batcher.scan_topk_one(*args)
                    elif op == "scan":                     # an explicit batch of one caller
                        out = self.db.scan_topk(*args)
                    elif op == "rescore":
                        q, ids, kw = args
                        out = self.db.rescore(q, ids, **kw)
                    elif op == "score_all":
                        out = self.db.score_all(*args)
                    elif op == "topk_from_scores":
                        sc, k, kw = args
                        out = self.db.topk_from_scores(sc, k, **kw)
                    elif op == "topk_from_order":
                        order, k, ex = args
                        out = self.db.topk_from_order(order, k, exclude=ex)
                    elif op == "info":
                        out = self.info()
                    elif op == "stats":
                        out = dict(batches_issued=self.batcher.batches_issued, queries_served=self.batcher.queries_served,
                                   connections_served=self.connections_served)
                    elif op == "shutdown":
                        conn.send(("ok", None))
                        self._stop.set()
                        try:                               # wake the accept loop
                            Client(self.address, family="AF_UNIX", authkey=self._authkey).close()
                        except OSError:
                            pass
                        return
                    else:
                        raise ValueError(f"unknown request {op!r}")
                    conn.send(("ok", out))
                except Exception as e:      # noqa: BLE001 - the session gets the error, the server lives on
                    conn.send(("error", f"{type(e).__name__}: {e}"))
        finally:
            conn.close()

    def serve_forever(self):
        while not self._stop.is_set():
            try:
                conn = self._listener.accept()
            except OSError:
                break
            if self._stop.is_set():
                conn.close()
                break
            self.connections_served += 1
            t = threading.Thread(target=self._serve_connection, args=(conn,), daemon=True)
            t.start()
            self._threads.append(t)
        self.close()

    def start(self):
        t = threading.Thread(target=self.serve_forever, name="ssw-scan-server", daemon=True)
        t.start()
        return t

    def close(self):
        self._stop.set()
        try:
            self._listener.close()
        except OSError:
            pass
        self.batcher.close()


class ScanClient:
    """A session process's handle on the server's database: the methods the index classes call on a
    :class:`seesaw_b200.engine.PatchDatabase`, answered over the socket.  One client per session process (or
    thread): requests on a connection are answered in order."""

    def __init__(self, address, *, authkey=b"seesaw_b200", connect_timeout_s=30.0):
        deadline = time.monotonic() + connect_timeout_s
        while True:
            try:
                self._conn = Client(address, family="AF_UNIX", authkey=authkey)
                break
            except (FileNotFoundError, ConnectionRefusedError):
                if time.monotonic() > deadline:
                    raise
                time.sleep(0.05)
        self._lock = threading.Lock()
        i = self._call("info", None)
        self.n_rows, self.n_images, self.dim, self.device = i["n_rows"], i["n_images"], i["dim"], i["device"]
        self.dtype = np.dtype(i["dtype"]).type

    def _call(self, op, args):
        with self._lock:
            self._conn.send((op, args))
            status, out = self._conn.recv()
        if status != "ok":
            raise RuntimeError(f"scan server: {out}")
        return out

    def scan_topk(self, queries, k, exclude=None):
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, self.dim))
        if q.shape[0] == 1:      # a session's own query: batched with the other sessions' on the server
            ex = None if exclude is None else exclude[0]
            r = self._call("scan1", (q[0], int(k), None if ex is None else np.asarray(ex)))
            n = len(r["dbidx"])
            out = dict(dbidx=np.full((1, k), -1, np.int32), score=np.full((1, k), -np.inf, np.float32),
                       row=np.full((1, k), -1, np.int64), count=np.array([n], np.int32))
            out["dbidx"][0, :n], out["score"][0, :n], out["row"][0, :n] = r["dbidx"], r["score"], r["row"]
            return out
        return self._call("scan", (q, int(k), None if exclude is None else [None if e is None else np.asarray(e) for e in exclude]))

    def rescore(self, query, cand_dbidx, *, query2=None, agg_method="avg_score", aug_larger="all"):
        return self._call("rescore", (np.asarray(query, np.float32), np.asarray(cand_dbidx),
                                      dict(query2=None if query2 is None else np.asarray(query2, np.float32),
                                           agg_method=agg_method, aug_larger=aug_larger)))

    def score_all(self, query):
        return self._call("score_all", (np.asarray(query, np.float32),))

    def topk_from_scores(self, scores, k, exclude=None, row_mask=None):
        return self._call("topk_from_scores", (np.asarray(scores, np.float32), int(k),
                                               dict(exclude=None if exclude is None else np.asarray(exclude), row_mask=row_mask)))

    def topk_from_order(self, row_order, k, exclude=None):
        return self._call("topk_from_order", (np.asarray(row_order, np.int64), int(k), None if exclude is None else np.asarray(exclude)))

    def set_boxes(self, *a, **k):
        """The server's database already holds the boxes (set once by the GPU-owning process)."""

    def exact_info(self):
        return dict(attached=False, rho=0.0, vmax=0.0, queries=0, rescans=0)

    def stats(self):
        return self._call("stats", None)

    def shutdown_server(self):
        try:
            self._call("shutdown", None)
        except (EOFError, OSError, RuntimeError):
            pass

    def close(self):
        try:
            self._conn.close()
        except OSError:
            pass
