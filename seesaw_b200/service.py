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
  ``topk_from_scores``, ``topk_from_order``) — so ``B200MultiscaleIndex.from_database(ScanClient(path), vector_meta)``
  in a session process is the whole integration.  A session has one request in flight and is answered in order:
  the serial per-session contract of the reference's actors.  Stage-1 scans of all connections are batched by
  the server's I/O thread and share tensor-core passes; the other calls go to the handle on a small pool (the
  library serialises them on its mutex).  The client side imports neither CUDA nor torch.

``scanner`` is anything with ``scan_topk(queries[nq, d], k, exclude=[ids_0, ...]) -> dict(dbidx, score,
row, count)`` — a :class:`seesaw_b200.engine.PatchDatabase` in production, a stub in the CPU tests."""
from __future__ import annotations

import os
import pickle
import queue
import selectors
import socket
import struct
import threading
import time
from concurrent.futures import Future, ThreadPoolExecutor

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
# Wire format (little endian), both directions:  magic "SSW1" | op or status u32 | a u32 | b u32 | payload bytes u64 | payload
#   request  op 1 SCAN1   a = k, b = number of excluded ids; payload = query float32[dim] + ids int32[b]
#            op 2 CALL    payload = pickle((method name, args)) — the rarer calls (rescore, score_all, ...)
#   response status 0 ok  SCAN1: a = count; payload = dbidx int32[a] + score float32[a] + row int64[a];  CALL: pickle(result)
#            status 1 error, payload = utf-8 message
# The hot request (one stage-1 scan per session step) is raw bytes on purpose: with pickle and one thread per
# connection the server spent ~100 us of interpreter time per request and the GPU idled two thirds of the time.
_HDR = struct.Struct("<4sIIIQ")
_MAGIC = b"SSW1"
_OP_SCAN1, _OP_CALL = 1, 2


def _recv_exact(sock, n):
    buf = bytearray(n)
    view, got = memoryview(buf), 0
    while got < n:
        r = sock.recv_into(view[got:], n - got)
        if r == 0:
            raise EOFError("connection closed")
        got += r
    return buf


class ScanServer:
    """Owns ``db`` (a PatchDatabase or anything with its host-buffer methods) and serves it on a Unix-domain
    socket.  One thread multiplexes all connections (selectors) and forms the batches, one thread runs the GPU passes
    and answers, a small pool serves the other calls.  ``serve_forever()`` blocks; ``start()`` runs it on a thread."""

    def __init__(self, db, address, *, max_batch=64, max_wait_s=0.002):
        """``max_wait_s``: how long the oldest request of a forming batch may wait for company while the GPU is idle.
        Measured with 64 closed-loop session processes on 10M x 512 (scripts/bench_service.py): 0 ms -> 8.4k queries/s,
        p99 15.6 ms; 0.5 ms -> 16.9k, p99 12 ms; 2 ms -> 24.4k, p50 2.3 / p99 2.7 ms (full 64-query passes keep the
        sessions in step; a pass takes 1.7 ms whatever it carries, so a short batch costs everyone behind it).  The wait
        only applies while requests arrive concurrently (the previous pass carried more than one): a lone session
        is answered at once."""
        self.db, self.address = db, address
        self.max_batch, self.max_wait_s = int(max_batch), float(max_wait_s)
        self._last_batch = 0
        if os.path.exists(address):
            os.unlink(address)
        self._lsock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._lsock.bind(address)
        self._lsock.listen(256)
        self._sel = selectors.DefaultSelector()
        self._sel.register(self._lsock, selectors.EVENT_READ, None)
        self._stop = threading.Event()
        self._ready = queue.Queue()            # batches for the GPU thread
        self._busy = 0                         # batches queued or running (I/O thread increments, GPU thread decrements)
        self._busy_lock = threading.Lock()
        self._wake_r, self._wake_w = socket.socketpair()      # the GPU thread pokes the I/O thread when it falls idle
        self._wake_r.setblocking(False)
        self._sel.register(self._wake_r, selectors.EVENT_READ, None)
        self._pool = ThreadPoolExecutor(4, thread_name_prefix="ssw-call")
        self._send_locks = {}
        self.batches_issued = self.queries_served = self.connections_served = 0
        self._gpu_thread = threading.Thread(target=self._gpu_loop, name="ssw-scan-gpu", daemon=True)
        self._gpu_thread.start()

    def info(self):
        db = self.db
        return dict(n_rows=int(db.n_rows), n_images=int(db.n_images), dim=int(db.dim), dtype=np.dtype(db.dtype).str,
                    device=int(getattr(db, "device", 0)))

    # ---- replies ------------------------------------------------------------------------------
    def _send(self, conn, status, a, payload):
        lock = self._send_locks.get(conn)
        if lock is None:
            return
        try:
            with lock:
                conn.sendall(_HDR.pack(_MAGIC, status, a, 0, len(payload)) + payload)
        except OSError:
            pass

    def _call(self, conn, blob):
        try:
            name, args = pickle.loads(blob)
            if name == "info":
                out = self.info()
            elif name == "stats":
                out = dict(batches_issued=self.batches_issued, queries_served=self.queries_served,
                           connections_served=self.connections_served)
            elif name == "shutdown":
                self._send(conn, 0, 0, pickle.dumps(None))
                self._stop.set()
                return
            elif name in ("scan_topk", "rescore", "score_all", "topk_from_scores", "topk_from_order"):
                a, kw = args
                out = getattr(self.db, name)(*a, **kw)
            else:
                raise ValueError(f"unknown request {name!r}")
            self._send(conn, 0, 0, pickle.dumps(out, protocol=pickle.HIGHEST_PROTOCOL))
        except Exception as e:      # noqa: BLE001 - the session gets the error, the server lives on
            self._send(conn, 1, 0, f"{type(e).__name__}: {e}".encode())

    # ---- the GPU thread: one pass per batch, then the answers ---------------------------------
    def _gpu_loop(self):
        dim = int(self.db.dim)
        Q = np.empty((self.max_batch, dim), np.float32)
        csr = getattr(self.db, "scan_topk_csr", None)
        while True:
            batch = self._ready.get()
            if batch is None:
                return
            n = len(batch)
            try:
                k = max(b[2] for b in batch)
                for i, b in enumerate(batch):
                    Q[i] = b[1]
                if csr is not None:
                    ids = [b[3] for b in batch]
                    off = np.zeros(n + 1, np.int64)
                    np.cumsum([len(x) for x in ids], out=off[1:])
                    r = csr(Q[:n], k, np.ascontiguousarray(np.concatenate(ids), np.int32), off)
                else:
                    r = self.db.scan_topk(Q[:n], k, exclude=[b[3] for b in batch])
                self.batches_issued += 1
                self.queries_served += n
                for i, (conn, _, ki, _) in enumerate(batch):
                    c = min(int(r["count"][i]), ki)
                    self._send(conn, 0, c, r["dbidx"][i, :c].tobytes() + r["score"][i, :c].tobytes() + r["row"][i, :c].tobytes())
            except Exception as e:      # noqa: BLE001
                msg = f"{type(e).__name__}: {e}".encode()
                for conn, *_ in batch:
                    self._send(conn, 1, 0, msg)
            with self._busy_lock:
                self._busy -= 1
            try:
                self._wake_w.send(b"x")
            except OSError:
                pass

    # ---- the I/O thread: accept, read, batch ---------------------------------------------------
    def serve_forever(self):
        dim = int(self.db.dim)
        bufs, pending, deadline = {}, [], None
        try:
            while not self._stop.is_set():
                # a batch goes out when it is full, or when its oldest request has waited max_wait_s AND the GPU is idle:
                # while a pass is running the next batch keeps growing (the pass takes longer than any sensible wait)
                if deadline is None or self._busy > 0:
                    timeout = 0.05
                else:
                    timeout = max(0.0, deadline - time.monotonic())
                for key, _ in self._sel.select(timeout):
                    sock = key.fileobj
                    if sock is self._wake_r:
                        try:
                            sock.recv(4096)
                        except OSError:
                            pass
                        continue
                    if sock is self._lsock:
                        conn, _ = sock.accept()
                        self._sel.register(conn, selectors.EVENT_READ, None)
                        bufs[conn] = bytearray()
                        self._send_locks[conn] = threading.Lock()
                        self.connections_served += 1
                        continue
                    try:
                        data = sock.recv(1 << 20)
                    except OSError:
                        data = b""
                    if not data:
                        self._sel.unregister(sock)
                        bufs.pop(sock, None)
                        self._send_locks.pop(sock, None)
                        sock.close()
                        continue
                    buf = bufs[sock]
                    buf += data
                    while len(buf) >= _HDR.size:
                        magic, op, a, b, nbytes = _HDR.unpack_from(buf)
                        if magic != _MAGIC:
                            raise RuntimeError("bad magic on the scan socket")
                        if len(buf) < _HDR.size + nbytes:
                            break
                        body = bytes(buf[_HDR.size:_HDR.size + nbytes])
                        del buf[:_HDR.size + nbytes]
                        if op == _OP_SCAN1:
                            if nbytes != dim * 4 + b * 4:
                                self._send(sock, 1, 0, b"ValueError: query width does not match the database")
                                continue
                            q = np.frombuffer(body, np.float32, dim)
                            ids = np.frombuffer(body, np.int32, b, dim * 4)
                            pending.append((sock, q, int(a), ids))
                            if deadline is None:
                                deadline = time.monotonic() + (self.max_wait_s if self._last_batch > 1 else 0.0)
                        else:
                            self._pool.submit(self._call, sock, body)
                if pending and (len(pending) >= self.max_batch or (self._busy == 0 and time.monotonic() >= deadline)):
                    with self._busy_lock:
                        self._busy += 1
                    self._last_batch = min(len(pending), self.max_batch)
                    self._ready.put(pending[: self.max_batch])
                    pending = pending[self.max_batch:]
                    deadline = time.monotonic() + self.max_wait_s if pending else None
        finally:
            self.close()

    def start(self):
        t = threading.Thread(target=self.serve_forever, name="ssw-scan-server", daemon=True)
        t.start()
        return t

    def close(self):
        self._stop.set()
        self._ready.put(None)
        for key in list(self._sel.get_map().values()):
            try:
                key.fileobj.close()
            except OSError:
                pass
        self._sel.close()
        self._wake_w.close()
        self._pool.shutdown(wait=False)
        if os.path.exists(self.address):
            try:
                os.unlink(self.address)
            except OSError:
                pass


class ScanClient:
    """A session process's handle on the server's database: the methods the index classes call on a
    :class:`seesaw_b200.engine.PatchDatabase`, answered over the socket.  One client per session process (or
    thread): requests on a connection are answered in order."""

    def __init__(self, address, *, connect_timeout_s=30.0):
        deadline = time.monotonic() + connect_timeout_s
        while True:
            self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                self._sock.connect(address)
                break
            except (FileNotFoundError, ConnectionRefusedError):
                self._sock.close()
                if time.monotonic() > deadline:
                    raise
                time.sleep(0.05)
        self._lock = threading.Lock()
        i = self._call("info")
        self.n_rows, self.n_images, self.dim, self.device = i["n_rows"], i["n_images"], i["dim"], i["device"]
        self.dtype = np.dtype(i["dtype"]).type

    def _request(self, op, a, b, payload):
        with self._lock:
            self._sock.sendall(_HDR.pack(_MAGIC, op, a, b, len(payload)) + payload)
            magic, status, ra, _, nbytes = _HDR.unpack(_recv_exact(self._sock, _HDR.size))
            body = _recv_exact(self._sock, nbytes) if nbytes else b""
        if magic != _MAGIC:
            raise RuntimeError("bad magic from the scan server")
        if status != 0:
            raise RuntimeError(f"scan server: {bytes(body).decode(errors='replace')}")
        return ra, body

    def _call(self, name, *a, **kw):
        _, body = self._request(_OP_CALL, 0, 0, pickle.dumps((name, (a, kw)), protocol=pickle.HIGHEST_PROTOCOL))
        return pickle.loads(body)

    def scan_topk(self, queries, k, exclude=None):
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, self.dim))
        if q.shape[0] != 1:      # an explicit batch of one caller: straight to the database
            return self._call("scan_topk", q, int(k), exclude=None if exclude is None else [None if e is None else np.asarray(e) for e in exclude])
        # a session's own query: raw bytes, batched with the other sessions' on the server
        ids = np.zeros(0, np.int32) if exclude is None or exclude[0] is None else np.ascontiguousarray(
            np.asarray(exclude[0] if isinstance(exclude[0], np.ndarray) else list(exclude[0])), np.int32).reshape(-1)
        n, body = self._request(_OP_SCAN1, int(k), len(ids), q.tobytes() + ids.tobytes())
        out = dict(dbidx=np.full((1, k), -1, np.int32), score=np.full((1, k), -np.inf, np.float32),
                   row=np.full((1, k), -1, np.int64), count=np.array([n], np.int32))
        out["dbidx"][0, :n] = np.frombuffer(body, np.int32, n)
        out["score"][0, :n] = np.frombuffer(body, np.float32, n, 4 * n)
        out["row"][0, :n] = np.frombuffer(body, np.int64, n, 8 * n)
        return out

    def rescore(self, query, cand_dbidx, *, query2=None, agg_method="avg_score", aug_larger="all"):
        return self._call("rescore", np.asarray(query, np.float32), np.asarray(cand_dbidx),
                          query2=None if query2 is None else np.asarray(query2, np.float32), agg_method=agg_method, aug_larger=aug_larger)

    def score_all(self, query):
        return self._call("score_all", np.asarray(query, np.float32))

    def topk_from_scores(self, scores, k, exclude=None, row_mask=None):
        return self._call("topk_from_scores", np.asarray(scores, np.float32), int(k),
                          exclude=None if exclude is None else np.asarray(exclude), row_mask=row_mask)

    def topk_from_order(self, row_order, k, exclude=None):
        return self._call("topk_from_order", np.asarray(row_order, np.int64), int(k), exclude=None if exclude is None else np.asarray(exclude))

    def set_boxes(self, *a, **k):
        """The server's database already holds the boxes (set once by the GPU-owning process)."""

    def exact_info(self):
        return dict(attached=False, rho=0.0, vmax=0.0, queries=0, rescans=0)

    def stats(self):
        return self._call("stats")

    def shutdown_server(self):
        try:
            self._call("shutdown")
        except (EOFError, OSError, RuntimeError):
            pass

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass
