"""CPU test of the batching front end with the oracle standing in for the GPU scanner."""
import os
import threading

import numpy as np
import pytest

import seesaw_oracle as orc
from seesaw_b200 import synth
from seesaw_b200.service import ScanBatcher


class OracleScanner:
    def __init__(self, vecs, dbidx):
        self.vecs, self.dbidx, self.calls = vecs, dbidx, []

    def scan_topk(self, queries, k, exclude=None):
        nq = queries.shape[0]
        self.calls.append(nq)
        out = dict(dbidx=np.full((nq, k), -1, np.int32), score=np.full((nq, k), -np.inf, np.float32),
                   row=np.full((nq, k), -1, np.int64), count=np.zeros(nq, np.int32))
        for i in range(nq):
            o = orc.query_prelim(self.vecs, self.dbidx, queries[i], k, exclude=None if exclude is None else exclude[i])
            n = len(o["dbidx"])
            out["dbidx"][i, :n], out["score"][i, :n], out["row"][i, :n], out["count"][i] = o["dbidx"], o["max_score"], o["best_row"], n
        return out


def test_concurrent_sessions_are_batched_and_get_their_own_results():
    counts = synth.patches_per_image(300, 1, 9, 3)
    dbidx = synth.dbidx_of_rows(counts)
    vecs = synth.synth_rows(0, int(counts.sum()), 256, 5, "lattice", np.float32)
    qs = synth.lattice_queries(40, 256, 6)
    sc = OracleScanner(vecs, dbidx)
    b = ScanBatcher(sc, max_batch=16, max_wait_s=0.2)
    results = [None] * len(qs)
    ks = [3 + (i % 5) for i in range(len(qs))]
    excl = [None if i % 3 else np.arange(i, 300, 7) for i in range(len(qs))]

    def session(i):
        results[i] = b.scan_topk_one(qs[i], ks[i], excl[i], timeout=60)

    threads = [threading.Thread(target=session, args=(i,)) for i in range(len(qs))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    b.close()
    assert b.queries_served == len(qs) and b.batches_issued < len(qs) and max(sc.calls) <= 16
    for i in range(len(qs)):
        o = orc.query_prelim(vecs, dbidx, qs[i], ks[i], exclude=excl[i])
        assert (results[i]["dbidx"] == o["dbidx"]).all() and (results[i]["row"] == o["best_row"]).all()


def test_scanner_failure_reaches_every_waiter():
    class Broken:
        def scan_topk(self, *a, **k):
            raise ValueError("boom")

    b = ScanBatcher(Broken(), max_wait_s=0.01)
    futs = [b.submit(np.zeros(8, np.float32), 3) for _ in range(3)]
    for f in futs:
        try:
            f.result(10)
            assert False
        except ValueError:
            pass
    b.close()


# ---- the process boundary: one server process, many session processes (CPU: the oracle stands in for the GPU database)
def _stub_database():
    import sys as _sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in _sys.path:
            _sys.path.insert(0, p)
    from fake_db import FakeDB
    counts = synth.patches_per_image(300, 1, 9, 3)
    dbidx = synth.dbidx_of_rows(counts)
    vecs = synth.synth_rows(0, int(counts.sum()), 256, 5, "lattice", np.float32)
    return FakeDB(vecs, dbidx), vecs, dbidx


def _server_process(address):
    from seesaw_b200.service import ScanServer
    db, _, _ = _stub_database()
    ScanServer(db, address, max_batch=16, max_wait_s=0.5).serve_forever()


def _session_process(address, i, out, gate):
    from seesaw_b200.service import ScanClient
    c = ScanClient(address)
    qs = synth.lattice_queries(24, 256, 6)
    ex = None if i % 3 else [np.arange(i, 300, 7)]
    try:
        gate.wait(90)        # all sessions connected: their requests now meet inside the server's waiting window
    except Exception:        # a broken barrier only costs the batching assertion, never a hang
        pass
    r = c.scan_topk(qs[i:i + 1], 3 + i % 5, exclude=ex)
    s = c.score_all(qs[i])
    out.put((i, r["dbidx"][0], r["row"][0], int(r["count"][0]), float(s.sum()), c.n_rows))
    c.close()


def test_session_processes_share_one_server_process(tmp_path):
    """12 session PROCESSES (the reference runs one Ray actor process per session, web_session_actor.py:13-16) send
    their single queries to one database-owning process; it batches them across connections and every session gets
    exactly what a call of its own would have returned."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    address = str(tmp_path / "ssw.sock")
    server = ctx.Process(target=_server_process, args=(address,), daemon=True)
    server.start()
    out = ctx.Queue()
    gate = ctx.Barrier(12)
    sessions = [ctx.Process(target=_session_process, args=(address, i, out, gate)) for i in range(12)]
    [p.start() for p in sessions]
    got = {}
    for _ in sessions:
        rec = out.get(timeout=120)
        got[rec[0]] = rec[1:]
    [p.join(60) for p in sessions]
    from seesaw_b200.service import ScanClient
    c = ScanClient(address)
    stats = c.stats()
    _, vecs, dbidx = _stub_database()
    qs = synth.lattice_queries(24, 256, 6)
    for i in range(12):
        k = 3 + i % 5
        o = orc.query_prelim(vecs, dbidx, qs[i], k, exclude=None if i % 3 else np.arange(i, 300, 7))
        d, row, cnt, ssum, n_rows = got[i]
        assert cnt == len(o["dbidx"]) and (d[:cnt] == o["dbidx"]).all() and (row[:cnt] == o["best_row"]).all(), i
        assert ssum == float((vecs @ qs[i]).sum()) and n_rows == len(vecs)
    assert stats["queries_served"] == 12 and stats["batches_issued"] < 12       # batched across processes
    with pytest.raises(RuntimeError):
        c._call("scan_topk", np.zeros((2, 255), np.float32), 3)         # wrong width -> server-side error, server survives
    assert c.stats()["queries_served"] == 12
    c.shutdown_server()
    server.join(30)
    assert not server.is_alive()
