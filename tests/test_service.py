"""CPU test of the batching front end with the oracle standing in for the GPU scanner."""
import threading

import numpy as np

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
