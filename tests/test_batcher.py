"""Cross-session batching scheduler (livetalking_b200/plugin/batcher.py): host logic with a fake engine."""
import threading
import time

import numpy as np
import pytest

import stubs  # noqa: E402

stubs.install()


class FakeMux:
    """infer_slots(requests) -> one 'frame' per request that encodes (avatar id, idx, mel[0,0]) so routing errors are visible."""

    def __init__(self, batch=8, latency=0.002):
        self.batch, self.latency, self.sizes = batch, latency, []

    def infer_slots(self, requests):
        assert 1 <= len(requests) <= self.batch
        self.sizes.append(len(requests))
        time.sleep(self.latency)
        return np.stack([np.array([av, idx, int(mel[0, 0])], np.int64) for av, idx, mel in requests])


def test_requests_of_many_sessions_are_packed_and_routed_back():
    from livetalking_b200.plugin.batcher import CrossSessionBatcher
    mux = FakeMux(batch=8)
    b = CrossSessionBatcher(mux, max_wait_ms=20.0)
    errors = []

    def session(sid, steps, per_step):
        try:
            for k in range(steps):
                reqs = [(sid, 100 * k + i, np.full((80, 16), 1000 * sid + k, np.float32)) for i in range(per_step)]
                got = b.submit(reqs)
                assert len(got) == per_step
                for i, f in enumerate(got):
                    assert tuple(f) == (sid, 100 * k + i, 1000 * sid + k), (sid, k, i, f)
        except Exception as e:                                  # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=session, args=(s, 25, 2 + s % 3)) for s in range(12)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=60)
    b.close()
    assert not errors, errors
    total = sum(25 * (2 + s % 3) for s in range(12))
    assert b.slots == total == sum(mux.sizes)
    assert max(mux.sizes) == 8                                   # full batches were formed ...
    assert b.batches < total / 3                                 # ... and far fewer engine calls than per-session steps x frames


def test_lone_request_is_flushed_after_max_wait_and_errors_propagate():
    from livetalking_b200.plugin.batcher import CrossSessionBatcher
    mux = FakeMux(batch=16)
    b = CrossSessionBatcher(mux, max_wait_ms=30.0)
    t0 = time.monotonic()
    got = b.submit([(7, 3, np.zeros((80, 16), np.float32))])
    dt = time.monotonic() - t0
    assert tuple(got[0]) == (7, 3, 0) and 0.02 <= dt < 0.5 and mux.sizes == [1]

    def boom(requests):
        raise RuntimeError("engine failure")

    mux.infer_slots = boom
    with pytest.raises(RuntimeError, match="engine failure"):
        b.submit([(1, 1, np.zeros((80, 16), np.float32))])
    b.close()
    with pytest.raises(RuntimeError, match="closed"):
        b.submit([(1, 1, np.zeros((80, 16), np.float32))])
