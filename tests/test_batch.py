"""The multi-restart / hyper-parameter batch scheduler (cgpcm_b200/batch.py; the reference's unit: one process per
resample index, src/experiment_toy.sh:7-11) on the CPU: fake sessions, no device.  The GPU side is
tests/test_gpu_batch.py."""
import threading
import time

import numpy as np
import pytest

from cgpcm_b200 import batch


class FakeSession(object):
    def __init__(self, device, seed):
        self.device, self.rng = device, np.random.RandomState(seed)


def _factory(device, seed):
    return FakeSession(device, seed)


def test_results_keep_task_order_and_every_device_works():
    seen = []
    lock = threading.Lock()

    def make(i):
        def task(sess):
            time.sleep(0.01)
            with lock:
                seen.append((i, sess.device, threading.current_thread().name))
            return i * i, float(sess.rng.randn())
        return task

    out = batch.run([make(i) for i in range(12)], devices=[0, 1, 2], session_factory=_factory)
    assert [o[0] for o in out] == [i * i for i in range(12)]
    assert {d for _, d, _ in seen} == {0, 1, 2}
    assert {nm for _, _, nm in seen} == {'cgpcm-batch-dev0', 'cgpcm-batch-dev1', 'cgpcm-batch-dev2'}
    st = batch.run.last_stats
    assert sorted(i for d in st['devices'].values() for i in d['tasks']) == list(range(12))
    # the private generators make every task reproducible whatever device ran it
    again = batch.run([make(i) for i in range(12)], devices=[1], session_factory=_factory)
    assert [o[1] for o in again] == [o[1] for o in out]


def test_costs_put_the_longest_task_first():
    order = []

    def make(i):
        def task(sess):
            order.append(i)
        return task

    batch.run([make(i) for i in range(5)], devices=[0], costs=[1, 5, 2, 9, 3], session_factory=_factory)
    assert order == [3, 1, 4, 2, 0]
    p = batch.plan(5, [0, 1], costs=[1, 5, 2, 9, 3])
    assert p == {0: [3, 0], 1: [1, 4, 2]}
    assert batch.plan(4, [0, 1]) == {0: [0, 2], 1: [1, 3]}


def test_failures_are_reported_after_all_workers_stop():
    done = []

    def ok(sess):
        done.append(1)
        return 'ok'

    def bad(sess):
        raise ValueError('boom')

    with pytest.raises(batch.TaskError) as ei:
        batch.run([ok, bad, ok, ok], devices=[0, 1], session_factory=_factory)
    assert ei.value.index == 1 and isinstance(ei.value.__cause__, ValueError)
    assert len(done) == 3
    out = batch.run([ok, bad], devices=[0], session_factory=_factory, return_exceptions=True)
    assert out[0] == 'ok' and isinstance(out[1], ValueError)
    with pytest.raises(ValueError):
        batch.run([ok], devices=[])
    with pytest.raises(ValueError):
        batch.run([ok, ok], devices=[0], seeds=[1])
