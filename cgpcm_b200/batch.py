"""Independent restarts / hyper-parameter batches spread across the GPUs of one box.

The reference's unit of this is one `controller.py -t <task> resample <i> ...` process per resample index and model
variant (``src/experiment_toy.sh:7-11``, ``src/controller.py:79-100``): every task builds its own session, seeds it,
trains its own model.  Nothing is exchanged between tasks, so here they are *replicas*: one handle (one CUDA stream,
all scratch) per task, one Python thread per GPU pulling tasks from a shared queue.  ``ctypes`` releases the GIL for
the duration of every C-ABI call, so the host threads only serialise on SciPy's L-BFGS bookkeeping.  No collective,
no NCCL: the aggregate throughput is the sum over devices.

    results = batch.run([lambda sess: train_one(sess, seed) for seed in range(16)], devices=[0, 1, 2, 3])

A task is a callable ``task(sess)``; ``sess`` is a :class:`cgpcm_b200.Session` bound to the worker's device that also
carries a private random generator (``sess.rng``, seeded per task) because numpy's global generator is shared by all
threads of the process.  ``config.reg`` is a module global exactly as in the reference (``src/config.py:3``): all tasks
of one batch must use the same value.
"""
import queue
import threading
import time

import numpy as np

from .cgpcm import Session


class TaskError(RuntimeError):
    """A task raised: ``.index`` is its position in the batch, ``.__cause__`` the original exception."""

    def __init__(self, index, exc):
        RuntimeError.__init__(self, 'task %d failed: %r' % (index, exc))
        self.index = index


def visible_devices():
    """Number of CUDA devices this process can use (no torch needed: asks the library)."""
    import ctypes
    from . import _lib
    n = ctypes.c_int(0)
    L = _lib.lib()
    if hasattr(L, 'cgpcm_device_count') and L.cgpcm_device_count(ctypes.byref(n)) == 0:
        return int(n.value)
    return 0


def plan(n_tasks, devices, costs=None):
    """Static assignment used for reporting and tests: the order in which a greedy scheduler hands ``n_tasks`` tasks
    with (optional) relative ``costs`` to ``devices`` -- longest first, always to the least-loaded device.  ``run``
    itself is dynamic (a device takes the next task when it becomes free); with equal costs both coincide."""
    devices = list(devices)
    if not devices:
        raise ValueError('no devices')
    costs = np.ones(n_tasks) if costs is None else np.asarray(costs, dtype=np.float64)
    if costs.shape[0] != n_tasks:
        raise ValueError('one cost per task')
    load = {d: 0.0 for d in devices}
    out = {d: [] for d in devices}
    for i in np.argsort(-costs, kind='stable'):
        d = min(devices, key=lambda k: (load[k], devices.index(k)))
        out[d].append(int(i))
        load[d] += float(costs[i])
    return out


def run(tasks, devices=None, seeds=None, costs=None, session_factory=None, return_exceptions=False):
    """Run ``tasks`` (callables ``task(sess) -> result``) on ``devices`` (list of CUDA device indices; default: all
    visible), one worker thread per device.  Tasks are handed out longest-first when ``costs`` are given, in order
    otherwise.  ``seeds[i]`` seeds task ``i``'s private generator (default ``i``).  Returns the results in task order;
    the first failure is re-raised as :class:`TaskError` once every worker has stopped (``return_exceptions=True``:
    exceptions are returned in place of results instead).

    ``run.last_stats`` afterwards holds per-device task lists and busy times."""
    tasks = list(tasks)
    if devices is None:
        devices = list(range(max(visible_devices(), 1)))
    devices = list(devices)
    if not devices:
        raise ValueError('no devices')
    n = len(tasks)
    seeds = list(range(n)) if seeds is None else list(seeds)
    if len(seeds) != n:
        raise ValueError('one seed per task')
    order = list(range(n)) if costs is None else [int(i) for i in np.argsort(-np.asarray(costs, dtype=np.float64),
                                                                               kind='stable')]
    todo = queue.Queue()
    for i in order:
        todo.put(i)
    results = [None] * n
    errors = [None] * n
    stats = {d: {'tasks': [], 'busy_s': 0.0} for d in devices}
    make = session_factory or (lambda device, seed: _task_session(device, seed))

    def worker(device):
        while True:
            try:
                i = todo.get_nowait()
            except queue.Empty:
                return
            t0 = time.perf_counter()
            try:
                results[i] = tasks[i](make(device, seeds[i]))
            except BaseException as exc:            # noqa: B902  (re-raised below, in the caller's thread)
                errors[i] = exc
            stats[device]['tasks'].append(i)
            stats[device]['busy_s'] += time.perf_counter() - t0

    threads = [threading.Thread(target=worker, args=(d,), name='cgpcm-batch-dev%d' % d, daemon=True) for d in devices]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    run.last_stats = {'devices': stats, 'wall_s': time.perf_counter() - t0}
    if return_exceptions:
        return [errors[i] if errors[i] is not None else results[i] for i in range(n)]
    for i in range(n):
        if errors[i] is not None:
            raise TaskError(i, errors[i]) from errors[i]
    return results


run.last_stats = None


def _task_session(device, seed):
    sess = Session(device=device, rank=0, world=1)
    sess.rng = np.random.RandomState(seed)
    return sess
