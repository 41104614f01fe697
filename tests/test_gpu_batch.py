"""Independent restarts spread over the visible GPUs (cgpcm_b200.batch.run): every task trains its own model on its own
handle; results do not depend on which device ran a task or on what ran beside it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import cgpcm_b200
from cgpcm_b200 import Data, batch, config, experiment
from tests.cases import make_case


def _task(seed):
    c = make_case('toy_small')

    def task(sess):
        e = Data(c['t'], c['y'] + 0.01 * sess.rng.randn(len(c['y'])))
        mod, rep = experiment.train(sess, e, nx=c['nx'], nh=c['nh'], tau_w=.1, tau_f=.05, causal=True, reg=c['reg'],
                                    iters_pre=8, iters=10, iters_post=6, iters_fpi_post=3)
        return sess.device, rep['elbo']['final'], mod._pack()
    return task


def test_batch_of_restarts():
    config.reg = 1e-6
    ndev = batch.visible_devices()
    assert ndev == torch.cuda.device_count() >= 1
    devices = list(range(ndev))
    tasks = [_task(s) for s in range(6)]
    out = batch.run(tasks, devices=devices)
    assert {d for d, _, _ in out} <= set(devices)
    if ndev > 1:
        assert len({d for d, _, _ in out}) > 1
    # the same batch on one device, tasks interleaved differently: bit-identical results per task
    again = batch.run(tasks, devices=[devices[-1]])
    for (d0, e0, p0), (d1, e1, p1) in zip(out, again):
        assert e0 == e1 and np.array_equal(p0, p1)
    # different seeds give different series and different optima
    assert len({e for _, e, _ in out}) == 6
    # training raised the bound
    c = make_case('toy_small')
    assert all(np.isfinite(e) for _, e, _ in out)


def test_two_handles_on_two_threads_do_not_interfere():
    """Two engines (same device when only one is visible) evaluated concurrently from two threads against the
    sequential results."""
    import threading
    c = make_case('toy_test')
    ndev = batch.visible_devices()
    engs = [cgpcm_b200.Engine(c['nh'], c['nx'], device=k % ndev) for k in range(2)]
    for e in engs:
        e.set_data(c['t'], c['y'], c['th'], c['tx'])
    want = [e.elbo_grad(c['params'], reg=c['reg']) for e in engs]
    got = [None, None]

    def work(k):
        for _ in range(5):
            got[k] = engs[k].elbo_grad(c['params'], reg=c['reg'])

    ths = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k in range(2):
        assert got[k][0] == want[k][0] and np.array_equal(got[k][2], want[k][2])
