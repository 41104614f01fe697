"""The N > 1 path on CPU: two gloo ranks.  Checks the host-side protocol the GPU path follows — contiguous
observation shards (shard_bounds), identical replicated variables, one SUM all-reduce of the packed
forward partials [sum_Axx, C1, Q, Y, n, sum_y2] — by computing the partials of each shard with the oracle,
reducing them over gloo and comparing with the whole series; and the Session / communicator bootstrap."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om
from tests.cases import make_case


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(c, lo, hi):
    """Packed forward partials of observations [lo, hi) at fixed H, iKx (what one GPU accumulates)."""
    hyp = [torch.tensor(v, dtype=torch.float64) for v in c['hyp']]
    k = om.prior_kernels(c['th'], c['tx'], *hyp, c['reg'])
    t, y = c['t'][lo:hi], c['y'][lo:hi]
    nh, nx = c['nh'], c['nx']
    if hi == lo:
        return np.zeros(nx * nx * 2 + nh * nh + nh * nx + 2)
    a, Ahh, Axx, Ahx = om.psi_closed(t, c['th'], c['tx'], *hyp)
    rng = np.random.default_rng(5)
    H = rng.standard_normal((nh, nh))
    H = torch.tensor(H + H.T)
    C1 = torch.sum(Ahx.transpose(-1, -2) @ (H @ Ahx), 0)
    Q = torch.sum(Ahx @ (k['iKx'] @ Ahx.transpose(-1, -2)), 0)
    Y = torch.sum(torch.tensor(y)[:, None, None] * Ahx, 0)
    return np.concatenate([Axx.sum(0).numpy().ravel(), C1.numpy().ravel(), Q.numpy().ravel(), Y.numpy().ravel(),
                           [hi - lo, float(np.sum(y ** 2))]])


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import cgpcm_b200
        from cgpcm_b200 import cgpcm as cg
        c = make_case('toy_small', n=41)          # odd length: shards of 21 and 20
        lo, hi = cg.shard_bounds(len(c['t']), rank, world)
        buf = torch.tensor(_partials(c, lo, hi))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        whole = _partials(c, 0, len(c['t']))
        err = float(np.abs(buf.numpy() - whole).max() / np.abs(whole).max())

        # Session picks the process-group facts up; the model broadcasts rank 0's communicator id
        class FakeEngine(object):
            ids = []

            def __init__(self, nh, nx, **kw):
                pass

            @staticmethod
            def unique_id():
                return b'id-from-rank-%d' % rank + b'\0' * 100

            def comm_init(self, uid, r, w):
                FakeEngine.ids.append((bytes(uid), r, w))

            def set_data(self, t, y, th, tx):
                self.n = len(t)

        cg.Engine = FakeEngine
        sess = cgpcm_b200.Session(device=0)
        np.random.seed(rank)
        mod = cg.VCGPCM.from_recipe(sess, cgpcm_b200.Data(c['t'], c['y']), nx=c['nx'], nh=c['nh'], tau_w=.1,
                                    tau_f=.05, causal=True)
        uid, r, w = FakeEngine.ids[0]
        # every host-side random draw must be identical on all ranks although they seed np.random differently
        draws = np.concatenate([mod.vars['mu_u'].value.ravel(), mod.vars['var_u'].value.ravel(),
                                mod.sample_q().ravel(), mod.sample_prior().ravel(), mod._rng.uniform(0, 1, 3)])
        q.put((rank, err, sess.rank, sess.world, uid[:15], r, w, mod.engine.n, mod.n, draws.tobytes()))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_and_bootstrap():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][9] == res[1][9], 'random draws differ between ranks'
    for rank, err, srank, sworld, uid, r, w, n_local, n, _ in res:
        assert err < 1e-13                       # reduced partials == partials of the whole series
        assert (srank, sworld, r, w) == (rank, world, rank, world)
        assert uid == b'id-from-rank-0\0'        # every rank joined with rank 0's id
        assert n == 41
    # contiguous shards of equal estimated cost (cgpcm.window_costs): a partition of the 41 observations, near-equal
    # sizes here because the toy windows cover the whole series
    sizes = [x[7] for x in res]
    assert sum(sizes) == 41 and abs(sizes[0] - sizes[1]) <= 3 and min(sizes) >= 1
