"""torch-CPU float64 wrapper of the oracle BVN CDF with its analytic gradient — the counterpart of
the gradient that the reference's external ``bvn-cdf`` op registers with TensorFlow (not in
``/root/reference``; call site ``src/core/exponentiated_quadratic.py:552``).  TEST INFRASTRUCTURE ONLY.
"""
import torch

from . import bvn as _bvn


class BvnCdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, rho):
        ctx.save_for_backward(x1, x2, rho)
        out = _bvn.bvn_cdf(x1.detach().numpy(), x2.detach().numpy(), rho.detach().numpy())
        return torch.as_tensor(out, dtype=torch.float64)

    @staticmethod
    def backward(ctx, g):
        x1, x2, rho = ctx.saved_tensors
        d1, d2, dr = _bvn.bvn_cdf_partials(x1.detach().numpy(), x2.detach().numpy(),
                                           rho.detach().numpy())
        t = lambda a: torch.as_tensor(a, dtype=torch.float64)
        return g * t(d1), g * t(d2), g * t(dr)


def bvn_cdf(x1, x2, rho):
    return BvnCdf.apply(x1.contiguous(), x2.contiguous(), rho.contiguous())
