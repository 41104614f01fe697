"""Stands where the reference expects `$BVN_CDF_REPO/bvn_cdf.py` (loader: src/core/tf_util.py:9-13; call site
src/core/exponentiated_quadratic.py:552).  The real `wesselb/bvn-cdf` (a TF custom op built from Genz's BVND) is not
under /root/reference and is unpinned; this op evaluates the oracle's restatement of Genz (2004) with its analytic
partial derivatives (oracle/bvn.py, pinned to scipy / mpmath in tests/test_oracle_bvn.py).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import torch
import tensorflow as tf

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', '..')))
from oracle import bvn_torch  # noqa: E402


def bvn_cdf(x1, x2, rho):
    def fn(a, b, r):
        if a.device.type == 'meta':
            return torch.empty_like(a)
        return bvn_torch.bvn_cdf(a, b, r)
    return tf.Tensor(fn, [tf.convert_to_tensor(x1), tf.convert_to_tensor(x2), tf.convert_to_tensor(rho)])
