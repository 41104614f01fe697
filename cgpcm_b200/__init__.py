"""cgpcm_b200 — B200-native evaluation of the VCGPCM evidence lower bound and its gradient.

Drop-in for ONE path of wesselb/cgpcm (``VCGPCM.from_recipe`` / ``precompute`` / ``elbo`` /
``learn.minimise_lbfgs``); all arithmetic runs in ``lib/libcgpcm_b200.so`` (hand-written sm_100a CUDA,
FP64) behind the C-ABI of ``include/cgpcm_b200.h``.  There is no CPU fallback.
"""
from . import batch, config, experiment, learn, sample, util
from .cgpcm import VCGPCM, CGPCM, AKM, Session, Var, Objective, shard_bounds, window_costs, window_radius, rebalance_costs
from .data import Data, UncertainData
from .engine import Engine, bvn_cdf, TERM_NAMES, n_params
from ._lib import (build, lib, LIB_PATH, CgpcmError, GRAD_ALL, GRAD_S2, GRAD_S2F, GRAD_ALPHA, GRAD_GAMMA,
                   GRAD_OMEGA, GRAD_MU_U, GRAD_VAR_U, MODE_FROZEN, MODE_FULL)

__all__ = ['VCGPCM', 'CGPCM', 'Session', 'Data', 'Engine', 'bvn_cdf', 'learn', 'config', 'util', 'build', 'lib']
