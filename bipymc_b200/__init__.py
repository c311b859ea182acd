"""bipymc_b200 -- B200-native DE-MC / DREAM population-MCMC engine.

Drop-in for the parallel samplers of wgurecky/bipymc (``DeMcMpi``, ``DreamMpi``,
``McmcChain``): same constructors, ``run_mcmc`` and ``param_est``; the per-generation
update runs as hand-written sm_100a CUDA kernels behind a C-ABI
(include/bipymc_b200.h).  Importing the package does not need a GPU; constructing a
sampler does, and fails loudly otherwise (no CPU fallback).
"""
from .chain import McmcChain  # noqa: F401
from .demc import DeMcMpi  # noqa: F401
from .dream import DreamMpi  # noqa: F401
from .samplers import DeMc  # noqa: F401
from . import targets  # noqa: F401

__all__ = ["McmcChain", "DeMcMpi", "DreamMpi", "DeMc", "targets"]
