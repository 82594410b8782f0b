"""The oracle's OWN transformer trees — TEST INFRASTRUCTURE ONLY (see oracle/ref.py).

The reference's tests build their programs by hand out of `Sample(:x, kernel, state -> (args...))`,
`Observe(...)`, `Sequence`, `Loop`, `Cond`, `Resample()`, `Move` (test/models.jl:70-254), without going through
`@model`.  So does the oracle: the classes below are plain records that `oracle/ref.py`'s `apply` / `score_walk`
interpret (by class name), their argument functions are NumPy closures over the oracle state, and
`oracle/models.py` writes out, statement by statement, the transformer programs that `@model` emits for the
benchmark configurations (derived from src/rewrites.jl:146-219,500-558,643-752).

Nothing here (or anywhere under oracle/) imports the product package: a wrong auto-`Resample()` insertion, depth
count, argument order or name mangling in the product's `@model` front-end shows up as a disagreement between
`wsb200.model(src)` run on the device and these trees run on the host.
"""
from __future__ import annotations

import numpy as np


class Kernel:
    """a `default_kernels` entry (src/default_kernels.jl:83-102), resolved BY NAME by oracle/ref.py"""

    def __init__(self, name, p=None):
        self.name = name
        self.p = p


Normal = Kernel("Normal")
MvNormal = Kernel("MvNormal")
Exponential = Kernel("Exponential")


def importance_kernel(pm, ps, tm, ts):
    """src/default_kernels.jl:57-64 with Normal proposal / target"""
    return Kernel("importance_kernel", (pm, ps, tm, ts))


class Proposal:
    def __init__(self, name):
        self.name = name


RW = Proposal("RW")
autoRW = Proposal("autoRW")


class Assign:                                   # transformers.jl:18-32
    def __init__(self, lhs, argfn):
        self.lhs, self.argfn = lhs, argfn


class Sample:                                   # transformers.jl:150-182 (lhs = (name, j): AccessorSample :95-131)
    def __init__(self, lhs, kernel, argfn):
        self.lhs, self.kernel, self.argfn = lhs, kernel, argfn


class Observe:                                  # transformers.jl:205-235
    def __init__(self, lhsfn, kernel, argfn):
        self.lhsfn, self.kernel, self.argfn = lhsfn, kernel, argfn


class Weight:                                   # transformers.jl:252-289
    def __init__(self, kernel, argfn):
        self.kernel, self.argfn = kernel, argfn


class Sequence:                                 # transformers.jl:318-334
    def __init__(self, *steps):
        self.steps = tuple(steps[0]) if len(steps) == 1 and isinstance(steps[0], (tuple, list)) else tuple(steps)


class Loop:                                     # transformers.jl:362-383
    def __init__(self, collfn, bodyfn):
        self.collfn, self.bodyfn = collfn, bodyfn


class Cond:                                     # transformers.jl:410-428
    def __init__(self, predfn, body):
        self.predfn, self.body = predfn, body


class Resample:                                 # transformers.jl:460-498
    pass


class Move:                                     # transformers.jl:535-623
    def __init__(self, targets, proposal, argfn=(), diversity_threshold=None):
        self.targets = list(targets)
        self.proposal = proposal
        self.argfn = argfn
        self.diversity_threshold = diversity_threshold


def col(name):
    """`getcol(state.store, name)` as an argument closure piece"""
    return lambda st: st.cols[name]


def vec(*xs):
    return np.asarray(xs, dtype=np.float64)
