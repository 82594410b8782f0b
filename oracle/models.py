"""Hand-built transformer programs of the benchmark configurations — TEST INFRASTRUCTURE ONLY.

Each builder writes out what the reference's `@model` macro emits for the cited model source (SURVEY.md
Appendix A; src/rewrites.jl: `x{e}` -> Symbol(x, :_, e) :500-558, `.=` -> Assign, `~` -> Sample / AccessorSample
followed by Resample() :572-573,707-711, `=>` -> Observe + Resample(), `<<` -> Move, `for` -> Loop, `if` -> Cond),
in the style of the reference's own hand-built test programs (test/models.jl:70-254).  Argument functions are
NumPy closures over the oracle state (`st.cols[name]`: (n,) for scalar columns, (n, d) for vector columns).

The product's `wsb200.model(source)` must produce, on the device, what these produce on the host.
"""
from __future__ import annotations

import math

import numpy as np

from .trees import (Assign, Cond, Exponential, Loop, Move, MvNormal, Normal, Observe, RW, Resample, Sample, Sequence,
                    Weight, autoRW, vec)

I2 = np.eye(2)


def ssm1d(obs):
    """examples/1D_ssm.jl:7-16 (C1).  depth: 2 + 4 T."""
    def body(to):
        t, o = to
        xt, xn = f"x_{t}", f"x_{t + 1}"
        return Sequence(
            Assign(xn, lambda st: st.cols[xt] + st.cols["v"]),           # uses the OLD v
            Sample("dv", Normal, lambda st: (0.0, 0.1)), Resample(),     # no-op Resample (weights unchanged)
            Assign("v", lambda st: st.cols["v"] + st.cols["dv"]),
            Observe(lambda st: o, Normal, lambda st: (st.cols[xn], 1.0)), Resample())
    return Sequence(Assign("x_1", lambda st: 0.0), Assign("v", lambda st: 0.0),
                    Loop(lambda st: list(enumerate(obs, 1)), body))


def ssm1d_filter(obs):
    """examples/1D_ssm.jl with `x .= x + v` (no history)."""
    def body(o):
        return Sequence(
            Assign("x", lambda st: st.cols["x"] + st.cols["v"]),
            Sample("dv", Normal, lambda st: (0.0, 0.1)), Resample(),
            Assign("v", lambda st: st.cols["v"] + st.cols["dv"]),
            Observe(lambda st: o, Normal, lambda st: (st.cols["x"], 1.0)), Resample())
    return Sequence(Assign("x", lambda st: 0.0), Assign("v", lambda st: 0.0), Loop(lambda st: list(obs), body))


def ssm2d(obs):
    """examples/2D_ssm.jl:7-17 (C2, history kept).  The MvNormal arguments are COVARIANCES (0.1 I, 0.5 I)."""
    def body(to):
        t, o = to
        xt, xn = f"x_{t}", f"x_{t + 1}"
        o = np.asarray(o, dtype=np.float64)
        return Sequence(
            Assign(xn, lambda st: st.cols[xt] + st.cols["v"]),
            Sample("dv", MvNormal, lambda st: (vec(0.0, 0.0), 0.1 * I2)), Resample(),
            Assign("v", lambda st: st.cols["v"] + st.cols["dv"]),
            Observe(lambda st: o, MvNormal, lambda st: (st.cols[xn], 0.5 * I2)), Resample())
    return Sequence(Assign("x_1", lambda st: vec(0.0, 0.0)), Assign("v", lambda st: vec(1.0, 0.0)),
                    Loop(lambda st: list(enumerate(obs, 1)), body))


def ssm2d_filter(obs, init=True):
    """the filter-only form of C2 (`x .= x + v`), the benchmark's step; init=False: a continuation (no initial assigns)"""
    def body(o):
        o = np.asarray(o, dtype=np.float64)
        return Sequence(
            Assign("x", lambda st: st.cols["x"] + st.cols["v"]),
            Sample("dv", MvNormal, lambda st: (vec(0.0, 0.0), 0.1 * I2)), Resample(),
            Assign("v", lambda st: st.cols["v"] + st.cols["dv"]),
            Observe(lambda st: o, MvNormal, lambda st: (st.cols["x"], 0.5 * I2)), Resample())
    loop = Loop(lambda st: list(obs), body)
    if not init:
        return Sequence(loop)
    return Sequence(Assign("x", lambda st: vec(0.0, 0.0)), Assign("v", lambda st: vec(1.0, 0.0)), loop)


def lgssm1d(data, a, q, r, x0_std, tail=()):
    """benchmarks/ssm/WeightedSampling/lgssm1d.jl:18-24 (the published-numbers model).  depth: 1 + 2 T.
    `tail`: extra transformers appended after the loop (e.g. a Move for the smoke test)."""
    def body(y):
        return Sequence(
            Sample("x", Normal, lambda st: (a * st.cols["x"], q)), Resample(),
            Observe(lambda st: y, Normal, lambda st: (st.cols["x"], r)), Resample())
    return Sequence(Sample("x", Normal, lambda st: (0.0, x0_std)), Resample(), Loop(lambda st: list(data), body), *tail)


def linear_regression(xs, ys):
    """examples/linear_regression.jl:17-27 (C3): α, β ~ N(0, 10); y_i => N(α + β x_i, 1); `if resampled` two autoRW moves."""
    def body(xy):
        x, y = xy
        return Sequence(
            Observe(lambda st: y, Normal, lambda st: (st.cols["α"] + st.cols["β"] * x, 1.0)), Resample(),
            Cond(lambda st: st.resampled, Sequence(Move(["α"], autoRW, ()), Move(["β"], autoRW, ()))))
    return Sequence(Sample("α", Normal, lambda st: (0.0, 10.0)), Resample(),
                    Sample("β", Normal, lambda st: (0.0, 10.0)), Resample(),
                    Loop(lambda st: list(zip(xs, ys)), body))


def eight_schools(J, y, sigma):
    """examples/eight_schools.jl:7-17 (C4): θ is a J-vector column; τ's move is bounded to (0, Inf)."""
    def body(j):            # j is 1-based as in the model source; plane j - 1 of θ
        return Sequence(
            Sample(("θ", j - 1), Normal, lambda st: (st.cols["μ"], st.cols["τ"])), Resample(),
            Observe(lambda st: y[j - 1], Normal, lambda st: (st.cols["θ"][:, j - 1], sigma[j - 1])), Resample(),
            Move(["μ"], autoRW, (), 0.9),
            Move(["τ"], autoRW, (1e-3, (0.0, math.inf)), 0.9))
    return Sequence(Sample("μ", Normal, lambda st: (0.0, 5.0)), Resample(),
                    Sample("τ", Exponential, lambda st: (5.0,)), Resample(),
                    Assign("θ", lambda st: np.zeros(J)),
                    Loop(lambda st: list(range(1, J + 1)), body))


def hierarchical_regression(J, groups):
    """benchmarks/multilevel/WeightedSampling/model.jl:20-41: nested loops, build-time `if j % 10 == 0`,
    dynamic-family move targets `alpha{j}`."""
    def obs_body(j):
        aj = f"alpha_{j}"

        def f(xy):
            x, y = xy
            return Sequence(
                Observe(lambda st: y, Normal, lambda st: (st.cols[aj] + st.cols["beta"] * x, st.cols["sigma"])), Resample(),
                Cond(lambda st: st.resampled, Move([aj], autoRW, (), 0.1)))
        return f

    def body(j):
        aj = f"alpha_{j}"
        steps = [Sample(aj, Normal, lambda st: (st.cols["mu_alpha"], st.cols["tau_alpha"])), Resample(),
                 Loop(lambda st: list(groups[j - 1]), obs_body(j))]
        if j % 10 == 0:     # evaluated when the body is built, as in the reference (rewrites.jl:720-731)
            steps += [Move(["mu_alpha"], autoRW, (), 0.1),
                      Move(["tau_alpha"], autoRW, (1e-3, (0.0, math.inf)), 0.1),
                      Move(["beta"], autoRW, (), 0.1),
                      Move(["sigma"], autoRW, (1e-3, (0.0, math.inf)), 0.1)]
        return Sequence(*steps)
    return Sequence(Sample("mu_alpha", Normal, lambda st: (0.0, 10.0)), Resample(),
                    Sample("tau_alpha", Exponential, lambda st: (1.0,)), Resample(),
                    Sample("beta", Normal, lambda st: (0.0, 10.0)), Resample(),
                    Sample("sigma", Exponential, lambda st: (1.0,)), Resample(),
                    Loop(lambda st: list(range(1, J + 1)), body))


def normal_normal(y, prior_sd, obs_sd, moves):
    """test/move_test.jl:69-98 shape: x ~ N(0, prior_sd); y => N(x, obs_sd); then `moves` (a list of Move)."""
    return Sequence(Sample("x", Normal, lambda st: (0.0, prior_sd)), Resample(),
                    Observe(lambda st: y, Normal, lambda st: (st.cols["x"], obs_sd)), Resample(), *moves)


__all__ = ["ssm1d", "ssm1d_filter", "ssm2d", "ssm2d_filter", "lgssm1d", "linear_regression", "eight_schools",
           "hierarchical_regression", "normal_normal", "RW", "autoRW", "Move", "Weight"]
