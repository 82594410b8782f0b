"""The product's statement lowering (csrc/ws_lowering.h) and micro-op interpreter (csrc/ws_vm.cuh), run on
the CPU by tests/host/harness.cpp, against the oracle — no GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import models
import wsb200 as ws
from hostlib import HostState, lib
from oracle import ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def strip_resample(t):
    """the harness has no resampler: run the statements between resamples"""
    k = type(t).__name__
    if k == "Sequence":
        return ws.Sequence(*[strip_resample(s) for s in t.steps if type(s).__name__ not in ("Resample", "Move")])
    if k == "Loop":
        return ws.Loop(t.collfn, lambda x, f=t.bodyfn: strip_resample(f(x)))
    if k == "Cond":
        return ws.Sequence()
    return t


@pytest.mark.parametrize("src,args,nn,ne", [
    (models.SSM1D, lambda r: (list(r.normal(size=9)),), 9, 0),
    (models.SSM2D, lambda r: ([r.normal(size=2) for _ in range(7)],), 14, 0),
    (models.LGSSM1D, lambda r: (list(r.normal(size=6)), 0.9, 1.0, 0.5, 2.0), 7, 0),
    (models.LINREG, lambda r: (r.uniform(0, 10, 5), r.normal(size=5)), 2, 0),
    (models.SCHOOLS, lambda r: (8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA), 9, 1),
])
def test_models_lower_and_execute_like_the_oracle(src, args, nn, ne):
    n = 300
    rng = np.random.default_rng(11)
    root = strip_resample(ws.model(src)(*args(rng)))
    normals, expon = rng.standard_normal(n * nn), rng.standard_exponential(n * ne)
    hs = HostState(n)
    hs.store.set_replay(normals=normals, exponentials=expon)
    root.apply(hs)
    ost = ref.OracleState(n, ref.Streams(normals, (), expon), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(root, ost)
    assert hs.store.colnames() == ost.names
    for name in ost.names:
        np.testing.assert_allclose(hs.store.getcol(name), ost.cols[name], rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(hs.store.logw(), ost.weights, rtol=1e-12, atol=1e-12)
    # score tape prefixes == the oracle's score! walk at every depth that ends on a scored statement
    ost.root = root
    d_full = ost.depth
    np.testing.assert_allclose(hs.store.score(hs.store.tape_len()), ref.score_logpdf(ost, [], d_full), rtol=1e-12, atol=1e-12)


def test_fusion_window_is_one_pass_for_an_ssm_step():
    n = 64
    hs = HostState(n)
    hs.store.setcol("x", np.zeros((n, 2)))
    hs.store.setcol("v", np.ones((n, 2)))
    step = ws.Sequence(ws.Assign("x", ws.col("x") + ws.col("v")),
                       ws.Sample("dv", "MvNormal", ([0.0, 0.0], 0.1 * np.eye(2))),
                       ws.Assign("v", ws.col("v") + ws.col("dv")),
                       ws.Observe([0.3, -0.2], "MvNormal", (ws.col("x"), 0.5 * np.eye(2))))
    step.apply(hs)
    L, h = hs.store.L, hs.store.h
    assert L.hh_n_flush(h) == 0                       # nothing executed yet: all four statements are queued
    assert L.hh_window_loads(h) == 4 and L.hh_window_stores(h) == 6
    assert L.hh_window_ops(h) <= 12 and L.hh_window_regs(h) <= 12
    hs.store.flush()
    assert L.hh_n_flush(h) == 1


SL_SIGS = {"ssm2d": 0, "ssm2d_hist": 1, "lgssm1d": 2, "ssm1d": 3, "linreg_obs": 4}


def _sig_of(setup, step):
    hs = HostState(32)
    for name, v in setup.items():
        hs.store.setcol(name, v)
    step.apply(hs)
    return hs.store.L.hh_window_signature(hs.store.h)


def test_benchmark_steps_match_their_straight_line_signatures():
    """csrc/ws_vm_sl.cuh: the windows that carry the benchmarks run on compile-time executors; the signature is the
    op pattern in canonical register numbering, so this pins the lowering AND the table (a changed lowering would
    silently fall back to the interpreter otherwise)."""
    n = 32
    z1, z2 = np.zeros(n), np.zeros((n, 2))
    I2 = np.eye(2)
    ssm2d = ws.Sequence(ws.Assign("x", ws.col("x") + ws.col("v")), ws.Sample("dv", "MvNormal", ([0.0, 0.0], 0.1 * I2)),
                        ws.Assign("v", ws.col("v") + ws.col("dv")),
                        ws.Observe([0.3, -0.2], "MvNormal", (ws.col("x"), 0.5 * I2)))
    assert _sig_of({"x": z2, "v": z2 + 1}, ssm2d) == SL_SIGS["ssm2d"]
    # the second and later steps (dv exists already) are the same window
    assert _sig_of({"x": z2, "v": z2 + 1, "dv": z2}, ssm2d) == SL_SIGS["ssm2d"]
    hist = ws.Sequence(ws.Assign("x_2", ws.col("x_1") + ws.col("v")), ws.Sample("dv", "MvNormal", ([0.0, 0.0], 0.1 * I2)),
                       ws.Assign("v", ws.col("v") + ws.col("dv")),
                       ws.Observe([0.3, -0.2], "MvNormal", (ws.col("x_2"), 0.5 * I2)))
    assert _sig_of({"x_1": z2, "v": z2 + 1}, hist) == SL_SIGS["ssm2d_hist"]
    lg = ws.Sequence(ws.Sample("x", "Normal", (0.9 * ws.col("x"), 1.0)), ws.Observe(0.3, "Normal", (ws.col("x"), 0.5)))
    assert _sig_of({"x": z1}, lg) == SL_SIGS["lgssm1d"]
    s1 = ws.Sequence(ws.Assign("x", ws.col("x") + ws.col("v")), ws.Sample("dv", "Normal", (0.0, 0.1)),
                     ws.Assign("v", ws.col("v") + ws.col("dv")), ws.Observe(0.3, "Normal", (ws.col("x"), 1.0)))
    assert _sig_of({"x": z1, "v": z1}, s1) == SL_SIGS["ssm1d"]
    obs = ws.Observe(0.3, "Normal", (ws.col("a") + ws.col("b") * 1.7, 1.0))
    assert _sig_of({"a": z1, "b": z1}, obs) == SL_SIGS["linreg_obs"]
    # anything else stays on the interpreter
    other = ws.Sequence(ws.Sample("x", "Normal", (0.9 * ws.col("x"), 1.0)), ws.Observe(0.3, "Normal", (ws.col("x"), ws.col("s"))))
    assert _sig_of({"x": z1, "s": z1 + 1}, other) == -1
    assert _sig_of({"x": z1}, ws.Assign("x", ws.col("x") * ws.col("x"))) == -1


def test_in_place_updates_and_register_recycling():
    n = 100
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=n), rng.normal(size=n)
    hs = HostState(n)
    hs.store.setcol("a", a)
    hs.store.setcol("b", b)
    ws.Assign("t", ws.exp(ws.col("a")) * 2.0 + 1.0).apply(hs)         # temporaries are freed afterwards ...
    ws.Assign("c", ws.col("b") * ws.col("b") - ws.col("a")).apply(hs)  # ... and `b` must still load into a fresh register
    ws.Assign("a", ws.col("a") * 3.0 + ws.col("c")).apply(hs)          # in place
    ws.Assign("a", ws.col("a") - 1.0).apply(hs)
    np.testing.assert_allclose(hs.store.getcol("t"), np.exp(a) * 2 + 1, rtol=1e-15)
    c = b * b - a
    np.testing.assert_allclose(hs.store.getcol("c"), c, rtol=1e-15)
    np.testing.assert_allclose(hs.store.getcol("a"), a * 3.0 + c - 1.0, rtol=1e-15)


def test_philox_known_answer_and_normals():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors) + Box-Muller moments."""
    L = lib()
    out = (C.c_uint32 * 4)()
    L.hh_philox(0, 0, 0, out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.hh_philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, out)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.hh_philox(0x85a308d3243f6a88, 0x0370734413198a2e, 0x299f31d0a4093822, out)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    z = np.empty(2)
    zs = []
    for i in range(20000):
        L.hh_randn2(i, 3, 99, z.ctypes.data_as(C.c_void_p))
        zs.extend(z.tolist())
    zs = np.array(zs)
    assert abs(zs.mean()) < 0.02 and abs(zs.var() - 1) < 0.03 and abs((zs ** 4).mean() - 3) < 0.15


def test_slot_count_function_equals_searchsorted():
    """ws_count_slots_le: F(C) = #{n : u_n <= C}, the per-particle form of icdf."""
    L = lib()
    rng = np.random.default_rng(4)
    for n in (1, 3, 1000, 4097):
        r = rng.random(n)
        r[rng.random(n) < 0.1] = 0.0
        us = ref.stratified_us(r)
        cs = np.concatenate([rng.random(500), us[rng.integers(0, n, 200)], [0.0, 1.0, us[-1], np.nextafter(us[0], 0)]])
        out = np.empty(cs.size, dtype=np.int64)
        L.hh_count_slots_le(cs.ctypes.data_as(C.c_void_p), cs.size, r.ctypes.data_as(C.c_void_p), n,
                            out.ctypes.data_as(C.c_void_p))
        np.testing.assert_array_equal(out, np.searchsorted(us, cs, side="right"))


def test_golden_ssm2d_without_resampling_prefix():
    g = np.load(os.path.join(GOLD, "ssm2d.npz"))
    n = g["weights"].shape[0]
    obs = [o for o in g["obs"]][:1]
    root = strip_resample(ws.model(models.SSM2D)(obs))
    hs = HostState(n)
    hs.store.set_replay(normals=g["normals"])
    root.apply(hs)
    ost = ref.OracleState(n, ref.Streams(g["normals"]), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(root, ost)
    np.testing.assert_allclose(hs.store.getcol("x_2"), ost.cols["x_2"], rtol=1e-14)
    np.testing.assert_allclose(hs.store.logw(), ost.weights, rtol=1e-12)


def test_score_tape_fuses_normal_terms_into_one_op():
    """alpha + beta * x_i with constant sigma (C3's likelihood): one ACC_SQLIN2 per observation on the score tape,
    the constant kept aside; the fold still equals the sum of Normal log-densities."""
    n = 200
    rng = np.random.default_rng(5)
    xs, ys = rng.uniform(0, 10, 7), rng.normal(size=7)
    root = strip_resample(ws.model(models.LINREG)(xs, ys))
    normals = rng.standard_normal(2 * n)
    hs = HostState(n)
    hs.store.set_replay(normals=normals)
    root.apply(hs)
    L, h = hs.store.L, hs.store.h
    L.hh_score_ops.argtypes = [C.c_void_p]
    assert L.hh_score_ops(h) == 2 + 7                      # two priors + one op per observation
    a, b = hs.store.getcol("α"), hs.store.getcol("β")
    want = ref.normal_logpdf(a, 0.0, 10.0) + ref.normal_logpdf(b, 0.0, 10.0)
    for x, y in zip(xs, ys):
        want = want + ref.normal_logpdf(y, a + b * x, 1.0)
    np.testing.assert_allclose(hs.store.score(hs.store.tape_len()), want, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hs.store.score(3), ref.normal_logpdf(a, 0.0, 10.0) + ref.normal_logpdf(b, 0.0, 10.0)
                               + ref.normal_logpdf(ys[0], a + b * xs[0], 1.0), rtol=1e-12, atol=1e-12)


HIER_SMALL = '''
@model function hier(J, groups)
    mu ~ Normal(0.0, 5.0)
    sigma ~ Exponential(1.0)
    beta ~ Normal(0.0, 2.0)
    for j in 1:J
        alpha{j} ~ Normal(mu, 1.5)
        for (x, y) in groups[j]
            y => Normal(alpha{j} + beta * x, sigma)
        end
    end
end
'''


@pytest.mark.parametrize("which", ["linreg", "hier"])
def test_device_order_fold_is_bit_identical(which):
    """The score fold as the kernels execute it (registers renumbered per launch, ops decoded to WsDop, runs of
    squared-residual entries over the same registers as one super-op, 128-entry chunks) must give the plain
    entry-by-entry fold bit for bit, at every tape prefix: a move's accept decision depends on it."""
    n = 64
    rng = np.random.default_rng(9)
    if which == "linreg":
        npts = 300                                             # > 2 chunks of 128 entries: runs are cut at chunk borders
        xs, ys = rng.uniform(0, 10, npts), rng.normal(size=npts)
        root = strip_resample(ws.model(models.LINREG)(xs, ys))
        normals, expo = rng.standard_normal(2 * n), None
    else:
        J, n_obs = 5, 40
        groups = [[(float(rng.uniform(0, 5)), float(rng.normal())) for _ in range(n_obs)] for _ in range(J)]
        root = strip_resample(ws.model(HIER_SMALL)(J, groups))
        normals, expo = rng.standard_normal((2 + J) * n), rng.exponential(size=n)
    hs = HostState(n)
    hs.store.autoflush = True
    hs.store.set_replay(normals=normals, exponentials=expo if expo is not None else ())
    root.apply(hs)
    T = hs.store.tape_len()
    assert not np.isnan(hs.store.score(T)).any()
    for k in sorted({0, 1, 2, 3, T // 3, T // 2, T - 1, T}):
        plain = hs.store.score(k)
        dev, runs, rows = hs.store.score_device_order(k)
        assert np.array_equal(plain, dev), (which, k, float(np.max(np.abs(plain - dev))))
    dev, runs, rows = hs.store.score_device_order(T)
    if which == "linreg":
        assert runs == 3 and rows <= 3          # 300 terms = 126 + 128 + 46 after the two priors; registers: alpha, beta (+0 temporaries)
    else:
        assert runs == 5 and rows <= 4 + 5 + 2  # one run per group; mu, sigma, beta, alpha_j and the 1/sigma, log sigma cache
