"""The custom FP64 routines of csrc/ws_math.cuh (exp for non-positive arguments, log / sqrt of positive
normals, octant sin/cos, Box-Muller from raw bits, Markstein division) — host instantiation against mpmath
and libm; the GPU test checks that the device instantiation (MUFU seeds instead of libm) gives the same
normals."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest
import scipy.stats as sst

from hostlib import HostState, lib

ULP = 2.0 ** -53
vp = C.c_void_p


def call(fn, *arrs):
    n = arrs[-1].size
    getattr(lib(), fn)(*[a.ctypes.data_as(vp) for a in arrs], C.c_int64(n))


def test_exp_nonpos():
    rng = np.random.default_rng(0)
    n = 100_000
    x = -np.abs(np.concatenate([rng.uniform(0, 750, n // 2), rng.exponential(3.0, n // 2)]))
    x[:6] = [0.0, -0.0, -707.99, -709.0, -np.inf, np.nan]
    y = np.empty(n)
    call("hh_exp_nonpos", x, y)
    assert y[0] == 1.0 and y[1] == 1.0 and y[3] == 0.0 and y[4] == 0.0 and np.isnan(y[5])
    ok = x > -708.0
    np.testing.assert_allclose(y[ok], np.exp(x[ok]), rtol=3 * ULP)
    assert np.all(y[~ok & ~np.isnan(x)] == 0.0)       # below the normal range: flushed (contributes < 2^-1021)
    mp.mp.dps = 40
    sub = rng.choice(np.where(ok)[0], 1500, replace=False)
    err = max(abs(mp.mpf(float(y[i])) / mp.exp(mp.mpf(float(x[i]))) - 1) for i in sub)
    assert float(err) < 2.0 * ULP


def test_log_pos_relative_accuracy_including_uniforms_next_to_one():
    rng = np.random.default_rng(1)
    n = 60_000
    v = np.concatenate([rng.integers(1, 2 ** 53, n // 3).astype(np.float64),
                        2.0 ** 53 - (2.0 * rng.integers(1, 10 ** 6, n // 3) - 1.0),
                        np.ldexp(1.0 + rng.random(n // 3), rng.integers(0, 53, n // 3))])
    kb = np.full(v.size, -53, dtype=np.int32)
    y = np.empty(v.size)
    call("hh_log_pos", v, kb, y)
    mp.mp.dps = 40
    sub = rng.choice(v.size, 3000, replace=False)
    l2 = mp.log(2)
    err = 0
    for i in sub:
        want = mp.log(mp.mpf(float(v[i]))) - 53 * l2
        if want != 0:
            err = max(err, abs(mp.mpf(float(y[i])) / want - 1))
    assert float(err) < 3 * ULP
    # plain log (kbias = 0) over the whole normal range
    x = np.ldexp(1.0 + rng.random(20000), rng.integers(-1000, 1000, 20000))
    y = np.empty(x.size)
    call("hh_log_pos", x, np.zeros(x.size, dtype=np.int32), y)
    np.testing.assert_allclose(y, np.log(x), rtol=4 * ULP, atol=1e-300)


def test_sqrt_div_sincos():
    rng = np.random.default_rng(2)
    n = 100_000
    t = np.concatenate([rng.uniform(0, 74, n // 2), 2.0 ** rng.uniform(-52, 7, n // 2)])
    y = np.empty(n)
    call("hh_sqrt_pos", t, y)
    np.testing.assert_allclose(y, np.sqrt(t), rtol=1.01 * ULP)
    a, b = rng.random(n), rng.uniform(1.0, 1e9, n)
    a[:3] = [0.0, 1.0, 2.0 ** -1000]
    q = np.empty(n)
    call("hh_div_pos", a, b, q)
    assert np.mean(q == a / b) > 0.9999 and np.max(np.abs(q - a / b) / np.maximum(a / b, 1e-300)) <= 2.3 * ULP
    f = np.concatenate([rng.random(n - 2), [0.0, 1.0]])
    s, c = np.empty(n), np.empty(n)
    call("hh_sincos_octant", f, s, c)
    np.testing.assert_allclose(s, np.sin(np.pi / 4 * f), rtol=0, atol=2.5 * ULP)
    np.testing.assert_allclose(c, np.cos(np.pi / 4 * f), rtol=0, atol=2.5 * ULP)
    assert s[-2] == 0.0 and c[-2] == 1.0


def test_box_muller_is_standard_normal_pair_uniform_on_the_circle():
    rng = np.random.default_rng(3)
    n = 400_000
    w1 = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    w2 = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    w1[:2] = [0, 2 ** 64 - 1]            # u1 = 2^-53 and 1 - 2^-53: both finite
    z0, z1 = np.empty(n), np.empty(n)
    call("hh_box_muller", w1, w2, z0, z1)
    assert np.all(np.isfinite(z0)) and np.all(np.isfinite(z1))
    assert abs(np.hypot(z0[0], z1[0]) - np.sqrt(2 * 53 * np.log(2))) < 1e-12 and np.hypot(z0[1], z1[1]) < 2e-8
    assert sst.kstest(z0, "norm").pvalue > 1e-3 and sst.kstest(z1, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z0, z1)[0, 1]) < 0.006
    assert sst.kstest((np.arctan2(z1, z0) + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3      # all eight octants
    assert sst.kstest(z0 ** 2 + z1 ** 2, sst.expon(scale=2).cdf).pvalue > 1e-3
    # exact reconstruction of radius and angle from the bits
    k = (w1 >> np.uint64(12)).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log((2.0 * k + 1.0) * 2.0 ** -53))
    np.testing.assert_allclose(np.hypot(z0, z1), rad, rtol=1e-14)


@pytest.mark.gpu
def test_device_normals_equal_host_instantiation(ws):
    """MUFU.RCP64H / MUFU.RSQ64H seeds + Newton steps on the device vs libm seeds on the host: same normals."""
    n = 300_000
    st = ws.SMCState(n, seed=123, device=0)
    step = ws.Sample("z", "MvNormal", ([0.0, 0.0], np.eye(2)))
    ws.run(ws.Sequence(step), st)
    hs = HostState(n, seed=123)
    step.apply(hs)
    np.testing.assert_allclose(st["z"], hs.store.getcol("z"), rtol=1e-14, atol=1e-300)
    assert np.mean(st["z"] == hs.store.getcol("z")) > 0.99
    sp = ws.SMCState(n, seed=5, device=0)
    ws.run(ws.Sequence(ws.Sample("e", "Exponential", (1.0,))), sp)
    hp = HostState(n, seed=5)
    ws.Sample("e", "Exponential", (1.0,)).apply(hp)
    np.testing.assert_allclose(sp["e"], hp.store.getcol("e"), rtol=1e-14)
