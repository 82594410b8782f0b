"""GPU parity for SURVEY §8(f)3 (wider op set): expression kernels, Bool-valued columns, ternaries, user
`WeightedKernel`s and helper functions — the reference's examples/fire_alarm.jl and
examples/damped_oscillator.jl run end to end on the device against the oracle under replayed streams."""
import math

import numpy as np
import pytest
import scipy.stats as sst

from oracle import ref
from test_expr_kernels import CASES, FIRE_ALARM, OSCILLATOR, REJ_CASES, SAMPLERS, _gamma_replay, half_normal, oscillator

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,args,dist", CASES, ids=[c[0] for c in CASES])
def test_device_logpdfs_match_scipy(ws, name, args, dist):
    n = 5000
    x = dist.rvs(size=n, random_state=np.random.default_rng(3))
    st = ws.SMCState(n, device=0)
    st.store.setcol("x", x)
    ws.Observe(ws.col("x"), name, args).apply(st)
    np.testing.assert_allclose(st.weights, dist.logpdf(x), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name,args,dist", SAMPLERS, ids=[c[0] for c in SAMPLERS])
def test_device_samplers(ws, name, args, dist):
    n = 200_000
    rng = np.random.default_rng(9)
    streams = dict(normals=rng.standard_normal(2 * n), uniforms=rng.random(2 * n),
                   exponentials=rng.standard_exponential(2 * n))
    step = ws.Sample("x", name, args)
    st = ws.SMCState(n, device=0)
    st.set_replay(**streams)
    ws.run(ws.Sequence(step), st)
    ost = ref.OracleState(n, ref.Streams(**streams))
    ost.expr_factory = ws.col
    ref.run(ws.Sequence(step), ost)
    np.testing.assert_allclose(st["x"], ost.cols["x"], rtol=1e-9, atol=1e-15)
    sp = ws.SMCState(n, seed=5, device=0)          # Philox
    ws.run(ws.Sequence(step), sp)
    x = sp["x"]
    assert sst.kstest(x, dist.cdf).pvalue > 1e-4
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(ws.score_logpdf(sp, ["x"], 1), dist.logpdf(x), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name,args,dist", REJ_CASES, ids=[c[0] for c in REJ_CASES])
def test_device_rejection_samplers(ws, name, args, dist):
    """Gamma / Beta / TDist / Chisq / InverseGamma on the device (Marsaglia-Tsang inside the fused pass): replayed
    accepted variates against the closed form and the oracle, Philox draws against scipy by KS over five seeds, the
    device draws equal the host instantiation of the same routine, and the `~` / `=>` symmetry through the score tape."""
    n = 100_000
    rng = np.random.default_rng(21)
    normals, variates, want = _gamma_replay(name, args, n, rng)
    step = ws.Sample("x", name, args)
    st = ws.SMCState(n, device=0)
    st.set_replay(normals=normals if len(normals) else None, variates=variates)
    ws.run(ws.Sequence(step), st)
    np.testing.assert_allclose(st["x"], want, rtol=1e-12)
    ost = ref.OracleState(n, ref.Streams(normals=normals, variates=variates))
    ost.expr_factory = ws.col
    ref.run(ws.Sequence(step), ost)
    np.testing.assert_allclose(st["x"], ost.cols["x"], rtol=1e-12)
    for seed in range(1, 6):
        sp = ws.SMCState(n, seed=seed, device=0)
        ws.run(ws.Sequence(step), sp)
        x = sp["x"]
        assert np.all(np.isfinite(x))
        assert sst.kstest(x, dist.cdf).pvalue > 1e-3 / 5, (name, seed)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(ws.score_logpdf(sp, ["x"], 1), dist.logpdf(x), rtol=1e-9, atol=1e-9)
    so = ws.SMCState(n, device=0)
    so.store.setcol("x", x)
    ws.Observe(ws.col("x"), name, args).apply(so)
    np.testing.assert_allclose(so.weights, ws.score_logpdf(sp, ["x"], 1), rtol=1e-12, atol=1e-12)
    from hostlib import HostState
    hp = HostState(2000, seed=5)
    step.apply(hp)
    np.testing.assert_allclose(x[:2000], hp.store.getcol("x"), rtol=1e-9)      # same counters, same routine (libm vs libdevice)


def test_device_poisson_sampler_in_a_model(ws):
    n = 200_000
    src = '''
    @model function counts(ys)
        lam ~ Gamma(2.0, 3.0)
        for y in ys
            y => Poisson(lam)
        end
        k ~ Poisson(lam)
    end
    '''
    ys = [4.0, 7.0, 5.0, 6.0]
    st = ws.SMCState(n, seed=8, ess_perc_min=0.5, device=0)
    ws.run(ws.model(src)(ys), st)
    # conjugate posterior: Gamma(2 + sum y, scale 1 / (1/3 + 4)); posterior predictive mean = posterior mean of lam
    a_post, scale_post = 2.0 + sum(ys), 1.0 / (1.0 / 3.0 + len(ys))
    assert abs(ws.E(lambda lam: lam, st) - a_post * scale_post) < 0.05
    assert abs(ws.E(lambda k: k, st) - a_post * scale_post) < 0.08
    k = st["k"]
    assert np.all(k >= 0) and np.all(k == np.floor(k))


def test_fire_alarm_end_to_end(ws):
    """examples/fire_alarm.jl:27-37 with the Resample after the observation firing (ESS% is about 3)."""
    n = 300_000
    rng = np.random.default_rng(0)
    u = rng.random(4 * n)
    root = ws.model(FIRE_ALARM)()
    st = ws.SMCState(n, device=0)
    st.set_replay(uniforms=u)
    ws.run(root, st)
    ost = ref.OracleState(n, ref.Streams(uniforms=u))
    ost.expr_factory = ws.col
    ref.run(root, ost)
    assert st.resampled and ost.resampled
    bad = 0
    for name in ("fire", "smoke", "lever"):
        bad = max(bad, int(np.sum(st[name] != ost.cols[name])))
    assert bad <= 2, bad   # ancestors are bit-exact except near-boundary uniforms
    assert abs(ws.log_evidence(st) - ref.log_evidence(ost)) <= 1e-9 * abs(ref.log_evidence(ost))
    # exact posterior P(fire | alarm) = 0.2469 (enumeration, tests/test_expr_kernels.py); @E with Bool logic
    assert abs(ws.E(lambda fire: fire, st) - 0.24686537568372155) < 0.01
    p_fire_no_smoke = ws.E(lambda fire, smoke: fire & ~smoke, st)   # examples/fire_alarm.jl:21
    assert abs(p_fire_no_smoke - float(np.mean(ost.cols["fire"] * (1 - ost.cols["smoke"])))) < 1e-9


def test_damped_oscillator_with_bounded_multi_target_moves(ws):
    """examples/damped_oscillator.jl:32-52: HalfNormal user kernel, Uniform prior, helper function in the
    likelihood, a 4-target autoRW with (0, Inf) bounds and a 1-target autoRW with (-pi, pi), both
    diversity-gated; replayed against the oracle, then statistically with Philox."""
    rng = np.random.default_rng(1)
    t_obs = np.linspace(0, 8, 12)
    y_obs = 3.0 * np.exp(-0.3 * t_obs) * np.cos(2.5 * t_obs + 0.5) + rng.normal(size=12)
    m = ws.model(OSCILLATOR, scope={"oscillator": oscillator})
    root = m(t_obs, y_obs, kernels={"HalfNormal": half_normal()})
    n = 3000
    T = len(t_obs)
    streams = dict(normals=rng.standard_normal(n * (4 + 5 * T)), uniforms=rng.random(n * (1 + 3 * T)))
    st = ws.SMCState(n, device=0)
    st.set_replay(**streams)
    ws.run(root, st)
    ost = ref.OracleState(n, ref.Streams(**streams))
    ost.expr_factory = ws.col
    ref.run(root, ost)
    for name in ("A", "ω", "γ", "ϕ", "σ"):
        d = np.abs(st[name] - ost.cols[name]) > 1e-8 * (1 + np.abs(ost.cols[name]))
        assert d.mean() < 0.01, (name, d.mean())
    assert abs(ws.log_evidence(st) - ref.log_evidence(ost)) < 1e-6
    # Philox, more particles: the posterior concentrates around the truth
    sp = ws.SMCState(100_000, seed=11, device=0)
    t2 = np.linspace(0, 8, 60)
    y2 = 3.0 * np.exp(-0.3 * t2) * np.cos(2.5 * t2 + 0.5) + np.random.default_rng(2).normal(size=60)
    ws.run(m(t2, y2, kernels={"HalfNormal": half_normal()}), sp)
    w = ws.exp_norm(sp)
    est = {k: float(np.sum(w * sp[k])) for k in ("A", "ω", "γ", "σ")}
    assert abs(est["ω"] - 2.5) < 0.2 and abs(est["A"] - 3.0) < 1.0 and abs(est["γ"] - 0.3) < 0.2 and abs(est["σ"] - 1.0) < 0.3


def test_conditionals_on_device(ws):
    n = 4096
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=n), rng.normal(size=n)
    m = ws.model('''
    @model function f()
        c .= a > b ? a : b
        d .= (a <= b) && (a > 0.0) ? 1.0 : -1.0
        h .= log1p(abs(a)) + expm1(b) - floor(a) + tan(b) + lgamma(abs(a) + 0.5) + tanh(a) * atan(b)
    end
    ''', particle_vars=("a", "b"))
    st = ws.SMCState(n, device=0)
    st.store.setcol("a", a)
    st.store.setcol("b", b)
    ws.run(m(), st)
    from scipy.special import gammaln
    np.testing.assert_array_equal(st["c"], np.maximum(a, b))
    np.testing.assert_array_equal(st["d"], np.where((a <= b) & (a > 0), 1.0, -1.0))
    np.testing.assert_allclose(st["h"], np.log1p(np.abs(a)) + np.expm1(b) - np.floor(a) + np.tan(b) + gammaln(np.abs(a) + 0.5)
                               + np.tanh(a) * np.arctan(b), rtol=1e-11, atol=1e-12)
