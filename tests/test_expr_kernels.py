"""SURVEY §8(f)3 — the wider op set: distributions written as device expressions (`WeightedKernel` with traced
sampler / weighter / logpdf, types.jl:226-230), per-particle ternary / `&&` / `||` / comparisons
(rewrites.jl:193-212), math in arguments.  Run on the CPU through the product's own lowering + micro-op
interpreter (tests/host/harness.cpp) and checked against the oracle and against scipy.stats (the formulas are
Distributions.jl's, un-vendored: scipy pins them)."""
import math

import numpy as np
import pytest
import scipy.stats as sst

import wsb200 as ws
from hostlib import HostState
from oracle import ref

FIRE_ALARM = '''
@model function fire_alarm()
    fire ~ Bernoulli(0.01)
    smoke ~ Bernoulli(fire ? 0.9 : 0.01)
    lever ~ Bernoulli(fire ? 0.7 : 0.01)
    true => Bernoulli(smoke || lever ? 0.98 : 0.01)
end
'''  # examples/fire_alarm.jl:27-32, verbatim

OSCILLATOR = '''
@model function damped_oscillator(t_obs, y_obs)
    A ~ HalfNormal(5)
    ω ~ HalfNormal(5)
    γ ~ HalfNormal(1)
    ϕ ~ Uniform(-π, π)
    σ ~ HalfNormal(1)

    for (t, y) in zip(t_obs, y_obs)
        y => Normal(oscillator(t, A, ω, γ, ϕ), σ)
        (A, ω, γ, σ) << autoRW(1e-3, (0.0, Inf), diversity=0.9)
        ϕ << autoRW(1e-3, (-π, π), diversity=0.9)
    end

end
'''  # examples/damped_oscillator.jl:32-46, verbatim


def oscillator(t, A, w, g, p):  # examples/damped_oscillator.jl:11
    return A * ws.exp(-g * t) * ws.cos(w * t + p)


def half_normal():
    """examples/damped_oscillator.jl:26-30: Truncated(Normal(0, σ), 0, Inf) as a device-expression kernel."""
    return ws.WeightedKernel(
        lambda s: abs(s * ws.randn()), None,
        lambda s, x: ws.where(x >= 0.0, math.log(2.0) - 0.5 * math.log(2 * math.pi) - ws.log(s) - 0.5 * (x / s) ** 2,
                              float("-inf")), "HalfNormal")


def strip(t):
    k = type(t).__name__
    if k == "Sequence":
        return ws.Sequence(*[strip(s) for s in t.steps if type(s).__name__ not in ("Resample", "Move")])
    if k == "Loop":
        return ws.Loop(t.collfn, lambda x, f=t.bodyfn: strip(f(x)))
    return t


# name, args (as particle-independent parameters), scipy frozen distribution, support sample
CASES = [
    ("Uniform", (-1.5, 2.0), sst.uniform(-1.5, 3.5)),
    ("LogNormal", (0.3, 0.7), sst.lognorm(0.7, scale=math.exp(0.3))),
    ("Laplace", (0.5, 1.3), sst.laplace(0.5, 1.3)),
    ("Cauchy", (-0.2, 0.8), sst.cauchy(-0.2, 0.8)),
    ("Logistic", (1.0, 0.6), sst.logistic(1.0, 0.6)),
    ("Gumbel", (0.4, 1.7), sst.gumbel_r(0.4, 1.7)),
    ("Rayleigh", (1.4,), sst.rayleigh(scale=1.4)),
    ("Weibull", (1.8, 2.2), sst.weibull_min(1.8, scale=2.2)),
    ("Pareto", (2.5, 1.2), sst.pareto(2.5, scale=1.2)),
    ("Gamma", (2.3, 1.6), sst.gamma(2.3, scale=1.6)),
    ("Beta", (2.0, 3.5), sst.beta(2.0, 3.5)),
    ("TDist", (4.5,), sst.t(4.5)),
    ("Chisq", (3.7,), sst.chi2(3.7)),
    ("InverseGamma", (3.2, 1.9), sst.invgamma(3.2, scale=1.9)),
]


@pytest.mark.parametrize("name,args,dist", CASES, ids=[c[0] for c in CASES])
def test_logpdf_formulas_match_scipy(name, args, dist):
    n = 2000
    rng = np.random.default_rng(3)
    x = dist.rvs(size=n, random_state=rng)
    x[:5] = [-3.0, 0.0, 1e-9, 0.5, 40.0]  # edge / out-of-support values flow as -Inf
    hs = HostState(n)
    hs.store.setcol("x", x)
    ws.Observe(ws.col("x"), name, args).apply(hs)
    with np.errstate(all="ignore"):
        want = dist.logpdf(x)
    got = hs.store.logw()
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-11, atol=1e-11)
    if name in ("Rayleigh", "Weibull", "Gamma"):
        fin[x == 0.0] = True  # density 0 at the boundary either way: -Inf on both sides or finite vs -Inf tolerated
    assert np.all(np.isneginf(got[~fin]) | np.isnan(got[~fin]) == True)  # noqa: E712


def test_poisson_and_bernoulli_logpdf():
    n = 500
    rng = np.random.default_rng(5)
    k = rng.poisson(3.0, n).astype(float)
    hs = HostState(n)
    hs.store.setcol("k", k)
    hs.store.setcol("b", (rng.random(n) < 0.3).astype(float))
    ws.Observe(ws.col("k"), "Poisson", (3.2,)).apply(hs)
    np.testing.assert_allclose(hs.store.logw(), sst.poisson(3.2).logpmf(k), rtol=1e-12)
    hs2 = HostState(n)
    hs2.store.setcol("b", hs.store.getcol("b"))
    ws.Observe(ws.col("b"), "Bernoulli", (0.3,)).apply(hs2)
    np.testing.assert_allclose(hs2.store.logw(), sst.bernoulli(0.3).logpmf(hs.store.getcol("b")), rtol=1e-13)


REJECTION = ("Gamma", "Beta", "TDist", "Chisq", "InverseGamma")      # built on the device's Gamma variate
SAMPLERS = [c for c in CASES if c[0] not in REJECTION]


@pytest.mark.parametrize("name,args,dist", SAMPLERS, ids=[c[0] for c in SAMPLERS])
def test_samplers_replay_equals_oracle_and_philox_passes_ks(name, args, dist):
    n = 20000
    rng = np.random.default_rng(9)
    streams = dict(normals=rng.standard_normal(2 * n), uniforms=rng.random(2 * n),
                   exponentials=rng.standard_exponential(2 * n))
    step = ws.Sample("x", name, args)
    hs = HostState(n)
    hs.store.set_replay(**streams)
    step.apply(hs)
    ost = ref.OracleState(n, ref.Streams(**streams), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(ws.Sequence(step), ost)
    np.testing.assert_allclose(hs.store.getcol("x"), ost.cols["x"], rtol=1e-10, atol=1e-15)  # tan near its poles is ill-conditioned
    assert sst.kstest(hs.store.getcol("x"), dist.cdf).pvalue > 1e-4
    # production RNG (Philox counters): distribution only
    hp = HostState(n, seed=77)
    step.apply(hp)
    assert sst.kstest(hp.store.getcol("x"), dist.cdf).pvalue > 1e-4
    # score!: the tape entry of the Sample is logpdf(args..., x)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(hp.store.score(1), dist.logpdf(hp.store.getcol("x")), rtol=1e-10, atol=1e-10)


def _gamma_replay(name, args, n, rng):
    """(normals, accepted variates) that reproduce `name(args...)` under replay, and the value they must give:
    the device takes the standard Gamma(a_i) variate from the replay stream where it would run Marsaglia-Tsang."""
    if name == "Gamma":
        g = rng.gamma(args[0], size=n)
        return (), g, args[1] * g
    if name == "Chisq":
        g = rng.gamma(args[0] / 2.0, size=n)
        return (), g, 2.0 * g
    if name == "InverseGamma":
        g = rng.gamma(args[0], size=n)
        return (), g, args[1] / g
    if name == "Beta":            # 1 / (1 + Y / X): Y ~ Gamma(b) is drawn first, then X ~ Gamma(a)
        y, x = rng.gamma(args[1], size=n), rng.gamma(args[0], size=n)
        return (), np.concatenate([y, x]), 1.0 / (1.0 + y / x)
    z, g = rng.standard_normal(n), rng.gamma(args[0] / 2.0, size=n)       # TDist: Z / sqrt(Chisq(v) / v)
    return z, g, z / np.sqrt(2.0 * g / args[0])


REJ_CASES = [c for c in CASES if c[0] in REJECTION]


@pytest.mark.parametrize("name,args,dist", REJ_CASES, ids=[c[0] for c in REJ_CASES])
def test_rejection_samplers_replay_philox_ks_and_symmetry(name, args, dist):
    """src/default_kernels.jl:83-102 entries whose sampler needs a rejection loop: replayed accepted
    variates reproduce the construction exactly (device lowering == oracle == closed form); under Philox the draws
    pass a KS test against scipy for five seeds; and `x ~ D(...)` followed by score! gives logpdf(D, x) — the same
    number `x => D(...)` adds to the weights (the `~` / `=>` symmetry of transformers.jl:172-182,228-235)."""
    n = 20000
    rng = np.random.default_rng(21)
    normals, variates, want = _gamma_replay(name, args, n, rng)
    step = ws.Sample("x", name, args)
    hs = HostState(n)
    hs.store.set_replay(normals=normals, variates=variates)
    step.apply(hs)
    np.testing.assert_allclose(hs.store.getcol("x"), want, rtol=1e-12)
    ost = ref.OracleState(n, ref.Streams(normals=normals, variates=variates), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(ws.Sequence(step), ost)
    np.testing.assert_allclose(hs.store.getcol("x"), ost.cols["x"], rtol=1e-12)
    for seed in range(1, 6):
        hp = HostState(n, seed=seed)
        step.apply(hp)
        x = hp.store.getcol("x")
        assert np.all(np.isfinite(x))
        assert sst.kstest(x, dist.cdf).pvalue > 1e-3 / 5, (name, seed)
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(hp.store.score(1), dist.logpdf(x), rtol=1e-9, atol=1e-9)
    ho = HostState(n)
    ho.store.setcol("x", x)
    ws.Observe(ws.col("x"), name, args).apply(ho)
    np.testing.assert_allclose(ho.store.logw(), hp.store.score(1), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("a", [0.05, 0.4, 1.0, 2.5, 40.0, 3000.0])
def test_gamma_variate_all_shapes(a):
    """Marsaglia-Tsang for a >= 1 and the U^(1/a) boost below 1, through the host instantiation of csrc/ws_math.cuh"""
    import ctypes as C
    from hostlib import lib
    n = 40000
    for seed in (1, 2, 3):
        out = np.empty(n)
        lib().hh_rand_gamma(a, 0, 7, seed, out.ctypes.data_as(C.c_void_p), n)
        assert np.all(out > 0) and np.all(np.isfinite(out))
        assert sst.kstest(out, sst.gamma(a).cdf).pvalue > 1e-3 / 3, (a, seed)
    out = np.empty(4)
    lib().hh_rand_gamma(-1.0, 0, 7, 1, out.ctypes.data_as(C.c_void_p), 4)
    assert np.all(np.isnan(out))


@pytest.mark.parametrize("lam", [0.0, 0.3, 4.0, 9.99, 10.0, 37.5, 1200.0, 3e6])
def test_poisson_variate_inversion_and_ptrs(lam):
    """Poisson: sequential inversion below 10, PTRS (Hormann 1993) above; chi-square against the exact pmf"""
    import ctypes as C
    from hostlib import lib
    n = 60000
    out = np.empty(n)
    lib().hh_rand_poisson(lam, 0, 11, 5, out.ctypes.data_as(C.c_void_p), n)
    assert np.all(out >= 0) and np.all(out == np.floor(out))
    if lam == 0.0:
        assert np.all(out == 0)
        return
    assert abs(out.mean() - lam) < 5 * math.sqrt(lam / n) and abs(out.var() / lam - 1.0) < 0.05
    lo, hi = int(max(0, lam - 4 * math.sqrt(lam))), int(lam + 4 * math.sqrt(lam)) + 1
    edges = np.unique(np.linspace(lo, hi, 25).astype(int))
    obs = np.histogram(out, bins=np.concatenate([[-0.5], edges + 0.5, [np.inf]]))[0]
    cdf = sst.poisson(lam).cdf(edges)
    exp = n * np.diff(np.concatenate([[0.0], cdf, [1.0]]))
    keep = exp > 5
    chi2 = float(np.sum((obs[keep] - exp[keep]) ** 2 / exp[keep]))
    assert chi2 < sst.chi2(int(keep.sum()) - 1).ppf(1 - 1e-4), (lam, chi2)


def test_poisson_kernel_sample_score_and_replay():
    n = 20000
    step = ws.Sample("k", "Poisson", (6.5,))
    hp = HostState(n, seed=3)
    step.apply(hp)
    k = hp.store.getcol("k")
    np.testing.assert_allclose(hp.store.score(1), sst.poisson(6.5).logpmf(k), rtol=1e-10)
    assert abs(k.mean() - 6.5) < 0.1
    kr = np.random.default_rng(2).poisson(6.5, n).astype(float)
    hs = HostState(n)
    hs.store.set_replay(variates=kr)
    step.apply(hs)
    np.testing.assert_array_equal(hs.store.getcol("k"), kr)
    # a per-particle rate: lam_i = exp(z_i)
    hv = HostState(n, seed=4)
    hv.store.setcol("lam", np.linspace(0.1, 60.0, n))
    ws.Sample("k", "Poisson", (ws.col("lam"),)).apply(hv)
    kk, lam = hv.store.getcol("k"), np.linspace(0.1, 60.0, n)
    zscore = (kk - lam) / np.sqrt(lam)
    assert abs(zscore.mean()) < 0.03 and abs(zscore.var() - 1.0) < 0.05


def test_fire_alarm_bayes_net_matches_oracle_and_exact_posterior():
    n = 400_000
    rng = np.random.default_rng(0)
    u = rng.random(3 * n)
    root = strip(ws.model(FIRE_ALARM)())
    hs = HostState(n)
    hs.store.set_replay(uniforms=u)
    root.apply(hs)
    ost = ref.OracleState(n, ref.Streams(uniforms=u), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(root, ost)
    for name in ("fire", "smoke", "lever"):
        got = hs.store.getcol(name)
        np.testing.assert_array_equal(got, ost.cols[name])
        assert set(np.unique(got)) <= {0.0, 1.0}
    np.testing.assert_allclose(hs.store.logw(), ost.weights, rtol=1e-13)
    # exact posterior by enumeration
    num = den = 0.0
    for f in (0, 1):
        for s in (0, 1):
            for lv in (0, 1):
                ps, pl = (0.9 if f else 0.01), (0.7 if f else 0.01)
                p = (0.01 if f else 0.99) * (ps if s else 1 - ps) * (pl if lv else 1 - pl) * (0.98 if (s or lv) else 0.01)
                den += p
                num += p * f
    w = np.exp(hs.store.logw())
    w /= w.sum()
    assert abs(float((w * hs.store.getcol("fire")).sum()) - num / den) < 0.015
    assert abs(math.log(den) - (np.log(np.mean(np.exp(hs.store.logw()))))) < 0.03   # evidence P(alarm)


def test_damped_oscillator_lowers_with_user_kernel_and_helper_function():
    n = 1000
    rng = np.random.default_rng(1)
    t_obs = np.linspace(0, 8, 6)
    y_obs = 3.0 * np.exp(-0.3 * t_obs) * np.cos(2.5 * t_obs + 0.5) + rng.normal(size=6)
    m = ws.model(OSCILLATOR, scope={"oscillator": oscillator})
    full = m(t_obs, y_obs, kernels={"HalfNormal": half_normal()})
    moves = [s for s in full.steps[-1].bodyfn((t_obs[0], y_obs[0])).steps if type(s).__name__ == "Move"]
    assert [mv.targets for mv in moves] == [["A", "ω", "γ", "σ"], ["ϕ"]] and moves[0].diversity_threshold == 0.9
    root = strip(full)
    streams = dict(normals=rng.standard_normal(4 * n), uniforms=rng.random(n))
    hs = HostState(n)
    hs.store.set_replay(**streams)
    root.apply(hs)
    ost = ref.OracleState(n, ref.Streams(**streams), ess_perc_min=0.0)
    ost.expr_factory = ws.col
    ref.run(root, ost)
    for name in ("A", "ω", "γ", "ϕ", "σ"):
        np.testing.assert_allclose(hs.store.getcol(name), ost.cols[name], rtol=1e-14)
    assert np.all(hs.store.getcol("A") >= 0) and np.all(np.abs(hs.store.getcol("ϕ")) <= math.pi)
    np.testing.assert_allclose(hs.store.logw(), ost.weights, rtol=1e-11, atol=1e-11)
    # independent restatement of the weights
    A, w_, g, p, s = (hs.store.getcol(k) for k in ("A", "ω", "γ", "ϕ", "σ"))
    want = sum(sst.norm(A * np.exp(-g * t) * np.cos(w_ * t + p), s).logpdf(y) for t, y in zip(t_obs, y_obs))
    np.testing.assert_allclose(hs.store.logw(), want, rtol=1e-10, atol=1e-10)
    # score tape = priors + likelihood
    prior = (sst.halfnorm(scale=5).logpdf(A) + sst.halfnorm(scale=5).logpdf(w_) + sst.halfnorm(scale=1).logpdf(g)
             + sst.uniform(-math.pi, 2 * math.pi).logpdf(p) + sst.halfnorm(scale=1).logpdf(s))
    np.testing.assert_allclose(hs.store.score(hs.store.tape_len()), prior + want, rtol=1e-10, atol=1e-10)


def test_conditionals_comparisons_and_math_in_assignments():
    n = 512
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=n), rng.normal(size=n)
    a[:3] = b[:3]
    m = ws.model('''
    @model function f()
        c .= a > b ? a : b
        d .= (a <= b) && (a > 0.0) ? 1.0 : -1.0
        e .= ifelse(a == b, 0.0, min(a, b)) + max(a, 0.5)
        g .= !(a < 0.0) || b != a ? tanh(a) : atan(b)
        h .= log1p(abs(a)) + expm1(b) - floor(a) + tan(b)
    end
    ''', particle_vars=("a", "b"))
    hs = HostState(n)
    hs.store.setcol("a", a)
    hs.store.setcol("b", b)
    m().apply(hs)
    np.testing.assert_array_equal(hs.store.getcol("c"), np.maximum(a, b))
    np.testing.assert_array_equal(hs.store.getcol("d"), np.where((a <= b) & (a > 0), 1.0, -1.0))
    np.testing.assert_allclose(hs.store.getcol("e"), np.where(a == b, 0.0, np.minimum(a, b)) + np.maximum(a, 0.5), rtol=1e-15)
    np.testing.assert_allclose(hs.store.getcol("g"), np.where(~(a < 0) | (b != a), np.tanh(a), np.arctan(b)), rtol=1e-14)
    np.testing.assert_allclose(hs.store.getcol("h"), np.log1p(np.abs(a)) + np.expm1(b) - np.floor(a) + np.tan(b), rtol=1e-12)


def test_untraceable_closures_are_rejected_not_run_on_the_host():
    hs = HostState(8)
    k = ws.WeightedKernel(lambda m: math.exp(m) + ws.randn(), None, lambda m, x: -x * x)
    with pytest.raises(ws.UnsupportedModelError):
        ws.Sample("x", k, (ws.col("x") + 1.0,)).apply(hs)   # math.exp on a particle expression
    with pytest.raises(ws.UnsupportedModelError):
        ws.Sample("y", "Wishart", (2.0, 1.0)).apply(hs)      # in the reference's table, outside the device-op set
    with pytest.raises(RuntimeError):
        ws.Observe(ws.col("x") + ws.randn(), "Normal", (0.0, 1.0)).apply(hs)  # variates outside a sampler


# ---- Julia Base scalar functions without a micro-op of their own (compositions in expr.py) -----------------------------
def _julia_round(x):
    return np.round(x)            # NumPy rounds half to even, as Julia's RoundNearest does


_BASE_FUNCTIONS = [
    # (model expression in x [, y], NumPy restatement of JULIA's function, rtol)
    ("sign(x)", lambda x, y: np.where(x == 0, x, np.sign(x)), 0.0),          # Julia: sign(-0.0) = -0.0
    ("ceil(x)", lambda x, y: np.ceil(x), 0.0),
    ("trunc(x)", lambda x, y: np.trunc(x), 0.0),
    ("round(x)", lambda x, y: _julia_round(x), 0.0),
    ("floor(Int, x) + round(Int, x)", lambda x, y: np.floor(x) + _julia_round(x), 0.0),
    ("clamp(x, -1.5, 2.0)", lambda x, y: np.where(x > 2.0, 2.0, np.where(x < -1.5, -1.5, x)), 0.0),
    ("isnan(x) ? 1.0 : 0.0", lambda x, y: np.isnan(x).astype(float), 0.0),
    ("isinf(x) ? 1.0 : 0.0", lambda x, y: np.isinf(x).astype(float), 0.0),
    ("isfinite(x) ? 1.0 : 0.0", lambda x, y: np.isfinite(x).astype(float), 0.0),
    ("log2(abs(x))", lambda x, y: np.log2(np.abs(x)), 4e-16),
    ("log10(abs(x))", lambda x, y: np.log10(np.abs(x)), 4e-16),
    ("log(3, abs(x))", lambda x, y: np.log(np.abs(x)) / np.log(3.0), 4e-16),
    ("exp2(x / 8)", lambda x, y: np.exp2(x / 8), 1e-14),
    ("exp10(x / 8)", lambda x, y: 10.0 ** (x / 8), 1e-14),
    ("sinh(x / 4)", lambda x, y: np.sinh(x / 4), 1e-14),
    ("cosh(x / 4)", lambda x, y: np.cosh(x / 4), 1e-14),
    ("asin(clamp(x / 8, -1, 1))", lambda x, y: np.arcsin(np.clip(x / 8, -1, 1)), 1e-14),
    ("acos(clamp(x / 8, -1, 1))", lambda x, y: np.arccos(np.clip(x / 8, -1, 1)), 1e-14),
    ("atan(y, x)", lambda x, y: np.arctan2(y, x), 1e-14),
    ("hypot(x, y)", lambda x, y: np.hypot(x, y), 1e-15),
    ("cbrt(x)", lambda x, y: np.cbrt(x), 5e-14),
    ("inv(x)", lambda x, y: 1.0 / x, 0.0),
    ("x ÷ 3", lambda x, y: np.trunc(x / 3), 0.0),
    ("x % 3", lambda x, y: x - 3 * np.trunc(x / 3), 0.0),
    ("mod(x, 3)", lambda x, y: x - 3 * np.floor(x / 3), 0.0),
    ("rem(x, y)", lambda x, y: x - y * np.trunc(x / y), 0.0),
    ("fld(x, 3) + cld(x, 3) + div(x, 3)", lambda x, y: np.floor(x / 3) + np.ceil(x / 3) + np.trunc(x / 3), 0.0),
    ("max(x, y, 0.5) - min(x, y, -0.5, 0.25)", lambda x, y: np.maximum(np.maximum(x, y), 0.5) - np.minimum(np.minimum(x, y), -0.5), 0.0),
    ("float(x) + Float64(y) + one(x) + zero(y)", lambda x, y: x + y + 1.0, 0.0),
    ("√(abs(x)) + √2", lambda x, y: np.sqrt(np.abs(x)) + math.sqrt(2.0), 0.0),
    ("iszero(x) ? 1.0 : (isone(x) ? 2.0 : 0.0)", lambda x, y: np.where(x == 0, 1.0, np.where(x == 1, 2.0, 0.0)), 0.0),
]


@pytest.mark.parametrize("src,want,rtol", _BASE_FUNCTIONS, ids=[c[0] for c in _BASE_FUNCTIONS])
def test_julia_base_functions_on_particle_values(src, want, rtol):
    """`vectorize` broadcasts any Julia function over particle columns (rewrites.jl:150-163); the device-op set is
    closed, so the common scalar Base functions are compositions of its micro-ops.  Run through the product's lowering
    and the micro-op interpreter (CPU harness) on ordinary values, ties, signed zeros, infinities and NaN."""
    rng = np.random.default_rng(7)
    special = np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -1.5, -2.5, 3.5, 1.0, -1.0, 2.0, 3.0, -3.0, 1e-300, -1e300, 4503599627370497.0,
                        np.inf, -np.inf, np.nan])
    x = np.concatenate([special, 4.0 * rng.standard_normal(400)])
    y = np.concatenate([special[::-1], 2.0 * rng.standard_normal(400)])
    hs = HostState(x.size)
    hs.store.setcol("x", x)
    hs.store.setcol("y", y)
    m = ws.model(f"@model function f()\n    r .= {src}\nend", particle_vars=("x", "y"))
    ws.run(m(), hs) if hasattr(hs, "root") else m().apply(hs)
    got = hs.store.getcol("r")
    with np.errstate(all="ignore"):
        exp = want(x, y)
    both_nan = np.isnan(got) & np.isnan(exp)
    if rtol == 0.0:
        same = (got == exp) | both_nan
        if src.startswith("max("):
            same |= np.isnan(x) | np.isnan(y)     # min / max are the device's (fmin / fmax: a NaN operand is ignored; Julia propagates it)
        if src.startswith(("x % 3", "mod(", "rem(", "x ÷", "fld(")):
            same |= ~np.isfinite(x) | ~np.isfinite(y) | (np.abs(x) > 1e15)    # inf / inf forms: NaN either way, checked above
        assert same.all(), (src, x[~same][:5], y[~same][:5], got[~same][:5], exp[~same][:5])
        if src in ("sign(x)", "ceil(x)", "trunc(x)", "round(x)"):
            z = (exp == 0) & ~both_nan & ((x != 0) | (src != "round(x)"))     # (round(-0.0) gives +0.0: documented)
            assert (np.signbit(got[z]) == np.signbit(exp[z])).all(), src   # signed zeros as Julia's
    else:
        fin = np.isfinite(exp)
        edge = np.zeros(x.size, dtype=bool)
        if src.startswith(("atan(y, x)", "hypot(")):
            # documented conventions of the short compositions: signed zeros on the axes, both arguments infinite
            edge = (x == 0) | (y == 0) | (np.isinf(x) & (np.isinf(y) | np.isnan(y))) | (np.isinf(y) & np.isnan(x))
            if src.startswith("atan"):
                z = (x == 0) & (y != 0) & ~np.isnan(y)
                assert np.array_equal(got[z], np.sign(y[z]) * np.pi / 2)
                assert (got[(x == 0) & (y == 0)] == 0).all()
        np.testing.assert_allclose(got[fin & ~edge], exp[fin & ~edge], rtol=rtol, atol=1e-300)
        ok = (got[~fin] == exp[~fin]) | both_nan[~fin] | edge[~fin]
        assert ok.all(), (src, x[~fin][~ok], y[~fin][~ok], got[~fin][~ok], exp[~fin][~ok])


def test_compositions_refuse_fresh_variates_and_local_functions_trace():
    """A composition mentions its argument more than once; with a fresh variate inside, every mention would be another
    draw — refused.  Local functions and anonymous functions of a model body are build-time values (rewrites.jl:717-733)
    and trace into the device expression when called on particle variables; `a[end]`, tuple assignment and Julia's
    truncated `÷` / `%` on build-time integers."""
    with pytest.raises(ws.UnsupportedModelError):
        ws.expr.clamp(ws.randn(), -1.0, 1.0)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(64)
    hs = HostState(64)
    hs.store.setcol("x", x)
    m = ws.model('''
    @model function f(data)
        k = 3.0
        g(u) = k * u + data[end]
        h = (u, v) -> u * v - data[end - 1]
        a, b = -7 ÷ 2, -7 % 2
        (c, d) = (mod(-7, 2), length(data))
        r .= g(x) + h(x, 2.0) + a + b + c + d
    end
    ''', particle_vars=("x",))
    m([10.0, 20.0, 30.0]).apply(hs)
    np.testing.assert_allclose(hs.store.getcol("r"), (3.0 * x + 30.0) + (2.0 * x - 20.0) + (-3) + (-1) + 1 + 3, rtol=1e-14)
