"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): ancestor indices bit-exact except where a uniform lies within
1e-12 of a CDF boundary (the count is printed); log-weights / log-evidence within 1e-9 relative.
"""
import math

import numpy as np
import pytest

from oracle import cref, ref
from oracle import models as om

pytestmark = pytest.mark.gpu

REL = 1e-9


@pytest.fixture(scope="module")
def ctx(ws):
    return ws.SMCState(1024, device=0)


def skewed_logw(n, s, seed=0x5EED):
    return s * np.random.default_rng(seed).standard_normal(n)


# ------------------------------------------------------------------------------------------------
# resampling numerics (src/resampling.jl)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 255, 256, 2047, 2048, 2049, 100_003, 1_000_000])
@pytest.mark.parametrize("s", [0.0, 0.5, 4.0])
def test_exp_norm_logsumexp_ess(ws, ctx, n, s):
    lw = skewed_logw(n, s) - 700.0  # large offset: the max-shift must make this safe
    w = ws.exp_norm(lw, ctx)
    w_ref = cref.exp_norm(lw)
    np.testing.assert_allclose(w, w_ref, rtol=1e-12, atol=0)
    assert abs(w.sum() - 1.0) < 1e-12
    lse = ws.logsumexp(lw, ctx)
    assert abs(lse - cref.logsumexp(lw)) <= 1e-12 * abs(lse)
    ess = ws.ess_perc(w, ctx)
    assert abs(ess - cref.ess_perc(w_ref)) <= 1e-11 * ess


def test_exp_norm_with_minus_inf(ws, ctx):
    lw = np.array([-np.inf, 0.0, -1.0, -np.inf, 2.0])
    w = ws.exp_norm(lw, ctx)
    np.testing.assert_allclose(w, ref.exp_norm(lw), rtol=1e-13)
    assert w[0] == 0.0 and w[3] == 0.0


def check_ancestors(a_dev, w, us, label):
    """bit-exact vs the reference's sequential-CDF icdf, except near-boundary uniforms"""
    n = w.shape[0]
    a_ref, clamped_ref = cref.icdf(w, us)
    a_dev = a_dev.astype(np.int64)
    assert a_dev.min() >= 0 and a_dev.max() < n
    assert np.all(np.diff(a_dev) >= 0), "ancestors must be sorted"
    bad = np.nonzero(a_dev != a_ref)[0]
    if bad.size:
        cdf = np.cumsum(w)
        # every mismatch must sit within 1e-12 of the CDF boundary that separates the two answers
        lo = np.minimum(a_dev[bad], a_ref[bad])
        dist = np.abs(us[bad] - cdf[lo])
        assert np.all(dist < 1e-12), f"{label}: mismatch not explained by a near-boundary uniform: {dist.max()}"
    print(f"[{label}] n={n} near-boundary ancestor mismatches: {bad.size}")
    return bad.size


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 2048, 2049, 6145, 100_000, 1_000_000])
@pytest.mark.parametrize("s", [0.5, 2.0, 4.0])
def test_stratified_resample_identical_uniforms(ws, ctx, n, s):
    w = cref.exp_norm(skewed_logw(n, s))
    r = np.random.default_rng(n + 17).random(n)
    a = ws.resample_indices(w, ctx, "stratified", uniforms=r)
    check_ancestors(a, w, cref.stratified_us(r), f"stratified s={s}")


def test_stratified_one_hot_and_zero_weights(ws, ctx):
    n = 50_000
    w = np.full(n, 0.01 / (n - 1))
    w[n // 3] = 0.99
    w[: n // 10] = 0.0  # leading zero-weight particles
    w = w / w.sum()
    r = np.random.default_rng(3).random(n)
    a = ws.resample_indices(w, ctx, "stratified", uniforms=r)
    check_ancestors(a, w, cref.stratified_us(r), "one-hot")
    assert (a == n // 3).sum() > 0.98 * n
    assert np.all(a >= n // 10)
    # uniform u = 0 on slot 1 with w[0] = 0: icdf keeps particle 1 (0-based 0)
    w2 = np.array([0.0, 0.5, 0.5, 0.0])
    r2 = np.array([0.0, 0.5, 0.5, 0.999])
    a2 = ws.resample_indices(w2, ctx, "stratified", uniforms=r2)
    np.testing.assert_array_equal(a2, ref.icdf_loop(w2, ref.stratified_us(r2)))


def test_icdf_arbitrary_sorted_uniforms(ws, ctx):
    n = 20_000
    w = cref.exp_norm(skewed_logw(n, 1.5))
    us = np.sort(np.random.default_rng(5).random(n))
    a = ws.icdf(w, us, ctx)
    check_ancestors(a, w, us, "icdf sorted-u")


def test_systematic_and_multinomial(ws, ctx):
    n = 30_000
    w = cref.exp_norm(skewed_logw(n, 2.0))
    a = ws.resample_indices(w, ctx, "systematic", uniforms=[0.37])
    check_ancestors(a, w, ref.systematic_us(0.37, n), "systematic")
    u = np.random.default_rng(9).random(n)
    a = ws.resample_indices(w, ctx, "multinomial", uniforms=u)
    check_ancestors(a, w, np.sort(u), "multinomial")


@pytest.mark.parametrize("n,s,scheme", [(1, 1.0, "stratified"), (257, 2.0, "stratified"), (4096, 0.5, "stratified"),
                                        (50_001, 3.0, "stratified"), (20_000, 2.0, "systematic")])
def test_production_integer_search_is_exact(ws, ctx, n, s, scheme):
    """The Philox (production) path evaluates the stratified search in exact integer arithmetic;
    reproduce it with Python big integers: every ancestor must match, no exceptions."""
    import ctypes as C
    w = cref.exp_norm(skewed_logw(n, s, seed=n))
    stream, seed = C.c_uint64(), C.c_uint64()
    ctx.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
    a, clamped = ws.resample_indices(w, ctx, scheme, return_clamped=True)
    a_ref, clamped_ref = ref.stratified_ancestors_fixed_point(w, seed.value, stream.value, scheme)
    np.testing.assert_array_equal(a, a_ref)
    assert clamped == clamped_ref


def test_production_search_one_hot_heavy_family(ws, ctx):
    import ctypes as C
    n = 80_000
    w = np.full(n, 0.02 / (n - 1))
    w[12345] = 0.98
    w = w / w.sum()
    stream, seed = C.c_uint64(), C.c_uint64()
    ctx.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
    a = ws.resample_indices(w, ctx, "stratified")
    a_ref, _ = ref.stratified_ancestors_fixed_point(w, seed.value, stream.value)
    np.testing.assert_array_equal(a, a_ref)
    assert (a == 12345).sum() > 0.97 * n


def test_resample_philox_statistics(ws, ctx):
    """native (Philox) uniforms: offspring counts of stratified resampling are within 1 of N*w"""
    n = 200_000
    w = cref.exp_norm(skewed_logw(n, 1.0))
    a = ws.resample_indices(w, ctx, "stratified")
    counts = np.bincount(a, minlength=n)
    assert counts.sum() == n
    assert np.all(np.abs(counts - n * w) < 2.0 + 1e-6)
    a2 = ws.resample_indices(w, ctx, "stratified")
    assert np.any(a2 != a), "successive resamples must use fresh uniforms"


def test_clamped_slots_reported(ws, ctx):
    # weights summing to < 1: the last slots run past the CDF; the reference throws, we clamp and count
    w = np.full(100, 0.009)
    r = np.full(100, 0.5)
    a, clamped = ws.resample_indices(w, ctx, "stratified", uniforms=r, return_clamped=True)
    assert clamped == 10 and np.all(a[-10:] == 99)


# ------------------------------------------------------------------------------------------------
# store + statements
# ------------------------------------------------------------------------------------------------
def test_store_interface(ws):
    st = ws.SMCState(100, device=0)
    assert st.store.nparticles() == 100 and st.store.colnames() == []
    assert repr(st.store) == "ColumnStore(n=100, columns=[])"
    v = np.arange(100.0)
    st.store.setcol("x", v)
    st.store.setcol("th", np.stack([v, -v], axis=1))
    assert st.store.hascol("x") and not st.store.hascol("y")
    np.testing.assert_array_equal(st["x"], v)
    np.testing.assert_array_equal(st["th"], np.stack([v, -v], axis=1))
    idx = np.random.default_rng(0).integers(0, 100, 100)
    st.store.resample(idx)
    np.testing.assert_array_equal(st["x"], v[idx])
    np.testing.assert_array_equal(st["th"][:, 1], -v[idx])
    np.testing.assert_array_equal(st.weights, np.zeros(100))
    assert st.store.colnames() == ["x", "th"]


def _replay_run(ws, root, n, normals=(), uniforms=(), exponentials=(), ess=0.5, resampler="stratified", oracle_root=None):
    """device: the product's tree (normally `ws.model(source)`); host: `oracle_root`, the oracle's own hand-built
    program of the same model (oracle/models.py), so the product's @model front-end is part of what is checked"""
    state = ws.SMCState(n, ess_perc_min=ess, device=0, resampler=resampler)
    state.set_replay(normals=normals if len(normals) else None, uniforms=uniforms if len(uniforms) else None,
                     exponentials=exponentials if len(exponentials) else None)
    ws.run(root, state)
    ost = ref.OracleState(n, ref.Streams(normals, uniforms, exponentials), ess_perc_min=ess, resampler=resampler)
    ref.run(oracle_root if oracle_root is not None else root, ost)
    return state, ost


def _compare_states(state, ost, rel=REL, max_bad=0):
    assert state.store.colnames() == ost.names
    total_bad = 0
    for name in ost.names:
        a, b = state[name], ost.cols[name]
        assert a.shape == b.shape, name
        bad = np.abs(a - b) > rel * (1.0 + np.abs(b))
        total_bad += int(bad.reshape(a.shape[0], -1).any(axis=1).sum())
    wa, wb = state.weights, ost.weights
    np.testing.assert_allclose(wa, wb, rtol=rel, atol=rel)
    assert total_bad <= max_bad, f"{total_bad} particles differ"
    return total_bad


SSM1D = '''
@model function ssm(obs)
    x{1} .= 0.0
    v .= 0.0
    for (t, o) in enumerate(obs)
        x{t + 1} .= x{t} + v
        dv ~ Normal(0.0, 0.1)
        v .= v + dv
        o => Normal(x{t + 1}, 1.0)
    end
end
'''

SSM2D = '''
@model function ssm(obs)
    I2 = [1.0 0.0; 0.0 1.0]
    x{1} .= [0.0, 0.0]
    v .= [1.0, 0.0]
    for (t, o) in enumerate(obs)
        x{t + 1} .= x{t} + v
        dv ~ MvNormal([0.0, 0.0], 0.1 * I2)
        v .= v + dv
        o => MvNormal(x{t + 1}, 0.5 * I2)
    end
end
'''


def test_c1_ssm1d_replay(ws):
    """BASELINE configs[0]: examples/1D_ssm.jl, N = 1000, T = 50, history kept (53 columns)."""
    n, T = 1000, 50
    rng = np.random.default_rng(7)
    x, v, obs = 0.0, 0.0, []
    for _ in range(T):
        obs.append(x + rng.standard_normal())
        x, v = x + v, v + 0.1 * rng.standard_normal()
    root = ws.model(SSM1D)(obs)
    state, ost = _replay_run(ws, root, n, normals=rng.standard_normal(n * T), uniforms=rng.random(n * T), oracle_root=om.ssm1d(obs))
    fired = [e for e in ost.log if e["resampled"]]
    assert len(fired) >= 3, "the ESS gate should fire several times in 50 steps"
    assert len(state.store.colnames()) == T + 3
    _compare_states(state, ost)
    assert abs(ws.log_evidence(state) - ref.log_evidence(ost)) <= REL * abs(ref.log_evidence(ost))
    assert state.stats()["resamples_done"] == len(fired)
    assert state.depth == ost.depth == 2 + 4 * T


def test_c2_ssm2d_replay(ws):
    n, T = 5000, 25
    rng = np.random.default_rng(42)
    obs = [rng.standard_normal(2) + np.array([t, 0.0]) for t in range(T)]
    root = ws.model(SSM2D)(obs)
    state, ost = _replay_run(ws, root, n, normals=rng.standard_normal(2 * n * T), uniforms=rng.random(n * T), oracle_root=om.ssm2d(obs))
    assert sum(e["resampled"] for e in ost.log) >= 3
    _compare_states(state, ost)
    # ancestors of the last firing resample are bit-identical
    last = [e for e in ost.log if e["resampled"]][-1]
    assert abs(ws.log_evidence(state) - ref.log_evidence(ost)) <= REL * abs(ref.log_evidence(ost))
    assert last["ancestors"].shape == (n,)


def test_resample_state_machine(ws):
    """transformers.jl:474-498: no-op without weights_changed (resampled untouched); gate on ess_perc_min."""
    n = 2000
    st = ws.SMCState(n, ess_perc_min=0.5, device=0)
    st.store.setcol("x", np.arange(n, dtype=float))
    r = ws.Resample()
    st.resampled = True
    r.apply(st)
    assert r.last.fired == 0 and st.resampled is True          # untouched
    # weighted, ESS high -> fired but not resampled
    ws.Observe(0.0, "Normal", (ws.col("x") * 1e-9, 1.0)).apply(st)
    assert st.weights_changed
    r.apply(st)
    assert r.last.fired == 1 and st.resampled is False and not st.weights_changed
    assert r.last.ess_perc > 0.99
    # strongly weighted -> resamples, weights reset to the log-mean, evidence preserved
    ws.Observe(0.0, "Normal", (ws.col("x"), 50.0)).apply(st)
    lw = st.weights
    le_before = ref.logsumexp(lw) - math.log(n)
    uniforms = np.random.default_rng(1).random(n)
    st.set_replay(uniforms=uniforms)
    r.apply(st)
    assert st.resampled is True
    np.testing.assert_allclose(st.weights, np.full(n, le_before), rtol=1e-12)
    assert abs(ws.log_evidence(st) - le_before) < 1e-12 * abs(le_before)
    anc = ref.icdf(ref.exp_norm(lw), ref.stratified_us(uniforms))
    np.testing.assert_array_equal(st["x"], np.arange(n, dtype=float)[anc])


def test_lgssm_kalman_native_rng(ws):
    """test/transformers_test.jl:158-186 with Philox: |dlogZ| <= 3.0 and |dmean| <= 1.0 at N = 1e4, T = 50
    (we use N = 2e5 and much tighter bounds)."""
    a, q, r = 0.9, 1.0, 0.5
    rng = np.random.default_rng(42)
    x, ys = rng.standard_normal(), []
    for _ in range(50):
        x = a * x + q * rng.standard_normal()
        ys.append(x + r * rng.standard_normal())
    lg = ws.model('''
    @model function lgssm1d(data, a, q, r, x0_std)
        x ~ Normal(0.0, x0_std)
        for y in data
            x ~ Normal(a * x, q)
            y => Normal(x, r)
        end
    end
    ''')
    state = ws.SMCState(200_000, ess_perc_min=0.5, seed=11, device=0)
    ws.run(lg(ys, a, q, r, 1.0), state)
    mean_exact, le_exact = ref.kalman_filter_evidence(ys, a, q, r)
    assert abs(ws.log_evidence(state) - le_exact) < 0.05
    assert abs(ws.E(lambda x: x, state) - mean_exact) < 0.02
    assert state.stats()["resamples_done"] > 5


def test_random_walk_moments_native_rng(ws):
    """test/transformers_test.jl:14-30: K = 4 walks, T = 10: mean 0, var T + 1."""
    K, T, n = 4, 10, 100_000
    steps = [ws.Sample(f"x{k}", "Normal", (0.0, 1.0)) for k in range(K)]
    loop = ws.Loop(lambda s: range(T), lambda t: ws.Sequence(*[ws.Sample(f"x{k}", "Normal", (ws.col(f"x{k}"), 1.0))
                                                                 for k in range(K)]))
    state = ws.SMCState(n, seed=5, device=0)
    ws.run(ws.Sequence(*steps, loop), state)
    for k in range(K):
        x = state[f"x{k}"]
        assert abs(x.mean()) < 0.05 and abs(x.var() - (T + 1)) < 0.05 * (T + 1)
    # distinct statements / particles draw distinct numbers
    assert abs(np.corrcoef(state["x0"], state["x1"])[0, 1]) < 0.02


def test_exponential_and_importance(ws):
    n = 200_000
    st = ws.SMCState(n, seed=3, device=0)
    ws.Sample("t", "Exponential", (5.0,)).apply(st)
    t = st["t"]
    assert t.min() >= 0 and abs(t.mean() - 5.0) < 0.05 and abs(t.var() - 25.0) < 0.6
    # importance_kernel(Normal(0,2), Normal(0,1)): IS mean of x^2 = 1, logZ = 0  (test/importance_kernel_test.jl:6-29)
    k = ws.importance_kernel(ws.NormalDist(0.0, 2.0), ws.NormalDist(0.0, 1.0))
    ws.Sample("x", k, ()).apply(st)
    assert st.weights_changed
    x = st["x"]
    np.testing.assert_allclose(st.weights, ref.normal_logpdf(x, 0, 1) - ref.normal_logpdf(x, 0, 2), rtol=1e-10, atol=1e-12)
    assert abs(ws.E(lambda x: x * x, st) - 1.0) < 0.05
    assert abs(ws.log_evidence(st)) < 0.05


def test_expression_ops(ws):
    n = 1000
    st = ws.SMCState(n, device=0)
    rng = np.random.default_rng(0)
    a, b = rng.uniform(0.5, 2.0, n), rng.uniform(0.5, 2.0, n)
    st.store.setcol("a", a)
    st.store.setcol("b", b)
    e = (ws.exp(ws.col("a")) * ws.col("b") + ws.log(ws.col("b")) / ws.col("a") - ws.sqrt(ws.col("a")) ** 3.0
         + ws.sin(ws.col("a")) * ws.cos(ws.col("b")) - abs(-ws.col("a")) + 2.0 * ws.col("a") ** 2 - 7.0 / ws.col("b"))
    ws.Assign("c", e).apply(st)
    expect = np.exp(a) * b + np.log(b) / a - np.sqrt(a) ** 3.0 + np.sin(a) * np.cos(b) - np.abs(-a) + 2.0 * a ** 2 - 7.0 / b
    np.testing.assert_allclose(st["c"], expect, rtol=1e-13)
    # weights through an arbitrary expression, and @E of products
    ws.Weight(None, (-(ws.col("a") - 1.0) ** 2,)).apply(st)
    np.testing.assert_allclose(st.weights, -(a - 1.0) ** 2, rtol=1e-13)
    w = ref.exp_norm(st.weights)
    got = ws.expectation([ws.col("a"), ws.col("a") * ws.col("b"), ws.col("c")], st)
    np.testing.assert_allclose(got, [np.sum(w * a), np.sum(w * a * b), np.sum(w * expect)], rtol=1e-11)
    assert abs(ws.E(lambda a, b: a + b, st) - np.sum(w * (a + b))) < 1e-11


def test_vector_assign_hazard_and_accessors(ws):
    n = 500
    st = ws.SMCState(n, device=0)
    v = np.random.default_rng(1).standard_normal((n, 2))
    st.store.setcol("p", v)
    ws.Assign("p", [ws.col("p")[1], ws.col("p")[0]]).apply(st)   # swap: RHS evaluated before the write
    np.testing.assert_array_equal(st["p"], v[:, ::-1])
    ws.Assign(("p", 0), ws.col("p")[1] * 2.0).apply(st)            # x[1] .= ...
    np.testing.assert_array_equal(st["p"][:, 0], v[:, 0] * 2.0)
    ws.Assign("q", np.zeros(3)).apply(st)
    assert st["q"].shape == (n, 3)
    with pytest.raises(KeyError):
        ws.Assign(("nope", 0), 1.0).apply(st)


# ------------------------------------------------------------------------------------------------
# API surface (test/api_test.jl)
# ------------------------------------------------------------------------------------------------
def test_api_surface(ws):
    n = 100_000
    st = ws.SMCState(n, seed=1, device=0)
    m = ws.model('''
    @model function m()
        x ~ Normal(2.0, 3.0)
    end
    ''')
    ws.run(m(), st)
    assert abs(ws.E(lambda x: x, st) - 2.0) < 0.05
    assert abs(ws.E(lambda x: x ** 2, st) - 13.0) < 0.2
    df = ws.sample(st, 1000)
    assert df.shape == (1000, 1) and abs(df["x"].mean() - 2.0) < 0.4
    df2 = ws.sample(st, 10, replace=False)
    assert len(set(df2["x"])) == 10
    with pytest.raises(ValueError):
        ws.sample(st, 0)
    with pytest.raises(ValueError):
        ws.sample(st, n + 1, replace=False)
    full = ws.to_dataframe(st)
    assert list(full.columns) == ["x", "log_weight"] and len(full) == n
    w = ws.exp_norm(st)
    assert abs(w.sum() - 1.0) < 1e-12 and abs(ws.ess_perc(st) - 1.0) < 1e-12
    assert "SMCState(n_particles=100000" in repr(st)


def test_multinomial_resampling_inside_run(ws):
    """multinomial = icdf over sort(N iid uniforms).  Replay against the oracle, then the
    Philox path: offspring counts are unbiased and noisier than stratified ones."""
    n, T = 5000, 8
    rng = np.random.default_rng(12)
    ys = list(rng.normal(size=T))
    import models
    root = ws.model(models.LGSSM1D)(ys, 0.9, 1.0, 0.5, 1.0)
    state, ost = _replay_run(ws, root, n, normals=rng.standard_normal(n * (T + 1)), uniforms=rng.random(n * T), ess=1.0,
                             resampler="multinomial", oracle_root=om.lgssm1d(ys, 0.9, 1.0, 0.5, 1.0))
    assert state.stats()["resamples_done"] == sum(1 for e in ost.log if e["resampled"]) == T
    _compare_states(state, ost, max_bad=3)
    assert abs(ws.log_evidence(state) - ref.log_evidence(ost)) <= REL * abs(ref.log_evidence(ost))
    # Philox draws: E[offspring of m] = N w_m; variance of the count of a particle with N w = 1 is ~1 (multinomial)
    # against < 1/4 for the stratified scheme
    n2 = 200_000
    lw = 0.7 * rng.normal(size=n2)
    counts = {}
    for scheme in ("multinomial", "stratified"):
        st = ws.SMCState(n2, ess_perc_min=float("inf"), seed=3, resampler=scheme, device=0)
        st.store.setcol("id", np.arange(n2, dtype=np.float64))
        st.weights = lw
        st.weights_changed = True
        ws.Resample().apply(st)
        ids = st["id"].astype(np.int64)
        assert np.all(np.diff(ids) >= 0)
        counts[scheme] = np.bincount(ids, minlength=n2)
    w = ref.exp_norm(lw)
    for scheme, c in counts.items():
        assert c.sum() == n2
        assert abs(np.sum(c * np.arange(n2)) / n2 - np.sum(w * np.arange(n2))) < 600.0     # weighted mean of the index
    resid_m = counts["multinomial"] - n2 * w
    resid_s = counts["stratified"] - n2 * w
    assert 0.7 < resid_m.var() / np.mean(n2 * w * (1 - w)) < 1.3                           # multinomial variance
    assert resid_s.var() < 0.5 * resid_m.var()


def test_device_against_reference_vectors_if_present(ws, ctx):
    """tests/golden/reference_vectors.bin is written by the REAL reference where a Julia toolchain exists
    (oracle/gen_from_reference.jl); with it, exp_norm / logsumexp / ess_perc / stratified_resample / icdf of the device
    are compared with Julia's own outputs (ancestors bit-exact except near-boundary uniforms, counted)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.bin")
    if not os.path.exists(path):
        pytest.skip("no Julia toolchain was available at build time: reference vectors not generated (parity unpinned)")
    buf = open(path, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    for _ in range(int(take("<i8", 1)[0])):
        n = int(take("<i8", 1)[0])
        logw, w = take("<f8", n).copy(), take("<f8", n).copy()
        lse, ess = float(take("<f8", 1)[0]), float(take("<f8", 1)[0])
        r, idx = take("<f8", n).copy(), take("<i8", n)
        us, idx2 = take("<f8", n).copy(), take("<i8", n)
        np.testing.assert_allclose(ws.exp_norm(logw, ctx), w, rtol=1e-12)
        assert abs(ws.logsumexp(logw, ctx) - lse) <= 1e-12 * abs(lse) and abs(ws.ess_perc(w, ctx) - ess) <= 1e-11 * ess
        a = ws.resample_indices(w, ctx, "stratified", uniforms=r).astype(np.int64)
        bad = np.nonzero(a != idx - 1)[0]
        cdf = np.cumsum(w)
        assert np.all(np.abs(ref.stratified_us(r)[bad] - cdf[np.minimum(a[bad], idx[bad] - 1)]) < 1e-12)
        a2 = ws.icdf(w, us, ctx).astype(np.int64)
        bad2 = np.nonzero(a2 != idx2 - 1)[0]
        assert np.all(np.abs(us[bad2] - cdf[np.minimum(a2[bad2], idx2[bad2] - 1)]) < 1e-12)
        print(f"[reference vectors] n={n}: near-boundary mismatches {bad.size} (stratified), {bad2.size} (icdf)")
