"""GPU parity tests for the score tape and the MH moves (ports of test/score_test.jl, test/move_test.jl,
test/move_macro_test.jl; replayed streams must reproduce the oracle's accept decisions exactly)."""
import math

import numpy as np
import pytest

from attribution import attribute
from oracle import models as om
from oracle import ref

pytestmark = pytest.mark.gpu


def test_score_logpdf_depth_cutoff(ws):
    """test/score_test.jl:20-54: theta ~ N(0,1); x .= theta; 1.5 => N(x, 0.5)."""
    n = 1000
    root = ws.Sequence(ws.Sample("θ", "Normal", (0.0, 1.0)), ws.Assign("x", ws.col("θ")),
                       ws.Observe(1.5, "Normal", (ws.col("x"), 0.5)))
    st = ws.SMCState(n, seed=42, device=0)
    ws.run(root, st)
    th, x = st["θ"], st["x"]
    np.testing.assert_array_equal(th, x)
    e1 = ref.normal_logpdf(th, 0.0, 1.0)
    e3 = e1 + ref.normal_logpdf(1.5, x, 0.5)
    assert np.all(ws.score_logpdf(st, ["θ"], 0) == 0.0)
    np.testing.assert_allclose(ws.score_logpdf(st, ["θ"], 1), e1, rtol=1e-12)
    np.testing.assert_allclose(ws.score_logpdf(st, ["θ"], 2), e1, rtol=1e-12)
    np.testing.assert_allclose(ws.score_logpdf(st, ["θ"], 3), e3, rtol=1e-12)
    np.testing.assert_allclose(st.weights, ref.normal_logpdf(1.5, x, 0.5), rtol=1e-12)


def _static_model(ws, y, tau0=1.0, sigma=1.0, extra=None):
    steps = [ws.Sample("θ", "Normal", (0.0, tau0))] + [ws.Observe(float(v), "Normal", (ws.col("θ"), sigma)) for v in y]
    if extra is not None:
        steps.append(extra)
    return ws.Sequence(*steps)


def test_move_cancellation_and_replay_exactness(ws):
    """test/move_test.jl:23-58: a target-independent factor does not change accept decisions; and with
    replayed streams the device move equals the oracle's move."""
    T, n, step = 3, 1000, 0.3
    rng = np.random.default_rng(1)
    y = rng.standard_normal(T)
    th0 = rng.standard_normal(n)
    normals, uniforms = rng.standard_normal(n), rng.random(n)
    out = []
    for with_extra in (False, True):
        extra = ws.Observe(0.7, "Normal", (0.0, 1.0)) if with_extra else None
        root = _static_model(ws, y, extra=extra)
        st = ws.SMCState(n, device=0)
        st.store.setcol("θ", th0)
        st.root = root
        st.depth = T + 1 + int(with_extra)
        st.set_replay(normals=normals, uniforms=uniforms)
        mv = ws.Move(["θ"], ws.RW, (step,))
        mv.apply(st)
        ost = ref.OracleState(n, ref.Streams(normals, uniforms))
        ost.setcol("θ", th0)
        ost.root, ost.depth = root, T + 1 + int(with_extra)
        ref.apply_move(mv, ost)
        np.testing.assert_array_equal(st["θ"], ost.cols["θ"])  # proposals are bit-exact, decisions identical
        assert mv.last.ran == 1 and mv.last.n_accepted == int(ost.last_accept.sum())
        assert 0.2 * n < mv.last.n_accepted < n
        out.append(st["θ"])
    np.testing.assert_allclose(out[0], out[1], atol=1e-9)


@pytest.mark.parametrize("proposal,args", [
    ("RW", (0.25,)), ("RW", (0.25, (0.0, math.inf))), ("RW", (0.4, (-1.0, 6.0))), ("autoRW", ()),
    ("autoRW", (1e-3, (0.0, math.inf))), ("autoRW", (1e-3, (-math.inf, 9.0))),
])
def test_move_replay_vs_oracle_bounds(ws, proposal, args):
    n = 4000
    rng = np.random.default_rng(3)
    tau0 = np.abs(rng.standard_normal(n)) + 0.2
    y = [0.8, 1.4, 2.2]
    root = ws.Sequence(ws.Sample("τ", "Exponential", (5.0,)),
                       *[ws.Observe(v, "Normal", (0.5 * ws.col("τ"), ws.col("τ"))) for v in y])
    lw = 0.3 * rng.standard_normal(n)  # non-uniform weights: exercises the weighted covariance of autoRW
    normals, uniforms = rng.standard_normal(n), rng.random(n)
    st = ws.SMCState(n, device=0)
    st.store.setcol("τ", tau0)
    st.weights = lw
    st.root, st.depth = root, 4
    st.set_replay(normals=normals, uniforms=uniforms)
    mv = ws.Move(["τ"], proposal, args)
    mv.apply(st)
    ost = ref.OracleState(n, ref.Streams(normals, uniforms))
    ost.setcol("τ", tau0)
    ost.weights = lw.copy()
    ost.root, ost.depth = root, 4
    ref.apply_move(mv, ost)
    a, b = st["τ"], ost.cols["τ"]
    bad = np.abs(a - b) > 1e-9 * (1 + np.abs(b))
    assert bad.sum() <= 1, f"{bad.sum()} particles differ"
    assert abs(mv.last.n_accepted - int(ost.last_accept.sum())) <= 1
    np.testing.assert_array_equal(st.weights, lw)  # a move never touches the weights


def test_joint_move_two_targets_replay(ws):
    n = 3000
    rng = np.random.default_rng(8)
    xs, ys = rng.uniform(0, 10, 6), rng.standard_normal(6)
    root = ws.Sequence(ws.Sample("α", "Normal", (0.0, 10.0)), ws.Sample("β", "Normal", (0.0, 10.0)),
                       *[ws.Observe(float(y), "Normal", (ws.col("α") + ws.col("β") * float(x), 1.0)) for x, y in zip(xs, ys)])
    a0, b0 = rng.standard_normal(n), 0.1 * rng.standard_normal(n) + 0.5 * rng.standard_normal(n)
    for proposal, args, nn in (("autoRW", (), 2 * n), ("RW", (0.05,), 2 * n)):
        normals, uniforms = rng.standard_normal(nn), rng.random(n)
        st = ws.SMCState(n, device=0)
        st.store.setcol("α", a0)
        st.store.setcol("β", b0)
        st.root, st.depth = root, 8
        st.set_replay(normals=normals, uniforms=uniforms)
        mv = ws.Move(["α", "β"], proposal, args)
        mv.apply(st)
        ost = ref.OracleState(n, ref.Streams(normals, uniforms))
        ost.setcol("α", a0)
        ost.setcol("β", b0)
        ost.root, ost.depth = root, 8
        ref.apply_move(mv, ost)
        for c in ("α", "β"):
            bad = np.abs(st[c] - ost.cols[c]) > 1e-9 * (1 + np.abs(ost.cols[c]))
            assert bad.sum() <= 1, (proposal, c, int(bad.sum()))


def test_move_invariance_native_rng(ws):
    """test/move_test.jl:69-98: RW sweeps leave the exact Normal-Normal posterior invariant."""
    T, tau0, sigma, n = 5, 2.0, 1.0, 200_000
    rng = np.random.default_rng(42)
    y = rng.standard_normal(T) * sigma + 1.3
    post_var = 1.0 / (1.0 / tau0 ** 2 + T / sigma ** 2)
    post_mean = post_var * y.sum() / sigma ** 2
    st = ws.SMCState(n, seed=9, device=0)
    st.store.setcol("θ", rng.standard_normal(n) * math.sqrt(post_var) + post_mean)
    st.root, st.depth = _static_model(ws, y, tau0, sigma), T + 1
    mv = ws.Move(["θ"], ws.RW, (0.3,))
    for _ in range(20):
        mv.apply(st)
    th = st["θ"]
    assert abs(th.mean() - post_mean) < 0.01 and abs(th.var() - post_var) < 0.01
    assert 0.5 * n < mv.last.n_accepted < 0.95 * n


def test_diversity_gating(ws):
    """test/move_test.jl:116-209."""
    n = 50_000
    rng = np.random.default_rng(3)
    th0 = rng.standard_normal(n)
    st = ws.SMCState(n, seed=2, device=0)
    st.store.setcol("θ", th0)
    mv = ws.Move(["θ"], ws.RW, (0.3,), 0.99)
    mv.apply(st)                                     # fully diverse: exact no-op, no root needed
    assert mv.last.ran == 0 and mv.last.diversity == 1.0
    np.testing.assert_array_equal(st["θ"], th0)
    # collapsed particles: runs until the gate is satisfied, then stops
    T, y = 5, rng.standard_normal(5) + 1.3
    st.store.setcol("θ", np.full(n, 0.7))
    st.root, st.depth = _static_model(ws, y, 2.0, 1.0), T + 1
    mv = ws.Move(["θ"], ws.RW, (0.3,), 0.9)
    assert abs(ws.marginal_diversity(st.store, ["θ"]) - 1.0 / n) < 1e-15
    ran = 0
    for _ in range(100):
        mv.apply(st)
        ran += mv.last.ran
    assert 0 < ran < 100
    assert ws.marginal_diversity(st.store, ["θ"]) >= 0.9
    snap = st["θ"]
    for _ in range(5):
        mv.apply(st)
    np.testing.assert_array_equal(st["θ"], snap)
    # marginal, not joint (move_test.jl:196-209); NaNs count once, -0.0 != 0.0 (isequal)
    a = np.repeat(np.arange(1.0, 6.0), n // 5)
    st.store.setcol("α", a)
    st.store.setcol("β", np.arange(1.0, n + 1.0))
    assert abs(ws.marginal_diversity(st.store, ["α", "β"]) - 5.0 / n) < 1e-15
    z = np.arange(n, dtype=float)
    z[:10] = np.nan
    z[10], z[11] = 0.0, -0.0
    st.store.setcol("z", z)
    assert abs(ws.marginal_diversity(st.store, ["z"]) - ref.marginal_diversity_bits(z)) < 1e-15


LINREG = '''
@model function linear_regression(xs, ys)
    α ~ Normal(0.0, 10.0)
    β ~ Normal(0.0, 10.0)
    for (x, y) in zip(xs, ys)
        y => Normal(α + β * x, 1.0)
        if resampled
            α << autoRW()
            β << autoRW()
        end
    end
end
'''


def test_c3_linear_regression_replay_and_recovery(ws):
    """BASELINE configs[2] shape at small N: replay parity with the oracle, then posterior recovery
    (test/move_macro_test.jl:26-61) with native RNG."""
    n, npts = 2000, 12
    rng = np.random.default_rng(42)
    xs = rng.uniform(0, 10, npts)
    ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(npts)
    root = ws.model(LINREG)(xs, ys)
    normals, uniforms = rng.standard_normal(n * (2 + 2 * npts)), rng.random(n * 3 * npts)
    state = ws.SMCState(n, ess_perc_min=0.5, device=0)
    state.set_replay(normals=normals, uniforms=uniforms)
    ws.run(root, state)
    ost = ref.OracleState(n, ref.Streams(normals, uniforms), ess_perc_min=0.5)
    ref.run(om.linear_regression(xs, ys), ost)       # the oracle's own hand-built program, not the product's tree
    assert sum(e["resampled"] for e in ost.log) >= 3 and state.stats()["moves_run"] >= 6
    rec = attribute("c3_linear_regression_replay", state, ost, ("α", "β"))
    if rec["differing_particles"] == 0:
        assert abs(ws.log_evidence(state) - ref.log_evidence(ost)) <= 1e-9 * abs(ref.log_evidence(ost))
    # native RNG recovery
    npts = 60
    xs = rng.uniform(0, 10, npts)
    ys = -1.0 + 2.0 * xs + rng.standard_normal(npts)
    st2 = ws.SMCState(20_000, seed=4, device=0)
    ws.run(ws.model(LINREG)(xs, ys), st2)
    assert abs(ws.E(lambda α: α, st2) + 1.0) < 0.3 and abs(ws.E(lambda β: β, st2) - 2.0) < 0.1


SCHOOLS = '''
@model function eight_schools(J, y, σ)
    μ ~ Normal(0.0, 5.0)
    τ ~ Exponential(5.0)
    θ .= zeros(J)
    for j in 1:J
        θ[j] ~ Normal(μ, τ)
        y[j] => Normal(θ[j], σ[j])
        μ << autoRW(; diversity=0.9)
        τ << autoRW(1e-3, (0.0, Inf); diversity=0.9)
    end
end
'''


def test_c4_eight_schools_replay(ws):
    """BASELINE configs[3] at small N (examples/eight_schools.jl:7-21), replayed against the oracle."""
    n, J = 3000, 8
    y = [28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0]
    sg = [15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0]
    rng = np.random.default_rng(42)
    root = ws.model(SCHOOLS)(J, y, sg)
    normals, uniforms = rng.standard_normal(n * (1 + J + 2 * J)), rng.random(n * 3 * J)
    expon = rng.standard_exponential(n)
    state = ws.SMCState(n, ess_perc_min=0.5, device=0)
    state.set_replay(normals=normals, uniforms=uniforms, exponentials=expon)
    ws.run(root, state)
    ost = ref.OracleState(n, ref.Streams(normals, uniforms, expon), ess_perc_min=0.5)
    ref.run(om.eight_schools(J, y, sg), ost)         # the oracle's own hand-built program
    assert state["θ"].shape == (n, J)
    rec = attribute("c4_eight_schools_replay", state, ost, ("μ", "τ", "θ"))
    assert np.all(state["τ"] > 0)
    if rec["differing_particles"] == 0:
        assert abs(ws.log_evidence(state) - ref.log_evidence(ost)) <= 1e-9 * abs(ref.log_evidence(ost))


def test_wide_tape_segmented_move_and_score(ws):
    """A tape that reads more planes than one register file holds (J = 260 group effects) is folded in
    segments; the move and score_logpdf must still reproduce the oracle under replay."""
    import models
    n, J = 1500, 260
    rng = np.random.default_rng(12)
    sig = rng.uniform(9, 18, J)
    y = 4.0 + 3.0 * rng.standard_normal(J) + sig * rng.standard_normal(J)
    src = '''
    @model function wide(J, y, σ)
        μ ~ Normal(0.0, 5.0)
        τ ~ Exponential(5.0)
        θ .= zeros(J)
        for j in 1:J
            θ[j] ~ Normal(μ, τ)
            y[j] => Normal(θ[j], σ[j])
        end
        μ << RW(0.5)
        τ << autoRW(1e-3, (0.0, Inf))
        (μ, τ) << RW(0.2)
    end
    '''
    root = ws.model(src)(J, list(y), list(sig))
    normals, uniforms = rng.standard_normal(n * (1 + J + 8)), rng.random(n * (J + 8))
    expon = rng.standard_exponential(n)
    state = ws.SMCState(n, ess_perc_min=0.5, device=0)
    state.set_replay(normals=normals, uniforms=uniforms, exponentials=expon)
    ws.run(root, state)
    ost = ref.OracleState(n, ref.Streams(normals, uniforms, expon), ess_perc_min=0.5)
    ref.run(root, ost)
    assert state.stats()["moves_run"] == 3
    attribute("wide_tape_segmented_move", state, ost, ("μ", "τ", "θ"))
    # full-depth score of the trace (2 + 2J scored statements over 262 planes)
    s_dev = ws.score_logpdf(state, ["μ"], state.depth)
    ost.root = root
    s_ref = ref.score_logpdf(ost, ["μ"], ost.depth)
    ok = np.abs(state["μ"] - ost.cols["μ"]) < 1e-9
    np.testing.assert_allclose(s_dev[ok], s_ref[ok], rtol=1e-10)


def test_hierarchical_regression_benchmark_model(ws):
    """benchmarks/multilevel/WeightedSampling/model.jl verbatim: nested loops, a build-time `if j % 10 == 0`,
    `if resampled`-gated moves on a dynamic family (alpha{j}), four diversity-gated global moves, two of them
    log-transformed.  Replayed against the oracle, then the benchmark's own quality metric under Philox."""
    import models
    J, n_obs, n = 10, 3, 3000
    groups, alpha_true = models.simulate_hier(J, n_obs)
    root = ws.model(models.HIER)(J, groups)
    rng = np.random.default_rng(9)
    streams = dict(normals=rng.standard_normal(n * (2 + J + J * n_obs + 4 + 8)),
                   uniforms=rng.random(n * (2 * J * n_obs + 4 + 8)), exponentials=rng.standard_exponential(2 * n))
    st = ws.SMCState(n, device=0)
    st.set_replay(**streams)
    ws.run(root, st)
    ost = ref.OracleState(n, ref.Streams(**streams))
    ref.run(om.hierarchical_regression(J, groups), ost)      # the oracle's own hand-built program
    assert st.store.colnames() == ost.names
    rec = attribute("hierarchical_regression_replay", st, ost, tuple(ost.names))
    assert rec["differing_particles"] <= 0.01 * n
    if rec["differing_particles"] == 0:
        assert abs(ws.log_evidence(st) - ref.log_evidence(ost)) <= 1e-9 * abs(ref.log_evidence(ost))
    else:
        assert abs(ws.log_evidence(st) - ref.log_evidence(ost)) < 1e-6 * abs(ref.log_evidence(ost))
    # benchmark protocol (run_ws.jl:41-75): rmse of the posterior-mean alpha_j against the simulated truth
    J2, n2 = 20, 10
    groups2, a_true = models.simulate_hier(J2, n2)
    sp = ws.SMCState(200_000, seed=1, device=0)
    ws.run(ws.model(models.HIER)(J2, groups2), sp)
    w = ws.exp_norm(sp)
    a_est = np.array([float(np.sum(w * sp[f"alpha_{j + 1}"])) for j in range(J2)])
    rmse = float(np.sqrt(np.mean((a_est - a_true) ** 2)))
    assert rmse < 0.6, rmse                                   # posterior sd of alpha_j is ~ sigma / sqrt(10) = 0.32
    assert abs(float(np.sum(w * sp["beta"])) - 3.0) < 0.3 and abs(float(np.sum(w * sp["sigma"])) - 1.0) < 0.3


@pytest.mark.parametrize("proposal", ["RW", "autoRW"])
def test_moves_ks_against_exact_posterior_over_seeds(ws, proposal):
    """north_star: MH moves under Philox must be statistically indistinguishable from the exact target.  Particles
    start as exact draws of the Normal-Normal posterior; after 10 sweeps every particle must still be an exact
    draw (invariance): one-sample KS against the analytic posterior CDF for five seeds (Bonferroni: p > 1e-3 / 5),
    and the pooled posterior moments to 4 standard errors."""
    import scipy.stats as sst
    T, tau0, sigma, n = 6, 2.0, 1.0, 100_000
    rng = np.random.default_rng(7)
    y = rng.standard_normal(T) * sigma + 0.8
    post_var = 1.0 / (1.0 / tau0 ** 2 + T / sigma ** 2)
    post_mean = post_var * y.sum() / sigma ** 2
    post = sst.norm(post_mean, math.sqrt(post_var))
    pooled = []
    for seed in range(1, 6):
        st = ws.SMCState(n, seed=seed, device=0)
        st.store.setcol("θ", rng.standard_normal(n) * math.sqrt(post_var) + post_mean)
        st.root, st.depth = _static_model(ws, y, tau0, sigma), T + 1
        mv = ws.Move(["θ"], ws.RW, (0.4,)) if proposal == "RW" else ws.Move(["θ"], ws.autoRW, ())
        for _ in range(10):
            mv.apply(st)
        th = st["θ"]
        assert 0.2 * n < mv.last.n_accepted < 0.95 * n
        assert sst.kstest(th, post.cdf).pvalue > 1e-3 / 5, (proposal, seed)
        pooled.append(th)
    th = np.concatenate(pooled)
    se = math.sqrt(post_var / th.size)
    assert abs(th.mean() - post_mean) < 4 * se * 3          # x3: successive sweeps leave particles autocorrelated, not biased
    assert abs(th.var() - post_var) < 4 * post_var * math.sqrt(2.0 / th.size) * 3


def test_c3_philox_matches_oracle_with_numpy_streams_over_seeds(ws):
    """The whole C3 pipeline (observe / resample / autoRW moves) under the device's Philox streams against the
    CPU restatement driven by NumPy streams: posterior-mean estimates over seeds must agree within their
    seed-to-seed scatter, and both with the exact conjugate posterior mean."""
    n, npts, seeds = 4000, 25, 6
    rng = np.random.default_rng(11)
    xs = rng.uniform(0, 10, npts)
    ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(npts)
    # exact posterior mean of (α, β): prior N(0, 10² I), noise sd 1 (examples/linear_regression.jl:17-27)
    X = np.stack([np.ones(npts), xs], axis=1)
    exact = np.linalg.solve(X.T @ X + np.eye(2) / 100.0, X.T @ ys)
    dev, cpu = [], []
    for seed in range(seeds):
        st = ws.SMCState(n, ess_perc_min=0.5, seed=100 + seed, device=0)
        ws.run(ws.model(LINREG)(xs, ys), st)
        dev.append([ws.E(lambda α: α, st), ws.E(lambda β: β, st)])
        r2 = np.random.default_rng(200 + seed)
        ost = ref.OracleState(n, ref.Streams(r2.standard_normal(n * (2 + 2 * npts)), r2.random(n * 3 * npts)), ess_perc_min=0.5)
        ref.run(om.linear_regression(xs, ys), ost)
        w = ref.exp_norm(ost.weights)
        cpu.append([float(np.sum(w * ost.cols["α"])), float(np.sum(w * ost.cols["β"]))])
    dev, cpu = np.array(dev), np.array(cpu)
    for k in range(2):
        scatter = math.sqrt(dev[:, k].var(ddof=1) / seeds + cpu[:, k].var(ddof=1) / seeds)
        assert abs(dev[:, k].mean() - cpu[:, k].mean()) < 5 * scatter + 1e-3, (k, dev[:, k], cpu[:, k])
        assert abs(dev[:, k].mean() - exact[k]) < 5 * math.sqrt(dev[:, k].var(ddof=1) / seeds) + 5e-3, (k, dev[:, k], exact)
