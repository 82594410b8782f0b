"""Port of the reference's test/macro_test.jl (its testsets, in order, with its model sources, sizes and tolerances):
`@model` output driven with USER kernels (`kernels=normal_kernels`: `NormalKernel` and the weighter-only
`NormalWeightKernel` of test/models.jl:16-34, here as device-expression `WeightedKernel`s), `Observe`, `Weight`
(`_ ~ NormalWeight(x, r, y)`), `Resample`, `Cond` / `if resampled`, against the exact Kalman filter and exact marginals.
The draws are Philox, not Julia's stream, so — as in the reference — these are statistical checks with the reference's
own tolerances (macro_test.jl:67-213)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOG_2PI = math.log(2.0 * math.pi)


def _normal_logpdf(ws):
    return lambda mu, s, x: -0.5 * ((x - mu) / s) ** 2 - ws.log(s) - 0.5 * LOG_2PI


def _normal_kernels(ws):
    """test/models.jl:16-34 — NormalKernel(sampler, nothing, logpdf); NormalWeightKernel(nothing, weighter, logpdf)."""
    lp = _normal_logpdf(ws)
    return {"Normal": ws.WeightedKernel(lambda mu, s: mu + s * ws.randn(), None, lp, "Normal"),
            "NormalWeight": ws.WeightedKernel(None, lp, lp, "NormalWeight")}


RANDOM_WALK1 = '''
@model function random_walk1(T::Int)
    x ~ Normal(0, 1)
    for t in 1:T
        x ~ Normal(x, 1)
    end
end
'''  # macro_test.jl:12-17

SSM_FILTER = '''
@model function ssm_filter(data, a, q, r)
    x ~ Normal(0, 1)
    for y in data
        x ~ Normal(a * x, q)
        y => Normal(x, r)
    end
end
'''  # macro_test.jl:19-25 (and :53-59, ssm_filter_resampled: the same body)

SSM_FILTER_WEIGHT = '''
@model function ssm_filter_weight(data, a, q, r)
    x ~ Normal(0, 1)
    for y in data
        x ~ Normal(a * x, q)
        _ ~ NormalWeight(x, r, y)
    end
end
'''  # macro_test.jl:27-33

SSM_FILTER_COND = '''
@model function ssm_filter_cond(data, a, q, r)
    x ~ Normal(0, 1)
    for y in data
        x ~ Normal(a * x, q)
        y => Normal(x, r)
        if resampled
            x .= x
        end
    end
end
'''  # macro_test.jl:41-51


def _generate_ssm_data(rng, T, a, q, r):   # macro_test.jl:86-95
    x_prev, data = rng.standard_normal(), []
    for _ in range(T):
        x = a * x_prev + q * rng.standard_normal()
        data.append(float(x + r * rng.standard_normal()))
        x_prev = x
    return data


def _kalman(data, a, q, r, x0_std=1.0):    # test/models.jl kalman_filter_evidence (x(0) ~ Normal(0, 1))
    mu, P, le = 0.0, x0_std ** 2, 0.0
    for y in data:
        mu_p, P_p = a * mu, a * a * P + q * q
        S = P_p + r * r
        le += -0.5 * (LOG_2PI + math.log(S) + (y - mu_p) ** 2 / S)
        K = P_p / S
        mu, P = mu_p + K * (y - mu_p), (1 - K) * P_p
    return mu, le


def test_macro_random_walk_k1(ws):
    """macro_test.jl:67-82: x(T) ~ Normal(0, sqrt(T + 1))."""
    T, n = 10, 100_000
    st = ws.SMCState(n, seed=42, device=0)
    ws.run(ws.model(RANDOM_WALK1)(T, kernels=_normal_kernels(ws)), st)
    xs = st["x"]
    assert abs(xs.mean()) < 0.15 and abs(xs.var() - (T + 1)) < 0.05 * (T + 1)


def _ssm_macro_correctness(ws, src, T, n, ess, max_abs_diff, mean_atol, seed=42):
    a, q, r = 0.8, 0.5, 0.5
    data = _generate_ssm_data(np.random.default_rng(seed), T, a, q, r)
    exact_mean, exact_evidence = _kalman(data, a, q, r)
    st = ws.SMCState(n, seed=seed, device=0, ess_perc_min=ess)
    ws.run(ws.model(src)(data, a, q, r, kernels=_normal_kernels(ws)), st)
    evidence = ws.log_evidence(st)
    est_mean = float(np.sum(ws.exp_norm(st) * st["x"]))
    assert abs(evidence - exact_evidence) < max_abs_diff and abs(est_mean - exact_mean) < mean_atol, (evidence, exact_evidence, est_mean, exact_mean)
    return st, evidence, est_mean


def test_macro_observe_against_exact_kalman_filter(ws):
    """macro_test.jl:97-117 (T = 5, N = 200 000; no resampling fires at the reference's default ess_perc_min = 0.5 ... or
    does: either way the estimate is unbiased)."""
    _ssm_macro_correctness(ws, SSM_FILTER, 5, 200_000, 0.5, 0.5, 0.3)


def test_macro_weight_against_exact_kalman_filter(ws):
    """macro_test.jl:119-121: `_ ~ NormalWeight(x, r, y)` adds the same factor as `y => Normal(x, r)`."""
    st_w, ev_w, mean_w = _ssm_macro_correctness(ws, SSM_FILTER_WEIGHT, 5, 200_000, 0.5, 0.5, 0.3)
    st_o, ev_o, mean_o = _ssm_macro_correctness(ws, SSM_FILTER, 5, 200_000, 0.5, 0.5, 0.3)
    # same seed, same statements in the same order: Weight and Observe are the same computation
    assert abs(ev_w - ev_o) <= 1e-9 * abs(ev_o) and abs(mean_w - mean_o) <= 1e-9 * (1 + abs(mean_o))


def test_macro_resample_against_exact_kalman_filter_t50(ws):
    """macro_test.jl:123-145 (T = 50, N = 10 000, ess_perc_min = 0.5)."""
    st, _, _ = _ssm_macro_correctness(ws, SSM_FILTER, 50, 10_000, 0.5, 3.0, 1.0)
    assert st.stats()["resamples_done"] >= 1


def test_macro_cond_noop_body_matches_resample_only_model(ws):
    """macro_test.jl:147-149: `if resampled; x .= x; end` does not change the model's distribution and plumbs
    `state.resampled` through."""
    st, _, _ = _ssm_macro_correctness(ws, SSM_FILTER_COND, 50, 10_000, 0.5, 3.0, 1.0)
    assert st.stats()["resamples_done"] >= 1


def test_cond_transformer_direct(ws):
    """macro_test.jl:157-176: Cond(predfn, body) runs body iff predfn(state)."""
    st = ws.SMCState(10, device=0)
    ws.Assign("x", lambda s: 0.0).apply(st)
    body = ws.Assign("x", lambda s: 1.0)
    ws.Cond(lambda s: False, body).apply(st)
    assert np.all(st["x"] == 0.0)
    ws.Cond(lambda s: True, body).apply(st)
    assert np.all(st["x"] == 1.0)


FIRE_ALARM = '''
@model function fire_alarm_macro()
    fire ~ Bernoulli(0.01)
    smoke ~ Bernoulli(fire ? 0.9 : 0.01)
    lever ~ Bernoulli(fire ? 0.7 : 0.01)
    alarm ~ Bernoulli(smoke || lever ? 0.98 : 0.01)
end
'''  # macro_test.jl:182-187


def test_macro_vectorized_ternary_and_short_circuit_or(ws):
    """macro_test.jl:189-213: forward-sampled marginals against the exact ones (a scalar collapse of `? :` / `||`
    would give every particle the same value)."""
    n, atol = 200_000, 0.004
    st = ws.SMCState(n, seed=42, device=0)
    ws.run(ws.model(FIRE_ALARM)(), st)
    exact_smoke = 0.01 * 0.9 + 0.99 * 0.01
    exact_lever = 0.01 * 0.7 + 0.99 * 0.01
    p_or = 0.01 * (1 - 0.1 * 0.3) + 0.99 * (1 - 0.99 * 0.99)
    exact_alarm = p_or * 0.98 + (1 - p_or) * 0.01
    assert abs(st["smoke"].mean() - exact_smoke) < atol
    assert abs(st["lever"].mean() - exact_lever) < atol
    assert abs(st["alarm"].mean() - exact_alarm) < atol


# ---- test/move_macro_test.jl ------------------------------------------------------------------------------------------
LINREG_IF_RESAMPLED = '''
@model function linear_regression(data)
    α ~ Normal(0.0, 5.0)
    β ~ Normal(0.0, 5.0)
    for (x, y) in data
        y => Normal(α + β * x, 0.5)
        if resampled
            (α, β) << RW(0.1)
        end
    end
end
'''  # move_macro_test.jl:40-50

LINREG_DIVERSITY = '''
@model function linear_regression(data)
    α ~ Normal(0.0, 5.0)
    β ~ Normal(0.0, 5.0)
    for (x, y) in data
        y => Normal(α + β * x, 0.5)
        (α, β) << RW(0.1; diversity=0.9)
    end
end
'''  # move_macro_test.jl:92-100


@pytest.mark.parametrize("src", [LINREG_IF_RESAMPLED, LINREG_DIVERSITY], ids=["if_resampled", "diversity_kwarg"])
def test_macro_linear_regression_with_mh_moves(ws, src):
    """move_macro_test.jl:27-66 / 82-116: `<<` end to end — a static-parameter RW move gated by `if resampled`, and the
    same move self-gated by the macro-level `diversity=` keyword; 10 points, N = 10 000, posterior means within 0.3 of
    the true (α, β) = (-1, 2).  Run with run (run!), user kernels and proposals handed in by name."""
    true_a, true_b, noise, n_points = -1.0, 2.0, 0.5, 10
    rng = np.random.default_rng(42)
    xs = np.linspace(0.0, 10.0, n_points)
    ys = true_a + true_b * xs + noise * rng.standard_normal(n_points)
    data = list(zip(xs.tolist(), ys.tolist()))
    st = ws.SMCState(10_000, seed=42, device=0)
    ws.run(ws.model(src)(data, kernels={"Normal": _normal_kernels(ws)["Normal"]}, proposals={"RW": ws.RW}), st)
    w = ws.exp_norm(st)
    est_a, est_b = float(np.sum(st["α"] * w)), float(np.sum(st["β"] * w))
    assert abs(est_a - true_a) < 0.3 and abs(est_b - true_b) < 0.3, (est_a, est_b)
    assert st.stats()["moves_run"] >= 1
