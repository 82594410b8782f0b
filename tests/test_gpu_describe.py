"""describe(state) (src/utils.jl:183-289) on the device against the restated StatsBase statistics."""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.gpu


def _check(ws, st, name, values, logw, rel=1e-9):
    w = ref.exp_norm(logw)
    want = ref.describe_column(values, w)
    df = ws.describe(st, cols=[name])
    row = df.iloc[0]
    for f in ("mean", "median", "std", "min", "max"):
        assert abs(row[f] - want[f]) <= rel * (1 + abs(want[f])), (f, row[f], want[f])
    assert abs(row["ess"] - st.store.n * ref.ess_perc(w)) <= 1e-6 * st.store.n
    return row, want


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 100_003])
@pytest.mark.parametrize("kind", ["normal", "duplicates", "uniform_weights"])
def test_describe_matches_statsbase_restatement(ws, n, kind):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    if kind == "duplicates":
        x = np.round(x, 1)                           # many ties (resampled particles look like this)
    lw = np.zeros(n) if kind == "uniform_weights" else 1.5 * rng.normal(size=n)
    st = ws.SMCState(n, device=0)
    st.store.setcol("x", x)
    if kind != "uniform_weights":
        st.weights = lw
    row, want = _check(ws, st, "x", x, lw)
    assert len(row["hist"]) == 8


def test_describe_edge_cases(ws):
    n = 5000
    rng = np.random.default_rng(1)
    st = ws.SMCState(n, device=0)
    x = rng.normal(size=n)
    st.store.setcol("x", x)
    st.store.setcol("c", np.full(n, 2.5))                        # constant column
    st.store.setcol("v", rng.normal(size=(n, 3)))                # vector column: component-wise, no histogram
    xn = x.copy()
    xn[7] = np.nan
    st.store.setcol("bad", xn)
    lw = rng.normal(size=n)
    lw[:100] = -np.inf                                           # zero weights are ignored by the median
    lw[200] = 12.0                                               # one dominant particle
    st.weights = lw
    _check(ws, st, "x", x, lw)
    df = ws.describe(st)
    assert list(df["variable"]) == ["x", "c", "v", "bad"]
    c = df.iloc[1]
    assert abs(c["mean"] - 2.5) < 1e-12 and c["median"] == 2.5 and c["std"] < 1e-12 and c["min"] == c["max"] == 2.5
    v = df.iloc[2]
    assert v["hist"] == "" and v["mean"].shape == (3,)
    w = ref.exp_norm(lw)
    vv = st["v"]
    for k in range(3):
        want = ref.describe_column(vv[:, k], w)
        assert abs(v["median"][k] - want["median"]) < 1e-9 and abs(v["std"][k] - want["std"]) < 1e-9
    b = df.iloc[3]
    assert np.isnan(b["mean"]) and np.isnan(b["median"]) and np.isnan(b["std"])
    with pytest.raises(ValueError):
        ws.describe(st, cols=["nope"])


def test_describe_after_a_run_with_history(ws):
    """columns several resampling events behind are read through the genealogy"""
    import models
    n, T = 20_000, 12
    rng = np.random.default_rng(5)
    obs = list(rng.normal(size=T))
    normals, uniforms = rng.standard_normal(n * T), rng.random(n * T)
    st = ws.SMCState(n, ess_perc_min=1.0, device=0)
    st.set_replay(normals=normals, uniforms=uniforms)
    root = ws.model(models.SSM1D)(obs)
    ws.run(root, st)
    df = ws.describe(st)
    lw = st.weights
    for name in ("x_2", "x_7", f"x_{T + 1}", "v"):
        want = ref.describe_column(st[name], ref.exp_norm(lw))
        row = df[df["variable"] == name].iloc[0]
        for f in ("mean", "median", "std", "min", "max"):
            assert abs(row[f] - want[f]) <= 1e-9 * (1 + abs(want[f])), (name, f)
