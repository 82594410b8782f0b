"""Two forms of the same kernel must give the same bits.

* the fused elementwise window: straight-line executors (csrc/ws_vm_sl.cuh: compile-time signatures, register file
  in registers) against the interpreter (ws_vm_kernel), selected with WSB200_VM=interp at context creation;
* CDF + ancestor search: the single-pass kernel (ticketed tiles + decoupled look-back, ws_scan_search_kernel,
  WSB200_SCAN=1pass) and the chain form (persistent CTAs, look-back deferred by one tile, ws_chain_kernel,
  WSB200_SCAN=chain) against the three-pass form (WSB200_SCAN=3pass, what sharded runs use).

Both pairs share their arithmetic by construction (the same ws_vm_exec_d / integer prefix sums), so the comparison
is exact: every particle, every ancestor (only the grid-shaped (m, S, Q) reduction may differ in the last place).
"""
import os

import numpy as np
import pytest

from models import LGSSM1D, LINREG, SSM1D, SSM2D, SSM2D_FILTER

pytestmark = pytest.mark.gpu


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(ws, src, args, n, seed, ess=1.0, spec_blocks=None, **envkv):
    with env(**envkv):
        st = ws.SMCState(n, ess_perc_min=ess, seed=seed, device=0)
    old = ws.core.SPEC_BLOCKS
    if spec_blocks is not None:      # loops of "observe; Resample; if resampled" element by element (False) or in blocks (True)
        ws.core.SPEC_BLOCKS = spec_blocks
    try:
        ws.run(ws.model(src)(*args), st)
    finally:
        ws.core.SPEC_BLOCKS = old
    return st


SSM1D_FILTER = '''
@model function ssm1d_filter(obs)
    x .= 0.0
    v .= 0.0
    for o in obs
        x .= x + v
        dv ~ Normal(0.0, 0.1)
        v .= v + dv
        o => Normal(x, 1.0)
    end
end
'''

CASES = [
    ("ssm2d", SSM2D_FILTER, lambda rng: ([rng.standard_normal(2) + np.array([t, 0.0]) for t in range(12)],), ["x", "v", "dv"]),
    ("ssm2d_hist", SSM2D, lambda rng: ([rng.standard_normal(2) + np.array([t, 0.0]) for t in range(6)],), ["x_7", "x_3", "v"]),
    ("lgssm1d", LGSSM1D, lambda rng: (list(rng.standard_normal(15)), 0.9, 1.0, 0.5, 1.0), ["x"]),
    ("ssm1d", SSM1D_FILTER, lambda rng: (list(rng.standard_normal(10)),), ["x", "v", "dv"]),
    ("ssm1d_hist", SSM1D, lambda rng: (list(rng.standard_normal(8)),), ["x_9", "v"]),
    ("linreg", LINREG, lambda rng: (list(rng.uniform(0, 10, 30)), list(1 - 0.5 * rng.uniform(0, 10, 30))), ["α", "β"]),
]


@pytest.mark.parametrize("name,src,mk,cols", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("n", [1000, 100_003])
@pytest.mark.parametrize("ess", [0.5, 1.0])
def test_straight_line_equals_interpreter(ws, name, src, mk, cols, n, ess):
    args = mk(np.random.default_rng(3))
    a = _run(ws, src, args, n, seed=21, ess=ess, spec_blocks=False)
    b = _run(ws, src, args, n, seed=21, ess=ess, spec_blocks=False, WSB200_VM="interp")
    sa, sb = a.stats(), b.stats()
    assert sb["sl_passes"] == 0
    if name != "ssm1d_hist":   # (its window is not in the signature table: x{t+1} is a new plane and v is stored too)
        assert sa["sl_passes"] >= sa["fused_passes"] - 2, (sa["sl_passes"], sa["fused_passes"])
    for c in cols:
        np.testing.assert_array_equal(a[c], b[c], err_msg=f"{name}: column {c}")
    # per-particle arithmetic is shared, so the columns agree to the last bit; the (m, S, Q) reduction is combined
    # per CTA and the two kernels use different grids, so log-sum-exp (hence the weights after a resampling step,
    # all equal to logsumexp - log N) may differ in the last place
    np.testing.assert_allclose(a.weights, b.weights, rtol=1e-13, atol=1e-13)
    assert abs(ws.log_evidence(a) - ws.log_evidence(b)) <= 1e-13 * abs(ws.log_evidence(b))
    assert sa["resamples_done"] == sb["resamples_done"]


SCAN_FORMS = ({"WSB200_SCAN": "3pass"}, {"WSB200_SCAN": "1pass"}, {"WSB200_SCAN": "chain"})


@pytest.mark.parametrize("n", [1, 2, 255, 2048, 2049, 100_003, 3_000_000])
@pytest.mark.parametrize("s", [0.5, 3.0])
@pytest.mark.parametrize("scheme", ["stratified", "systematic"])
def test_single_pass_scan_search_equals_three_pass(ws, n, s, scheme):
    w = np.exp(s * np.random.default_rng(n).standard_normal(n))
    w /= w.sum()
    out = []
    for kv in SCAN_FORMS:
        with env(**kv):
            st = ws.SMCState(max(n, 2), seed=5, device=0)
        out.append(ws.resample_indices(w, st, scheme))
        r = np.random.default_rng(1).random(n if scheme == "stratified" else 1)
        out.append(ws.resample_indices(w, st, scheme, uniforms=r))
    for f in range(1, len(SCAN_FORMS)):
        np.testing.assert_array_equal(out[0], out[2 * f], err_msg=str(SCAN_FORMS[f]))       # Philox (integer grid)
        np.testing.assert_array_equal(out[1], out[2 * f + 1], err_msg=str(SCAN_FORMS[f]))   # replayed uniforms (reference's floating-point grid)


@pytest.mark.parametrize("n", [5000, 100_003, 2_000_000])
@pytest.mark.parametrize("ess", [0.5, 1.0])
@pytest.mark.parametrize("form", ["1pass", "chain"])
def test_scan_forms_inside_a_filter(ws, n, ess, form):
    """the whole filter (log-weights -> exp_norm weights -> fixed point inside the scan kernels, mode 0), every column"""
    args = ([np.random.default_rng(2).standard_normal(2) + np.array([t, 0.0]) for t in range(10)],)
    a = _run(ws, SSM2D_FILTER, args, n, seed=17, ess=ess, WSB200_SCAN="3pass", WSB200_SMALL_RESAMPLE="0")
    b = _run(ws, SSM2D_FILTER, args, n, seed=17, ess=ess, WSB200_SCAN=form, WSB200_SMALL_RESAMPLE="0")
    for c in ("x", "v"):
        np.testing.assert_array_equal(a[c], b[c], err_msg=c)
    np.testing.assert_array_equal(a.weights, b.weights)
    assert ws.log_evidence(a) == ws.log_evidence(b)
    assert a.stats()["resamples_done"] == b.stats()["resamples_done"] > 0


def test_single_pass_one_hot_and_zero_weights(ws):
    n = 300_000
    w = np.zeros(n)
    w[123_456] = 0.97
    w[::7] += 0.03 / len(w[::7])
    w /= w.sum()
    res = []
    for kv in SCAN_FORMS:
        with env(**kv):
            st = ws.SMCState(n, seed=9, device=0)
        res.append(ws.resample_indices(w, st, "stratified"))
    for r in res[1:]:
        np.testing.assert_array_equal(res[0], r)
    assert (res[0] == 123_456).sum() > 0.96 * n
    assert np.all(w[res[0]] > 0)


@pytest.mark.parametrize("n,extra", [(50_001, 0), (50_001, 14), (4096, 20), (100_000, 16)])
def test_integer_slot_grid_small_shift(ws, n, extra):
    """N > 2^29 particles (eight GPUs) leave fewer than 32 fractional bits per slot and ws_slot_split shifts left;
    WSB200_FX_EXTRA_BITS shrinks the scale so that this arithmetic runs at test sizes.  Against the big-integer
    specification, every ancestor, both kernel forms."""
    import ctypes as C
    from oracle import ref
    w = np.exp(2.0 * np.random.default_rng(n + extra).standard_normal(n))
    w /= w.sum()
    for kv in SCAN_FORMS:
        with env(WSB200_FX_EXTRA_BITS=str(extra), **kv):
            st = ws.SMCState(n, seed=13, device=0)
        stream, seed = C.c_uint64(), C.c_uint64()
        st.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
        for scheme in ("stratified", "systematic"):
            a, clamped = ws.resample_indices(w, st, scheme, return_clamped=True)
            a_ref, clamped_ref = ref.stratified_ancestors_fixed_point(w, seed.value, stream.value, scheme, extra_bits=extra)
            np.testing.assert_array_equal(a, a_ref)
            assert clamped == clamped_ref
            stream.value += 1
    with env(WSB200_FX_EXTRA_BITS="0"):
        ws.SMCState(2, device=0)      # back to the default scale for the tests that follow


# ------------------------------------------------------------------------------------------------
# Resample.apply! queued without waiting for its outcome (ws_resample_async) == the synchronous state machine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,src,mk,cols", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ess", [0.3, 0.5, 1.0])
def test_async_resample_equals_sync(ws, name, src, mk, cols, ess):
    """transformers.jl:474-498 with the decision left on the device: gated scan / search, identity ancestors when
    the step does not fire, log-weight base chosen by the device flag.  Every particle, the weights, the evidence and
    the counters must equal the run in which the host waits for every decision (WSB200_ASYNC_RESAMPLE=0)."""
    args = mk(np.random.default_rng(5))
    n = 20_011
    a = _run(ws, src, args, n, seed=33, ess=ess, spec_blocks=False)
    b = _run(ws, src, args, n, seed=33, ess=ess, spec_blocks=False, WSB200_ASYNC_RESAMPLE="0")
    for c in cols:
        np.testing.assert_array_equal(a[c], b[c], err_msg=f"{name}: column {c}")
    np.testing.assert_array_equal(a.weights, b.weights)
    assert ws.log_evidence(a) == ws.log_evidence(b)
    sa, sb = a.stats(), b.stats()
    for k in ("resamples_fired", "resamples_done", "moves_run", "fused_passes"):
        assert sa[k] == sb[k], k
    assert a.resampled == b.resampled
    if ess == 0.3 and name != "linreg":
        assert sa["resamples_done"] < sa["resamples_fired"]        # some steps did not fire: the identity path ran


def test_async_resample_state_machine(ws):
    """the flags and `last` of a queued step; a no-op step leaves `resampled` untouched; untouched planes of a
    step that did not fire stay correct (they are read through identity ancestors)."""
    n = 5000
    st = ws.SMCState(n, ess_perc_min=0.5, seed=2, device=0)
    st.store.setcol("x", np.arange(n, dtype=float))
    st.store.setcol("keep", np.arange(n, dtype=float) * 2.0)
    r = ws.Resample()
    st.resampled = True
    r.apply(st)
    assert r.last.fired == 0 and st.resampled is True
    ws.Observe(0.0, "Normal", (ws.col("x") * 1e-9, 1.0)).apply(st)
    r.apply(st)                                   # queued; ESS is ~1: does not fire
    ws.Observe(0.0, "Normal", (ws.col("x") * 1e-9, 1.0)).apply(st)   # next weighting pass runs behind the pending step
    r.apply(st)
    assert r.last.fired == 1 and r.last.resampled == 0 and r.last.ess_perc > 0.99
    assert st.resampled is False and st.stats()["resamples_done"] == 0 and st.stats()["resamples_fired"] == 2
    np.testing.assert_array_equal(st["keep"], np.arange(n) * 2.0)
    lw_expected = 2 * (-0.5 * (np.arange(n) * 1e-9) ** 2 - 0.5 * np.log(2 * np.pi))
    np.testing.assert_allclose(st.weights, lw_expected, rtol=1e-12)
    ws.Observe(0.0, "Normal", (ws.col("x"), 40.0)).apply(st)
    le = ws.log_evidence(st)
    r.apply(st)                                   # fires
    ws.Assign("y", ws.col("x") + 1.0).apply(st)   # a pass without a weight term behind a pending step
    assert st.resampled is True and r.last.resampled == 1 and st.stats()["resamples_done"] == 1
    np.testing.assert_allclose(st.weights, np.full(n, le), rtol=1e-12)
    x = st["x"]
    assert np.all(np.diff(x) >= 0) and len(np.unique(x)) < n
    np.testing.assert_array_equal(st["keep"], 2.0 * x)
    np.testing.assert_array_equal(st["y"], x + 1.0)


# ------------------------------------------------------------------------------------------------
# multinomial resampling without a sort (exponential spacings regenerated blockwise from Philox)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,s", [(1, 1.0), (2, 1.0), (255, 0.5), (256, 2.0), (257, 2.0), (2048, 1.0), (2049, 3.0), (50_001, 0.5),
                                 (50_001, 3.0), (300_000, 1.5)])
def test_multinomial_spacings_match_the_big_integer_specification(ws, n, s):
    """every ancestor against oracle/ref.py: multinomial_ancestors_fixed_point, with the spacings regenerated by the
    host instantiation of the device routine (csrc/ws_math.cuh: ws_spacing_of_slot)"""
    import ctypes as C
    from hostlib import lib
    from oracle import ref
    w = np.exp(s * np.random.default_rng(n).standard_normal(n))
    w /= w.sum()
    st = ws.SMCState(max(n, 2), seed=19, device=0)
    stream, seed = C.c_uint64(), C.c_uint64()
    st.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
    a, clamped = ws.resample_indices(w, st, "multinomial", return_clamped=True)
    e = np.empty(n + 1, dtype=np.uint64)
    lib().hh_spacings(n + 1, ref.multinomial_mn_shift(n), seed.value, stream.value, e.ctypes.data_as(C.c_void_p))
    a_ref, clamped_ref = ref.multinomial_ancestors_fixed_point(w, e)
    np.testing.assert_array_equal(a, a_ref)
    assert clamped == clamped_ref


def test_multinomial_one_hot_and_statistics(ws):
    """a heavy particle (its slots span many blocks: the per-lane fallback and the heavy-tile expansion) and the
    multinomial variance of the offspring counts"""
    import ctypes as C
    from hostlib import lib
    from oracle import ref
    n = 120_000
    w = np.full(n, 0.03 / (n - 1))
    w[77_777] = 0.97
    w /= w.sum()
    st = ws.SMCState(n, seed=23, device=0)
    stream, seed = C.c_uint64(), C.c_uint64()
    st.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
    a = ws.resample_indices(w, st, "multinomial")
    e = np.empty(n + 1, dtype=np.uint64)
    lib().hh_spacings(n + 1, ref.multinomial_mn_shift(n), seed.value, stream.value, e.ctypes.data_as(C.c_void_p))
    a_ref, _ = ref.multinomial_ancestors_fixed_point(w, e)
    np.testing.assert_array_equal(a, a_ref)
    assert (a == 77_777).sum() > 0.96 * n
    # statistics over fresh draws: E[count] = N w, Var[count] = N w (1 - w) (multinomial), unlike stratified (< 1/4)
    n2 = 200_000
    w2 = np.exp(0.7 * np.random.default_rng(4).standard_normal(n2))
    w2 /= w2.sum()
    st2 = ws.SMCState(n2, seed=5, device=0)
    counts = np.bincount(ws.resample_indices(w2, st2, "multinomial"), minlength=n2)
    assert counts.sum() == n2
    resid = counts - n2 * w2
    assert abs(resid.mean()) < 1e-9 and 0.9 < resid.var() / np.mean(n2 * w2 * (1 - w2)) < 1.1
    again = ws.resample_indices(w2, st2, "multinomial")
    assert np.any(np.bincount(again, minlength=n2) != counts)


def test_ess_knife_edge_is_counted(ws):
    """exactly equal weights with ess_perc_min = 1.0: the reference's 1/(N sum w^2) lands on either side of 1.0
    depending on N; the device's S^2/(N Q) is exactly 1.0 (no resample).  Such steps are counted, not hidden."""
    n = 2000
    st = ws.SMCState(n, ess_perc_min=1.0, seed=1, device=0)
    st.store.setcol("x", np.zeros(n))
    ws.Observe(0.5, "Normal", (ws.col("x"), 1.0)).apply(st)     # identical log-weights
    ws.Resample().apply(st)
    assert st.ess_ties() == 1 and st.stats()["resamples_done"] == 0
    st.store.setcol("x", np.arange(n) * 1e-3)
    ws.Observe(0.5, "Normal", (ws.col("x"), 1.0)).apply(st)
    ws.Resample().apply(st)
    assert st.ess_ties() == 1 and st.stats()["resamples_done"] == 1


# ------------------------------------------------------------------------------------------------
# loop bodies described once and replayed per element (ws_exec) == the per-element rebuild
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,src,mk,cols", [c for c in CASES if c[0] in ("ssm2d", "lgssm1d", "ssm1d")], ids=["ssm2d", "lgssm1d", "ssm1d"])
@pytest.mark.parametrize("ess", [0.5, 1.0])
def test_loop_template_equals_per_element_rebuild(ws, name, src, mk, cols, ess):
    import subprocess
    import sys
    args = mk(np.random.default_rng(8))
    n = 30_011
    a = _run(ws, src, args, n, seed=44, ess=ess)
    # the switch is read when the package is imported: run the reference configuration in a fresh interpreter
    code = f"""
import sys, pickle, numpy as np
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r}); sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
import wsb200 as ws
from wsb200 import core
assert core.LOOP_TEMPLATES is False
import test_gpu_kernel_forms as t
case = [c for c in t.CASES if c[0] == {name!r}][0]
st = t._run(ws, case[1], case[2](np.random.default_rng(8)), {n}, seed=44, ess={ess})
pickle.dump(({{c: st[c] for c in case[3]}}, st.weights, ws.log_evidence(st), st.stats()["fused_passes"]), sys.stdout.buffer)
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, env=dict(os.environ, WSB200_LOOP_TEMPLATE="0"), check=True)
    import pickle
    cols_b, w_b, le_b, passes_b = pickle.loads(out.stdout)
    for c in cols:
        np.testing.assert_array_equal(a[c], cols_b[c], err_msg=f"{name}: column {c}")
    np.testing.assert_array_equal(a.weights, w_b)
    assert ws.log_evidence(a) == le_b and a.stats()["fused_passes"] == passes_b


# ------------------------------------------------------------------------------------------------
# small particle sets: the Resample step as ONE kernel == the multi-kernel resampler
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,src,mk,cols", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("n", [1, 2, 300, 2049, 4096])
@pytest.mark.parametrize("ess", [0.5, 1.0])
def test_single_kernel_resample_equals_multi_kernel(ws, name, src, mk, cols, n, ess):
    args = mk(np.random.default_rng(6))
    a = _run(ws, src, args, n, seed=55, ess=ess)
    b = _run(ws, src, args, n, seed=55, ess=ess, WSB200_SMALL_RESAMPLE="0")
    for c in cols:
        np.testing.assert_array_equal(a[c], b[c], err_msg=f"{name}: column {c}")
    np.testing.assert_array_equal(a.weights, b.weights)
    assert ws.log_evidence(a) == ws.log_evidence(b)
    assert a.stats()["resamples_done"] == b.stats()["resamples_done"]
    assert a.stats()["kernel_launches"] < b.stats()["kernel_launches"] or a.stats()["resamples_fired"] == 0


@pytest.mark.parametrize("scheme", ["stratified", "systematic"])
def test_single_kernel_resample_one_hot_and_spec(ws, scheme):
    """ancestors of the one-kernel step against the big-integer specification, inside a run (ids through the gather)"""
    import ctypes as C
    from oracle import ref
    n = 4000
    for s in (0.5, 3.0, None):
        st = ws.SMCState(n, ess_perc_min=float("inf"), seed=3, resampler=scheme, device=0)
        st.store.setcol("id", np.arange(n, dtype=np.float64))
        lw = s * np.random.default_rng(2).standard_normal(n) if s is not None else np.where(np.arange(n) == 1234, 0.0, -12.0)
        st.weights = lw
        st.weights_changed = True
        stream, seed = C.c_uint64(), C.c_uint64()
        st.store._call("ws_next_philox_stream", C.byref(stream), C.byref(seed))
        ws.Resample().apply(st)
        ids = st["id"].astype(np.int64)
        a_ref, _ = ref.stratified_ancestors_fixed_point(ref.exp_norm(lw), seed.value, stream.value, scheme)
        np.testing.assert_array_equal(ids, a_ref)


# ------------------------------------------------------------------------------------------------
# Loop steps run in speculative blocks (ws_exec_spec: K observations in one pass, the ESS after each from the
# pass's checkpoints, roll-back to the first step that resamples) == the same steps run one by one
# ------------------------------------------------------------------------------------------------
OBS_ONLY = '''
@model function obs_only(xs, ys)
    α ~ Normal(0.0, 10.0)
    β ~ Normal(0.0, 10.0)
    for (x, y) in zip(xs, ys)
        y => Normal(α + β * x, 1.0)
        if resampled
            α .= α + 0.0
        end
    end
end
'''


@pytest.mark.parametrize("vm", ["sl", "interp", "interp16"])   # register-resident block kernel / the interpreter's checkpoints (8 or 16 steps per block)
@pytest.mark.parametrize("src", ["linreg", "obs_only"])
@pytest.mark.parametrize("n", [20_011, 300_000])
@pytest.mark.parametrize("ess", [0.5, 0.9, 1.0])     # 1.0: every step resamples, the blocks shrink to single steps
def test_speculative_blocks_equal_stepwise(ws, src, n, ess, vm):
    rng = np.random.default_rng(5)
    xs = rng.uniform(0, 10, 150)
    ys = 1 - 0.5 * xs + rng.standard_normal(150)
    kv = {"WSB200_VM": "interp"} if vm != "sl" else {}
    steps = ws.core.SPEC_BLOCK_STEPS
    ws.core.SPEC_BLOCK_STEPS = 16 if vm == "interp16" else steps
    try:
        a, b = [_run(ws, LINREG if src == "linreg" else OBS_ONLY, (list(xs), list(ys)), n, seed=31, ess=ess, spec_blocks=spec, **kv)
                for spec in (False, True)]
    finally:
        ws.core.SPEC_BLOCK_STEPS = steps
    sa, sb = a.stats(), b.stats()
    assert sa["resamples_done"] == sb["resamples_done"] > 0
    assert sa["resamples_fired"] == sb["resamples_fired"] >= 150
    assert sa["moves_run"] == sb["moves_run"]
    if ess < 1.0:
        assert sb["fused_passes"] < sa["fused_passes"], (sa["fused_passes"], sb["fused_passes"])
    else:
        assert sb["fused_passes"] < sa["fused_passes"] + 16, (sa["fused_passes"], sb["fused_passes"])   # only the first blocks speculate
    if vm == "sl":
        assert sb["sl_passes"] > 0
    # same association of the log-weight sums and the same Philox stream numbering: only the grouping of the
    # (m, S, Q) partials differs (different kernel shapes), i.e. log-sum-exp in the last place
    for c in ("α", "β"):
        np.testing.assert_allclose(a[c], b[c], rtol=1e-9, atol=1e-12, err_msg=c)
    np.testing.assert_allclose(a.weights, b.weights, rtol=1e-10, atol=1e-10)
    assert abs(ws.log_evidence(a) - ws.log_evidence(b)) <= 1e-11 * abs(ws.log_evidence(b))
    assert a.depth == b.depth
