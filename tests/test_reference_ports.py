"""Ports of reference tests that were still missing: test/dynamic_vars_test.jl:47-70,141-153
(macro-expansion errors of dynamic-variable families), test/dynamic_move_test.jl:15-75 (`<<` on a dynamic family
member, accessor targets rejected), test/accessors_test.jl:23-60 (`[]` reads / writes on an array-valued column; the
struct-valued `.` cases are outside the device-op set and must be REJECTED, not run), test/default_kernels_test.jl:14-64
(`~` / `<<` resolve the default tables with no kernels= / proposals= argument)."""
import numpy as np
import pytest

import wsb200 as ws
from wsb200.model import ModelSyntaxError


def _model(body, args=""):
    return ws.model(f"@model function f({args})\n{body}\nend")


# ---- test/dynamic_vars_test.jl:47-70 "Dynamic-variable error paths" ------------------------------------------------
@pytest.mark.parametrize("body", [
    "y ~ Normal(0, 1)\n x{y} ~ Normal(0, 1)",          # bad_dyn_idx: the index must not depend on a particle variable
    "x ~ Normal(0, 1)\n x{1} ~ Normal(0, 1)",          # bad_collision1: plain variable, then family
    "x{1} ~ Normal(0, 1)\n x ~ Normal(0, 1)",          # bad_collision2: family, then plain variable
    "y .= x{1} + 1",                                   # bad_unregistered: family read before it was ever assigned
])
def test_dynamic_variable_error_paths(body):
    with pytest.raises(ModelSyntaxError):
        _model(body)


# ---- test/dynamic_vars_test.jl:141-153 "Chained dynamic-variable accessor error paths" --------------------------------
@pytest.mark.parametrize("body", [
    "v{1}[1] .= 1.0",                                                    # bad_dyn_chain_unregistered
    "y ~ Normal(0, 1)\n v{1} .= [1.0, 2.0]\n v{y}[1] .= 1.0",            # bad_dyn_chain_idx: index purity inside a chain
])
def test_chained_dynamic_accessor_error_paths(body):
    with pytest.raises(ModelSyntaxError):
        _model(body)


# ---- test/dynamic_move_test.jl:56-75 "<< Move rejects value-level accessor targets" -----------------------------------
@pytest.mark.parametrize("body", [
    "v .= [1.0, 2.0]\n v[1] << RW(0.1)",                                 # bad_move_ref
    "p .= 1.0\n p.x << RW(0.1)",                                         # bad_move_prop
    "a ~ Normal(0, 1)\n v .= [1.0, 2.0]\n (a, v[1]) << RW(0.1)",         # bad_move_tuple
])
def test_move_rejects_accessor_targets(body):
    with pytest.raises((ModelSyntaxError, ws.UnsupportedModelError)):
        _model(body)


# ---- test/accessors_test.jl:35-60: struct-valued columns are outside the device-op set ---------------------------------
@pytest.mark.parametrize("body", [
    "a ~ Normal(0, 1)\n b ~ Normal(10, 1)\n p .= Point(a, b)\n s .= p.x + p.y",          # prop_model
    "a ~ Normal(0, 1)\n b ~ Normal(10, 1)\n bag .= Bag([a, b])\n s .= bag.v[1] + bag.v[2]",   # chained_model
])
def test_struct_valued_columns_are_rejected_at_macro_expansion(body):
    with pytest.raises((ws.UnsupportedModelError, ModelSyntaxError)):
        _model(body)()


# ---- test/accessors_test.jl:23-33,62-80: `[]` reads and writes on an array-valued column --------------------------------
@pytest.mark.gpu
def test_idx_accessor_reads_and_writes():
    m = _model("a ~ Normal(0, 1)\n b ~ Normal(10, 1)\n v .= [a, b]\n s .= v[1] + v[2]\n v[1] .= v[1] + 100.0")
    st = ws.SMCState(1000, seed=42, device=0)
    ws.run(m(), st)
    a, b, s, v = st["a"], st["b"], st["s"], st["v"]
    np.testing.assert_array_equal(s, a + b)               # read_ok
    np.testing.assert_array_equal(v[:, 0], a + 100.0)     # write_ok (the pre-write value of a)
    np.testing.assert_array_equal(v[:, 1], b)             # other_untouched


# ---- test/dynamic_move_test.jl:15-50: `(α, β{1}) << RW(0.1)` -----------------------------------------------------------
@pytest.mark.gpu
def test_linear_regression_with_a_dynamic_family_move_target():
    rng = np.random.default_rng(42)
    xs = np.linspace(0, 10, 10)
    ys = -1.0 + 2.0 * xs + 0.5 * rng.standard_normal(10)
    m = ws.model('''
    @model function linear_regression_dyn(data)
        α ~ Normal(0.0, 5.0)
        β{1} ~ Normal(0.0, 5.0)
        for (x, y) in data
            y => Normal(α + β{1} * x, 0.5)
            if resampled
                (α, β{1}) << RW(0.1)
            end
        end
    end
    ''')
    st = ws.SMCState(10_000, seed=42, device=0)
    ws.run(m(list(zip(xs, ys))), st)
    w = ws.exp_norm(st)
    assert abs(float(np.sum(st["α"] * w)) + 1.0) < 0.3 and abs(float(np.sum(st["β_1"] * w)) - 2.0) < 0.3
    assert st.stats()["moves_run"] >= 1


# ---- test/default_kernels_test.jl:14-64: default tables with no kernels= / proposals= ---------------------------------
@pytest.mark.gpu
def test_default_kernels_and_proposals_resolve_without_tables():
    T, n = 10, 100_000
    rw = ws.model('''
    @model function random_walk_default(T)
        x ~ Normal(0, 1)
        for t in 1:T
            x ~ Normal(x, 1)
        end
    end
    ''')
    st = ws.SMCState(n, seed=42, device=0)
    ws.run(rw(T), st)
    x = st["x"]
    assert abs(x.mean()) < 0.15 and abs(x.var() / (T + 1) - 1.0) < 0.05
    rng = np.random.default_rng(42)
    xs = np.linspace(0, 10, 10)
    ys = -1.0 + 2.0 * xs + 0.5 * rng.standard_normal(10)
    lr = ws.model('''
    @model function linear_regression_default(data)
        α ~ Normal(0.0, 5.0)
        β ~ Normal(0.0, 5.0)
        for (x, y) in data
            y => Normal(α + β * x, 0.5)
            if resampled
                (α, β) << RW(0.1)
            end
        end
    end
    ''')
    st = ws.SMCState(10_000, seed=42, device=0)
    ws.run(lr(list(zip(xs, ys))), st)
    w = ws.exp_norm(st)
    assert abs(float(np.sum(st["α"] * w)) + 1.0) < 0.3 and abs(float(np.sum(st["β"] * w)) - 2.0) < 0.3
