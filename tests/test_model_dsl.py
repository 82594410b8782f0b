"""@model front-end (rewrites.jl): statement forms, auto-Resample, dynamic families, error cases — CPU only
(building a transformer tree needs no device)."""
import numpy as np
import pytest

import models
import wsb200 as ws
from wsb200.model import ModelSyntaxError


def kinds(t, state=None):
    out = []
    for s in t.steps:
        k = type(s).__name__
        if k == "Loop":
            coll = list(s.collfn(None))
            out.append(("Loop", len(coll), kinds(s.bodyfn(coll[0]))))
        elif k == "Cond":
            out.append(("Cond", kinds(s.body)))
        elif k == "Sequence":
            out.append(kinds(s))
        else:
            out.append(k)
    return out


def test_expansions_match_survey_appendix_a():
    t = ws.model(models.SSM1D)([0.1, 0.2, 0.3])
    assert kinds(t) == ["Assign", "Assign", ("Loop", 3, ["Assign", "Sample", "Resample", "Assign", "Observe", "Resample"])]
    body = t.steps[2].bodyfn((2, 0.5))
    assert body.steps[0].lhs == "x_3" and body.steps[1].lhs == "dv"
    t = ws.model(models.LINREG)(range(1, 4), [1.0, 2.0, 3.0])
    assert kinds(t) == ["Sample", "Resample", "Sample", "Resample",
                        ("Loop", 3, ["Observe", "Resample", ("Cond", ["Move", "Move"])])]
    t = ws.model(models.SCHOOLS)(8, models.SCHOOLS_Y, models.SCHOOLS_SIGMA)
    assert kinds(t) == ["Sample", "Resample", "Sample", "Resample", "Assign",
                        ("Loop", 8, ["Sample", "Resample", "Observe", "Resample", "Move", "Move"])]
    b = t.steps[5].bodyfn(3)
    assert b.steps[0].lhs == ("θ", 2)                      # 1-based θ[3] -> plane 2
    assert b.steps[4].diversity_threshold == 0.9 and b.steps[4].proposal is ws.autoRW
    assert b.steps[5].argfn == (1e-3, (0.0, float("inf")))
    t = ws.model(models.LGSSM1D)([0.1], 0.9, 1.0, 0.5, 1.0)
    assert kinds(t) == ["Sample", "Resample", ("Loop", 1, ["Sample", "Resample", "Observe", "Resample"])]
    assert getattr(t, "_has_moves") is False


def test_build_time_locals_and_matrix_literals():
    m = ws.model('''
    @model function f(n)
        I2 = [1.0 0.0; 0.0 1.0]
        s = 0.0
        for k in 1:n
            s += k
            if k % 2 == 0
                z{k} .= s
            end
        end
        w ~ MvNormal([0.0, 0.0], 0.5 * I2)
    end
    ''')
    t = m(4)
    loop = t.steps[0]
    bodies = [loop.bodyfn(k) for k in loop.collfn(None)]
    conds = [b.steps[0] for b in bodies]
    assert [c.predfn(None) for c in conds] == [False, True, False, True]
    assert conds[1].body.steps[0].lhs == "z_2" and conds[1].body.steps[0].argfn == 3.0
    assert conds[3].body.steps[0].argfn == 10.0
    np.testing.assert_array_equal(t.steps[1].argfn[1], 0.5 * np.eye(2))
    assert m.dynamic_families == frozenset({"z"})


@pytest.mark.parametrize("body,exc", [
    ("x .= 1.0\n x = 2.0", ModelSyntaxError),                       # plain = on a particle variable
    ("x .= 1.0\n y = x + 1", ModelSyntaxError),                     # particle variable on a plain = RHS
    ("x .= 1.0\n x{1} .= 2.0", ModelSyntaxError),                   # base both plain and dynamic
    ("x{1} .= 1.0\n x .= 2.0", ModelSyntaxError),
    ("i .= 1.0\n x{i} .= 2.0", ModelSyntaxError),                   # particle-dependent column name
    ("y .= x{3}", ModelSyntaxError),                                # unregistered family
    ("x .= 1.0\n x .+= 1.0", ModelSyntaxError),                     # dotted compound
    ("θ .= zeros(3)\n θ[1] << RW(0.1)", ModelSyntaxError),          # accessor move target (dynamic_move_test.jl:56-75)
    ("z << RW(0.1)", ModelSyntaxError),                             # unknown move target
    ("x .= 1.0\n if x\n x .= 2.0\n end", ModelSyntaxError),         # particle variable in a condition
    ("b ~ Dirichlet([1.0, 2.0])", ws.UnsupportedModelError),        # outside the device-op set
    ("p .= 1.0\n q .= p.x", ws.UnsupportedModelError),              # struct columns
])
def test_macro_expansion_errors(body, exc):
    with pytest.raises(exc):
        ws.model("@model function f()\n" + body + "\nend")


def test_expression_lowering_rejects_non_device_values():
    with pytest.raises(ws.UnsupportedModelError):
        ws.Assign("x", "a string")
        ws.expr.wrap("a string")
    with pytest.raises(ws.UnsupportedModelError):
        bool(ws.col("x") + 1.0)
    with pytest.raises(ws.UnsupportedModelError):
        ws.core.resolve_kernel("Wishart")
    with pytest.raises(ws.UnsupportedModelError):
        ws.Move(["x"], lambda state, targets: None)


def test_model_signature_is_a_julia_signature():
    """`@model function name(args...; kwargs...)` splices both parameter lists into the generated function verbatim
    (rewrites.jl:776-806), so type annotations (test/macro_test.jl:12 `T::Int`), optional positional arguments, keyword
    arguments with defaults (evaluated left to right with the earlier parameters in scope) and required keywords work
    as in Julia; `kernels` / `proposals` stay the two extra keywords of every model."""
    src = '''
    @model function f(data::Vector{Float64}, a=0.5, b::Float64=2a; q=a + b, T::Int)
        x ~ Normal(0.0, q)
        for t in 1:T
            x ~ Normal(a * x, b)
        end
    end
    '''
    m = ws.model(src)
    tree = m([1.0], T=2)
    assert [type(s).__name__ for s in tree.steps] == ["Sample", "Resample", "Loop"]
    assert len(list(tree.steps[2].collfn(None))) == 2
    assert len(list(m([1.0], 0.1, 0.2, T=3, q=1.0).steps[2].collfn(None))) == 3
    with pytest.raises(TypeError):
        m([1.0])                       # required keyword T
    with pytest.raises(TypeError):
        m([1.0], T=1, z=3)             # unknown keyword
    with pytest.raises(TypeError):
        m(T=1)                         # required positional argument
    with pytest.raises(TypeError):
        m([1.0], 1, 2, 3, T=1)         # too many positional arguments


def test_loop_variables_destructure_as_julia_tuples():
    """`for (i, (x, y)) in enumerate(data)`: the loop variable is the body closure's single destructuring argument
    (rewrites.jl:652-664), so patterns nest."""
    src = '''
    @model function g(data)
        a ~ Normal(0.0, 1.0)
        for (i, (x, y)) in enumerate(data)
            z{i} .= a * x + y
        end
    end
    '''
    tree = ws.model(src)([(1.0, 2.0), (3.0, 4.0)])
    loop = tree.steps[2]
    bodies = [loop.bodyfn(el) for el in loop.collfn(None)]
    assert [b.steps[0].lhs for b in bodies] == ["z_1", "z_2"]
