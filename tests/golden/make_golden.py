"""Generates tests/golden/*.npz from the oracle (oracle/ref.py) and the oracle's OWN hand-built transformer
programs (oracle/models.py) on fixed seeds.  Nothing of the product is imported.

The reference is Julia and cannot run here, so these are NOT reference outputs: they freeze the
oracle's answers (so that a later edit of the oracle or of the product is visible) and give the GPU
tests size-independent fixtures.  Run:  python tests/golden/make_golden.py
(oracle/gen_from_reference.jl writes the same files from the real reference where a Julia toolchain exists.)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import models as om  # noqa: E402
from oracle import ref  # noqa: E402

SCHOOLS_Y = [28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0]          # examples/eight_schools.jl:19-21
SCHOOLS_SIGMA = [15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0]


def resampling_case():
    rng = np.random.default_rng(2024)
    n = 4096
    logw = 2.0 * rng.standard_normal(n) - 50.0
    r = rng.random(n)
    w = ref.exp_norm(logw)
    return dict(logw=logw, r=r, w=w, lse=ref.logsumexp(logw), ess=ref.ess_perc(w), us=ref.stratified_us(r),
                ancestors=ref.icdf(w, ref.stratified_us(r)), ancestors_sys=ref.icdf(w, ref.systematic_us(r[0], n)))


def run_case(root, n, n_normals, n_uniforms, n_expon=0, ess=0.5, seed=1):
    rng = np.random.default_rng(seed)
    normals, uniforms = rng.standard_normal(n_normals), rng.random(n_uniforms)
    expon = rng.standard_exponential(n_expon)
    st = ref.OracleState(n, ref.Streams(normals, uniforms, expon), ess_perc_min=ess)
    ref.run(root, st)
    out = dict(normals=normals, uniforms=uniforms, exponentials=expon, weights=st.weights,
               log_evidence=ref.log_evidence(st), n_resampled=sum(e["resampled"] for e in st.log), depth=st.depth)
    for name in st.names:
        out["col_" + name] = st.cols[name]
    return out


def abi_smoke_fixture(path):
    """inputs and expected outputs of tests/host/abi_smoke.c (a plain C consumer of include/wsb200.h), little-endian:
    int64 n, T, n_normals, n_uniforms | doubles a q r x0_std step ess_perc_min | obs[T] | normals | uniforms |
    log_evidence | x[n] | weights[n] | int64 resamples, accepted, depth"""
    from oracle.trees import Move, RW
    n, T = 512, 6
    a, q, r, x0, step, ess = 0.9, 1.0, 0.5, 1.0, 0.3, 1.0
    rng = np.random.default_rng(31)
    obs = rng.standard_normal(T)
    normals, uniforms = rng.standard_normal(n * (T + 2)), rng.random(n * (T + 1))
    root = om.lgssm1d(list(obs), a, q, r, x0, tail=(Move(["x"], RW, (step,)),))
    st = ref.OracleState(n, ref.Streams(normals, uniforms), ess_perc_min=ess)
    ref.run(root, st)
    with open(path, "wb") as f:
        np.array([n, T, normals.size, uniforms.size], dtype="<i8").tofile(f)
        np.array([a, q, r, x0, step, ess], dtype="<f8").tofile(f)
        for arr in (obs, normals, uniforms, [ref.log_evidence(st)], st.cols["x"], st.weights):
            np.asarray(arr, dtype="<f8").tofile(f)
        np.array([sum(e["resampled"] for e in st.log), int(st.last_accept.sum()), st.depth], dtype="<i8").tofile(f)


def main():
    abi_smoke_fixture(os.path.join(HERE, "abi_smoke.bin"))
    np.savez_compressed(os.path.join(HERE, "resampling.npz"), **resampling_case())
    rng = np.random.default_rng(7)
    obs1 = np.cumsum(rng.standard_normal(12)) * 1.5
    np.savez_compressed(os.path.join(HERE, "ssm1d.npz"), obs=obs1,
                        **run_case(om.ssm1d(list(obs1)), 256, 256 * 12, 256 * 12))
    obs2 = rng.standard_normal((8, 2)) + np.arange(8)[:, None] * np.array([1.0, 0.0])
    np.savez_compressed(os.path.join(HERE, "ssm2d.npz"), obs=obs2,
                        **run_case(om.ssm2d([o for o in obs2]), 256, 256 * 2 * 8, 256 * 8))
    xs = rng.uniform(0, 10, 8)
    ys = 1.0 - 0.5 * xs + 0.5 * rng.standard_normal(8)
    np.savez_compressed(os.path.join(HERE, "linreg.npz"), xs=xs, ys=ys,
                        **run_case(om.linear_regression(xs, ys), 512, 512 * (2 + 16), 512 * 24))
    np.savez_compressed(os.path.join(HERE, "schools.npz"),
                        **run_case(om.eight_schools(8, SCHOOLS_Y, SCHOOLS_SIGMA), 512, 512 * 25, 512 * 24, 512))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
