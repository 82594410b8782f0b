"""Derives the polynomial coefficients of the custom FP64 kernels in csrc/ws_math.cuh (ws_exp_nonpos,
ws_log_u, ws_sincos_octant) by interpolation at Chebyshev nodes in 60-digit arithmetic (near-minimax),
rounds them to double and reports the worst error of the double-precision Horner evaluation.
Run: python scripts/fit_math_polys.py   (needs mpmath; prints C initialisers)."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    """degree n-1 polynomial (n coefficients, ascending) interpolating f at n Chebyshev nodes of [a, b]"""
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[j] for j in range(n)]


def horner(c, x):
    acc = np.full_like(x, c[-1])
    for k in c[-2::-1]:
        acc = acc * x + k
    return acc


def report(name, c):
    print(f"// {name}")
    print("{" + ", ".join(float(v).hex() for v in c) + "}")
    print("  = {" + ", ".join(repr(float(v)) for v in c) + "}")


# ---- sin / cos of (pi/4) f, f in [0, 1]:  sin = f * S(f^2), cos = C(f^2)
q = mp.pi / 4
for ns in (7, 8):
    S = cheb_fit(lambda t: mp.sin(q * mp.sqrt(t)) / mp.sqrt(t) if t > 0 else q, mp.mpf(0), mp.mpf(1), ns)
    Sd = [float(v) for v in S]
    f = np.linspace(0.0, 1.0, 200001)
    got = f * horner(Sd, f * f)
    want = np.array([float(mp.sin(q * mp.mpf(float(v)))) for v in f[::40]])
    err = np.max(np.abs(got[::40] - want) / np.maximum(want, 1e-300)[...].clip(1e-300))
    abserr = np.max(np.abs(got[::40] - want))
    print(f"sin terms={ns}: max rel err {err:.3e} abs {abserr:.3e}")
    if ns == 7:
        report("S (sin((pi/4) f) = f * S(f^2))", S)
for nc in (8, 9):
    Cc = cheb_fit(lambda t: mp.cos(q * mp.sqrt(t)), mp.mpf(0), mp.mpf(1), nc)
    Cd = [float(v) for v in Cc]
    got = horner(Cd, f * f)
    want = np.array([float(mp.cos(q * mp.mpf(float(v)))) for v in f[::40]])
    print(f"cos terms={nc}: max abs err {np.max(np.abs(got[::40] - want)):.3e}")
    if nc == 8:
        report("C (cos((pi/4) f) = C(f^2))", Cc)

# ---- exp(r), |r| <= ln2/2
h = mp.log(2) / 2
for ne in (11, 12, 13):
    E = cheb_fit(mp.exp, -h, h, ne)
    Ed = [float(v) for v in E]
    r = np.linspace(-float(h), float(h), 100001)
    got = horner(Ed, r)
    want = np.array([float(mp.exp(mp.mpf(float(v)))) for v in r[::20]])
    print(f"exp terms={ne}: max rel err {np.max(np.abs(got[::20] - want) / want):.3e}")
    if ne == 12:
        report("E (exp(r), |r| <= ln2/2)", E)

# ---- log1p(r) = r + r^2 * L(r), |r| <= 2^-7
b = mp.mpf(2) ** -7
for nl in (6, 7, 8):
    Lc = cheb_fit(lambda t: (mp.log1p(t) - t) / (t * t) if abs(t) > mp.mpf(10) ** -20 else mp.mpf(-0.5) + t / 3, -b, b, nl)
    Ld = [float(v) for v in Lc]
    r = np.linspace(-float(b), float(b), 100001)
    got = r + r * r * horner(Ld, r)
    want = np.array([float(mp.log1p(mp.mpf(float(v)))) for v in r[::20]])
    nz = np.abs(want) > 0
    print(f"log1p terms={nl}: max rel err {np.max(np.abs(got[::20][nz] - want[nz]) / np.abs(want[nz])):.3e}")
    if nl == 6:
        report("L (log1p(r) = r + r^2 L(r), |r| <= 2^-7)", Lc)

# ---- log(z) = 2 atanh(s), s = (z-1)/(z+1), z in [sqrt(1/2), sqrt(2)]  =>  |s| <= 0.171573
# atanh(s) = s + s^3 * A(s^2)
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
for na in (6, 7, 8):
    A = cheb_fit(lambda t: (mp.atanh(mp.sqrt(t)) - mp.sqrt(t)) / (t * mp.sqrt(t)) if t > mp.mpf(10) ** -30 else mp.mpf(1) / 3 + t / 5,
                 mp.mpf(0), smax * smax * mp.mpf("1.02"), na)
    Ad = [float(v) for v in A]
    s = np.linspace(-float(smax), float(smax), 100001)
    got = 2.0 * (s + s * (s * s) * horner(Ad, s * s))
    want = np.array([float(2 * mp.atanh(mp.mpf(float(v)))) for v in s[::20]])
    nz = np.abs(want) > 0
    print(f"atanh terms={na}: max rel err {np.max(np.abs(got[::20][nz] - want[nz]) / np.abs(want[nz])):.3e}")
    if na == 7:
        report("A (atanh(s) = s + s^3 A(s^2), |s| <= 0.1716)", A)
