"""Fits the two branch-free polynomials the GEMM epilogue uses for the exact (erf) GELU and checks the float32 evaluation
against float64:   |z| < 1:  erf(z) = z + z * P(z^2)            (deg-6 P, weighted least squares on Chebyshev nodes)
                   |z| >= 1: erf(z) = sign(z) (1 - exp(Q(|z|)))   (deg-7 Q ~ log erfc on [0.9, 4], weight erfc)
Prints the coefficients pasted into rohm_b200/csrc/gemm.cu (gelu_erf) and the max abs error of the whole GELU."""
import numpy as np
from scipy.special import erf, erfc

f32 = np.float32
B = 1.0
z = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5 * B + 0.5 * B
z = z[z > 1e-3]
small, *_ = np.linalg.lstsq(np.vander(z * z, 7, increasing=True) * z[:, None], (erf(z) / z - 1) * z, rcond=None)
lo, hi = B - 0.1, 4.0
t = np.cos(np.pi * (np.arange(6000) + 0.5) / 6000) * 0.5 * (hi - lo) + 0.5 * (hi + lo)
w = erfc(t)
large, *_ = np.linalg.lstsq(np.vander(t, 8, increasing=True) * w[:, None], np.log(erfc(t)) * w, rcond=None)
small32, large32 = [f32(c) for c in small], [f32(c) for c in large]


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(f32)


def gelu32(x, rng):
    x = x.astype(f32)
    zz = (x * f32(0.70710678118654752440)).astype(f32)
    ss = (zz * zz).astype(f32)
    r = np.full_like(zz, small32[-1])
    for c in small32[-2::-1]:
        r = fma(r, ss, c)
    es = fma(r, zz, zz)
    tt = np.minimum(np.abs(zz), f32(4.0)).astype(f32)
    q = np.full_like(zz, large32[-1])
    for c in large32[-2::-1]:
        q = fma(q, tt, c)
    e = np.exp(q.astype(np.float64)) * (1 + rng.uniform(-1, 1, size=q.shape) * 2.0 ** -21)  # ex2.approx error model
    el = np.copysign((f32(1.0) - e.astype(f32)).astype(f32), zz)
    er = np.where(np.abs(zz) < f32(B), es, el).astype(f32)
    h = (f32(0.5) * x).astype(f32)
    return fma(h, er, h)


rng = np.random.default_rng(0)
x = np.concatenate([np.linspace(-12, 12, 2000001), rng.standard_normal(1000000) * 3])
x64 = x.astype(f32).astype(np.float64)
ref = 0.5 * x64 * (1 + erf(x64 / np.sqrt(2)))
err = np.abs(gelu32(x, rng).astype(np.float64) - ref)
z32 = (x.astype(f32) * f32(0.70710678118654752440)).astype(f32)
torch_like = (f32(0.5) * x.astype(f32) * (f32(1) + erf(z32.astype(np.float64)).astype(f32))).astype(f32)
print("max |gelu32 - gelu64|           :", err.max())
print("max |fp32 exact-erf path - gelu64|:", np.abs(torch_like.astype(np.float64) - ref).max())
print("small:", ", ".join(f"{float(c):.9e}f" for c in small32))
print("large:", ", ".join(f"{float(c):.9e}f" for c in large32))
