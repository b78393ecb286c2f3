"""Transcendentals with PORTABLE arithmetic: every operation is an individually rounded IEEE
float64 add / subtract / multiply (plus rint), so the numpy code below and the CUDA code in
csrc/classic_env.cu (same constants, __dmul_rn / __dadd_rn / __dsub_rn) give bit-identical
results -- which the libm / CUDA math library pair does not guarantee.  Accuracy ~1 ulp
(argument reduction by Cody-Waite in two terms, polynomial kernels with the classic fdlibm
coefficients); used by the closed-form classic-control tasks (environments/classic.py).
"""

import numpy as np

INV_PIO2 = 6.36619772367581382433e-01     # 2 / pi
PIO2_HI = 1.57079632673412561417e+00      # first 33 bits of pi / 2
PIO2_LO = 6.07710050650619224932e-11      # pi / 2 - PIO2_HI
S = (-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
     2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10)
C = (4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
     -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11)


def sincos(x):
    """(sin x, cos x) for float64 scalars / arrays, |x| < 1e5."""
    x = np.asarray(x, np.float64)
    k = np.rint(x * INV_PIO2)
    r = (x - k * PIO2_HI) - k * PIO2_LO
    z = r * r
    ps = S[5]
    for coef in (S[4], S[3], S[2], S[1], S[0]):
        ps = coef + z * ps
    s = r + (r * z) * ps
    pc = C[5]
    for coef in (C[4], C[3], C[2], C[1], C[0]):
        pc = coef + z * pc
    c = (1.0 - 0.5 * z) + (z * z) * pc
    q = k.astype(np.int64) & 3
    sin = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c)))
    cos = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s)))
    return sin, cos


def fmix32(h):
    """murmur3 finaliser on uint64 arrays holding 32-bit values (csrc/common.cuh::fmix32)."""
    m = np.uint64(0xFFFFFFFF)
    h = np.asarray(h, np.uint64) & m
    h = h ^ (h >> np.uint64(16))
    h = (h * np.uint64(0x85EBCA6B)) & m
    h = h ^ (h >> np.uint64(13))
    h = (h * np.uint64(0xC2B2AE35)) & m
    h = h ^ (h >> np.uint64(16))
    return h


def reset_uniform(seed, episode, coordinate):
    """float64 in [0, 1): 24 hashed bits of (seed, episode, coordinate) -- the counter-based
    reset stream shared with the device kernels (csrc/env_dynamics.cuh: reset_key /
    reset_coordinate)."""
    m = np.uint64(0xFFFFFFFF)
    seed = np.asarray(seed, np.uint64) & m
    episode = np.asarray(episode, np.uint64) & m
    key = fmix32((seed + np.uint64(0x9E3779B9) * ((episode + np.uint64(1)) & m)) & m)
    h = fmix32(key ^ ((np.uint64(0x85EBCA6B) * np.uint64(coordinate + 1)) & m))
    return (h >> np.uint64(8)).astype(np.float64) * 2.0 ** -24
