#!/usr/bin/env python
"""Coefficients of the device exp_nonpos() polynomial (csrc/smpc_math.cuh): exp(r) = 1 + r + r^2 h(r) on the reduced
range |r| <= ln2/2, h = degree-9 Chebyshev interpolant (60-digit mpmath), and the error of the whole double-precision
evaluation (fma emulated exactly) against mpmath over x in [-708, 0]. Run: python tools/fit_exp.py"""
import random

import mpmath as mp

mp.mp.dps = 60
H = mp.log(2) / 2 * mp.mpf('1.0002')


def h(r):
    return (mp.exp(r) - 1 - r) / (r * r) if r != 0 else mp.mpf(1) / 2


def fit(n):  # n coefficients (degree n-1) of h on [-H, H]
    nodes = [H * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        for j in range(n):
            A[i, j] = z ** j
        b[i] = h(z)
    c = mp.lu_solve(A, b)
    return [float(c[j]) for j in range(n)]


def fma(a, b, c):
    return float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c))


L2E = 1.4426950408889634
LN2_HI = 0.6931471805599453
LN2_LO = float(mp.log(2) - mp.mpf(LN2_HI))
MAGIC = 6755399441055744.0


def exp_nonpos(x, c):
    t = fma(x, L2E, MAGIC)
    nf = t - MAGIC
    n = int(nf)
    r = fma(nf, -LN2_HI, x)
    r = fma(nf, -LN2_LO, r)
    q = c[-1]  # Horner, the device function's exact operation order
    for k in reversed(c[:-1]):
        q = fma(q, r, k)
    q = fma(q, r, 1.0)  # c1
    q = fma(q, r, 1.0)  # c0
    return q * 2.0 ** n


if __name__ == '__main__':
    c = fit(10)
    print('LN2_LO', repr(LN2_LO))
    print('h coefficients (c2..c11):')
    for k in c:
        print('   ', repr(k))
    random.seed(1)
    worst = 0
    for i in range(40000):
        x = -random.random() * (708.0 if i % 2 else 40.0)
        ref = mp.exp(mp.mpf(x))
        got = exp_nonpos(x, c)
        worst = max(worst, abs((mp.mpf(got) - ref) / ref))
    print('max relative error', mp.nstr(worst, 5), '=', mp.nstr(worst / mp.mpf(2) ** -53, 4), 'x 2^-53')
