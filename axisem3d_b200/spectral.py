"""nPol=4 spectral constants computed from first principles (double precision).

The reference hard-codes them to 12 decimals (S/preloop/spectral/SpectralConstants.cpp:43-49)
and hands G_GLL/G_GLJ to the hot path through `Gradient::setGMat` (Gradient.cpp:324-329).
Here they are derived numerically: GLL nodes = roots of (1-x^2) P_N'(x); GLJ(0,1) nodes =
roots of (1-x^2) d/dx[(P_N + P_{N+1})/(1+x)]; weights from exactness of the quadrature with
weight 1 (GLL) or (1+x) (GLJ); G(i,j) = l_i'(x_j) (row-major, SpectralConstants.cpp:104-110).
tests/test_spectral.py checks them against the reference's decimals when /root/reference exists.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as L
from numpy.polynomial import polynomial as Pn

N = 4


def _newton_polish(p, x):
    dp = Pn.polyder(p)
    for _ in range(50):
        x = x - Pn.polyval(x, p) / Pn.polyval(x, dp)
    return x


def gll_nodes(n=N):
    pN = L.leg2poly([0] * n + [1])
    inner = np.sort(np.real(Pn.polyroots(Pn.polyder(pN))))
    inner = _newton_polish(Pn.polyder(pN), inner)
    return np.concatenate([[-1.0], inner, [1.0]])


def glj_nodes(n=N):
    s = Pn.polyadd(L.leg2poly([0] * n + [1]), L.leg2poly([0] * (n + 1) + [1]))
    q, r = Pn.polydiv(s, np.array([1.0, 1.0]))          # (P_N + P_{N+1}) / (1 + x), exact
    assert np.max(np.abs(r)) < 1e-12
    inner = np.sort(np.real(Pn.polyroots(Pn.polyder(q))))
    inner = _newton_polish(Pn.polyder(q), inner)
    return np.concatenate([[-1.0], inner, [1.0]])


def quad_weights(x, jacobi01):
    """Interpolatory weights: exact for monomials 0..n with weight 1 or (1+x)."""
    n = len(x)
    V = np.vander(x, n, increasing=True).T              # V[k, i] = x_i^k
    k = np.arange(n)
    m0 = (1.0 - (-1.0) ** (k + 1)) / (k + 1)            # int x^k
    m1 = (1.0 - (-1.0) ** (k + 2)) / (k + 2)            # int x^(k+1)
    return np.linalg.solve(V, m0 + m1 if jacobi01 else m0)


def lagrange_deriv_matrix(x):
    """G[i, j] = l_i'(x_j)."""
    n = len(x)
    c = np.array([1.0 / np.prod([x[i] - x[k] for k in range(n) if k != i]) for i in range(n)])
    G = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                G[i, j] = (c[i] / c[j]) / (x[j] - x[i])
    for j in range(n):
        G[j, j] = -np.sum(G[:, j]) + G[j, j]            # rows of D sum to zero: sum_i l_i'(x_j) = 0
    return G


P_GLL = gll_nodes()
P_GLJ = glj_nodes()
W_GLL = quad_weights(P_GLL, False)
W_GLJ = quad_weights(P_GLJ, True)
G_GLL = lagrange_deriv_matrix(P_GLL)
G_GLJ = lagrange_deriv_matrix(P_GLJ)


def next_lucky_number(n, force_odd=False):
    """PreloopFFTW::nextLuckyNumber / isLuckyNumber (S/preloop/utilities/PreloopFFTW.cpp:59-109):
    prime factors <= 13, at most one factor of 11 or 13 in total; odd if force_odd, otherwise
    even whenever n > 10."""
    def lucky(m):
        if force_odd and m % 2 == 0:
            return False
        if (not force_odd) and m % 2 != 0 and m > 10:
            return False
        num = m
        for p in range(2, m + 1):
            while num % p == 0:
                num //= p
                if p > 13:
                    return False
            if num == 1:
                break
        e = f = 0
        num = m
        while num % 11 == 0:
            num //= 11
            e += 1
        num = m
        while num % 13 == 0:
            num //= 13
            f += 1
        return e + f <= 1
    n = int(n)
    while not lucky(n):
        n += 1
    return n
