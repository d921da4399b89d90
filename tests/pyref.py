"""Tiny pure-Python big-int reference (independent of both the oracle C code and the CUDA code)."""
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
ZETA = pow(pow(7, (R_MOD - 1) // 3, R_MOD), 2, R_MOD)


def omega(k):
    return pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)


def dft(a, w):
    n = len(a)
    return [sum(a[j] * pow(w, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


def ec_add(p, q):
    """affine BN254 G1 addition; None = identity"""
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


def ec_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = ec_add(r, p)
        p = ec_add(p, p)
        k >>= 1
    return r


def msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = ec_add(acc, ec_mul(p, s % R_MOD))
    return acc


def aff_tuple(t):
    """(x,y) ints with halo2's (0,0) identity -> None"""
    return None if t == (0, 0) else t
