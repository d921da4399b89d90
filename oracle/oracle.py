"""ctypes binding of the CPU oracle (oracle/halo2_oracle.c). TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (spectre_b200/) must never import this module.

All field elements cross this boundary as numpy uint64 arrays of shape (..., 4): the Montgomery limbs
halo2curves keeps in memory (SURVEY.md 8b "Data conventions"). Helpers convert to / from Python ints.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libhalo2_oracle.so")

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
MONT_R = 1 << 256


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "halo2_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_init()
        _lib.orc_domain_new.restype = ctypes.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def default_threads():
    return os.cpu_count() or 1


# ---- int <-> Montgomery limb conversion (pure Python, independent of the C code) ----------------
def to_mont(vals, mod):
    """list/iterable of Python ints -> (n,4) uint64 Montgomery limbs."""
    vals = list(vals)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % mod) * MONT_R % mod
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_mont(arr, mod):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(MONT_R, -1, mod)
    out = []
    for row in arr:
        m = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        out.append(m * rinv % mod)
    return out


def fr(vals):
    return to_mont(vals, R_MOD)


def fq(vals):
    return to_mont(vals, P_MOD)


def fr_ints(arr):
    return from_mont(arr, R_MOD)


def fq_ints(arr):
    return from_mont(arr, P_MOD)


def affine_ints(arr):
    """(n,8) uint64 G1Affine -> list of (x,y) ints; identity -> (0,0)."""
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 8)
    xs = fq_ints(arr[:, :4]); ys = fq_ints(arr[:, 4:])
    return list(zip(xs, ys))


# ---- wrapped entry points --------------------------------------------------------------------------
def fr_random_chacha(n, seed, first=0):
    """n Fr elements as `Fr::random(&mut ChaCha20Rng::from_seed(seed))` would draw them (from the `first`-th draw on).
    seed: int or 32 bytes."""
    if isinstance(seed, int):
        seed = seed.to_bytes(32, "little")
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_fr_random_chacha_from(_p(out), ctypes.c_size_t(n), ctypes.c_char_p(bytes(seed)), ctypes.c_uint64(first))
    return out


def best_fft(a, omega, log_n, threads=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    assert a.shape == (1 << log_n, 4)
    lib().orc_best_fft(_p(a), _p(omega), ctypes.c_uint32(log_n), ctypes.c_int(threads or default_threads()))
    return a


def best_multiexp(coeffs, bases, threads=None):
    """-> (12,) uint64 Jacobian point."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64); bases = np.ascontiguousarray(bases, dtype=np.uint64)
    n = coeffs.shape[0]
    assert bases.shape[0] == n, "best_multiexp: coeffs.len() != bases.len()"
    out = np.empty(12, dtype=np.uint64)
    lib().orc_best_multiexp(_p(coeffs), _p(bases), ctypes.c_size_t(n), ctypes.c_int(threads or default_threads()), _p(out))
    return out


def g1_to_affine(j):
    j = np.ascontiguousarray(j, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_to_affine(_p(out), _p(j))
    return out


def g1_generator():
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_mul(base, scalar):
    base = np.ascontiguousarray(base, dtype=np.uint64); scalar = np.ascontiguousarray(scalar, dtype=np.uint64)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_mul(_p(out), _p(base), _p(scalar))
    return out


def g1_add(a, b):
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add(_p(out), _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)))
    return out


def g1_on_curve(p):
    return bool(lib().orc_g1_on_curve(_p(np.ascontiguousarray(p, dtype=np.uint64))))


def g1_fixed_base_mul(scalars, threads=None):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = scalars.shape[0]
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_g1_fixed_base_mul(_p(scalars), ctypes.c_size_t(n), ctypes.c_int(threads or default_threads()), _p(out))
    return out


def srs_tau():
    out = np.empty(4, dtype=np.uint64)
    lib().orc_srs_tau(_p(out))
    return out


def srs_g(k, start, count, threads=None):
    out = np.empty((count, 8), dtype=np.uint64)
    lib().orc_srs_g(ctypes.c_uint32(k), ctypes.c_size_t(start), ctypes.c_size_t(count), ctypes.c_int(threads or default_threads()), _p(out))
    return out


def srs_g_lagrange(k, start, count, threads=None):
    out = np.empty((count, 8), dtype=np.uint64)
    lib().orc_srs_g_lagrange(ctypes.c_uint32(k), ctypes.c_size_t(start), ctypes.c_size_t(count), ctypes.c_int(threads or default_threads()), _p(out))
    return out


def srs_s_g2():
    out = np.empty((4, 4), dtype=np.uint64)
    lib().orc_srs_s_g2(_p(out))
    return out


def commit_known_tau(coeffs):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_commit_known_tau(_p(coeffs), ctypes.c_size_t(coeffs.shape[0]), _p(out))
    return out


def commit_lagrange_known_tau(k, evals):
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_commit_lagrange_known_tau(ctypes.c_uint32(k), _p(evals), ctypes.c_size_t(evals.shape[0]), _p(out))
    return out


class Domain:
    """EvaluationDomain::new(j, k) restated (oracle side)."""

    def __init__(self, j, k):
        self.j, self.k = j, k
        self.h = ctypes.c_void_p(lib().orc_domain_new(ctypes.c_uint32(j), ctypes.c_uint32(k)))
        ek = ctypes.c_uint32(0)
        self.omega = np.empty(4, dtype=np.uint64); self.extended_omega = np.empty(4, dtype=np.uint64)
        consts = np.empty((6, 4), dtype=np.uint64)
        lib().orc_domain_describe(self.h, ctypes.byref(ek), _p(self.omega), _p(self.extended_omega), _p(consts), None)
        self.extended_k = ek.value
        (self.omega_inv, self.extended_omega_inv, self.g_coset, self.g_coset_inv,
         self.ifft_divisor, self.extended_ifft_divisor) = [consts[i].copy() for i in range(6)]
        self.t_evaluations = np.empty((1 << (self.extended_k - k), 4), dtype=np.uint64)
        lib().orc_domain_describe(self.h, ctypes.byref(ek), _p(self.omega), _p(self.extended_omega), _p(consts), _p(self.t_evaluations))

    def __del__(self):
        try:
            lib().orc_domain_free(self.h)
        except Exception:
            pass

    def lagrange_to_coeff(self, a, threads=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_lagrange_to_coeff(self.h, _p(a), ctypes.c_int(threads or default_threads()))
        return a

    def coeff_to_extended(self, a, threads=None):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty((1 << self.extended_k, 4), dtype=np.uint64)
        lib().orc_coeff_to_extended(self.h, _p(a), _p(out), ctypes.c_int(threads or default_threads()))
        return out

    def extended_to_coeff(self, a, threads=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        out = np.empty(((1 << self.k) * (self.j - 1), 4), dtype=np.uint64)
        lib().orc_extended_to_coeff(self.h, _p(a), _p(out), ctypes.c_int(threads or default_threads()))
        return out

    def divide_by_vanishing_poly(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_divide_by_vanishing_poly(self.h, _p(a))
        return a


def eval_polynomial(poly, point):
    poly = np.ascontiguousarray(poly, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_eval_polynomial(_p(out), _p(poly), ctypes.c_size_t(poly.shape[0]), _p(np.ascontiguousarray(point, dtype=np.uint64)))
    return out


def kate_division(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    q = np.empty((a.shape[0] - 1, 4), dtype=np.uint64)
    lib().orc_kate_division(_p(q), _p(a), ctypes.c_size_t(a.shape[0]), _p(np.ascontiguousarray(b, dtype=np.uint64)))
    return q


def batch_invert(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_batch_invert(_p(a), ctypes.c_size_t(a.shape[0]))
    return a


def fr_binop(name, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    getattr(lib(), "orc_" + name)(_p(out), _p(a), _p(b))
    return out


def fr_seq(n):
    """Montgomery limbs of 0, 1, ..., n-1 (range-table column), built by repeated addition in the oracle."""
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_fr_seq(_p(out), ctypes.c_size_t(n))
    return out


# ---- quotient numerator (evaluate_h pieces) ----------------------------------------------------------
def _ptr_array(arrs):
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in arrs]
    return (ctypes.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs]), arrs


def fr_delta():
    out = np.empty(4, dtype=np.uint64)
    lib().orc_fr_delta(_p(out))
    return out


def graph_evaluate(prog, ncalc, n_inter, constants, rotations, fixed, advice, instance, challenges, bgty, values, rot_scale):
    prog = np.ascontiguousarray(prog, dtype=np.uint32); rotations = np.ascontiguousarray(rotations, dtype=np.int32)
    constants = np.ascontiguousarray(constants, dtype=np.uint64).reshape(-1, 4)
    challenges = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4); bgty = np.ascontiguousarray(bgty, dtype=np.uint64).reshape(4, 4)
    values = np.ascontiguousarray(values, dtype=np.uint64).copy()
    pf, kf = _ptr_array(fixed); pa, ka = _ptr_array(advice); pi, ki = _ptr_array(instance)
    lib().orc_graph_evaluate(_p(prog), ctypes.c_uint32(ncalc), ctypes.c_uint32(n_inter), _p(constants), _p(rotations), ctypes.c_uint32(len(rotations)),
                             pf, pa, pi, _p(challenges), _p(bgty), _p(values), ctypes.c_uint64(values.shape[0]), ctypes.c_int32(rot_scale))
    return values


def permutation_constraints(values, rot_scale, last_rotation, chunk_len, z, col_values, sigma, l0, l_last, l_active, beta, gamma, y, extended_omega):
    values = np.ascontiguousarray(values, dtype=np.uint64).copy()
    pz, kz = _ptr_array(z); pc, kc = _ptr_array(col_values); ps, ks = _ptr_array(sigma)
    d = fr_delta()
    args = [np.ascontiguousarray(a, dtype=np.uint64) for a in (l0, l_last, l_active, beta, gamma, y, d, extended_omega)]
    lib().orc_permutation_constraints(_p(values), ctypes.c_uint64(values.shape[0]), ctypes.c_int32(rot_scale), ctypes.c_int32(last_rotation), ctypes.c_uint32(len(z)),
                                      ctypes.c_uint32(chunk_len), pz, ctypes.c_uint32(len(col_values)), pc, ps, *[_p(a) for a in args])
    return values


def lookup_constraints(values, rot_scale, product, permuted_input, permuted_table, table_value, l0, l_last, l_active, beta, gamma, y):
    values = np.ascontiguousarray(values, dtype=np.uint64).copy()
    args = [np.ascontiguousarray(a, dtype=np.uint64) for a in (product, permuted_input, permuted_table, table_value, l0, l_last, l_active, beta, gamma, y)]
    lib().orc_lookup_constraints(_p(values), ctypes.c_uint64(values.shape[0]), ctypes.c_int32(rot_scale), *[_p(a) for a in args])
    return values


def write_params_file(path, k, threads=None):
    """The file `ParamsKZG::<Bn256>::setup(k, ChaCha20Rng::from_seed([0;32])).write(..)` produces (SerdeFormat::RawBytes):
    k u32 LE | g | g_lagrange | g2 | s_g2, Montgomery limbs -- i.e. halo2-base gen_srs's params/kzg_bn254_{k}.srs."""
    n = 1 << k
    g2 = np.empty((4, 4), dtype=np.uint64); s_g2 = np.empty((4, 4), dtype=np.uint64)
    lib().orc_srs_g2_raw(_p(g2), _p(s_g2))
    with open(path, "wb") as f:
        f.write(np.uint32(k).tobytes())
        f.write(srs_g(k, 0, n, threads).tobytes())
        f.write(srs_g_lagrange(k, 0, n, threads).tobytes())
        f.write(g2.tobytes()); f.write(s_g2.tobytes())
    return g2, s_g2


def permute_expression_pair(input_expr, table_expr):
    """-> (permuted_input, permuted_table) over the usable rows given; raises ValueError like upstream's
    Error::ConstraintSystemFailure when an input value is missing from the table."""
    a = np.ascontiguousarray(input_expr, dtype=np.uint64); t = np.ascontiguousarray(table_expr, dtype=np.uint64)
    assert a.shape == t.shape
    pi = np.empty_like(a); pt = np.empty_like(a)
    rc = lib().orc_permute_expression_pair(_p(a), _p(t), ctypes.c_size_t(a.shape[0]), _p(pi), _p(pt))
    if rc != 0:
        raise ValueError("permute_expression_pair: ConstraintSystemFailure (%d)" % rc)
    return pi, pt


# ---- argument provers (permutation / lookup grand products, SHPLONK) -----------------------------------------
def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def permutation_product(k, values, sigma, first_col, beta, gamma, blinds, last_z):
    """One permutation set -> (z over all 2^k rows, new last_z)."""
    n = 1 << k
    pv, kv = _ptr_array(values); ps, ks = _ptr_array(sigma)
    blinds = _c(blinds).reshape(-1, 4); z = np.empty((n, 4), dtype=np.uint64); lz = _c(last_z).copy()
    lib().orc_permutation_product(ctypes.c_uint32(k), pv, ps, ctypes.c_uint32(len(values)), ctypes.c_uint32(first_col), _p(_c(beta)), _p(_c(gamma)),
                                  _p(blinds), ctypes.c_uint32(blinds.shape[0]), _p(lz), _p(z))
    return z, lz


def lookup_product(compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma, blinds):
    a = [_c(x) for x in (compressed_input, compressed_table, permuted_input, permuted_table)]
    n = a[0].shape[0]
    blinds = _c(blinds).reshape(-1, 4); z = np.empty((n, 4), dtype=np.uint64)
    lib().orc_lookup_product(ctypes.c_size_t(n), *[_p(x) for x in a], _p(_c(beta)), _p(_c(gamma)), _p(blinds), ctypes.c_uint32(blinds.shape[0]), _p(z))
    return z


class _RotationSet(ctypes.Structure):
    _fields_ = [("points", ctypes.c_void_p), ("n_points", ctypes.c_uint32), ("polys", ctypes.c_void_p), ("n_polys", ctypes.c_uint32), ("evals", ctypes.c_void_p)]


def _rotation_sets(sets):
    """sets: list of (points (m,4), [poly arrays], evals (n_polys, m, 4)) -> ctypes array + keep-alive list"""
    keep = []
    arr = (_RotationSet * len(sets))()
    for i, (points, polys, evals) in enumerate(sets):
        points = _c(points).reshape(-1, 4); evals = _c(evals).reshape(len(polys), points.shape[0], 4)
        pp, kp = _ptr_array(polys)
        keep += [points, evals, pp, kp]
        arr[i] = _RotationSet(points.ctypes.data, points.shape[0], ctypes.cast(pp, ctypes.c_void_p).value, len(polys), evals.ctypes.data)
    return arr, keep


def shplonk_quotient(n, sets, y, v):
    arr, keep = _rotation_sets(sets)
    h = np.empty((n, 4), dtype=np.uint64)
    lib().orc_shplonk_quotient(ctypes.c_size_t(n), arr, ctypes.c_uint32(len(sets)), _p(_c(y)), _p(_c(v)), _p(h))
    return h


def shplonk_linearisation(n, sets, y, v, u, h_x):
    arr, keep = _rotation_sets(sets)
    out = np.empty((n - 1, 4), dtype=np.uint64)
    rc = lib().orc_shplonk_linearisation(ctypes.c_size_t(n), arr, ctypes.c_uint32(len(sets)), _p(_c(y)), _p(_c(v)), _p(_c(u)), _p(_c(h_x)), _p(out))
    if rc != 0:
        raise ValueError("shplonk: linearisation polynomial does not vanish at u (%d): evaluations inconsistent with the polynomials" % rc)
    return out


def vec_scale(a, alpha):
    a = _c(a).copy()
    lib().orc_vec_scale(_p(a), _p(_c(alpha)), ctypes.c_size_t(a.shape[0]))
    return a


def vec_fold(polys, y):
    pp, keep = _ptr_array(polys)
    out = np.empty_like(keep[0])
    lib().orc_vec_fold(pp, ctypes.c_size_t(len(polys)), _p(_c(y)), _p(out), ctypes.c_size_t(out.shape[0]))
    return out


def compute_inner_product(a, b):
    """arithmetic::compute_inner_product: sum_i a_i * b_i (Montgomery in, Montgomery out)"""
    a = _c(a); b = _c(b)
    assert a.shape == b.shape
    out = np.empty(4, dtype=np.uint64)
    lib().orc_compute_inner_product(_p(out), _p(a), _p(b), ctypes.c_size_t(a.shape[0]))
    return out
