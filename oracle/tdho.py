"""CPU oracle for the two-dimensional harmonic-oscillator family (TEST INFRASTRUCTURE ONLY).

* ``coulomb_ho`` / ``get_coulomb_elements`` / ``get_indices_nm`` / ``get_index_p``: ctypes front of
  ``oracle/tdho_oracle.c``, the plain-C restatement of reference
  ``quantum_dots/two_dim/coulomb_elements.py:6-152`` and ``two_dim_helper.py:111-166,250-268``.
* ``coulomb_ho_exact``: the same matrix element evaluated in exact rational arithmetic
  (``fractions.Fraction``), an independent ground truth used to bound the rounding of BOTH the
  reference's alternating sums and the CUDA kernel's.  Small samples only (pure Python).
* ``get_one_body_elements``, ``shell_energy_B``, ``energy_sorted_quantum_numbers``: the O(l) / O(l^2)
  host pieces of ``two_dim_helper.py:169-182,271-280,380-415``.

Pinned by ``tests/test_oracle_golden.py`` (reference golden table + reference-generated vectors).
All citations are relative to ``/root/reference/quantum_systems/``.
"""

import ctypes
from fractions import Fraction
from math import comb, factorial, pi, sqrt

import numpy as np

from . import build_oracle

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_oracle.build())
        i64, dp, ip = ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
        lib.tdho_coulomb_ho.argtypes = [i64] * 8
        lib.tdho_coulomb_ho.restype = ctypes.c_double
        lib.tdho_indices_nm.argtypes = [i64, ip, ip]
        lib.tdho_indices_nm.restype = None
        lib.tdho_index_p.argtypes = [i64, i64]
        lib.tdho_index_p.restype = i64
        lib.tdho_coulomb_elements.argtypes = [i64, dp]
        lib.tdho_coulomb_elements.restype = None
        lib.tdho_coulomb_elements_nm.argtypes = [ip, ip, i64, dp]
        lib.tdho_coulomb_elements_nm.restype = None
        lib.tdho_coulomb_sample.argtypes = [ip, ip, ip, i64, dp]
        lib.tdho_coulomb_sample.restype = None
        _LIB = lib
    return _LIB


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def coulomb_ho(n_i, m_i, n_j, m_j, n_l, m_l, n_k, m_k):
    """One matrix element, argument order of coulomb_elements.py:7 (third pair is "l", fourth "k")."""
    return _lib().tdho_coulomb_ho(n_i, m_i, n_j, m_j, n_l, m_l, n_k, m_k)


def get_indices_nm(p):
    """two_dim_helper.py:136-166."""
    n, m = ctypes.c_int64(), ctypes.c_int64()
    _lib().tdho_indices_nm(int(p), ctypes.byref(n), ctypes.byref(m))
    return n.value, m.value


def get_index_p(n, m):
    """two_dim_helper.py:111-133."""
    return _lib().tdho_index_p(int(n), int(m))


def quantum_numbers(num_orbitals):
    nm = np.array([get_indices_nm(p) for p in range(num_orbitals)], dtype=np.int64).reshape(num_orbitals, 2)
    return nm[:, 0].copy(), nm[:, 1].copy()


def get_coulomb_elements(num_orbitals, n=None, m=None):
    """``u[p,q,r,s] = coulomb_ho(nm_p, nm_q, nm_r, nm_s)`` -- two_dim_helper.py:250-268 (or :283-300 when
    explicit quantum-number arrays are given).  OpenMP over (p, q)."""
    u = np.zeros((num_orbitals,) * 4)
    out = u.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if n is None:
        _lib().tdho_coulomb_elements(num_orbitals, out)
    else:
        n, n_p = _i64(n)
        m, m_p = _i64(m)
        _lib().tdho_coulomb_elements_nm(n_p, m_p, num_orbitals, out)
    return u


def coulomb_sample(n, m, pqrs):
    """Individual elements ``u[p,q,r,s]`` for rows of ``pqrs`` (count, 4)."""
    n, n_p = _i64(n)
    m, m_p = _i64(m)
    pqrs, idx_p = _i64(pqrs)
    out = np.zeros(len(pqrs))
    _lib().tdho_coulomb_sample(n_p, m_p, idx_p, len(pqrs), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def get_shell_energy(n, m):
    """two_dim_helper.py:169-171."""
    return 2 * n + abs(m) + 1


def get_one_body_elements(num_orbitals):
    """Diagonal of shell energies -- two_dim_helper.py:174-182."""
    h = np.zeros((num_orbitals, num_orbitals))
    for p in range(num_orbitals):
        h[p, p] = get_shell_energy(*get_indices_nm(p))
    return h


# --------------------------------------------------------------------------------------------
# exact rational evaluation (independent ground truth)
# --------------------------------------------------------------------------------------------


def _signed_binomial_product(g_plus, g_minus):
    """Integer coefficients of (1 + x)^g_plus (1 - x)^g_minus."""
    out = [0] * (g_plus + g_minus + 1)
    for a in range(g_plus + 1):
        for b in range(g_minus + 1):
            out[a + b] += comb(g_plus, a) * comb(g_minus, b) * (-1) ** b
    return out


def coulomb_ho_exact(n_i, m_i, n_j, m_j, n_l, m_l, n_k, m_k):
    """The Anisimovas-Matulis element with every sum carried in ``Fraction``; only the final
    ``sqrt(rational) * sqrt(pi/2)`` is rounded.  Uses that the inner sum of coulomb_elements.py:60-82
    depends on (j_1 + j_4, j_2 + j_3) only and that l_1 + l_2 = l_3 + l_4 =: Lambda there, so
    ``Gamma(1 + L/2) = Lambda!`` and ``Gamma((G - L + 1)/2) = Gamma(S - Lambda + 1/2)``, S = g_1 + g_2."""
    n_i, m_i, n_j, m_j, n_l, m_l, n_k, m_k = (int(x) for x in (n_i, m_i, n_j, m_j, n_l, m_l, n_k, m_k))
    if m_i + m_j != m_k + m_l:
        return 0.0

    def up(m):
        return (abs(m) + m) // 2

    def down(m):
        return (abs(m) - m) // 2

    def weight(n, m, j):
        return Fraction((-1) ** j * factorial(n + abs(m)), factorial(j) * factorial(n - j) * factorial(j + abs(m)))

    total = Fraction(0)
    for s14 in range(n_i + n_l + 1):
        a14 = sum(
            (weight(n_i, m_i, j) * weight(n_l, m_l, s14 - j) for j in range(max(0, s14 - n_l), min(n_i, s14) + 1)),
            Fraction(0),
        )
        for s23 in range(n_j + n_k + 1):
            a23 = sum(
                (weight(n_j, m_j, j) * weight(n_k, m_k, s23 - j) for j in range(max(0, s23 - n_k), min(n_j, s23) + 1)),
                Fraction(0),
            )
            g1 = s14 + up(m_i) + down(m_l)
            g2 = s23 + up(m_j) + down(m_k)
            g3 = s23 + up(m_k) + down(m_j)
            g4 = s14 + up(m_l) + down(m_i)
            big_s = g1 + g2
            assert big_s == g3 + g4
            left = _signed_binomial_product(g1, g2)
            right = _signed_binomial_product(g4, g3)
            inner = Fraction(0)
            for lam in range(big_s + 1):
                k = big_s - lam  # Gamma(k + 1/2) = (2k)! / (4^k k!) sqrt(pi)
                inner += factorial(lam) * Fraction(factorial(2 * k), 4**k * factorial(k)) * left[lam] * right[lam]
            total += a14 * a23 * inner * (-1) ** (g2 + g3) * Fraction(1, 2**big_s)
    norm = Fraction(1)
    for n, m in ((n_i, m_i), (n_j, m_j), (n_k, m_k), (n_l, m_l)):
        norm *= Fraction(factorial(n), factorial(n + abs(m)))
    # exact value = total * sqrt(norm) * sqrt(pi) / sqrt(2)
    return float(total) * sqrt(float(norm)) * sqrt(pi / 2)


# --------------------------------------------------------------------------------------------
# magnetic-field ordering
# --------------------------------------------------------------------------------------------


def shell_energy_B(n, m, omega_c=0, omega=1):
    """two_dim_helper.py:271-272."""
    return omega * (2 * n + abs(m) + 1) - (omega_c * m) / 2
