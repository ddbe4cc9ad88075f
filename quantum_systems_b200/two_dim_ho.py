"""Two-dimensional quantum dots in the polar harmonic-oscillator basis (mirror of the reference's
``quantum_systems/quantum_dots/two_dim/two_dim_ho.py`` and ``two_dim_helper.py``).

Device side: the l^4 Coulomb elements (reference ``coulomb_ho``, coulomb_elements.py:6-92, evaluated for
every index tuple by two_dim_helper.py:250-268 / :283-300) run in one CUDA kernel (``ops.tdho_coulomb``,
csrc/tdho.cu).  Host side (O(l^2) scalars, numpy): shell bookkeeping, one-body elements, position
integrals and the grid orbitals.  The reference integrates the radial functions symbolically with sympy
(two_dim_helper.py:53-70); here the same integrals are the closed form

    int_0^inf r^(1+k) R_p(r) R_q(r) dr = 1 / (2 a^(2+k)) sum_ij c_i(n_p,|m_p|) c_j(n_q,|m_q|) Gamma(b + i + j + 1),
    b = (k + |m_p| + |m_q|) / 2,   L_n^alpha(x) = sum_i c_i x^i,   c_i = (-1)^i binom(n + alpha, n - i) / i!,

which is what sympy returns, evaluated in a few microseconds instead of seconds.
"""

import math

import numpy as _numpy
import scipy.special

from . import _arrays, ops
from .basis_set import BasisSet


# ------------------------------------------------------------------------------------------------
# shell bookkeeping (reference two_dim_helper.py:111-182)
# ------------------------------------------------------------------------------------------------
def get_indices_nm(p):
    """Orbital index -> (n, m).  Shell R = 2n + |m| holds R + 1 states, listed with m ascending."""
    p = int(p)
    shell = (math.isqrt(8 * p + 1) - 1) // 2
    m = -shell + 2 * (p - shell * (shell + 1) // 2)
    return (shell - abs(m)) // 2, m


def get_index_p(n, m):
    """(n, m) -> orbital index; inverse of :func:`get_indices_nm`."""
    shell = 2 * int(n) + abs(int(m))
    return shell * (shell + 1) // 2 + (int(m) + shell) // 2


def get_shell_energy(n, m):
    return 2 * n + abs(m) + 1


def get_one_body_elements(num_orbitals):
    """Diagonal matrix of shell energies in units of omega (two_dim_helper.py:174-182)."""
    h = _numpy.zeros((num_orbitals, num_orbitals))
    for p in range(num_orbitals):
        h[p, p] = get_shell_energy(*get_indices_nm(p))
    return h


def get_coulomb_elements(num_orbitals, np=None, scale=1.0, quantum_numbers=None):
    """``u[p,q,r,s] = scale * <pq|1/r_12|rs>`` in the oscillator basis, computed on the GPU
    (two_dim_helper.py:185-268).  ``quantum_numbers=(n, m)`` overrides the shell ordering
    (the B-field variant, :283-300).  Returns an array of the module ``np`` (device tensor by default)."""
    if quantum_numbers is None:
        nm = _numpy.array([get_indices_nm(p) for p in range(num_orbitals)], dtype=_numpy.int64).reshape(-1, 2)
        n, m = nm[:, 0], nm[:, 1]
    else:
        n, m = quantum_numbers
    u = ops.tdho_coulomb(n, m, scale)
    return _arrays.to_module(u, _arrays.default_module() if np is None else np)


# ------------------------------------------------------------------------------------------------
# single-particle functions and their integrals (reference two_dim_helper.py:15-108)
# ------------------------------------------------------------------------------------------------
def bohr_radius(mass, omega):
    return math.sqrt(mass * omega)


def spf_norm(n, m, mass, omega):
    return bohr_radius(mass, omega) * math.sqrt(math.factorial(n) / (math.pi * math.factorial(n + abs(m))))


def spf_state(r, theta, n, m, mass, omega):
    """Normalised Fock-Darwin orbital on a polar mesh (two_dim_helper.py:15-50)."""
    a = bohr_radius(mass, omega)
    x = (a * r) ** 2
    radial = (a * r) ** abs(m) * scipy.special.eval_genlaguerre(n, abs(m), x) * _numpy.exp(-x / 2.0)
    return spf_norm(n, m, mass, omega) * _numpy.exp(1j * m * theta) * radial


def _laguerre_coefficients(n, alpha):
    return [(-1) ** i * math.comb(n + alpha, n - i) / math.factorial(i) for i in range(n + 1)]


def radial_integral(n_p, m_p, n_q, m_q, mass, omega, order=1):
    """``int_0^inf r r^order R_p(r) R_q(r) dr`` for the un-normalised radial functions
    ``R(r) = (a r)^|m| L_n^|m|(a^2 r^2) exp(-a^2 r^2 / 2)`` (two_dim_helper.py:53-70)."""
    a = bohr_radius(mass, omega)
    beta = 0.5 * (order + abs(m_p) + abs(m_q))
    total = 0.0
    for i, c_i in enumerate(_laguerre_coefficients(n_p, abs(m_p))):
        for j, c_j in enumerate(_laguerre_coefficients(n_q, abs(m_q))):
            total += c_i * c_j * math.gamma(beta + i + j + 1)
    return total / (2.0 * a ** (2 + order))


def theta_1_integral(m_p, m_q):
    """``int_0^2pi cos(theta) exp(i (m_q - m_p) theta) dtheta``."""
    return math.pi if abs(m_p - m_q) == 1 else 0


def theta_2_integral(m_p, m_q):
    """``int_0^2pi sin(theta) exp(i (m_q - m_p) theta) dtheta``."""
    return -(m_p - m_q) * 1j * math.pi if abs(m_p - m_q) == 1 else 0


def theta_1_tilde_integral(m_p, m_q):
    """``int_0^2pi |cos(theta)| exp(i (m_q - m_p) theta) dtheta`` (two_dim_helper.py:95-101)."""
    d = m_p - m_q
    if d % 2:
        return 0
    sign = 1 if (abs(d) // 2) % 2 == 0 else -1
    return sign * 4 / (1 - d**2)


def theta_2_tilde_integral(m_p, m_q):
    """``int_0^2pi |sin(theta)| exp(i (m_q - m_p) theta) dtheta`` (two_dim_helper.py:104-108)."""
    d = m_p - m_q
    if d % 2:
        return 0
    return 4 / (1 - d**2)


def smooth_theta_integral_1(m_p, m_q):
    return (-1) ** (abs(m_p - m_q) % 2) * (3 * math.pi / 4)


def smooth_theta_integral_2(m_p, m_q):
    return (-1) ** (abs(m_p - m_q) % 2) * math.pi


def get_double_well_one_body_elements(num_orbitals, omega, mass, barrier_strength, dtype=_numpy.float64, axis=0,
                                      indices_nm=get_indices_nm):
    """Oscillator + ``omega^2 (b^2/4 - b |x_axis|) / 2`` barrier in the oscillator basis
    (two_dim_helper.py:303-338)."""
    h = _numpy.zeros((num_orbitals, num_orbitals), dtype=dtype)
    theta_tilde = theta_1_tilde_integral if axis == 0 else theta_2_tilde_integral
    for p in range(num_orbitals):
        n_p, m_p = indices_nm(p)
        h[p, p] += omega * get_shell_energy(n_p, m_p) + omega**2 * barrier_strength**2 / 8.0
        for q in range(num_orbitals):
            n_q, m_q = indices_nm(q)
            if abs(m_p - m_q) == 1:
                continue
            angular = theta_tilde(m_p, m_q)
            if angular == 0:
                continue
            h[p, q] -= (
                0.5 * omega**2 * barrier_strength
                * spf_norm(n_p, m_p, mass, omega) * spf_norm(n_q, m_q, mass, omega)
                * radial_integral(n_p, m_p, n_q, m_q, mass, omega)
                * angular
            )
    return h


def get_smooth_double_well_one_body_elements(num_orbitals, omega, mass, a=2, b=2, dtype=_numpy.float64,
                                             indices_nm=get_indices_nm):
    """Quartic smooth double well in the oscillator basis (two_dim_helper.py:341-377)."""
    h = _numpy.zeros((num_orbitals, num_orbitals), dtype=dtype)
    prefactor = omega**2 / 4
    for p in range(num_orbitals):
        n_p, m_p = indices_nm(p)
        h[p, p] += omega * get_shell_energy(n_p, m_p) + omega**2 * a**2 / 64
        for q in range(num_orbitals):
            n_q, m_q = indices_nm(q)
            norms = spf_norm(n_p, m_p, mass, omega) * spf_norm(n_q, m_q, mass, omega)
            h[p, q] += (
                prefactor * (1 / a**2) * norms
                * radial_integral(n_p, m_p, n_q, m_q, mass, omega, order=4) * smooth_theta_integral_1(m_p, m_q)
            )
            h[p, q] -= (
                prefactor * ((5 * b) / 2) * norms
                * radial_integral(n_p, m_p, n_q, m_q, mass, omega, order=2) * smooth_theta_integral_2(m_p, m_q)
            )
    return h


# ------------------------------------------------------------------------------------------------
# magnetic-field level ordering (reference two_dim_helper.py:271-280, :380-415)
# ------------------------------------------------------------------------------------------------
def get_shell_energy_B(n, m, omega_c=0, omega=1):
    return omega * (2 * n + abs(m) + 1) - (omega_c * m) / 2


def energy_sorted_levels(n_array, m_array, omega_c=0, omega=1):
    """Fock-Darwin levels ``(n, m, E)`` for all n in ``n_array``, m in ``m_array``, sorted by (E, m) and
    cut after the last complete degenerate level beyond ``len(n_array) (len(n_array) - 1) / 2`` states --
    the table the reference keeps in a pandas frame (two_dim_helper.py:380-415).  Returns int arrays
    ``n``, ``m`` and the float array ``E``."""
    nn, mm = _numpy.meshgrid(_numpy.asarray(n_array), _numpy.asarray(m_array), indexing="ij")
    nn, mm = nn.ravel(), mm.ravel()
    energy = _numpy.array([get_shell_energy_B(n, m, omega_c=omega_c, omega=omega) for n, m in zip(nn, mm)], dtype=float)
    order = _numpy.lexsort((mm, energy))  # primary key E, ties by m; stable
    nn, mm, energy = nn[order], mm[order], energy[order]

    level = _numpy.zeros(len(energy), dtype=int)
    seen = []
    for e in _numpy.round(energy, 8):
        if e not in seen:
            seen.append(e)
    for i, e in enumerate(seen):
        level[_numpy.abs(energy - e) < 1e-6] = i

    keep = len(n_array) * (len(n_array) - 1) // 2
    cap = level[keep]
    while level[keep] == cap:
        keep += 1
    return nn[:keep].astype(_numpy.int64), mm[:keep].astype(_numpy.int64), energy[:keep]


# ------------------------------------------------------------------------------------------------
# systems
# ------------------------------------------------------------------------------------------------
class TwoDimensionalHarmonicOscillator(BasisSet):
    """Two-dimensional harmonic oscillator in polar coordinates (two_dim_ho.py:23-139).

    Parameters
    ----------
    l : int
        Number of (spatial) basis functions.
    radius_length : float
        Extent of the radial grid the orbitals are tabulated on.
    num_grid_points : int
        Number of radial and of angular grid points.
    omega : float, default 1
        Oscillator frequency.
    mass : float, default 1
        Particle mass.
    """

    def __init__(self, l, radius_length, num_grid_points, omega=1, mass=1, verbose=False, **kwargs):
        super().__init__(l, dim=2, **kwargs)
        self.omega = omega
        self.mass = mass
        self.verbose = verbose
        self.radius_length = radius_length
        self.num_grid_points = num_grid_points
        self.radius = _numpy.linspace(0, self.radius_length, self.num_grid_points)
        self.theta = _numpy.linspace(0, 2 * _numpy.pi, self.num_grid_points)
        self.setup_basis()

    def get_indices_nm(self, p):
        return get_indices_nm(p)

    def _quantum_numbers(self):
        nm = _numpy.array([self.get_indices_nm(p) for p in range(self.l)], dtype=_numpy.int64).reshape(-1, 2)
        return nm[:, 0].copy(), nm[:, 1].copy()

    def setup_basis(self):
        """``h``, ``u``, ``s``, ``spf``, ``position`` (two_dim_ho.py:84-98)."""
        self.h = self.omega * get_one_body_elements(self.l)
        self.u = _arrays.to_module(
            ops.tdho_coulomb(*self._quantum_numbers(), scale=math.sqrt(self.omega)), self.np
        )
        self.s = _numpy.eye(self.l)
        self.setup_spf()
        self.construct_position_integrals()

    def setup_spf(self):
        """Orbitals on the (theta, r) mesh, ``spf[p, i_theta, i_r]`` (two_dim_ho.py:100-111)."""
        self.R, self.T = _numpy.meshgrid(self.radius, self.theta)
        spf = _numpy.zeros((self.l, self.num_grid_points, self.num_grid_points), dtype=_numpy.complex128)
        for p in range(self.l):
            spf[p] = spf_state(self.R, self.T, *self.get_indices_nm(p), self.mass, self.omega)
        self.spf = spf

    def construct_position_integrals(self):
        """``<p| x |q>`` and ``<p| y |q>``: non-zero for ``|m_p - m_q| = 1`` only (two_dim_ho.py:116-139)."""
        position = _numpy.zeros((2, self.l, self.l), dtype=_numpy.complex128)
        for p in range(self.l):
            n_p, m_p = self.get_indices_nm(p)
            for q in range(self.l):
                n_q, m_q = self.get_indices_nm(q)
                if abs(m_p - m_q) != 1:
                    continue
                scale = (
                    spf_norm(n_p, m_p, self.mass, self.omega) * spf_norm(n_q, m_q, self.mass, self.omega)
                    * radial_integral(n_p, m_p, n_q, m_q, self.mass, self.omega)
                )
                position[0, p, q] = scale * theta_1_integral(m_p, m_q)
                position[1, p, q] = scale * theta_2_integral(m_p, m_q)
        self.position = position


class TwoDimensionalDoubleWell(TwoDimensionalHarmonicOscillator):
    """Oscillator basis with a double-well barrier along ``axis`` (two_dim_ho.py:142-189)."""

    def __init__(self, *args, barrier_strength=1, axis=0, **kwargs):
        self.barrier_strength = barrier_strength
        self.axis = axis
        super().__init__(*args, **kwargs)

    def setup_basis(self):
        super().setup_basis()
        self.h = get_double_well_one_body_elements(
            self.l, self.omega, self.mass, self.barrier_strength, dtype=_numpy.complex128, axis=self.axis
        )


class TwoDimSmoothDoubleWell(TwoDimensionalHarmonicOscillator):
    """Oscillator basis with the smooth quartic double well (two_dim_ho.py:192-211).  The reference sets
    ``a`` and ``b`` only after the parent constructor has already needed them (``AttributeError`` at
    two_dim_ho.py:206-208); here they are set first, which is evidently what was meant."""

    def __init__(self, *args, a=2, b=2, **kwargs):
        self.a = a
        self.b = b
        super().__init__(*args, **kwargs)

    def setup_basis(self):
        super().setup_basis()
        self.h = get_smooth_double_well_one_body_elements(
            self.l, self.omega, self.mass, a=self.a, b=self.b, dtype=_numpy.complex128
        )


class TwoDimHarmonicOscB(TwoDimensionalHarmonicOscillator):
    """Two-dimensional oscillator in a homogeneous magnetic field of cyclotron frequency ``omega_c``:
    Fock-Darwin levels ordered by energy (two_dim_ho.py:214-281)."""

    def __init__(self, *args, omega_c=0, **kwargs):
        self.omega_c = omega_c
        super().__init__(*args, **kwargs)

    def setup_basis(self):
        self.omega = math.sqrt(self.omega**2 + self.omega_c**2 / 4)
        n_array = _numpy.arange(self.l)
        m_array = _numpy.arange(-self.l - 5, self.l + 6)
        self.level_n, self.level_m, self.level_energy = energy_sorted_levels(
            n_array, m_array, omega_c=self.omega_c, omega=self.omega
        )
        self.h = _numpy.diag(self.level_energy[: self.l])
        self.s = _numpy.eye(self.l)
        self.u = _arrays.to_module(
            ops.tdho_coulomb(self.level_n[: self.l], self.level_m[: self.l], scale=math.sqrt(self.omega)), self.np
        )
        self.setup_spf()
        self.construct_position_integrals()
        self.cast_to_complex()

    def get_indices_nm(self, p):
        return int(self.level_n[p]), int(self.level_m[p])
