"""``ODQD`` -- one-dimensional quantum dot on a grid (mirror of reference
quantum_dots/one_dim/one_dim_qd.py:169-289).

Host side (numpy/scipy, O(G l)): grid, potential, tridiagonal eigenproblem -- exactly the reference's
recipe.  Device side: the O(l^2 G^2 + l^4 G) shielded-Coulomb build ``u_abcd`` runs as two chained
FP64 tensor-core GEMMs (``ops.odqd_coulomb``) instead of ``np.einsum``.
"""

import numpy as _numpy
import scipy.linalg
import scipy.special

import torch

from . import _arrays, ops, potentials
from .basis_set import BasisSet


def grid_orbitals(l, grid_length, num_grid_points, potential):
    """Host part of the ODQD recipe (one_dim_qd.py:258-266): the grid, and the ``l`` lowest eigenpairs of the
    finite-difference Hamiltonian on the interior points.  Returns ``(grid, eps, C)`` with ``C`` of shape
    ``(num_grid_points - 2, l)``; O(G l) work on the host."""
    grid = _numpy.linspace(-grid_length, grid_length, num_grid_points)
    inner = grid[1:-1]
    dx = grid[1] - grid[0]
    diagonal = 1.0 / (dx**2) + potential(inner)
    off_diagonal = -1.0 / (2 * dx**2) * _numpy.ones(num_grid_points - 3)
    eps, C = scipy.linalg.eigh_tridiagonal(diagonal, off_diagonal, select="i", select_range=(0, l - 1))
    return grid, eps, C


class ODQD(BasisSet):
    """Create a 1-D quantum-dot basis of the ``l`` lowest eigenfunctions of ``potential`` on
    ``linspace(-grid_length, grid_length, num_grid_points)``.

    Parameters (one_dim_qd.py:172-186)
    ----------
    l : int
        Number of basis functions.
    grid_length : int or float
        Half-width of the grid.
    num_grid_points : int
        Number of grid points (the two end points carry zero amplitude).
    a : float, default 0.25
        Screening parameter of the shielded Coulomb interaction.
    alpha : float, default 1.0
        Strength of the shielded Coulomb interaction.
    beta : float, default 0.0
        Strength of the non-dipole ``x^2`` term in the position operator.
    potential : callable
        Confinement; defaults to ``HOPotential(omega=0.25)``.
    """

    HOPotential = potentials.HOPotential
    DWPotential = potentials.DWPotential
    DWPotentialSmooth = potentials.DWPotentialSmooth
    SymmetricDWPotential = potentials.SymmetricDWPotential
    AsymmetricDWPotential = potentials.AsymmetricDWPotential
    GaussianPotential = potentials.GaussianPotential
    AtomicPotential = potentials.AtomicPotential

    def __init__(self, l, grid_length, num_grid_points, a=0.25, alpha=1.0, beta=0, potential=None, **kwargs):
        super().__init__(l, dim=1, **kwargs)
        self.a = a
        self.alpha = alpha
        self.grid_length = grid_length
        self.num_grid_points = num_grid_points
        self.grid = _numpy.linspace(-self.grid_length, self.grid_length, self.num_grid_points)
        self.beta = beta
        if potential is None:
            potential = potentials.HOPotential(0.25)  # Zanghellini et al. frequency, one_dim_qd.py:248-252
        self.potential = potential
        self.setup_basis()

    def setup_basis(self):
        """Fill ``h, s, u, spf, position`` (one_dim_qd.py:258-289)."""
        inner = self.grid[1:-1]
        dx = self.grid[1] - self.grid[0]

        # finite-difference Hamiltonian on the interior points; l lowest eigenpairs (host, O(G l))
        _, eps, C = grid_orbitals(self.l, self.grid_length, self.num_grid_points, self.potential)
        self.eigen_energies = eps

        spf = _numpy.zeros((self.l, self.num_grid_points), dtype=_numpy.complex128)
        spf[:, 1:-1] = C.T / _numpy.sqrt(dx)
        self.spf = spf
        self.h = _numpy.diag(eps).astype(_numpy.complex128)
        self.s = _numpy.eye(self.l)

        # shielded-Coulomb two-body elements: two chained DMMA GEMMs on the device
        u = ops.odqd_coulomb(_arrays.to_device(C), _arrays.to_device(inner), self.alpha, self.a)
        self.u = _arrays.to_module(u, self.np)

        position = _numpy.zeros((1, self.l, self.l), dtype=_numpy.complex128)
        position[0] = (C.T * (inner + self.beta * inner**2)) @ C  # <a| x + beta x^2 |b>, O(G l^2) host
        self.position = position


class ODHO(BasisSet):
    """Harmonic-oscillator eigenfunctions on a grid with trapezoid-rule Coulomb integrals (mirror of reference
    quantum_dots/one_dim/one_dim_qd.py:71-166; not exported by the reference package either).

    ``u[p,q,r,s] = trapz_i( spf_p spf_r(x_i) trapz_j( spf_q spf_s(x_j) W(x_i, x_j) ) )`` -- the reference's two numba
    loop nests (``_compute_inner_integral`` :35-51, ``_compute_orbital_integrals`` :54-68).  With the trapezoid
    weights ``w`` (dx, end points halved, :19-27) folded into the orbital rows, ``L[i, p] = sqrt(w_i) spf_p(x_i)``,
    this is exactly the two-GEMM grid build of ``ODQD`` on the FULL grid,
    ``u_abcd = sum_ij L_ia L_ic W_ij L_jb L_jd`` (``ops.odqd_coulomb``): the inner integral is never stored.

    >>> # odho = ODHO(20, 11, 201, omega=1); odho.l == 20; abs(0.5 - odho.h[0, 0]) == 0
    """

    def __init__(self, l, grid_length, num_grid_points, omega=0.25, a=0.25, alpha=1.0, beta=0, **kwargs):
        super().__init__(l, dim=1, **kwargs)
        self.omega = omega
        self.a = a
        self.alpha = alpha
        self.grid_length = grid_length
        self.num_grid_points = num_grid_points
        self.grid = _numpy.linspace(-self.grid_length, self.grid_length, self.num_grid_points)
        self.beta = beta
        self.setup_basis()

    def setup_basis(self):
        """Fill ``h, s, spf, u, position`` (one_dim_qd.py:115-134)."""
        dx = self.grid[1] - self.grid[0]
        self.eigen_energies = self.omega * (_numpy.arange(self.l) + 0.5)
        self.h = _numpy.diag(self.eigen_energies).astype(_numpy.complex128)
        self.s = _numpy.eye(self.l)
        spf = _numpy.stack([self.ho_function(self.grid, p) for p in range(self.l)])
        self.spf = spf

        weights = _numpy.full(self.num_grid_points, dx)
        weights[0] *= 0.5
        weights[-1] *= 0.5
        L = _numpy.ascontiguousarray((spf * _numpy.sqrt(weights)).T)  # (G, l): trapezoid weights in the rows
        u = ops.odqd_coulomb(_arrays.to_device(L), _arrays.to_device(self.grid), self.alpha, self.a)
        self.u = _arrays.to_module(u.to(torch.complex128), self.np)  # the reference allocates u as complex128

        self.construct_position_integrals()

    def ho_function(self, x, n):
        """``N_n exp(-omega x^2 / 2) H_n(sqrt(omega) x)`` (one_dim_qd.py:136-141)."""
        return (
            self.normalization(n)
            * _numpy.exp(-0.5 * self.omega * x**2)
            * scipy.special.eval_hermite(n, _numpy.sqrt(self.omega) * x)
        )

    def normalization(self, n):
        """``(omega / pi)^(1/4) / sqrt(2^n n!)`` (one_dim_qd.py:143-148)."""
        return 1.0 / _numpy.sqrt(2.0**n * scipy.special.factorial(n)) * (self.omega / _numpy.pi) ** 0.25

    def construct_position_integrals(self):
        """Analytic ``<n| x |n+1> = N_n N_{n+1} (n+1) sqrt(pi) 2^n n! / omega`` (one_dim_qd.py:150-166)."""
        position = _numpy.zeros((1, self.l, self.l))
        for n in range(self.l - 1):
            pos = (
                self.normalization(n) * self.normalization(n + 1) * (n + 1) * _numpy.sqrt(_numpy.pi) * 2.0**n
                * scipy.special.factorial(n) / self.omega
            )
            position[0, n, n + 1] = position[0, n + 1, n] = pos
        self.position = position
