"""``ODSincDVR`` -- one-dimensional sinc-DVR basis (mirror of the reference's
``quantum_systems/sinc_dvr/one_dim/sinc_dvr.py``).

In a DVR basis the two-body operator is diagonal, ``u[a,b,c,d] = W[a,b] delta_ac delta_bd``; the class keeps it as
the (l, l) matrix ``W`` (``u_repr = "2d"``) and overrides the four static/instance hooks of ``BasisSet`` that see
``u`` -- the reference's own demonstration that the override seam is the plugin interface.
Device side: a basis change of the 2-D form runs as two chained FP64 tensor-core GEMMs
(``ops.transform_two_body_diagonal``, csrc/structured.cu), O(l'^4 l) instead of the dense O(l^5) transform.
Host side (numpy, O(l^2)): kinetic matrix, potential, sinc functions on the grid, the shielded-Coulomb matrix.
"""

import warnings

import numpy as _numpy
import torch

from . import _arrays, ops, potentials
from .basis_set import BasisSet


def _ndim(a):
    return a.dim() if isinstance(a, torch.Tensor) else _numpy.ndim(a)


class ODSincDVR(BasisSet):
    """Sinc-DVR functions on ``linspace(-grid_length, grid_length, l)`` (sinc_dvr.py:23-91).

    Parameters
    ----------
    l : int
        Number of sinc-DVR functions (= grid points).
    grid_length : int or float
        Half-width of the grid.
    a : float, default 0.25
        Screening parameter of the shielded Coulomb interaction.
    alpha : float, default 1.0
        Strength of the shielded Coulomb interaction.
    beta : float, default 0.0
        Strength of the non-dipole ``x^2`` term of the position operator.
    potential : callable
        Confinement; defaults to ``HOPotential(omega=0.25)``.
    u_repr : "2d" or "4d"
        Storage of the two-body operator: its (l, l) diagonal, or the dense (l, l, l, l) tensor.
    """

    HOPotential = potentials.HOPotential
    DWPotential = potentials.DWPotential
    DWPotentialSmooth = potentials.DWPotentialSmooth
    SymmetricDWPotential = potentials.SymmetricDWPotential
    AsymmetricDWPotential = potentials.AsymmetricDWPotential
    GaussianPotential = potentials.GaussianPotential
    AtomicPotential = potentials.AtomicPotential

    def __init__(self, l, grid_length, a=0.25, alpha=1.0, beta=0, potential=None, u_repr="2d", **kwargs):
        if u_repr not in ("2d", "4d"):
            raise ValueError("Invalid u_repr value: '{}'".format(u_repr))
        super().__init__(l, dim=1, **kwargs)
        self.alpha = alpha
        self.a = a
        self.beta = beta
        self.grid_length = grid_length
        self.grid = _numpy.linspace(-self.grid_length, self.grid_length, self.l)
        self.num_grid_points = self.l
        if potential is None:
            potential = potentials.HOPotential(0.25)  # Zanghellini et al. frequency, sinc_dvr.py:83-87
        self.potential = potential
        self.setup_basis(u_repr)

    @property
    def sparse_repr(self):
        return self.u_repr == "2d"

    @property
    def u_repr(self):
        if self.u is None:
            return "unknown"
        return {2: "2d", 4: "4d"}.get(_ndim(self.u), "unknown")

    def setup_basis(self, u_repr):
        """``h``, ``s``, ``spf``, ``u``, ``position`` (sinc_dvr.py:99-127)."""
        self.dx = self.grid[1] - self.grid[0]
        ind = _numpy.arange(self.l)
        diff = ind[:, None] - ind
        h = _numpy.zeros((self.l, self.l), dtype=_numpy.complex128)
        off = diff != 0
        h[off] = (-1.0) ** diff[off] / (self.dx**2 * diff[off] ** 2)  # kinetic energy between grid points
        h[ind, ind] = _numpy.pi**2 / (6 * self.dx**2) + self.potential(self.grid)
        self.h = h
        self.s = self.construct_s()
        self.spf = self.construct_sinc_grid()
        self.u = self.construct_coulomb_elements(u_repr)
        self.construct_position_integrals()
        self.cast_to_complex()

    def set_u_repr(self, new_repr):
        """Reference behaviour, kept on purpose: the converted array is computed and then DROPPED
        (sinc_dvr.py:129-144 never assigns ``new_u``), so the stored representation does not change.  Use
        :meth:`converted_u` to obtain the other representation."""
        if new_repr == self.u_repr:
            print("u repr is already {}, doing nothing".format(new_repr))
        elif new_repr in ("2d", "4d"):
            self.converted_u(new_repr)
        else:
            raise ValueError("'{}' is not a valid representation".format(new_repr))

    def converted_u(self, new_repr):
        """``u`` in the requested representation (a new array of the basis' module)."""
        u = _arrays.to_device(self.u)
        ind = torch.arange(self.l, device=u.device)
        p, q = ind[:, None], ind[None, :]
        if new_repr == self.u_repr:
            out = u.clone()
        elif new_repr == "4d":
            out = torch.zeros((self.l,) * 4, dtype=u.dtype, device=u.device)
            out[p, q, p, q] = u
        elif new_repr == "2d":
            out = u[p, q, p, q].contiguous()
        else:
            raise ValueError("'{}' is not a valid representation".format(new_repr))
        return _arrays.to_module(out, self.np)

    def construct_sinc_grid(self):
        x = self.grid
        return 1 / _numpy.sqrt(self.dx) * _numpy.sinc((x - x[:, None]) / self.dx)

    def construct_position_integrals(self):
        position = _numpy.zeros((1, self.l, self.l), dtype=_numpy.complex128)
        position[0] = _numpy.diag(self.grid + self.beta * self.grid**2)
        self.position = position

    def construct_coulomb_elements(self, u_repr="4d"):
        """Shielded-Coulomb values between grid points, as the (l, l) diagonal or scattered into the dense
        tensor ``u[p,q,p,q]`` (sinc_dvr.py:154-176)."""
        x = self.grid
        w = self.alpha / _numpy.sqrt((x[:, None] - x[None, :]) ** 2 + self.a**2)
        if u_repr == "2d":
            self.u = w
        else:
            dense = _numpy.zeros((self.l,) * 4)
            ind = _numpy.arange(self.l)
            dense[ind[:, None], ind[None, :], ind[:, None], ind[None, :]] = w
            self.u = dense
        return self.u

    def construct_s(self):
        return _numpy.eye(self.l)

    def change_to_general_orbital_basis(self, anti_symmetrize=True):
        if anti_symmetrize and self.u_repr == "2d":
            if self.l > 100:
                warnings.warn("Warning, l large. Change to gos with anti_symmetrize=True forces 4d u.")
            self.set_u_repr("4d")
        return super().change_to_general_orbital_basis(anti_symmetrize=anti_symmetrize)

    def change_module(self, np):
        if self.sparse_repr:
            self.np = np
            warnings.warn("change_module not implemented for sparse u, doing nothing")
        else:
            return super().change_module(np)

    @staticmethod
    def add_spin_two_body(u, np):
        """2-D form: spin symmetry equals the DVR symmetry, every element is doubled along both axes
        (``kron(u, ones((2, 2)))``, sinc_dvr.py:200-208); 4-D form: the ``BasisSet`` kernel."""
        if _ndim(u) == 2:
            dev = _arrays.to_device(u)
            return _arrays.to_module(dev.repeat_interleave(2, dim=0).repeat_interleave(2, dim=1), np)
        return BasisSet.add_spin_two_body(u, np)

    @staticmethod
    def anti_symmetrize_u(_u):
        if _ndim(_u) == 2:
            return _u  # the 2-D form cannot hold the exchange term (sinc_dvr.py:210-215)
        return BasisSet.anti_symmetrize_u(_u)

    def transform_two_body_elements(self, u, C, np, anti_symmetrize=False, C_tilde=None):
        """Basis change of ``u``.  From the 2-D form the result is the dense 4-D tensor, optionally
        anti-symmetrised on the fly (the 2-D form cannot be); from the 4-D form it is the plain four-index
        transform (sinc_dvr.py:217-260)."""
        if self.u_repr == "2d":
            out = ops.transform_two_body_diagonal(
                _arrays.to_device(u), _arrays.to_device(C), _arrays.to_device(C_tilde), anti_symmetrize=anti_symmetrize
            )
            return _arrays.to_module(out, np)
        assert not anti_symmetrize, "antisymmetrize only valid for sparse storage of u"
        return BasisSet.transform_two_body_elements(u, C, np, C_tilde)

    def change_basis(self, *args, **kwargs):
        super().change_basis(*args, **kwargs)
