"""``SpatialOrbitalSystem`` -- closed-shell spatial orbitals (mirror of reference spatial_orbital_system.py)."""

import copy

from .general_orbital_system import GeneralOrbitalSystem, _construct_fock, _scalar
from .system import QuantumSystem


class SpatialOrbitalSystem(QuantumSystem):
    r"""Spatial orbitals, doubly occupied: ``n`` particles occupy ``n // 2`` basis functions
    (spatial_orbital_system.py:40-50).

    >>> # spas = SpatialOrbitalSystem(4, RandomBasisSet(20, 2)); spas.n == 2
    """

    def __init__(self, n, basis_set, **kwargs):
        assert n % 2 == 0, "n must be divisable by 2 to be a closed-shell system"
        assert not basis_set.includes_spin, (
            f"{self.__class__.__name__} only supports basis sets without spin-dependence."
        )
        super().__init__(n // 2, basis_set, **kwargs)

    def construct_general_orbital_system(self, a=[1, 0], b=[0, 1], anti_symmetrize=True):
        r"""Spin-double a COPY of the basis set and wrap it in a ``GeneralOrbitalSystem`` with
        ``2 n`` occupied spin-orbitals (spatial_orbital_system.py:52-104)."""
        gos = GeneralOrbitalSystem(
            self.n * 2, self._basis_set.copy_basis(), a=a, b=b, anti_symmetrize=anti_symmetrize
        )
        if self._time_evolution_operator is not None:
            gos.set_time_evolution_operator(copy.deepcopy(self._time_evolution_operator))
        return gos

    def compute_reference_energy(self, h=None, u=None):
        r"""``E_0 = 2 h_ii + 2 u_ijij - u_ijji + E_n`` (spatial_orbital_system.py:106-148)."""
        h = self.h if h is None else h
        u = self.u if u is None else u
        tr_h, direct, exchange = self._occupied_trace_terms(h, u)
        return _scalar(2 * tr_h + 2 * direct - exchange + self.nuclear_repulsion_energy)

    def construct_fock_matrix(self, h, u, f=None):
        r"""Restricted Fock matrix ``f_pq = h_pq + 2 u_piqi - u_piiq`` over occupied ``i``; ``u`` is NOT
        anti-symmetrised.  ``f`` is filled in place when given (spatial_orbital_system.py:150-190)."""
        return _construct_fock(self, h, u, f, spatial=True)

    def change_to_hf_basis(self, *args, **kwargs):
        raise NotImplementedError("There is currently no RHF implementation")
