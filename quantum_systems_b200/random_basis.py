"""``RandomBasisSet`` -- random matrix elements with the symmetries of second-quantised integrals
(mirror of reference random_basis.py:4-69).  The universal synthetic fixture of the test-suite."""

import numpy as _numpy

from .basis_set import BasisSet


class RandomBasisSet(BasisSet):
    """Random Hermitian ``h``, ``s``; ``u`` with ``u_pqrs = u_qpsr``; Hermitian position components.

    Random numbers are drawn on the host from numpy's global stream in the same order as the
    reference (h, s, u, position, nuclear repulsion, charge), then stored in the basis set's module.
    """

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.setup_basis()

    def setup_basis(self):
        l, dim = self.l, self.dim
        self.h = self.make_hermitian(self.get_random_elements((l, l), _numpy))
        self.s = self.make_hermitian(self.get_random_elements((l, l), _numpy))
        self.u = self.make_two_body_symmetry(self.get_random_elements((l, l, l, l), _numpy))
        self.position = self.make_position_elements_hermitian(self.get_random_elements((dim, l, l), _numpy))
        self.nuclear_repulsion_energy = _numpy.random.random()
        self.charge = _numpy.random.choice([-1, 1])

    @staticmethod
    def make_hermitian(h):
        return 0.5 * (h + h.conj().T)

    @staticmethod
    def make_position_elements_hermitian(position):
        for i in range(len(position)):
            position[i] = RandomBasisSet.make_hermitian(position[i])
        return position

    @staticmethod
    def make_two_body_symmetry(u):
        return 0.5 * (u + u.transpose(1, 0, 3, 2))

    @staticmethod
    def get_random_elements(shape, np):
        """Complex array ``random(shape) + 1j random(shape)`` from ``np.random`` (random_basis.py:53-69)."""
        return np.random.random(shape) + 1j * np.random.random(shape)
