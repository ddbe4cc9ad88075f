"""One-dimensional confinement potentials for ``ODQD`` (host side, O(G) work).

Same classes, constructor arguments and formulas as the reference's
``quantum_systems/quantum_dots/one_dim/one_dim_potentials.py`` (:14-159); they run on the host in
numpy because the finite-difference eigenproblem they feed is solved on the host as well.
"""

import abc

import numpy as _np


class OneDimPotential(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def __call__(self, x):
        """Potential energy on the grid points ``x``."""

    def derivative(self, x):
        raise NotImplementedError()


class HOPotential(OneDimPotential):
    """Harmonic oscillator ``omega^2 x^2 / 2`` (one_dim_potentials.py:14-22)."""

    def __init__(self, omega):
        self.omega = omega

    def __call__(self, x):
        return 0.5 * self.omega**2 * x**2

    def derivative(self, x):
        return self.omega**2 * x


class DWPotential(HOPotential):
    """Cusped double well: HO plus ``omega^2 (l^2/4 - l|x|) / 2`` (one_dim_potentials.py:25-43)."""

    def __init__(self, omega, l):
        super().__init__(omega)
        self.l = l

    def __call__(self, x):
        well = 0.25 * self.l**2 - self.l * abs(x)
        return super().__call__(x) + 0.5 * self.omega**2 * well

    def derivative(self, x):
        return super().derivative(x) - self.l * self.omega**2 * (_np.heaviside(x, 0.5) - 0.5)


class DWPotentialSmooth(OneDimPotential):
    """Quartic double well ``(x + a/2)^2 (x - a/2)^2 / (2 a^2)`` (one_dim_potentials.py:46-74)."""

    def __init__(self, a=4):
        self.a = a

    def __call__(self, x):
        half = 0.5 * self.a
        return (1.0 / (2 * self.a**2)) * (x + half) ** 2 * (x - half) ** 2

    def derivative(self, x):
        half = 0.5 * self.a
        return ((x + half) * (x - half) ** 2 + (x - half) * (x + half) ** 2) / self.a**2


class SymmetricDWPotential(OneDimPotential):
    """``a x^6 + b x^4 + c x^2`` (one_dim_potentials.py:77-92)."""

    def __init__(self, a=0.5, b=1, c=-7):
        self.a, self.b, self.c = a, b, c

    def __call__(self, x):
        return self.a * x**6 + self.b * x**4 + self.c * x**2

    def derivative(self, x):
        # as written in the reference (:91-92), including its 3 b x^3 term
        return 6 * self.a * x**5 + 3 * self.b * x**3 + 2 * self.c * x


class AsymmetricDWPotential(OneDimPotential):
    """``a x^4 + b x^3 + c x^2`` (one_dim_potentials.py:95-110)."""

    def __init__(self, a=1, b=1, c=-2.5):
        self.a, self.b, self.c = a, b, c

    def __call__(self, x):
        return self.a * x**4 + self.b * x**3 + self.c * x**2

    def derivative(self, x):
        return 4 * self.a * x**3 + 3 * self.b * x**2 + 2 * self.c * x


class GaussianPotential(OneDimPotential):
    """``-weight exp(-(x - center)^2 / (2 deviation^2))`` (one_dim_potentials.py:113-127)."""

    def __init__(self, weight, center, deviation, np=None):
        self.weight = weight
        self.center = center
        self.deviation = deviation
        self.np = _np if np is None else np

    def __call__(self, x):
        return -self.weight * self.np.exp(-((x - self.center) ** 2) / (2.0 * self.deviation**2))

    def derivative(self, x):
        return -(x - self.center) / self.deviation**2 * self(x)


class GaussianPotentialHardWall(OneDimPotential):
    """Gaussian well plus a 1e5 wall beyond ``|x| > x_wall`` (one_dim_potentials.py:130-149)."""

    def __init__(self, weight, center, deviation, x_wall):
        self.weight = weight
        self.center = center
        self.deviation = deviation
        self.x_wall = x_wall

    def __call__(self, x):
        wall = _np.where(_np.abs(x) > self.x_wall, 1e5, 0.0)
        return -self.weight * _np.exp(-((x - self.center) ** 2) / (2.0 * self.deviation**2)) + wall


class AtomicPotential(OneDimPotential):
    """Soft-Coulomb atom ``-Za / sqrt(x^2 + c)`` (one_dim_potentials.py:152-159)."""

    def __init__(self, Za=2, c=0.54878464):
        self.Za = Za
        self.c = c

    def __call__(self, x):
        return -self.Za / _np.sqrt(x**2 + self.c)

    def derivative(self, x):
        return self.Za * x / (x**2 + self.c) ** (3 / 2)
