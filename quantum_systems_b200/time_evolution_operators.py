"""Time-dependent pieces of the Hamiltonian (mirror of reference time_evolution_operators/operator.py:4-220).

An operator is attached with ``QuantumSystem.set_time_evolution_operator`` and contributes ``h_t(t)`` and/or
``u_t(t)`` to ``QuantumSystem.h_t`` / ``u_t`` (reference system.py:144-215).  Every contribution here is a scalar
weight times a stored array, so besides the reference's ``h_t`` / ``u_t`` (which return the scaled array -- one pass
of ``qs_scale_add`` for device, host and sharded storage alike) each operator exposes the pair itself through
``h_t_scaled`` / ``u_t_scaled``; ``QuantumSystem`` uses it to fold ``u_0 + f(t) u`` into ONE pass over the n^4 tensor.
User-defined operators that only implement the reference interface keep working.
"""

import abc


def _as_function(value):
    return value if callable(value) else (lambda t, _v=value: _v)


class TimeEvolutionOperator(metaclass=abc.ABCMeta):
    """Base class (operator.py:4-84): no contribution to either part unless overridden."""

    @property
    def is_one_body_operator(self):
        return False

    @property
    def is_two_body_operator(self):
        return False

    def set_system(self, system):
        self._system = system
        return self

    def h_t(self, current_time):
        return 0

    def u_t(self, current_time):
        return 0


class _ScaledOneBody(TimeEvolutionOperator):
    @property
    def is_one_body_operator(self):
        return True

    def h_t(self, current_time):
        from .system import scaled_sum

        return scaled_sum(self._system.np, [self.h_t_scaled(current_time)])


class DipoleFieldInteraction(_ScaledOneBody):
    r"""``h_I(t) = -E(t) eps(t) . d`` in the length gauge, ``+E(t) eps(t) . p (+ E(t)^2 / 2)`` in the velocity
    gauge (operator.py:87-178).  ``polarization_vector`` defaults to the first axis."""

    def __init__(self, field_strength, polarization_vector=None, gauge="length", quadratic_term=True):
        assert gauge in ["length", "velocity"], "gauge must be either length or velocity."
        self._length_gauge = gauge == "length"
        self._quadratic_term = quadratic_term
        self._field_strength = _as_function(field_strength)
        self._polarization = polarization_vector

    def _terms(self, current_time):
        import numpy

        vector = self._system.dipole_moment if self._length_gauge else self._system.momentum
        if self._polarization is None:
            eps = numpy.zeros(vector.shape[0])
            eps[0] = 1
            self._polarization = eps
        eps = numpy.asarray(_as_function(self._polarization)(current_time))
        field = self._field_strength(current_time)
        sign = -1.0 if self._length_gauge else 1.0
        return [(sign * field * e.item(), vector[i]) for i, e in enumerate(eps) if e != 0], field

    def h_t_scaled(self, current_time):
        terms, _ = self._terms(current_time)
        if len(terms) == 1 and (self._length_gauge or not self._quadratic_term):
            return terms[0]
        return 1.0, self.h_t(current_time)

    def h_t(self, current_time):
        from .system import scaled_sum

        np = self._system.np
        terms, field = self._terms(current_time)
        if not self._length_gauge and self._quadratic_term:
            terms.append((0.5 * field**2, np.eye(self._system.l)))
        if not terms:
            return np.zeros_like(self._system.h)
        return scaled_sum(np, terms)


class AdiabaticSwitching(TimeEvolutionOperator):
    """``u(t) = f(t) u`` (operator.py:181-196)."""

    def __init__(self, switching_function):
        self._switching_function = _as_function(switching_function)

    @property
    def is_two_body_operator(self):
        return True

    def u_t_scaled(self, current_time):
        return self._switching_function(current_time), self._system.u

    def u_t(self, current_time):
        from .system import scaled_sum

        return scaled_sum(self._system.np, [self.u_t_scaled(current_time)])


class CustomOneBodyOperator(_ScaledOneBody):
    """``h(t) = w(t) O`` for a stored one-body matrix ``O`` (operator.py:199-217)."""

    def __init__(self, weight, operator):
        self._weight = _as_function(weight)
        self._operator = operator

    def h_t_scaled(self, current_time):
        return self._weight(current_time), self._operator
