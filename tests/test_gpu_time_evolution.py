"""``QuantumSystem.h_t`` / ``u_t`` with time-evolution operators (reference system.py:144-215,
time_evolution_operators/operator.py), re-expressing the reference's tests/test_time_evolution_operators.py on the
drop-in API in both storage modes.  Every sum runs through ``qs_scale_add``; results are compared with the plain
numpy expressions the reference evaluates."""

import numpy as np
import pytest

from conftest import assert_close_scaled

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


@pytest.fixture(params=["numpy", "xp"])
def module(request):
    from quantum_systems_b200 import xp

    return {"numpy": np, "xp": xp}[request.param]


def systems(module, n=4, l=10, dim=3):
    from quantum_systems_b200 import GeneralOrbitalSystem, RandomBasisSet, SpatialOrbitalSystem

    np.random.seed(2024)
    spas = SpatialOrbitalSystem(n, RandomBasisSet(l, dim, np=module))
    gos = GeneralOrbitalSystem(n, RandomBasisSet(l, dim, np=module))
    return spas, gos


def test_no_operators(module):
    spas, gos = systems(module)
    for system in (spas, gos):
        assert not system.has_one_body_time_evolution_operator and not system.has_two_body_time_evolution_operator
        np.testing.assert_array_equal(host(system.h_t(10)), host(system.h))
        np.testing.assert_array_equal(host(system.u_t(10)), host(system.u))
        system.set_time_evolution_operator([], add_h_0=False, add_u_0=False)
        np.testing.assert_array_equal(host(system.h_t(0)), np.zeros_like(host(system.h)))
        np.testing.assert_array_equal(host(system.u_t(0)), np.zeros_like(host(system.u)))


def test_single_one_body_operator(module):
    from quantum_systems_b200.time_evolution_operators import CustomOneBodyOperator

    spas, gos = systems(module)
    spas.set_time_evolution_operator(CustomOneBodyOperator(2, spas.h), add_h_0=False)
    gos.set_time_evolution_operator(CustomOneBodyOperator(3, gos.h), add_u_0=False)
    assert spas.has_one_body_time_evolution_operator and not spas.has_two_body_time_evolution_operator
    assert_close_scaled(host(spas.h_t(0)), host(spas.h) * 2, rel=1e-15)
    np.testing.assert_array_equal(host(spas.u_t(0)), host(spas.u))
    assert_close_scaled(host(gos.h_t(0)), host(gos.h) + host(gos.h) * 3, rel=1e-15)
    np.testing.assert_array_equal(host(gos.u_t(0)), np.zeros_like(host(gos.u)))


@pytest.mark.parametrize("gauge", ["length", "velocity"])
def test_dipole_field_interaction(module, gauge):
    from quantum_systems_b200.time_evolution_operators import DipoleFieldInteraction

    spas, gos = systems(module)
    field = lambda t: np.sin(0.25 * 2 + t)  # noqa: E731
    polarization = np.array([0.0, 1.0, 0.5])
    for system in (spas, gos):
        if gauge == "velocity":
            system._basis_set.momentum = system.np.asarray(np.random.default_rng(1).standard_normal((3, system.l, system.l)))
        system.set_time_evolution_operator(DipoleFieldInteraction(field, polarization, gauge=gauge))
        assert system.has_one_body_time_evolution_operator and not system.has_two_body_time_evolution_operator
        for t in [0, 0.1, 1.3]:
            if gauge == "length":
                ref = host(system.h) - field(t) * np.tensordot(polarization, host(system.dipole_moment), axes=(0, 0))
            else:
                ref = host(system.h) + field(t) * np.tensordot(polarization, host(system.momentum), axes=(0, 0))
                ref = ref + 0.5 * field(t) ** 2 * np.eye(system.l)
            got = system.h_t(t)
            assert isinstance(got, np.ndarray) == (module is np)
            assert_close_scaled(host(got), ref, rel=1e-14)
            np.testing.assert_array_equal(host(system.u_t(t)), host(system.u))
    # default polarization: the first axis (operator.py:150-153)
    spas.set_time_evolution_operator(DipoleFieldInteraction(0.3))
    assert_close_scaled(host(spas.h_t(0)), host(spas.h) - 0.3 * host(spas.dipole_moment)[0], rel=1e-15)


def test_multiple_operators_and_adiabatic_switching(module):
    from quantum_systems_b200.time_evolution_operators import AdiabaticSwitching, CustomOneBodyOperator

    spas, gos = systems(module)
    spas.set_time_evolution_operator(
        [CustomOneBodyOperator(2, spas.h), CustomOneBodyOperator(3, spas.s), AdiabaticSwitching(2)], add_u_0=False
    )
    gos.set_time_evolution_operator(
        (CustomOneBodyOperator(1, gos.h), CustomOneBodyOperator(3, gos.s), CustomOneBodyOperator(-2, gos.position[0])),
        add_h_0=False,
    )
    assert spas.has_two_body_time_evolution_operator and not gos.has_two_body_time_evolution_operator
    assert_close_scaled(host(spas.h_t(0)), host(spas.h) + host(spas.h) * 2 + host(spas.s) * 3, rel=1e-15)
    np.testing.assert_array_equal(host(spas.u_t(0)), 2 * host(spas.u))
    assert_close_scaled(host(gos.h_t(0)), host(gos.h) + host(gos.s) * 3 - host(gos.position[0]) * 2, rel=1e-15)
    np.testing.assert_array_equal(host(gos.u_t(0)), host(gos.u))
    # u_0 + f(t) u folded into one pass; time-dependent and complex switching functions
    gos.set_time_evolution_operator(AdiabaticSwitching(lambda t: 0.25 * t))
    u_t = gos.u_t(2.0)
    assert isinstance(u_t, np.ndarray) == (module is np)
    assert_close_scaled(host(u_t), 1.5 * host(gos.u), rel=1e-15)
    spas.set_time_evolution_operator([AdiabaticSwitching(1j), AdiabaticSwitching(0.5)])
    assert_close_scaled(host(spas.u_t(0)), (1.5 + 1j) * host(spas.u), rel=1e-15)
    # the operator alone returns the scaled array like the reference (operator.py:193-196)
    assert_close_scaled(host(AdiabaticSwitching(3).set_system(spas).u_t(0)), 3 * host(spas.u), rel=1e-15)


def test_reference_style_user_operator(module):
    """An operator that only implements the reference interface (arrays out of h_t / u_t, the base class's
    scalar 0 for the other part) is summed the same way."""
    from quantum_systems_b200.time_evolution_operators import TimeEvolutionOperator

    class Kick(TimeEvolutionOperator):
        is_one_body_operator = True
        is_two_body_operator = True

        def h_t(self, t):
            return t * self._system.s

        def u_t(self, t):
            return -t * self._system.u

    spas, _ = systems(module)
    spas.set_time_evolution_operator([Kick(), TimeEvolutionOperator()])
    assert_close_scaled(host(spas.h_t(0.5)), host(spas.h) + 0.5 * host(spas.s), rel=1e-15)
    assert_close_scaled(host(spas.u_t(0.25)), 0.75 * host(spas.u), rel=1e-15)
