"""The staged asynchronous epilogue of the quarter GEMM (cp.async.bulk shared -> global, include/qsb200.h
``qs_set_bulk_epilogue_mode``) stores exactly what the register epilogue stores: every launch shape is run in both
modes and the results are compared BIT FOR BIT (same accumulators, only the way out of the SM differs), and against
the oracle."""

import numpy as np
import pytest

from conftest import assert_close_scaled
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


@pytest.fixture
def modes():
    from quantum_systems_b200 import _native

    lib = _native.load()
    previous = lib.qs_set_bulk_epilogue_mode(-1)  # out of range: query only

    def run(mode, fn):
        lib.qs_set_bulk_epilogue_mode(mode)
        try:
            return fn()
        finally:
            lib.qs_set_bulk_epilogue_mode(previous)

    yield run
    lib.qs_set_bulk_epilogue_mode(previous)


@pytest.mark.parametrize("n,m,u_complex,c_complex", [
    (12, 8, False, False),    # NT = 1, rows per block (12) < rows per warp (32): three runs per staged column
    (4, 6, False, False),     # tiny blocks: eight runs per column
    (20, 24, False, False),   # NT = 3: the second column group of the last pass is absent
    (30, 40, False, False),   # NT = 5
    (64, 64, False, False),   # NT = 8, whole tiles
    (66, 70, False, False),   # two tile groups (NT = 5 and 4), ragged last row tile
    (13, 13, False, False),   # odd real extent: padded pitch, the staged path must step aside
    (10, 12, True, True),     # complex output: 16-byte elements, 8 complex columns per pass
    (24, 18, False, True),    # real u x complex C
    (24, 20, True, False),    # complex u x real C: the split (2M) kernel keeps its register stores
])
def test_four_index_transform_is_bit_identical_in_both_epilogue_modes(modes, n, m, u_complex, c_complex):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(1000 * n + m)
    u, C = rand(rng, (n,) * 4, u_complex), rand(rng, (n, m), c_complex)
    u_dev, C_dev = dev(u), dev(C)
    registers = modes(0, lambda: ops.transform_two_body(u_dev, C_dev, symmetry=0))
    staged = modes(2, lambda: ops.transform_two_body(u_dev, C_dev, symmetry=0))
    assert torch.equal(registers, staged)
    assert_close_scaled(staged.cpu().numpy(), oracle.transform_two_body_elements(u, C))


@pytest.mark.parametrize("kind", ["antisym", "exchange"])
@pytest.mark.parametrize("n,m,complex_", [(48, 48, False), (50, 56, False), (48, 50, True)])
def test_masked_launches_with_row_tables_in_both_epilogue_modes(modes, n, m, complex_, kind):
    """Symmetry-aware transform: step 3 stores through a block table (staged when its entries are even), step 4
    places rows one by one (always register stores)."""
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(n + m)
    u = rand(rng, (n,) * 4, complex_)
    u = u - u.transpose(0, 1, 3, 2) if kind == "antisym" else 0.5 * (u + u.transpose(1, 0, 3, 2))
    C = rand(rng, (n, m), complex_)
    symmetry = 1 if kind == "antisym" else 2
    u_dev, C_dev = dev(u), dev(C)
    registers = modes(0, lambda: ops.transform_two_body(u_dev, C_dev, symmetry=symmetry))
    staged = modes(2, lambda: ops.transform_two_body(u_dev, C_dev, symmetry=symmetry))
    assert torch.equal(registers, staged)
    assert_close_scaled(staged.cpu().numpy(), oracle.transform_two_body_elements(u, C))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("n,complex_,antisymmetric", [(24, False, False), (26, True, False), (48, False, True),
                                                       (27, False, False)])
def test_scattering_store_in_both_epilogue_modes(modes, world, n, complex_, antisymmetric):
    """Fused re-partition (emulated ranks): three-level row split, column dealing, destinations per column; the
    anti-symmetric schedule adds the packed pair layout.  Default mode 1 stages exactly these launches."""
    from quantum_systems_b200 import sharded

    rng = np.random.default_rng(n + world)
    u = rand(rng, (n,) * 4, complex_)
    if antisymmetric:
        u = u - u.transpose(0, 1, 3, 2)
    C = rand(rng, (n, n), complex_)
    expected = oracle.transform_two_body_elements(u, C)

    def transform():
        ctx = sharded.EmulatedContext(world)
        basis = sharded.ShardedBasisSet.from_global(ctx, np.eye(n), np.eye(n), u)
        return sharded.transform_two_body_sharded(basis.u, dev(C)).gather()

    results = [modes(mode, transform) for mode in (0, 1, 2)]
    assert torch.equal(results[0], results[1]) and torch.equal(results[0], results[2])
    assert_close_scaled(results[1].cpu().numpy(), expected)


def test_one_body_and_grid_functions_in_both_epilogue_modes(modes):
    from quantum_systems_b200 import ops

    rng = np.random.default_rng(3)
    h, C = rand(rng, (18, 18), True), rand(rng, (18, 10), True)
    a = modes(0, lambda: ops.transform_one_body(dev(h), dev(C)))
    b = modes(2, lambda: ops.transform_one_body(dev(h), dev(C)))
    assert torch.equal(a, b)
    assert_close_scaled(b.cpu().numpy(), oracle.transform_one_body_elements(h, C))
    spf = rand(rng, (18, 40), False)
    a = modes(0, lambda: ops.transform_functions(dev(spf), dev(C), bra=False))
    b = modes(2, lambda: ops.transform_functions(dev(spf), dev(C), bra=False))
    assert torch.equal(a, b)
    assert_close_scaled(b.cpu().numpy(), oracle.transform_spf(spf, C))
