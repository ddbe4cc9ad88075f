"""The streamed host path (uploads and downloads hidden behind the quarter GEMMs, quantum_systems_b200/streamed.py)
gives the same numbers as the plain path -- bit for bit, since every output element is accumulated by the same
kernel in the same order -- and the oracle's to 1e-12."""

import numpy as np
import pytest

from conftest import assert_close_scaled
from oracle import qs_oracle as oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rand(rng, shape, complex_):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if complex_ else x


@pytest.fixture()
def small_threshold(monkeypatch):
    from quantum_systems_b200 import streamed

    monkeypatch.setattr(streamed, "MIN_BYTES", 0)


@pytest.mark.parametrize("u_complex,c_complex,biorth", [(False, False, False), (True, False, False), (False, True, True),
                                                        (True, True, True), (True, True, False)])
@pytest.mark.parametrize("n,m", [(8, 8), (20, 20), (18, 26), (26, 18), (40, 40), (13, 13)])
def test_streamed_path_matches_plain_path_and_oracle(small_threshold, n, m, u_complex, c_complex, biorth):
    from quantum_systems_b200 import BasisSet, ops, streamed

    rng = np.random.default_rng(31 * n + m)
    u, C = rand(rng, (n,) * 4, u_complex), rand(rng, (n, m), c_complex)
    Ct = rand(rng, (m, n), c_complex) if biorth else None
    C_dev = torch.from_numpy(C).cuda()
    Ct_dev = None if Ct is None else torch.from_numpy(Ct).cuda()
    expected_streamed = not (n % 2 and not u_complex)  # odd real extents take the plain path
    assert streamed.applicable(u, C_dev) == expected_streamed
    got = BasisSet.transform_two_body_elements(u, C, np, C_tilde=Ct)
    assert isinstance(got, np.ndarray)
    plain = ops.transform_two_body(torch.from_numpy(u).cuda(), C_dev, Ct_dev).cpu().numpy()
    np.testing.assert_array_equal(got, plain)
    assert_close_scaled(got, oracle.transform_two_body_elements(u, C, Ct))


def test_streamed_path_with_pinned_input_and_repeated_calls(small_threshold):
    from quantum_systems_b200 import BasisSet

    n = 32
    rng = np.random.default_rng(3)
    pinned = torch.empty((n,) * 4, dtype=torch.float64, pin_memory=True)
    pinned.numpy()[...] = rng.standard_normal((n,) * 4)
    C = np.linalg.qr(rng.standard_normal((n, n)))[0]
    bs = BasisSet(n, 1, np=np)
    bs.u, bs.h, bs.s = pinned.numpy(), np.eye(n), np.eye(n)
    u0 = bs.u.copy()
    for _ in range(3):  # side streams and staging buffers are reused across calls
        bs.change_basis(C)
        bs.change_basis(C.T.copy())
    assert_close_scaled(bs.u, u0, rel=1e-12)
    np.testing.assert_array_equal(pinned.numpy(), u0)  # the input was never modified
