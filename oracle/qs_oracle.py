"""CPU oracle for the two-body integral pipeline of HyQD/quantum-systems.

TEST INFRASTRUCTURE ONLY.  This module restates, in plain numpy, the algorithm the reference
runs for the hot path.  It exists to *check* the CUDA path; nothing
in ``quantum_systems_b200`` imports it.  The only permitted importers are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs.

Where the arithmetic lives: the reference is pure Python and delegates every contraction to
numpy (-> OpenBLAS ``dgemm``/``zgemm``).  numpy is an *unpinned* third-party dependency of the
reference (``setup.py:7-14`` lists bare ``numpy``); this container and the GPU box ship numpy 2.3.5 /
OpenBLAS 0.3.30.  The oracle therefore calls the same numpy primitives in the same order as the
reference call sites cited on each function.

Parity pin: every function below is checked in ``tests/test_oracle_golden.py`` against vectors
produced by importing the reference itself (``tests/golden/make_golden.py``, run in the build
container where ``/root/reference`` is mounted) and against the reference's own golden files
(``tests/dat/od*_{h,u,spf,dipole_moment}.npy``, re-packed by the same script).

All citations are relative to ``/root/reference/quantum_systems/``.
"""

import numpy as np
import scipy.linalg
import scipy.special


# --------------------------------------------------------------------------------------------
# basis changes
# --------------------------------------------------------------------------------------------


def default_bra_coefficients(C):
    """``C_tilde = C^dagger`` when the caller gives none (basis_set.py:331-332, :338-339, :451-452)."""
    return np.conj(C).T


def transform_one_body_elements(h, C, C_tilde=None):
    """``C_tilde @ (h @ C)`` -- basis_set.py:329-334 (right product first, then left)."""
    bra = default_bra_coefficients(C) if C_tilde is None else C_tilde
    right = np.dot(h, C)
    return np.dot(bra, right)


def transform_two_body_elements(u, C, C_tilde=None):
    """Four-index transform ``u'_pqrs = sum C~[p,a] C~[q,b] u[a,b,c,d] C[c,r] C[d,s]``.

    Follows basis_set.py:336-350: four ``tensordot`` quarter steps contracting, in this order,
    the 4th index with C, the 3rd with C, the 2nd with C_tilde, the 1st with C_tilde, each followed
    by the axis permutation that puts the new index back in the contracted slot.
    """
    bra = default_bra_coefficients(C) if C_tilde is None else C_tilde
    step = np.tensordot(u, C, axes=([3], [0]))  # (a,b,c,s)            :342
    step = np.tensordot(step, C, axes=([2], [0]))  # (a,b,s,r)
    step = np.swapaxes(step, 2, 3)  # (a,b,r,s)            :344
    step = np.tensordot(step, bra, axes=([1], [1]))  # (a,r,s,q)
    step = np.moveaxis(step, 3, 1)  # (a,q,r,s)            :346
    return np.tensordot(bra, step, axes=([1], [0]))  # (p,q,r,s)            :348


def transform_spf(spf, C):
    """Ket single-particle functions, basis_set.py:321-323."""
    return np.tensordot(C, spf, axes=([0], [0]))


def transform_bra_spf(bra_spf, C_tilde):
    """Bra single-particle functions, basis_set.py:325-327."""
    return np.tensordot(C_tilde, bra_spf, axes=([1], [0]))


# --------------------------------------------------------------------------------------------
# spin doubling and anti-symmetrisation
# --------------------------------------------------------------------------------------------


def spin_delta(p, q):
    """1 when p and q have the same parity (same spin), else 0 -- system_helper.py:9-11."""
    return 1 - ((p ^ q) & 1)


def add_spin_one_body(h):
    """``kron(h, I_2)`` -- basis_set.py:768-770.  Result dtype follows numpy promotion with float eye."""
    return np.kron(h, np.eye(2))


def add_spin_two_body(u):
    """``kron(u, delta_pr delta_qs)`` on the 2x2x2x2 spin block -- basis_set.py:772-774.

    ``U[2p+s1, 2q+s2, 2r+s3, 2s+s4] = u[p,q,r,s] * [s1==s3] * [s2==s4]`` (spin index fastest).
    """
    eye = np.eye(2)
    spin_block = np.einsum("pr, qs -> pqrs", eye, eye)
    return np.kron(u, spin_block)


def add_spin_spf(spf):
    """Row interleave ``new[2p] = new[2p+1] = spf[p]`` -- basis_set.py:751-759."""
    out = np.zeros((2 * spf.shape[0],) + tuple(spf.shape[1:]), dtype=spf.dtype)
    out[0::2] = spf
    out[1::2] = spf
    return out


def anti_symmetrize_u(u):
    """``u - u.transpose(0,1,3,2)`` -- basis_set.py:776-778."""
    return u - np.transpose(u, (0, 1, 3, 2))


def add_spin_anti_symmetrize_loop(u_spatial):
    """Brute-force element loop, the independent check the reference's tests use
    (tests/test_helper.py:108-133, tests/conftest.py:21-45).  Small l only."""
    l = 2 * u_spatial.shape[0]
    out = np.zeros((l, l, l, l), dtype=u_spatial.dtype)
    for p in range(l):
        for q in range(l):
            for r in range(l):
                for s in range(l):
                    direct = spin_delta(p, r) * spin_delta(q, s) * u_spatial[p // 2, q // 2, r // 2, s // 2]
                    exchange = spin_delta(p, s) * spin_delta(q, r) * u_spatial[p // 2, q // 2, s // 2, r // 2]
                    out[p, q, r, s] = direct - exchange
    return out


def setup_pauli_matrices(a, b):
    """Pauli matrices in the spin basis {a, b} (column vectors) -- basis_set.py:638-697."""
    a = np.asarray(a, dtype=np.complex128).reshape(-1, 1)
    b = np.asarray(b, dtype=np.complex128).reshape(-1, 1)
    cartesian = [
        np.array([[0, 1], [1, 0]], dtype=np.complex128),
        np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
        np.array([[1, 0], [0, -1]], dtype=np.complex128),
    ]
    basis = [a, b]
    out = []
    for sigma in cartesian:
        m = np.zeros((2, 2), dtype=np.complex128)
        for i, bra in enumerate(basis):
            for j, ket in enumerate(basis):
                m[i, j] = np.dot(bra.conj().T, np.dot(sigma, ket))[0, 0]
        out.append(m)
    return tuple(out)


def setup_spin_squared_operator(spin_x, spin_y, spin_z, overlap):
    """One- and two-body parts of S^2 -- basis_set.py:699-749."""
    l = len(spin_x)
    spin_2 = np.zeros_like(spin_x)
    spin_2_tb = np.zeros((l, l, l, l), dtype=spin_2.dtype)
    for s_i in (spin_x, spin_y, spin_z):
        spin_2 += s_i @ overlap @ s_i
        spin_2_tb += np.einsum("pr, qs -> pqrs", s_i, s_i)
    return spin_2, spin_2_tb


def change_to_general_orbital_basis(basis, a=(1, 0), b=(0, 1), anti_symmetrize=True):
    """Spin-double a spatial basis given as a dict with keys ``h, s, u`` and optional
    ``position, momentum, spf, bra_spf``.  Returns a new dict with every array complex128.

    Order of operations follows basis_set.py:568-636: one-body kron; two-body kron; Pauli and spin
    matrices from the *spatial* overlap (:572, :597-599); S^2 with the *doubled* overlap (:601-603);
    anti-symmetrise u and spin_2_tb (:605-606); position/momentum/spf; cast to complex (:632-634).
    """
    out = {}
    overlap = basis["s"].copy()
    out["h"] = add_spin_one_body(basis["h"])
    out["s"] = add_spin_one_body(basis["s"])
    out["u"] = add_spin_two_body(basis["u"])

    sx, sy, sz = setup_pauli_matrices(a, b)
    out["sigma_x"], out["sigma_y"], out["sigma_z"] = sx, sy, sz
    out["spin_x"] = 0.5 * np.kron(overlap, sx)
    out["spin_y"] = 0.5 * np.kron(overlap, sy)
    out["spin_z"] = 0.5 * np.kron(overlap, sz)
    out["spin_2"], out["spin_2_tb"] = setup_spin_squared_operator(
        out["spin_x"], out["spin_y"], out["spin_z"], out["s"]
    )
    if anti_symmetrize:
        out["u"] = anti_symmetrize_u(out["u"])
        out["spin_2_tb"] = anti_symmetrize_u(out["spin_2_tb"])
    for name in ("position", "momentum"):
        if basis.get(name) is not None:
            out[name] = np.array([add_spin_one_body(x) for x in basis[name]])
    if basis.get("spf") is not None:
        out["spf"] = add_spin_spf(basis["spf"])
        if basis.get("bra_spf") is not None:
            out["bra_spf"] = add_spin_spf(basis["bra_spf"])
    return {k: (v.astype(np.complex128) if k not in ("sigma_x", "sigma_y", "sigma_z") else v) for k, v in out.items()}


def change_basis(basis, C, C_tilde=None):
    """``BasisSet.change_basis`` on a dict of arrays -- basis_set.py:413-464.

    h, s, u, spin_2_tb, position, momentum, spf/bra_spf are transformed; spin_x/y/z/spin_2 are
    returned UNCHANGED, reproducing the reference's discarded loop variable (:368-372).
    """
    bra = default_bra_coefficients(C) if C_tilde is None else C_tilde
    out = dict(basis)
    out["h"] = transform_one_body_elements(basis["h"], C, bra)
    if basis.get("s") is not None:
        out["s"] = transform_one_body_elements(basis["s"], C, bra)
    out["u"] = transform_two_body_elements(basis["u"], C, bra)
    if basis.get("spin_2_tb") is not None:
        out["spin_2_tb"] = transform_two_body_elements(basis["spin_2_tb"], C, bra)
    for name in ("position", "momentum"):
        if basis.get(name) is not None:
            out[name] = np.asarray([transform_one_body_elements(x, C, bra) for x in basis[name]])
    if basis.get("spf") is not None:
        bra_spf = basis.get("bra_spf")
        if bra_spf is None:
            bra_spf = basis["spf"].conj()  # basis_set.py:246-251
        out["bra_spf"] = transform_bra_spf(bra_spf, bra)
        out["spf"] = transform_spf(basis["spf"], C)
    return out


# --------------------------------------------------------------------------------------------
# Fock matrices and reference energies
# --------------------------------------------------------------------------------------------


def construct_fock_matrix_general(h, u, n_occ, f=None):
    """``f = h + sum_i u[p,i,q,i]`` over occupied i -- general_orbital_system.py:119-159."""
    o = slice(0, n_occ)
    if f is None:
        f = np.zeros_like(h)
    f.fill(0)
    f += h
    f += np.einsum("piqi -> pq", u[:, o, :, o])
    return f


def construct_fock_matrix_spatial(h, u, n_occ, f=None):
    """``f = h + 2 u[p,i,q,i] - u[p,i,i,q]`` -- spatial_orbital_system.py:150-190."""
    o = slice(0, n_occ)
    if f is None:
        f = np.zeros_like(h)
    f.fill(0)
    f += h
    f += 2 * np.einsum("piqi -> pq", u[:, o, :, o])
    f -= np.einsum("piiq -> pq", u[:, o, o, :])
    return f


def reference_energy_general(h, u, n_occ, nuclear_repulsion_energy=0.0):
    """``h_ii + 1/2 u_ijij + E_n`` -- general_orbital_system.py:75-117."""
    o = slice(0, n_occ)
    return (
        np.trace(h[o, o])
        + 0.5 * np.trace(np.trace(u[o, o, o, o], axis1=1, axis2=3))
        + nuclear_repulsion_energy
    )


def reference_energy_spatial(h, u, n_occ, nuclear_repulsion_energy=0.0):
    """``2 h_ii + 2 u_ijij - u_ijji + E_n`` -- spatial_orbital_system.py:106-148."""
    o = slice(0, n_occ)
    return (
        2 * np.trace(h[o, o])
        + 2 * np.trace(np.trace(u[o, o, o, o], axis1=1, axis2=3))
        - np.trace(np.trace(u[o, o, o, o], axis1=1, axis2=2))
        + nuclear_repulsion_energy
    )


# --------------------------------------------------------------------------------------------
# one-dimensional quantum dot on a grid
# --------------------------------------------------------------------------------------------


def shielded_coulomb(x_1, x_2, alpha, a):
    """``alpha / sqrt((x1-x2)^2 + a^2)`` -- quantum_dots/one_dim/one_dim_qd.py:29-32."""
    return alpha / np.sqrt((x_1 - x_2) ** 2 + a**2)


def odqd_orbitals(l, grid_length, num_grid_points, potential):
    """Finite-difference eigenproblem of ``ODQD.setup_basis`` -- one_dim_qd.py:258-266.

    Returns ``(grid, eps, C)`` with ``C`` of shape ``(G-2, l)``: the dx-normalised eigenvectors on
    the interior grid points.
    """
    grid = np.linspace(-grid_length, grid_length, num_grid_points)
    dx = grid[1] - grid[0]
    diag = 1.0 / (dx**2) + potential(grid[1:-1])
    off = -1.0 / (2 * dx**2) * np.ones(num_grid_points - 3)
    eps, C = scipy.linalg.eigh_tridiagonal(diag, off, select="i", select_range=(0, l - 1))
    return grid, eps, C


def odqd_coulomb_elements(C, grid, alpha, a):
    """``u_abcd = sum_pq C_pa C_qb C_pc C_qd W_pq`` -- one_dim_qd.py:275-280.

    Same einsum call as the reference (``optimize=True``), so the contraction path, and hence the
    rounding, is numpy's own two-GEMM path.  The result is a permuted view, as in the reference.
    """
    inner = grid[1:-1]
    w = shielded_coulomb(inner[None, :], inner[:, None], alpha, a)
    return np.einsum("pa, qb, pc, qd, pq -> abcd", C, C, C, C, w, optimize=True)


def odqd_position_elements(C, grid, beta):
    """``<a| x + beta x^2 |b>`` on the interior grid -- one_dim_qd.py:282-289."""
    inner = grid[1:-1]
    return np.einsum("pa, p, pb -> ab", C, inner + beta * inner**2, C, optimize=True)


def odqd_setup_basis(l, grid_length, num_grid_points, potential, a=0.25, alpha=1.0, beta=0.0):
    """Everything ``ODQD.setup_basis`` stores -- one_dim_qd.py:258-289.  Returns a dict."""
    grid, eps, C = odqd_orbitals(l, grid_length, num_grid_points, potential)
    dx = grid[1] - grid[0]
    spf = np.zeros((l, num_grid_points), dtype=np.complex128)
    spf[:, 1:-1] = C.T / np.sqrt(dx)
    position = np.zeros((1, l, l), dtype=np.complex128)
    position[0] = odqd_position_elements(C, grid, beta)
    return {
        "grid": grid,
        "eigen_energies": eps,
        "C": C,
        "spf": spf,
        "h": np.diag(eps).astype(np.complex128),
        "s": np.eye(l),
        "u": odqd_coulomb_elements(C, grid, alpha, a),
        "position": position,
    }


def odho_functions(l, grid, omega):
    """Harmonic-oscillator eigenfunctions on the grid, ``N_n exp(-omega x^2 / 2) H_n(sqrt(omega) x)`` with
    ``N_n = (omega / pi)^(1/4) / sqrt(2^n n!)`` -- ``ODHO.ho_function`` / ``normalization``, one_dim_qd.py:136-148
    (the reference evaluates ``scipy.special.hermite(n)``; same call here)."""
    spf = np.zeros((l, grid.shape[0]))
    for n in range(l):
        norm = 1.0 / np.sqrt(2**n * scipy.special.factorial(n)) * (omega / np.pi) ** 0.25
        spf[n] = norm * np.exp(-0.5 * omega * grid**2) * scipy.special.hermite(n)(np.sqrt(omega) * grid)
    return spf


def _trapz_prep(vec, dx):
    """Trapezoid rule as a weight vector: times dx, end points halved -- one_dim_qd.py:19-27."""
    out = vec * dx
    out[0] *= 0.5
    out[-1] *= 0.5
    return out


def odho_coulomb_elements(spf, grid, alpha, a):
    """Two nested trapezoid integrals of ``ODHO.setup_basis`` -- ``_compute_inner_integral`` one_dim_qd.py:35-51 and
    ``_compute_orbital_integrals`` :54-68, the same loops with the innermost dot products written as matrix
    products: ``inner[q,s,i] = trapz_j(conj(spf_q) W(x_i, .) spf_s)``, ``u[p,q,r,s] = trapz_i(conj(spf_p) inner[q,s] spf_r)``."""
    l, G = spf.shape
    dx = grid[1] - grid[0]
    inner = np.zeros((l, l, G), dtype=np.complex128)
    for i in range(G):
        prepped = _trapz_prep(shielded_coulomb(grid[i], grid, alpha, a), dx)
        inner[:, :, i] = (np.conjugate(spf) * prepped) @ spf.T
    u = np.zeros((l, l, l, l), dtype=np.complex128)
    for q in range(l):
        for s in range(l):
            prepped = _trapz_prep(inner[q, s].copy(), dx)
            u[:, q, :, s] = (np.conjugate(spf) * prepped) @ spf.T.astype(np.complex128)
    return u


def odho_position_elements(l, omega, dtype=np.float64):
    """Analytic ``<n| x |n+1>`` of ``ODHO.construct_position_integrals`` -- one_dim_qd.py:150-166."""
    position = np.zeros((1, l, l), dtype=dtype)
    for n in range(l - 1):
        nn = 1.0 / np.sqrt(2**n * scipy.special.factorial(n)) * (omega / np.pi) ** 0.25
        nn_up = 1.0 / np.sqrt(2 ** (n + 1) * scipy.special.factorial(n + 1)) * (omega / np.pi) ** 0.25
        pos = nn * nn_up * (n + 1) * np.sqrt(np.pi) * 2**n * scipy.special.factorial(n) / omega
        position[0, n, n + 1] = pos
        position[0, n + 1, n] = pos
    return position


def odho_setup_basis(l, grid_length, num_grid_points, omega=0.25, a=0.25, alpha=1.0):
    """Everything ``ODHO.setup_basis`` stores -- one_dim_qd.py:115-134.  Returns a dict."""
    grid = np.linspace(-grid_length, grid_length, num_grid_points)
    eps = omega * (np.arange(l) + 0.5)
    spf = odho_functions(l, grid, omega)
    return {
        "grid": grid,
        "eigen_energies": eps,
        "spf": spf,
        "h": np.diag(eps).astype(np.complex128),
        "s": np.eye(l),
        "u": odho_coulomb_elements(spf, grid, alpha, a),
        "position": odho_position_elements(l, omega, spf.dtype),
    }


# --------------------------------------------------------------------------------------------
# one-dimensional sinc-DVR (two-body operator diagonal in the basis)
# --------------------------------------------------------------------------------------------


def sinc_dvr_setup_basis(l, grid_length, potential, a=0.25, alpha=1.0, beta=0.0, u_repr="2d"):
    """Everything ``ODSincDVR.setup_basis`` stores -- sinc_dvr/one_dim/sinc_dvr.py:99-127, :146-176.
    ``u`` is the (l, l) matrix of shielded-Coulomb values for ``u_repr="2d"`` or the (l,l,l,l) tensor
    with ``u[p,q,p,q] = W[p,q]`` for ``"4d"``; every array is cast to complex128 at the end (:126)."""
    grid = np.linspace(-grid_length, grid_length, l)
    dx = grid[1] - grid[0]
    ind = np.arange(l)
    diff = ind[:, None] - ind
    h = np.zeros((l, l), dtype=np.complex128)
    off = ~np.eye(l, dtype=bool)
    h[off] = (-1.0) ** diff[off] / (dx**2 * diff[off] ** 2)  # :108-113
    h[ind, ind] = np.pi**2 / (6 * dx**2) + potential(grid)  # :115-116
    spf = 1 / np.sqrt(dx) * np.sinc((grid - grid[:, None]) / dx)  # :146-148
    w = shielded_coulomb(grid[:, None], grid[None, :], alpha, a)  # :160-165
    if u_repr == "2d":
        u = w
    else:
        u = np.zeros((l, l, l, l))
        u[ind[:, None], ind[None, :], ind[:, None], ind[None, :]] = w  # :171-173
    position = np.zeros((1, l, l), dtype=np.complex128)
    position[0] = np.diag(grid + beta * grid**2)  # :150-152
    return {
        "grid": grid,
        "h": h,
        "s": np.eye(l).astype(np.complex128),
        "spf": spf.astype(np.complex128),
        "u": u.astype(np.complex128),
        "position": position,
    }


def sinc_dvr_transform_two_body_elements(u2d, C, C_tilde=None, anti_symmetrize=False):
    """``u'_pqrs = sum_ab C~_pa C_ar C~_qb C_bs u_ab`` (minus the r <-> s exchange) --
    sinc_dvr/one_dim/sinc_dvr.py:225-252, the same einsum strings."""
    bra = default_bra_coefficients(C) if C_tilde is None else C_tilde
    out = np.einsum("bs,ar,qb,pa,ab->pqrs", C, C, bra, bra, u2d, optimize=True)
    if anti_symmetrize:
        out = out - np.einsum("br,as,qb,pa,ab->pqrs", C, C, bra, bra, u2d, optimize=True)
    return out


def sinc_dvr_add_spin_two_body(u2d):
    """``kron(u, ones(2, 2))`` for the 2-D representation -- sinc_dvr.py:200-208."""
    return np.kron(u2d, np.ones((2, 2)))
