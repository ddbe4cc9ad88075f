"""``GeneralOrbitalSystem`` -- spin-orbital system (mirror of reference general_orbital_system.py)."""

import torch

from . import _arrays, ops
from .system import QuantumSystem


def _scalar(x):
    return x.item() if isinstance(x, torch.Tensor) else x


class GeneralOrbitalSystem(QuantumSystem):
    r"""General spin-orbitals ``psi(x) = psi^a(r) alpha(m_s) + psi^b(r) beta(m_s)``.

    ``GeneralOrbitalSystem(n, basis_set, a=[1, 0], b=[0, 1], anti_symmetrize=True)`` spin-doubles a
    spatial ``basis_set`` in place (fused add_spin + anti-symmetrise + cast kernel) and
    anti-symmetrises an already spin-doubled one if needed (general_orbital_system.py:39-53).
    """

    def __init__(self, n, basis_set, a=[1, 0], b=[0, 1], anti_symmetrize=True, **kwargs):
        if not basis_set.includes_spin:
            basis_set = basis_set.change_to_general_orbital_basis(a=a, b=b, anti_symmetrize=anti_symmetrize)
        if anti_symmetrize:
            basis_set.anti_symmetrize_two_body_elements()
        super().__init__(n, basis_set, **kwargs)

    @property
    def spin_x(self):
        return self._basis_set.spin_x

    @property
    def spin_y(self):
        return self._basis_set.spin_y

    @property
    def spin_z(self):
        return self._basis_set.spin_z

    @property
    def spin_2(self):
        return self._basis_set.spin_2

    @property
    def spin_2_tb(self):
        return self._basis_set.spin_2_tb

    def compute_reference_energy(self, h=None, u=None):
        r"""``E_0 = h_ii + 1/2 u_ijij + E_n`` over occupied ``i, j`` (general_orbital_system.py:75-117)."""
        h = self.h if h is None else h
        u = self.u if u is None else u
        tr_h, direct, _ = self._occupied_trace_terms(h, u)
        return _scalar(tr_h + 0.5 * direct + self.nuclear_repulsion_energy)

    def construct_fock_matrix(self, h, u, f=None):
        r"""``f_pq = h_pq + sum_i u_piqi`` over occupied ``i``; ``u`` is assumed anti-symmetrised.
        If ``f`` is given it is overwritten in place and returned (general_orbital_system.py:119-159)."""
        return _construct_fock(self, h, u, f, spatial=False)

    def change_to_hf_basis(self, *args, **kwargs):
        raise NotImplementedError("There is currently no GHF implementation")


def _construct_fock(system, h, u, f, spatial):
    """Shared Fock driver.  Device arrays go straight to the kernel; for host (numpy) arrays only the
    ``n_occ n^2`` needed elements of ``u`` are gathered (pure indexing) and staged into HBM."""
    n_occ = system.n
    host = not isinstance(h, torch.Tensor)
    if isinstance(u, torch.Tensor) and u.is_cuda:
        h_dev = _arrays.to_device(h)
        if isinstance(f, torch.Tensor) and f.dtype != h_dev.dtype:
            h_dev = h_dev.to(f.dtype)  # f += h promotes a real h into a complex f
        f_dev = f if isinstance(f, torch.Tensor) and f.is_cuda and f.is_contiguous() else None
        out = (ops.fock_spatial if spatial else ops.fock_general)(h_dev, u, n_occ, f=f_dev)
    else:
        import numpy

        idx = numpy.arange(n_occ)
        # qs_fock_gathered reads both blocks as (n_occ, n, n), ordered [i, p, q].  Advanced indices split by a
        # slice put the gathered axis first: u[:, idx, :, idx][i, p, q] = u[p, i, q, i].  ADJACENT advanced
        # indices stay in place, u[:, idx, idx, :][p, i, q] = u[p, i, i, q], so that block is transposed.
        direct = numpy.ascontiguousarray(u[:, idx, :, idx])
        exchange = numpy.ascontiguousarray(u[:, idx, idx, :].transpose(1, 0, 2)) if spatial else None
        h_dev = _arrays.to_device(h)
        if h_dev.dtype == torch.float64 and numpy.iscomplexobj(direct):
            # the reference's `f.fill(0); f += h; f += ...` accumulates a real h into a caller's complex f
            if f is not None and (f.is_complex() if isinstance(f, torch.Tensor) else numpy.iscomplexobj(f)):
                h_dev = h_dev.to(torch.complex128)
            else:
                raise TypeError("complex u cannot be accumulated into a real Fock matrix")
        out = ops.fock_gathered(
            h_dev, _arrays.to_device(direct), _arrays.to_device(exchange) if spatial else None, n_occ,
            2.0 if spatial else 1.0, -1.0 if spatial else 0.0,
        )
    if f is None:
        return _arrays.to_host(out) if host else out
    if isinstance(f, torch.Tensor):
        if out is not f:
            f.copy_(out)
        return f
    f[...] = _arrays.to_host(out)  # in-place semantics for a host array
    return f
