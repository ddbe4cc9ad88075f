"""``QuantumSystem`` -- a ``BasisSet`` plus a particle number (mirror of reference system.py:6-250).

The class is deliberately thin: it forwards the matrix elements of its basis set and keeps the
occupied/virtual bookkeeping; the heavy methods (``change_basis``, ``construct_fock_matrix``) end in
the CUDA kernels behind ``BasisSet`` / ``ops``.
"""

import abc
import copy
import types
import typing


def _is_scalar(x):
    return isinstance(x, (int, float, complex)) or getattr(x, "ndim", None) == 0


def _zeros_like(module, a):
    from .sharded import ShardedTwoBody

    if isinstance(a, ShardedTwoBody):
        return a.scaled(0.0)
    return module.zeros_like(a)


def scaled_sum(module, terms):
    """``sum_k w_k X_k`` for ``terms = [(w_k, X_k), ...]`` of equally shaped arrays, with ``qs_scale_add``: the
    first two terms in one pass, every further term one in-place pass.  ``X_k`` may be CUDA tensors, host arrays
    (staged; the result is returned in the storage of ``module``) or ``ShardedTwoBody`` handles (shard-local, no
    communication).  Mixed real / complex operands and complex weights give a complex128 result."""
    import torch

    from . import _arrays, ops
    from .sharded import ShardedTwoBody

    (w0, x0), rest = terms[0], terms[1:]
    if isinstance(x0, ShardedTwoBody):
        acc = x0.scaled(w0, rest[0][1], rest[0][0]) if rest else x0.scaled(w0)
        for w, x in rest[1:]:
            acc.axpby_(1.0, x, w)
        return acc
    device_in = isinstance(x0, torch.Tensor) and x0.is_cuda
    xs = [_arrays.to_device(x) for _, x in terms]
    ws = [w for w, _ in terms]
    if len(xs) == 1:
        acc = ops.scale_add(xs[0], ws[0])
    else:
        acc = ops.scale_add(xs[0], ws[0], xs[1], ws[1])
        for w, x in zip(ws[2:], xs[2:]):
            if x.dtype != acc.dtype or complex(w).imag != 0:
                acc = ops.scale_add(acc, 1.0, x, w)  # promotion: a fresh complex result
            else:
                ops.scale_add(acc, 1.0, x, w, out=acc)
    return acc if device_in and not _arrays.is_host_module(module) else _arrays.to_module(acc, module)


class QuantumSystem(metaclass=abc.ABCMeta):
    """Abstract base: ``n`` occupied basis functions out of ``basis_set.l`` (system.py:6-30)."""

    def __init__(self, n, basis_set):
        self._basis_set = basis_set
        assert n <= self._basis_set.l
        self.np = self._basis_set.np
        self.set_system_size(n, self._basis_set.l)
        self._time_evolution_operator = []
        self._add_h_0 = True
        self._add_u_0 = True

    def set_system_size(self, n, l):
        """Set ``n, l, m = l - n`` and the occupied/virtual slices ``o, v`` (system.py:32-51)."""
        assert n <= l
        self.n = n
        self.l = l
        self.m = self.l - self.n
        self.o = slice(0, self.n)
        self.v = slice(self.n, self.l)

    @abc.abstractmethod
    def construct_fock_matrix(self, h, u, f=None):
        pass

    def change_module(self, np):
        """Move the system (and its basis set) to another array module (system.py:57-67)."""
        self.np = np
        self._basis_set.change_module(self.np)

    def change_basis(self, C, C_tilde=None):
        """Basis change of every matrix element; ``o``/``v`` follow the new ``l`` (system.py:69-71)."""
        self._basis_set.change_basis(C, C_tilde)
        self.set_system_size(self.n, self._basis_set.l)

    @abc.abstractmethod
    def change_to_hf_basis(self, *args, **kwargs):
        pass

    @abc.abstractmethod
    def compute_reference_energy(self, h=None, u=None):
        pass

    def compute_particle_density(self, rho_qp, C=None, C_tilde=None):
        return self._basis_set.compute_particle_density(rho_qp, C=C, C_tilde=C_tilde)

    # forwarding properties (system.py:86-142)
    @property
    def dim(self):
        return self._basis_set.dim

    @property
    def grid(self):
        return self._basis_set.grid

    @property
    def h(self):
        """One-body Hamiltonian."""
        return self._basis_set.h

    @property
    def u(self):
        """Two-body Hamiltonian."""
        return self._basis_set.u

    @property
    def s(self):
        """Overlap matrix."""
        return self._basis_set.s

    @property
    def position(self):
        return self._basis_set.position

    @property
    def momentum(self):
        return self._basis_set.momentum

    @property
    def dipole_moment(self):
        return self._basis_set.dipole_moment

    @property
    def spf(self):
        return self._basis_set.spf

    @property
    def bra_spf(self):
        return self._basis_set.bra_spf

    @property
    def nuclear_repulsion_energy(self):
        return self._basis_set.nuclear_repulsion_energy

    @property
    def particle_charge(self):
        return self._basis_set.particle_charge

    # time-dependent operators (system.py:144-215)
    def set_time_evolution_operator(self, time_evolution_operator, add_h_0=True, add_u_0=True):
        if not isinstance(time_evolution_operator, typing.Iterable):
            time_evolution_operator = [time_evolution_operator]
        self._add_h_0 = add_h_0
        self._add_u_0 = add_u_0
        self._time_evolution_operator = [op.set_system(self) for op in time_evolution_operator]

    @property
    def has_one_body_time_evolution_operator(self):
        return any(op.is_one_body_operator for op in self._time_evolution_operator)

    @property
    def has_two_body_time_evolution_operator(self):
        return any(op.is_two_body_operator for op in self._time_evolution_operator)

    def h_t(self, current_time):
        """``h_0 + sum_op h_op(t)`` (system.py:189-201), accumulated by ``qs_scale_add``."""
        return self._hamiltonian_part(
            self._basis_set.h, self._add_h_0, self.has_one_body_time_evolution_operator, "h_t", current_time
        )

    def u_t(self, current_time):
        """``u_0 + sum_op u_op(t)`` (system.py:203-215).  Every term costs one ``qs_scale_add`` pass over the n^4
        elements -- ``u_0 + f(t) u`` of an adiabatic switching a single one -- in HBM-resident, host and sharded
        storage alike."""
        return self._hamiltonian_part(
            self._basis_set.u, self._add_u_0, self.has_two_body_time_evolution_operator, "u_t", current_time
        )

    def _hamiltonian_part(self, base, add_base, any_operator, method, current_time):
        if not any_operator:
            return base if add_base else _zeros_like(self.np, base)
        terms = [(1.0, base)] if add_base else []
        constant = 0
        for op in self._time_evolution_operator:
            scaled = getattr(op, method + "_scaled", None)
            weight, operand = scaled(current_time) if scaled is not None else (1.0, getattr(op, method)(current_time))
            if _is_scalar(operand):  # the base class contributes the number 0 (operator.py:66, :84)
                constant = constant + weight * operand
            else:
                terms.append((weight, operand))
        if not terms:
            return _zeros_like(self.np, base) + constant
        out = scaled_sum(self.np, terms)
        return out if _is_scalar(constant) and constant == 0 else out + constant

    def transform_one_body_elements(self, h, C, C_tilde=None):
        return self._basis_set.transform_one_body_elements(h, C, np=self.np, C_tilde=C_tilde)

    def transform_two_body_elements(self, u, C, C_tilde=None):
        return self._basis_set.transform_two_body_elements(u, C, np=self.np, C_tilde=C_tilde)

    def copy_system(self):
        """Deep copy of the system; array modules are shared (system.py:227-250)."""
        memo = {id(self.np): self.np, id(self._basis_set.np): self._basis_set.np}
        for holder in (self, self._basis_set, getattr(self._basis_set, "potential", None)):
            if holder is None:
                continue
            for value in vars(holder).values():
                if isinstance(value, types.ModuleType):
                    memo[id(value)] = value
        new_system = copy.deepcopy(self, memo)
        assert new_system.np is self.np
        return new_system

    # helpers shared by the concrete systems -----------------------------------------------
    def _occupied_trace_terms(self, h, u):
        """``tr h[o,o]``, ``sum_ij u[i,j,i,j]``, ``sum_ij u[i,j,j,i]`` as Python scalars, reduced on the GPU
        (``ops.occupied_traces``).  Host arrays: only the occupied corner (n_occ^4 elements) is staged."""
        import numpy
        import torch

        from . import _arrays, ops

        n_occ = self.n
        if n_occ == 0:
            return 0.0, 0.0, 0.0
        if isinstance(u, torch.Tensor) and u.is_cuda:
            terms = ops.occupied_traces(_arrays.to_device(h), u, n_occ)
            is_complex = u.is_complex() or (h.is_complex() if isinstance(h, torch.Tensor) else numpy.iscomplexobj(h))
        else:
            o = self.o
            corner = _arrays.to_device(numpy.ascontiguousarray(u[o, o, o, o]))
            h_oo = _arrays.to_device(numpy.ascontiguousarray(h[o, o]))
            terms = ops.occupied_traces(h_oo, corner, n_occ)
            is_complex = numpy.iscomplexobj(u) or numpy.iscomplexobj(h)
        tr_h, direct, exchange = (complex(x) for x in terms.cpu().tolist())
        if not is_complex:
            tr_h, direct, exchange = tr_h.real, direct.real, exchange.real
        return tr_h, direct, exchange
