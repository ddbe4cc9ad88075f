"""TEST INFRASTRUCTURE: closed-form results of the sharded pipeline at sizes no CPU reference can hold
(structured inputs whose exact transform is known in O(n^2) pieces).

A spatial tensor of the ODQD form  u[a,b,c,d] = sum_pq L[p,a] L[p,c] W[p,q] L[q,b] L[q,d]  (reference
quantum_dots/one_dim/one_dim_qd.py:275-280 with L = eigenvectors on the grid, W = shielded Coulomb; any L and W
are allowed, e.g. a low-rank synthetic one) is spin-doubled (basis_set.py:772-774), anti-symmetrised
(basis_set.py:776-778) and basis-changed (basis_set.py:336-350) into

    u'[P,Q,R,S] = G[P,Q,R,S] - G[P,Q,S,R],      G[P,Q,R,S] = sum_pq W[p,q] F_p[P,R] F_q[Q,S],
    F_p[P,R]    = sum_sigma alpha_sigma[p,P] beta_sigma[p,R],
    alpha_sigma[p,P] = sum_a L[p,a] C~[P, 2a+sigma],   beta_sigma[p,R] = sum_c L[p,c] C[2c+sigma, R]

so a whole (R,S) plane costs O(G'^2 n) and one element O(G'^2) on the host.  Only numpy is used; this module is
imported by tests/ and by bench.py's in-run parity checks, never by the product package.
"""

import numpy as np


def shielded_coulomb(inner_grid, alpha, a):
    """``W[p,q] = alpha / sqrt((x_p - x_q)^2 + a^2)`` -- reference one_dim_qd.py:30-32."""
    x = np.asarray(inner_grid)
    return alpha / np.sqrt((x[:, None] - x[None, :]) ** 2 + a**2)


class SpinDoubledClosedForm:
    """``L`` (G', l), ``W`` (G', G'); ``C`` (2l, m), ``C_tilde`` (m, 2l) or None for ``C^dagger`` -- the NET basis
    change (a sequence of ``change_basis`` calls composes into ``C_1 C_2 ...`` and ``... C~_2 C~_1``)."""

    def __init__(self, L, W, C, C_tilde=None, anti_symmetrize=True):
        L, W, C = np.asarray(L), np.asarray(W), np.asarray(C)
        Ct = C.conj().T if C_tilde is None else np.asarray(C_tilde)
        self.W = W
        self.anti = anti_symmetrize
        self.alpha = [L @ Ct[:, s::2].T for s in (0, 1)]  # (G', m)
        self.beta = [L @ C[s::2, :] for s in (0, 1)]      # (G', m)
        self.m = C.shape[1]

    def _F(self, P):
        """F[p, R] for a fixed left index P: (G', m)."""
        return self.alpha[0][:, P, None] * self.beta[0] + self.alpha[1][:, P, None] * self.beta[1]

    def plane(self, P, Q):
        """u'[P, Q, :, :] as an (m, m) array."""
        G = self._F(P).T @ (self.W @ self._F(Q))
        return G - G.T if self.anti else G

    def elements(self, P, Q, R, S):
        """u' at index arrays (P, Q, R, S) of equal length."""
        P, Q, R, S = (np.asarray(x) for x in (P, Q, R, S))
        out = np.empty(P.shape[0], dtype=np.result_type(self.alpha[0], self.beta[0], self.W))
        a0, a1, b0, b1 = self.alpha[0], self.alpha[1], self.beta[0], self.beta[1]
        for lo in range(0, P.shape[0], 256):
            sl = slice(lo, lo + 256)
            p, q, r, s = P[sl], Q[sl], R[sl], S[sl]
            FPR = a0[:, p] * b0[:, r] + a1[:, p] * b1[:, r]  # (G', chunk)
            FQS = a0[:, q] * b0[:, s] + a1[:, q] * b1[:, s]
            val = np.einsum("pi,pi->i", FPR, self.W @ FQS)
            if self.anti:
                FPS = a0[:, p] * b0[:, s] + a1[:, p] * b1[:, s]
                FQR = a0[:, q] * b0[:, r] + a1[:, q] * b1[:, r]
                val = val - np.einsum("pi,pi->i", FPS, self.W @ FQR)
            out[sl] = val
        return out

    def fock_two_body(self, n_occ):
        """``sum_{i < n_occ} u'[P, i, R, i]`` as an (m, m) array -- the two-body part of the general Fock matrix
        (reference general_orbital_system.py:119-159) -- in O(n_occ G'^2 m)."""
        a0, a1, b0, b1 = self.alpha[0], self.alpha[1], self.beta[0], self.beta[1]
        occ = slice(0, n_occ)
        trace = (a0[:, occ] * b0[:, occ] + a1[:, occ] * b1[:, occ]).sum(axis=1)  # t_q = sum_i F_q[i, i]
        v = self.W @ trace
        direct = (a0 * v[:, None]).T @ b0 + (a1 * v[:, None]).T @ b1            # sum_p v_p F_p[P, R]
        if not self.anti:
            return direct
        exchange = np.zeros_like(direct)
        for i in range(n_occ):
            left = a0 * b0[:, i, None] + a1 * b1[:, i, None]    # F_p[P, i]  as (G', m) over P
            right = a0[:, i, None] * b0 + a1[:, i, None] * b1   # F_q[i, R]  as (G', m) over R
            exchange += left.T @ (self.W @ right)
        return direct - exchange

    def dense(self):
        """The whole (m, m, m, m) tensor -- small sizes only (tests of this module itself)."""
        F = np.stack([self._F(P) for P in range(self.m)])  # (P, p, R)
        G = np.einsum("apr,pq,bqs->abrs", F, self.W, F, optimize=True)
        return G - G.transpose(0, 1, 3, 2) if self.anti else G


def check_shard(local, p0, form, rng, samples=2000, planes=2):
    """Compare a rank's slab ``local`` = u'[p0 : p0 + local.shape[0]] (any array with numpy-style indexing that
    returns host arrays via ``fetch``) with the closed form.  ``local`` is accessed through two callables so that
    device tensors need no full copy: ``local.plane(i, Q)`` -> (m, m) ndarray and ``local.gather(i, Q, R, S)``.
    Returns ``(max|err| on planes, max|err| on samples, anti-symmetry defect, max|ref| seen)``."""
    count, m = local.count, form.m
    if count == 0:
        return 0.0, 0.0, 0.0, 0.0
    scale = 0.0
    worst_plane = 0.0
    rows = sorted({0, count - 1})[:planes]
    for i in rows:
        for Q in sorted({0, int(rng.integers(m))}):
            ref = form.plane(p0 + i, Q)
            got = local.plane(i, Q)
            scale = max(scale, float(np.abs(ref).max()))
            worst_plane = max(worst_plane, float(np.abs(got - ref).max()))
    I = rng.integers(0, count, size=samples)
    Q, R, S = (rng.integers(0, m, size=samples) for _ in range(3))
    ref = form.elements(p0 + I, Q, R, S)
    got = local.gather(I, Q, R, S)
    scale = max(scale, float(np.abs(ref).max()))
    worst_sample = float(np.abs(got - ref).max())
    plane = local.plane(0, min(1, m - 1))
    defect = float(np.abs(plane + plane.T).max()) if form.anti else 0.0
    return worst_plane, worst_sample, defect, scale


class TorchSlab:
    """Accessors of ``check_shard`` over a (count, m, m, m) torch tensor (CUDA or CPU)."""

    def __init__(self, tensor):
        self.t = tensor
        self.count = tensor.shape[0]

    def plane(self, i, Q):
        return self.t[i, Q].cpu().numpy()

    def gather(self, I, Q, R, S):
        import torch

        dev = self.t.device
        idx = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (I, Q, R, S)]
        return self.t[idx[0], idx[1], idx[2], idx[3]].cpu().numpy()
