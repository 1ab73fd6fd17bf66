"""Host-side model construction (runs once per call; not a throughput kernel).

Mirrors the reference's setup interface so callers keep passing the same
closures: ``ss_modulators_nmf(w_subband,w_modulator,kernel1,kernel2)``
(matlab/ss_modulators_nmf.m:1), ``lti_disc(F,L,Qc,dt)``
(matlab/unifying_prob_tf/lti_disc.m:1) and the ``balance`` stanza
(matlab/ihgp_ep_modulator_nmf.m:81-87).  What goes to the GPU is the *block
form*: the reference's n-by-n matrices are block diagonal with one small block
per latent and one observation row per block, so only the
per-block A, Q, Pinf and h are packed and uploaded (``BlockModel``).
"""
import dataclasses
import math

import numpy as np
import scipy.linalg as sla

_SQRT = {"matern32": 3.0, "matern52": 5.0, "matern72": 7.0}


def kernel_order(kernel):
    """State dimension of one latent of the given kernel family (cf_*_to_ss.m)."""
    try:
        return {"exp": 1, "matern32": 2, "matern52": 3, "matern72": 4}[kernel]
    except KeyError:
        raise ValueError("unsupported kernel %r (exp, matern32, matern52, matern72)" % (kernel,))


def kernel_sde(kernel, magnSigma2, lengthScale):
    """Continuous-time SDE (F, L, Qc, H, Pinf) of a Matern-family GP prior;
    same parametrisation as matlab/unifying_prob_tf/cf_{exp,matern32,matern52,
    matern72}_to_ss.m (companion form, white-noise density Qc, stationary Pinf)."""
    s2, ell = float(magnSigma2), float(lengthScale)
    tau = kernel_order(kernel)
    if kernel == "exp":
        lam = 1.0 / ell
        Qc = 2.0 * s2 / ell
    else:
        lam = math.sqrt(_SQRT[kernel]) / ell
    F = np.zeros((tau, tau))
    for i in range(tau - 1):
        F[i, i + 1] = 1.0
    # last row: -binom(tau, i) * lam^(tau-i)
    for i in range(tau):
        F[tau - 1, i] = -math.comb(tau, i) * lam ** (tau - i)
    L = np.zeros((tau, 1)); L[-1, 0] = 1.0
    H = np.zeros((1, tau)); H[0, 0] = 1.0
    if kernel == "matern32":
        Qc = 12.0 * math.sqrt(3.0) / ell ** 3 * s2
        Pinf = np.diag([s2, 3.0 * s2 / ell ** 2])
    elif kernel == "matern52":
        Qc = s2 * 400.0 * math.sqrt(5.0) / 3.0 / ell ** 5
        kap = 5.0 / 3.0 * s2 / ell ** 2
        Pinf = np.array([[s2, 0.0, -kap], [0.0, kap, 0.0], [-kap, 0.0, 25.0 * s2 / ell ** 4]])
    elif kernel == "matern72":
        Qc = s2 * 10976.0 * math.sqrt(7.0) / 5.0 / ell ** 7
        kap = 7.0 / 5.0 * s2 / ell ** 2
        kap2 = 9.8 * s2 / ell ** 4
        Pinf = np.array([[s2, 0.0, -kap, 0.0], [0.0, kap, 0.0, -kap2],
                         [-kap, 0.0, kap2, 0.0], [0.0, -kap2, 0.0, 343.0 * s2 / ell ** 6]])
    else:
        Pinf = np.array([[s2]])
    return F, L, np.array([[Qc]]), H, Pinf


def kernel_sde_derivs(kernel, magnSigma2, lengthScale):
    """Derivatives of ``kernel_sde`` with respect to (magnSigma2, lengthScale): three arrays with a trailing axis
    of length 2 (dF, dQc, dPinf), the values of cf_*_to_ss.m's 6th-8th outputs.  With lam = c / ell every entry is
    a monomial in ell -- F[last, i] ~ ell^-(tau-i), Pinf[i, j] ~ s2 ell^-(i+j), Qc ~ s2 ell^-(2 tau - 1) -- so the
    derivatives are the entries times (-power / ell) and (1 / s2)."""
    s2, ell = float(magnSigma2), float(lengthScale)
    F, _, Qc, _, Pinf = kernel_sde(kernel, s2, ell)
    tau = F.shape[0]
    dF = np.zeros((tau, tau, 2)); dQc = np.zeros((1, 1, 2)); dPinf = np.zeros((tau, tau, 2))
    for i in range(tau):
        dF[tau - 1, i, 1] = -(tau - i) * F[tau - 1, i] / ell
    dQc[0, 0, 0] = Qc[0, 0] / s2
    dQc[0, 0, 1] = -(2 * tau - 1) * Qc[0, 0] / ell
    dPinf[:, :, 0] = Pinf / s2
    ij = np.add.outer(np.arange(tau), np.arange(tau))
    dPinf[:, :, 1] = -ij * Pinf / ell
    return dF, dQc, dPinf


class DerivStack:
    """One of the derivative stacks dF / dQc / dPinf of ``ss_modulators_nmf`` (its 6th-8th outputs).  Every slice is
    zero outside ONE latent's diagonal block (ss_modulators_nmf.m:25-129), so the stack is kept as
    ``(latent, block)`` pairs; ``np.asarray(stack)`` gives the reference's dense n x n x (3D+2N) array."""

    def __init__(self, n, starts, items):
        self.n, self.starts, self.items = int(n), np.asarray(starts, int), list(items)

    @property
    def shape(self):
        return (self.n, self.n, len(self.items))

    def __len__(self):
        return len(self.items)

    def __array__(self, dtype=None, copy=None):
        out = np.zeros(self.shape)
        for p, (lat, blk) in enumerate(self.items):
            o = self.starts[lat]
            out[o:o + blk.shape[0], o:o + blk.shape[1], p] = blk
        return out if dtype is None else out.astype(dtype)


def ss_modulators_nmf(w_subband, w_modulator, kernel1, kernel2):
    """``[F,L,Qc,H,Pinf,dF,dQc,dPinf] = ss_modulators_nmf(...)``.

    D quasi-periodic subbands (kernel1 x cosine(omega), block 2*tau1) followed by
    N modulators (kernel2, block tau3).  The derivative stacks (slices ordered
    [var1 (D), len1 (D), omega (D), var2 (N), len2 (N)], ss_modulators_nmf.m:76-78,127-129) come back as
    ``DerivStack`` objects: only gf_giekf_modulator_nmf with GradObj = 'on' reads them
    (the EP entry points discard them, gf_ep_modulator_nmf.m:78)."""
    ws = np.asarray(w_subband, float).ravel()
    wm = np.asarray(w_modulator, float).ravel()
    if ws.size % 3 or wm.size % 2:
        raise ValueError("w_subband must hold [var;len;omega], w_modulator [var;len]")
    D, N = ws.size // 3, wm.size // 2
    I2 = np.eye(2)
    Fs, Ls, Qs, Hs, Ps = [], [], [], [], []
    for d in range(D):
        F1, L1, Qc1, H1, P1 = kernel_sde(kernel1, ws[d], ws[D + d])
        om = ws[2 * D + d]
        rot = np.array([[0.0, -om], [om, 0.0]])
        tau1 = F1.shape[0]
        Fs.append(np.kron(F1, I2) + np.kron(np.eye(tau1), rot))
        Ls.append(np.kron(L1, I2))
        Qs.append(Qc1[0, 0] * I2)
        Hs.append(np.kron(H1, np.array([[1.0, 0.0]])))
        Ps.append(np.kron(P1, I2))
    for j in range(N):
        F2, L2, Qc2, H2, P2 = kernel_sde(kernel2, wm[j], wm[N + j])
        Fs.append(F2); Ls.append(L2); Qs.append(Qc2); Hs.append(H2); Ps.append(P2)
    F = sla.block_diag(*Fs)
    n = F.shape[0]
    tau1, tau3 = kernel_order(kernel1), kernel_order(kernel2)
    starts = np.concatenate([2 * tau1 * np.arange(D), 2 * tau1 * D + tau3 * np.arange(N)])
    qstarts = np.concatenate([2 * np.arange(D), 2 * D + np.arange(N)])
    dFs, dQs, dPs = [None] * (3 * D + 2 * N), [None] * (3 * D + 2 * N), [None] * (3 * D + 2 * N)
    J2 = np.array([[0.0, -1.0], [1.0, 0.0]])
    for d in range(D):
        dF1, dQ1, dP1 = kernel_sde_derivs(kernel1, ws[d], ws[D + d])
        for which, p in ((0, d), (1, D + d)):
            dFs[p] = (d, np.kron(dF1[:, :, which], I2))
            dQs[p] = (d, dQ1[0, 0, which] * I2)
            dPs[p] = (d, np.kron(dP1[:, :, which], I2))
        p = 2 * D + d                                                   # omega enters the rotation only
        dFs[p] = (d, np.kron(np.eye(tau1), J2)); dQs[p] = (d, np.zeros((2, 2))); dPs[p] = (d, np.zeros((2 * tau1, 2 * tau1)))
    for j in range(N):
        dF2, dQ2, dP2 = kernel_sde_derivs(kernel2, wm[j], wm[N + j])
        for which, p in ((0, 3 * D + j), (1, 3 * D + N + j)):
            dFs[p] = (D + j, dF2[:, :, which]); dQs[p] = (D + j, dQ2[:, :, which]); dPs[p] = (D + j, dP2[:, :, which])
    return (F, sla.block_diag(*Ls), sla.block_diag(*Qs), sla.block_diag(*Hs), sla.block_diag(*Ps),
            DerivStack(n, starts, dFs), DerivStack(2 * D + N, qstarts, dQs), DerivStack(n, starts, dPs))


def ss_modulators(w, kernel1, kernel2):
    """``[F,L,Qc,H,Pinf,dF,dQc,dPinf] = ss_modulators(w,kernel1,kernel2)`` (matlab/ss_modulators.m): D carrier x
    modulator pairs with ``w = [var1; len1; omega; var2; len2]`` -- the construction of ss_modulators_nmf with N = D."""
    w = np.asarray(w, float).ravel()
    if w.size % 5:
        raise ValueError("w must hold [var_fast; len_fast; omega; var_slow; len_slow] for D pairs")
    D = w.size // 5
    return ss_modulators_nmf(w[:3 * D], w[3 * D:], kernel1, kernel2)


def lti_disc(F, L=None, Qc=None, dt=1.0):
    """``[A,Q] = lti_disc(F,L,Qc,dt)``: A = expm(F dt), Q by matrix-fraction
    decomposition (lti_disc.m:73-82)."""
    F = np.asarray(F, float)
    n = F.shape[0]
    L = np.eye(n) if L is None else np.asarray(L, float)
    Qc = np.zeros((L.shape[1], L.shape[1])) if Qc is None else np.asarray(Qc, float)
    A = sla.expm(F * dt)
    Phi = np.zeros((2 * n, 2 * n))
    Phi[:n, :n] = F
    Phi[:n, n:] = L @ Qc @ L.T
    Phi[n:, n:] = -F.T
    E = sla.expm(Phi * dt)
    # [0;I] selects the right block column; Q = E12 / E22
    Q = sla.solve(E[n:, n:].T, E[:n, n:].T).T
    return A, Q


def balance(F, L, H, Pinf, return_T=False):
    """``[T,F]=balance(F); L=T\\L; H=H*T; LL=T\\chol(Pinf,'lower'); Pinf=LL*LL'``."""
    Fb, T = sla.matrix_balance(np.asarray(F, float), permute=True, scale=True, separate=False)
    LL = sla.solve(T, np.linalg.cholesky(Pinf))
    res = (Fb, sla.solve(T, L), np.asarray(H, float) @ T, LL @ LL.T)
    return res + (T,) if return_T else res


@dataclasses.dataclass
class BlockModel:
    """Discrete model in block form, the layout the C ABI takes (include/nsagp.h).

    D subband blocks of size bz followed by N modulator blocks of size bg.  All
    blocks are packed column-major, one after another: A, Q, Pinf hold
    D*bz*bz + N*bg*bg doubles, h holds D*bz + N*bg."""
    D: int
    N: int
    bz: int
    bg: int
    A: np.ndarray
    Q: np.ndarray
    Pinf: np.ndarray
    h: np.ndarray

    @property
    def M(self):
        return self.D + self.N

    @property
    def n(self):
        return self.D * self.bz + self.N * self.bg

    def block_sizes(self):
        return [self.bz] * self.D + [self.bg] * self.N

    def starts(self):
        return np.concatenate([[0], np.cumsum(self.block_sizes())]).astype(int)

    def blocks(self, packed):
        """Unpack a packed array of b-by-b blocks into a list of matrices."""
        out, off = [], 0
        for b in self.block_sizes():
            out.append(np.asarray(packed[off:off + b * b]).reshape((b, b), order="F"))
            off += b * b
        return out

    def hrows(self):
        out, off = [], 0
        for b in self.block_sizes():
            out.append(self.h[off:off + b]); off += b
        return out

    def dense(self):
        """Dense (A, Q, H, Pinf) -- for inspection and tests."""
        A = sla.block_diag(*self.blocks(self.A)); Q = sla.block_diag(*self.blocks(self.Q))
        P = sla.block_diag(*self.blocks(self.Pinf))
        H = np.zeros((self.M, self.n)); st = self.starts()
        for i, hr in enumerate(self.hrows()):
            H[i, st[i]:st[i + 1]] = hr
        return A, Q, H, P


def to_block_model(A, Q, H, Pinf, D, N):
    """Extract the block form from dense model matrices, checking that they have
    the structure the GPU path relies on: M = D+N diagonal blocks (sizes taken
    from the observation rows, ihgp_ep_modulator_nmf.m:104), equal size within
    the subband group and within the modulator group, nothing off the blocks."""
    A = np.asarray(A, float); Q = np.asarray(Q, float)
    H = np.asarray(H, float); Pinf = np.asarray(Pinf, float)
    M, n = H.shape
    if M != D + N:
        raise ValueError("H has %d rows, expected D+N = %d" % (M, D + N))
    first = [int(np.flatnonzero(H[i])[0]) if np.any(H[i]) else -1 for i in range(M)]
    if first[0] != 0 or any(f < 0 for f in first) or any(np.diff(first) <= 0):
        raise ValueError("observation rows do not delimit consecutive state blocks")
    st = np.array(first + [n])
    sizes = np.diff(st)
    bz, bg = int(sizes[0]), int(sizes[-1])
    if np.any(sizes[:D] != bz) or np.any(sizes[D:] != bg):
        raise ValueError("blocks within the subband / modulator groups must have equal size")
    mask = np.zeros((n, n), bool)
    hmask = np.zeros((M, n), bool)
    for i in range(M):
        mask[st[i]:st[i + 1], st[i]:st[i + 1]] = True
        hmask[i, st[i]:st[i + 1]] = True
    for name, X in (("A", A), ("Q", Q), ("Pinf", Pinf)):
        if np.any(X[~mask] != 0):
            raise ValueError("%s is not block diagonal; the GPU EP path needs one independent block per latent" % name)
    if np.any(H[~hmask] != 0):
        raise ValueError("H couples a site to more than one block")
    pk = lambda X: np.concatenate([X[st[i]:st[i + 1], st[i]:st[i + 1]].reshape(-1, order="F") for i in range(M)])
    h = np.concatenate([H[i, st[i]:st[i + 1]] for i in range(M)])
    return BlockModel(D, N, bz, bg, pk(A), pk(Q), pk(Pinf), h)
