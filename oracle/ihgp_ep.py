"""Oracle: infinite-horizon (steady-state gain) Power-EP (test infrastructure only).

Restates matlab/ihgp_ep_modulator_nmf.m (table setup :90-134,150-191; predict
mode :195-526; nlZ mode :533-624) and the differences of
matlab/ihgp_ep_modulator_nmf_constraints.m (running site vectors in nlZ mode,
:568-651), plus matlab/apxGrid.m:555-566 (``neqinterp``: piecewise-linear,
inverse-distance weights -- the grid is not equispaced so the "cubic" request
falls through, apxGrid.m:461-472,695-707).

``dare`` (Control System Toolbox, not under /root/reference) is restated with
scipy.linalg.solve_discrete_are / solve_discrete_lyapunov.
"""
import numpy as np
import scipy.linalg as sla

from . import ssmodel
from .gf_ep import merge_inputs, _max0


def neqinterp(s, t):
    """apxGrid.m:555-566.  Returns the dense (nt, ns) interpolation matrix."""
    s = np.asarray(s, float).ravel(); t = np.asarray(t, float).ravel()
    ns = s.size
    order = np.argsort(s, kind="stable"); s = s[order]
    if np.min(np.diff(s)) < 1e-10:
        raise ValueError("Some source points are equal.")
    edges = np.concatenate([[-np.inf], s[1:-1], [np.inf]])
    ii = np.searchsorted(edges, t, side="right") - 1       # histc bin (0-based)
    ii = np.minimum(ii, ns - 2)
    d0 = t - s[ii]; d1 = s[ii + 1] - t
    d0 = np.where(d0 < 0, 0.0, d0); d1 = np.where(d1 < 0, 0.0, d1)
    U = np.zeros((t.size, ns))
    rows = np.arange(t.size)
    U[rows, order[ii]] += d1 / (d1 + d0)
    U[rows, order[ii + 1]] += d0 / (d1 + d0)
    return U


def block_starts(H):
    """ihgp_ep_modulator_nmf.m:104  ilist = [find(sum(H,1)) size(H,2)+1] (0-based)."""
    return np.concatenate([np.flatnonzero(H.sum(axis=0)), [H.shape[1]]]).astype(int)


def forward_tables(A, Q, H, ilist):
    """ihgp_ep_modulator_nmf.m:106-134.  PPlist[n] is (200, b*b), rows = PP(:)'."""
    M = H.shape[0]
    r = np.logspace(-2, 4, 200)
    ro = np.logspace(-2, 4, 32)
    PPlisto, PPlist = [], []
    for n in range(M):
        ii = slice(ilist[n], ilist[n + 1])
        Ab = A[ii, ii]; Qb = Q[ii, ii]; hb = H[n, ii]
        tab = np.empty((ro.size, Ab.size))
        for j in range(ro.size):
            # dare(A',H',Q,ro): A X A' - X - A X H'(H X H'+ro)^-1 H X A' + Q = 0
            PP = sla.solve_discrete_are(Ab.T, hb.reshape(-1, 1), Qb, np.array([[ro[j]]]))
            tab[j] = PP.reshape(-1, order="F")
        U = neqinterp(ro, r)
        PPlisto.append(tab)
        PPlist.append(U @ tab)
    return r, ro, PPlisto, PPlist


def smoother_tables(A, Q, H, ilist, r, ro, PPlisto):
    """ihgp_ep_modulator_nmf.m:150-191.  PGlist[n] is (200, 2*b*b): [PS2(:)' G(:)']."""
    M = H.shape[0]
    PGlist = []
    for n in range(M):
        ii = slice(ilist[n], ilist[n + 1])
        Ab = A[ii, ii]; Qb = Q[ii, ii]; hb = H[n, ii]
        b = Ab.shape[0]
        tab = np.empty((ro.size, 2 * b * b))
        for j in range(ro.size):
            PP = PPlisto[n][j].reshape((b, b), order="F")
            S = hb @ PP @ hb + ro[j]
            K = PP @ hb / S
            P = PP - np.outer(K * ro[j], K)
            L = np.linalg.cholesky(Ab @ P @ Ab.T + Qb)      # failure branch :166-170 has a typo; never reached
            G = np.linalg.solve(L.T, np.linalg.solve(L, (P @ Ab.T).T)).T
            QQ = P - G @ PP @ G.T; QQ = (QQ + QQ.T) / 2
            DD, V = np.linalg.eigh(QQ)
            ind = DD > 0
            QQ = (V[:, ind] * DD[ind][None, :]) @ V[:, ind].T
            # dare(G',0*G,QQ): with B = 0 this is the Lyapunov equation X = G X G' + QQ
            PS2 = sla.solve_discrete_lyapunov(G, QQ)
            tab[j] = np.concatenate([PS2.reshape(-1, order="F"), G.reshape(-1, order="F")])
        U = neqinterp(ro, r)
        PGlist.append(U @ tab)
    return PGlist


def _lookup_filter(r, Rn):
    """[~,ind] = min(abs(r-R))  (:239): first index on ties; all-Inf/NaN -> index 0."""
    with np.errstate(invalid="ignore"):
        d = np.abs(r - Rn)
    if np.all(np.isnan(d)):
        return 0
    return int(np.nanargmin(d)) if np.any(np.isnan(d)) else int(np.argmin(d))


def ihgp_setup(A, Q, H, want_smoother=True):
    Q = (Q + Q.T) / 2                                         # :97
    ilist = block_starts(H)
    r, ro, PPlisto, PPlist = forward_tables(A, Q, H, ilist)
    PGlist = smoother_tables(A, Q, H, ilist, r, ro, PPlisto) if want_smoother else None
    return dict(Q=Q, ilist=ilist, r=r, ro=ro, PPlist=PPlist, PGlist=PGlist)


def _filter_pass(A, H, Pinf, tabs, lik_param, Wnmf, yall, mom, ep_damp, ttau, tnu, R, MS, m,
                 first_pass, running_sites=False):
    """One forward pass: ihgp_ep_modulator_nmf.m:233-310 (predict) / :551-613 (nlZ).
    ``running_sites``: the _constraints nlZ variant keeps (M,) site vectors that are
    carried from step to step (ihgp_ep_modulator_nmf_constraints.m:568-615)."""
    ilist, r, PPlist = tabs["ilist"], tabs["r"], tabs["PPlist"]
    M, T = H.shape[0], yall.size
    n = A.shape[0]
    lZ = 0.0
    lZk_all = np.zeros(T)
    HA = H @ A
    if running_sites:
        tt = np.zeros(M); tn = np.zeros(M); Rrun = np.zeros(M)
    for k in range(T):
        if k > 0:
            PP = np.zeros((n, n))
            for nn in range(M):
                Rprev = Rrun[nn] if running_sites else R[nn, k - 1]
                ind = _lookup_filter(r, Rprev)
                ii = slice(ilist[nn], ilist[nn + 1]); b = ilist[nn + 1] - ilist[nn]
                PP[ii, ii] = PPlist[nn][ind].reshape((b, b), order="F")
        else:
            PP = Pinf
        fmu = HA @ m; W = PP @ H.T; HPH = np.diag(H @ W).copy()          # :250
        if running_sites:
            tt_k, tn_k = tt, tn
        else:
            tt_k, tn_k = ttau[:, k], tnu[:, k]
        if first_pass or k == T - 1:
            lZ_k, dlZ, d2lZ = mom(lik_param, fmu, HPH, Wnmf, 1, yall, k)  # :256
            lZ += lZ_k; lZk_all[k] = lZ_k
            with np.errstate(all="ignore"):
                tt_new = (1 - ep_damp) * tt_k + ep_damp * (-d2lZ / (1 + d2lZ * HPH))
                tn_new = (1 - ep_damp) * tn_k + ep_damp * ((dlZ - fmu * d2lZ) / (1 + d2lZ * HPH))
                Rk = 1.0 / tt_new                                        # before the clamp (:269)
            tt_k, tn_k = tt_new, tn_new
        else:
            Rk = R[:, k].copy()
        tt_k = _max0(tt_k)                                               # :274
        with np.errstate(all="ignore"):
            ys = tn_k / tt_k                                             # :277
        for nn in range(M):
            ii = slice(ilist[nn], ilist[nn + 1])
            if tt_k[nn] == 0:
                Rk[nn] = np.inf                                          # :287
                m[ii] = A[ii, ii] @ m[ii]
            else:
                K = W[ii, nn] / (HPH[nn] + Rk[nn])                       # :293
                AKHA = A[ii, ii] - np.outer(K, H[nn, ii]) @ A[ii, ii]    # :296
                m[ii] = AKHA @ m[ii] + K * ys[nn]                        # :299
        if running_sites:
            # the constraints nlZ variant writes Inf into column k of a vector R
            # (…_constraints.m:626), so the *next* look-up sees the pre-clamp
            # 1/ttau (negative / Inf / NaN) -- all of which select index 1 (0 here).
            tt, tn = tt_k, tn_k
            Rrun = np.where(tt_k == 0, np.inf, Rk)
        else:
            ttau[:, k] = tt_k; tnu[:, k] = tn_k; R[:, k] = Rk
        if MS is not None:
            MS[:, k] = m
    return lZ, lZk_all


def ihgp_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction, ep_damping, ep_itts,
                 return_ind, tabs=None):
    """Predict mode, ihgp_ep_modulator_nmf.m:146-526.  Returns (Eft, Varft, lb, ub, out)."""
    if tabs is None:
        tabs = ihgp_setup(A, Q, H)
    ilist, r, PGlist = tabs["ilist"], tabs["r"], tabs["PGlist"]
    n = A.shape[0]; M = H.shape[0]; T = yall.size
    ep_damping = np.atleast_1d(np.asarray(ep_damping, float))
    m = np.zeros(n)
    P = Pinf.copy()
    MS = np.zeros((n, T))
    ttau = np.zeros((M, T)); tnu = np.zeros((M, T))
    R = np.exp(float(np.asarray(lik_param).ravel()[0])) * np.ones((M, T))     # :209
    nlZ = np.zeros(ep_itts)
    out = {}
    n_negcav = 0
    maxDiffM_hist = []
    ep_damp = ep_damping[0]
    for itt in range(1, ep_itts + 1):
        maxDiffM = 0.0
        MSP = MS.copy()
        # NB: m carries over from the previous iteration's smoother (k=1) -- :198 is outside the EP loop
        lZ, _ = _filter_pass(A, H, Pinf, tabs, lik_param, Wnmf, yall, mom, ep_damp, ttau, tnu, R, MS, m,
                             first_pass=(itt == 1))
        if itt == 1:
            nlZ[0] = -lZ
        out.update(tnu=tnu.copy(), ttau=ttau.copy(), lZ=lZ, R=R.copy(), MF=MS.copy())
        P = np.zeros((n, n)); G = np.zeros((n, n))
        if itt < ep_itts:
            ep_damp = ep_damping[itt]
        for k in range(T - 2, -1, -1):
            for nn in range(M):                                          # :379-388
                if np.isinf(R[nn, k]):
                    ind = r.size - 1
                else:
                    ind = _lookup_filter(r, R[nn, k])
                ii = slice(ilist[nn], ilist[nn + 1]); b = ilist[nn + 1] - ilist[nn]
                PG = PGlist[nn][ind]
                P[ii, ii] = PG[:b * b].reshape((b, b), order="F")
                G[ii, ii] = PG[b * b:].reshape((b, b), order="F")
            m = MS[:, k] + G @ (m - A @ MS[:, k])                        # :391
            MS[:, k] = m
            if itt < ep_itts and not np.isnan(yall[k]):
                m_marginal = H @ m
                v_marginal = np.diag(H @ P @ H.T).copy()
                with np.errstate(all="ignore"):
                    v_cav = 1.0 / (1.0 / v_marginal - ep_fraction * ttau[:, k])
                    m_cav = v_cav * (m_marginal / v_marginal - ep_fraction * tnu[:, k])
                    upd = v_cav > 0
                    n_negcav += int(np.sum(~upd))
                    lZ_k, dlZ, d2lZ = mom(lik_param, m_cav, v_cav, Wnmf, ep_fraction, yall, k)
                    if itt > 1:
                        lZ += lZ_k                                       # :420
                    ttau[upd, k] = (1 - ep_damp * ep_fraction) * ttau[upd, k] + \
                        ep_damp * (-d2lZ[upd] / (1 + d2lZ[upd] * v_cav[upd]))
                    tnu[upd, k] = (1 - ep_damp * ep_fraction) * tnu[upd, k] + \
                        ep_damp * ((dlZ[upd] - m_cav[upd] * d2lZ[upd]) / (1 + d2lZ[upd] * v_cav[upd]))
                    R[upd, k] = 1.0 / ttau[upd, k]                       # :434 (no clamp here)
            maxDiffM = max(maxDiffM, np.max(np.abs(H @ MSP[:, k] - H @ m)))
        if itt < ep_itts:
            nlZ[itt] = -lZ
        maxDiffM_hist.append(maxDiffM)
    out.update(tnu=tnu, ttau=ttau, R=R, MS=MS, nlZ=nlZ, n_negcav=n_negcav, maxDiffM=np.array(maxDiffM_hist),
               r=r, ro=tabs["ro"], PPlist=tabs["PPlist"])
    Eft = H @ MS[:, return_ind]
    Varft = np.tile(np.diag(H @ P @ H.T)[:, None], (1, len(return_ind)))   # :492 (P of the last look-up, k=1)
    Varft = np.abs(Varft)                                                # :493-496
    lb = Eft - 1.96 * np.sqrt(Varft); ub = Eft + 1.96 * np.sqrt(Varft)
    return Eft, Varft, lb, ub, out


def ihgp_nlz_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_damping, tabs=None, running_sites=False):
    """nlZ mode: one ADF sweep (ihgp_ep_modulator_nmf.m:533-624; _constraints :568-651)."""
    if tabs is None:
        tabs = ihgp_setup(A, Q, H, want_smoother=False)
    M, T = H.shape[0], yall.size
    ep_damp = np.atleast_1d(np.asarray(ep_damping, float))[0]
    ttau = np.zeros((M, T)); tnu = np.zeros((M, T)); R = np.zeros((M, T))
    m = np.zeros(A.shape[0])
    lZ, lZk = _filter_pass(A, H, Pinf, tabs, lik_param, Wnmf, yall, mom, ep_damp, ttau, tnu, R, None, m,
                           first_pass=True, running_sites=running_sites)
    out = dict(ttau=ttau, tnu=tnu, R=R, lZ=lZk)
    return (-lZ if running_sites else -np.sum(lZk)), out


def _model(lik_param, param1, param2, ss, x, kernel1, kernel2):
    F, L, Qc, H, Pinf = ss(x, param1, param2, kernel1, kernel2)[:5]
    F, L, H, Pinf, _ = ssmodel.balance_ss(F, L, H, Pinf)                 # :81-87 (balance ON)
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    return A, Q, H, Pinf


def ihgp_ep_modulator_nmf(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                          ep_fraction, ep_damping, ep_itts):
    """ihgp_ep_modulator_nmf.m:1."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_log(w, num_lik_params, D, N)
    A, Q, H, Pinf = _model(lik_param, param1, param2, ss, x, kernel1, kernel2)
    if xt is not None and np.size(xt) > 0:
        Eft, Varft, lb, ub, out = ihgp_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction,
                                               ep_damping, ep_itts, return_ind)
        return Eft, Varft, None, lb, ub, out
    edata, _ = ihgp_nlz_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_damping)
    return edata, np.zeros(np.size(w))


def ihgp_ep_modulator_nmf_constraints(w, x, y, ss, mom, xt, kernel1, kernel2, num_lik_params, D, N,
                                      ep_fraction, ep_damping, ep_itts, constraints, w_fixed, tune_hypers):
    """ihgp_ep_modulator_nmf_constraints.m:1."""
    yall, return_ind = merge_inputs(x, y, xt)
    lik_param, param1, param2, Wnmf = ssmodel.unpack_constraints(w, num_lik_params, D, N, constraints,
                                                                 w_fixed, tune_hypers)
    A, Q, H, Pinf = _model(lik_param, param1, param2, ss, x, kernel1, kernel2)
    if xt is not None and np.size(xt) > 0:
        Eft, Varft, lb, ub, out = ihgp_ep_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_fraction,
                                               ep_damping, ep_itts, return_ind)
        return Eft, Varft, None, lb, ub, out
    edata, _ = ihgp_nlz_core(A, Q, H, Pinf, lik_param, Wnmf, yall, mom, ep_damping, running_sites=True)
    return edata, np.zeros(np.size(w))
