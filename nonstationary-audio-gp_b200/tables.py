"""Host-side steady-state tables for the infinite-horizon path (setup, once per call).

Interface follows matlab/ihgp_ep_modulator_nmf.m:90-134 (forward predictive
covariances ``PPlist``) and :150-191 (smoother ``PGlist`` = [smoothed cov, gain]):
per block, 32 Riccati solutions on ``ro = logspace(-2,4,32)`` are interpolated
piecewise-linearly (in r, not log r -- apxGrid's non-equispaced branch,
matlab/apxGrid.m:461-472,555-566) onto ``r = logspace(-2,4,200)``.  The GPU path
takes the tables as inputs (include/nsagp.h: nsagp_tables).
"""
import dataclasses

import numpy as np
import scipy.linalg as sla

N_COARSE = 32
N_FINE = 200


def _linear_interp_matrix(src, dst):
    """Rows: weights of the two bracketing source nodes, clamped at the ends."""
    src = np.asarray(src, float); dst = np.asarray(dst, float)
    seg = np.clip(np.searchsorted(src, dst, side="right") - 1, 0, src.size - 2)
    lo = np.maximum(dst - src[seg], 0.0)
    hi = np.maximum(src[seg + 1] - dst, 0.0)
    U = np.zeros((dst.size, src.size))
    rows = np.arange(dst.size)
    U[rows, seg] = hi / (hi + lo)
    U[rows, seg + 1] += lo / (hi + lo)
    return U


@dataclasses.dataclass
class IhgpTables:
    """r: (200,) grid of equivalent noise variances.  PP[n]: (200, b*b) predictive
    covariances, PG[n]: (200, 2*b*b) [smoothed covariance, smoother gain]; each row
    is a column-major flattened b-by-b matrix.  PG is None for nlZ-mode calls."""
    r: np.ndarray
    ro: np.ndarray
    PP: list
    PG: list

    def packed(self):
        pp = np.concatenate([t.ravel() for t in self.PP])
        pg = None if self.PG is None else np.concatenate([t.ravel() for t in self.PG])
        return pp, pg


# Which Riccati solver build_tables uses by default (see its docstring).  The parity tests switch this off so
# that kernels and oracle run on bitwise identical tables; the solvers are compared separately.
DEFAULT_NATIVE = True


def build_tables(model, want_smoother=True, native=None):
    """Solve the per-block DAREs and interpolate (``model``: ssmodel.BlockModel,
    with Q already symmetrised as in ihgp_ep_modulator_nmf.m:97).

    native=True: the library's own host routine (C ABI ``nsagp_ihgp_tables``: doubling algorithm
    in long double, a few milliseconds).  native=False: the same tables through SciPy's generic
    Riccati / Lyapunov solvers (what the oracle uses; ~0.3 s) -- kept as the cross-check."""
    if native is None:
        native = DEFAULT_NATIVE
    if native:
        return _build_tables_native(model, want_smoother)
    r = np.logspace(-2, 4, N_FINE)
    ro = np.logspace(-2, 4, N_COARSE)
    U = _linear_interp_matrix(ro, r)
    PP, PG = [], []
    for Ab, Qb, hb in zip(model.blocks(model.A), model.blocks(model.Q), model.hrows()):
        b = Ab.shape[0]
        fwd = np.empty((N_COARSE, b * b))
        bwd = np.empty((N_COARSE, 2 * b * b))
        for j, rj in enumerate(ro):
            try:
                # dare(A',h',Q,r): predictive covariance of the steady-state filter
                Pp = sla.solve_discrete_are(Ab.T, hb.reshape(b, 1), Qb, np.array([[rj]]))
            except Exception as e:      # the reference drops the node; we refuse instead
                raise RuntimeError("forward DARE failed at ro[%d]=%g: %s" % (j, rj, e))
            fwd[j] = Pp.reshape(-1, order="F")
            if not want_smoother:
                continue
            K = Pp @ hb / (hb @ Pp @ hb + rj)
            Pf = Pp - rj * np.outer(K, K)
            C = np.linalg.cholesky(Ab @ Pf @ Ab.T + Qb)
            G = sla.cho_solve((C, True), Ab @ Pf.T).T          # Pf A' / (A Pf A' + Q)
            QQ = Pf - G @ Pp @ G.T
            QQ = 0.5 * (QQ + QQ.T)
            ev, V = np.linalg.eigh(QQ)
            keep = ev > 0
            QQ = (V[:, keep] * ev[keep]) @ V[:, keep].T
            # dare(G',0,QQ) == discrete Lyapunov  X = G X G' + QQ
            Ps = sla.solve_discrete_lyapunov(G, QQ)
            bwd[j] = np.concatenate([Ps.reshape(-1, order="F"), G.reshape(-1, order="F")])
        PP.append(U @ fwd)
        if want_smoother:
            PG.append(U @ bwd)
    return IhgpTables(r, ro, PP, PG if want_smoother else None)


def _build_tables_native(model, want_smoother):
    import ctypes as C

    from . import _lib
    arrs = [_lib.as_f64(a) for a in (model.A, model.Q, model.Pinf, model.h)]
    cm = _lib.Model()
    cm.D, cm.N, cm.bz, cm.bg = model.D, model.N, model.bz, model.bg
    cm.A, cm.Q, cm.Pinf, cm.h = [_lib.dptr(a) for a in arrs]
    sizes = model.block_sizes()
    r = np.empty(N_FINE)
    pp = np.empty(N_FINE * int(sum(b * b for b in sizes)))
    pg = np.empty(2 * pp.size) if want_smoother else None
    _lib.check(_lib.lib().nsagp_ihgp_tables(C.byref(cm), int(want_smoother), N_COARSE, N_FINE, -2.0, 4.0, _lib.dptr(r),
                                            _lib.dptr(pp), _lib.dptr(pg) if want_smoother else None))
    PP, PG, o1, o2 = [], [], 0, 0
    for b in sizes:
        PP.append(pp[o1:o1 + N_FINE * b * b].reshape(N_FINE, b * b)); o1 += N_FINE * b * b
        if want_smoother:
            PG.append(pg[o2:o2 + N_FINE * 2 * b * b].reshape(N_FINE, 2 * b * b)); o2 += N_FINE * 2 * b * b
    # the grid the look-up thresholds are derived from is the host's own logspace (bitwise what the
    # reference's logspace gives); the library's pow()-based copy may differ from it in the last ulp
    grid = np.logspace(-2, 4, N_FINE)
    assert np.allclose(r, grid, rtol=1e-14, atol=0)
    return IhgpTables(grid, np.logspace(-2, 4, N_COARSE), PP, PG if want_smoother else None)
