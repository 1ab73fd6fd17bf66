/*
 * nsagp_oracle.c -- second, independent CPU restatement (plain C, FP64) of the
 * infinite-horizon Power-EP hot path of AaltoML/nonstationary-audio-gp.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke
 * check and bench.py's cpu_baseline / --impl reference legs may load the shared
 * object built from it (oracle/c/Makefile -> oracle/c/libnsagp_oracle.so).  The
 * product library (csrc/libnsagp.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (MATLAB) ships no golden vectors for this path and
 * cannot be executed here; this restatement is pinned against the NumPy oracle
 * (the oracle/ Python modules, written separately, different operation layout) and the committed
 * fixtures under tests/golden/.
 *
 * It follows the reference's own operation sequence with DENSE n-by-n matrices, as
 * the MATLAB code executes it:
 *   oracle_mom            matlab/likModulatorNMFPower.m:28-87,
 *                         matlab/experiments/likModulatorPreCalcwn.m:28-86
 *   oracle_ihgp_predict   matlab/ihgp_ep_modulator_nmf.m:195-526
 *   oracle_ihgp_nlz       matlab/ihgp_ep_modulator_nmf.m:533-624 and the running-site
 *                         variant of matlab/ihgp_ep_modulator_nmf_constraints.m:568-651
 * The steady-state tables (dare, :90-191) are inputs, as for the product library.
 * All matrices are column-major (MATLAB layout).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define JITTER 1e-10                      /* likModulatorNMFPower.m:28 */
#define PI 3.14159265358979323846

typedef struct oracle_problem {
  int D, N, n;                            /* M = D + N sites, state dimension n */
  const int *ilist;                       /* [M+1] block starts (0-based), ihgp_ep_modulator_nmf.m:104 */
  const double *A, *H, *Pinf;             /* n-by-n, M-by-n, n-by-n dense */
  int nr;                                 /* table rows (200) */
  const double *r;                        /* [nr] */
  const double *PP;                       /* per block nr rows of b*b (PP(:)'), blocks concatenated (:131-133) */
  const double *PG;                       /* per block nr rows of 2*b*b ([PS2(:)' G(:)']), may be NULL (:183-189) */
  int kind;                               /* 0 likModulatorNMFPower, 1 likModulatorPreCalcwn */
  double lik_param;                       /* log noise variance */
  double shift;                           /* link = log(1+exp(g-shift)) */
  const double *W;                        /* D-by-N */
  int S;
  const double *wn, *xn;                  /* [S], N-by-S unit sigma points */
  double alpha;                           /* ep_fraction */
  const double *damping;                  /* [ep_itts] */
  int ep_itts;
} oracle_problem;

static double max_nan(double a, double b) { return (a > b || isnan(b)) ? a : b; }   /* MATLAB max(a, b) for a never NaN: NaN loses */

/* [lZ, dlZ, d2lZ] = mom(...)  -- one call of the likelihood file. */
int oracle_mom(int kind, double lik_param, double shift, int D, int N, const double *W, int S, const double *wn,
               const double *xn_unscaled, double alpha, double y, const double *mu, const double *s2,
               double *lZ, double *dlZ, double *d2lZ) {
  const int M = D + N;
  const double sn2 = exp(lik_param);
  const double *mu_z = mu, *mu_g = mu + D, *s2_z = s2, *s2_g = s2 + D;
  double *buf = (double *)malloc(sizeof(double) * ((size_t)S * (N + D + 1) + 2 * M));
  if (!buf) return -1;
  double *xn = buf;                       /* S-by-N  (:34) */
  double *a = xn + (size_t)S * N;         /* S-by-D  link(xn)*W' (:44) */
  double *pdf = a + (size_t)S * D;        /* S */
  double *s1 = pdf + S, *s2acc = s1 + M;
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < N; ++j) xn[s + (size_t)j * S] = mu_g[j] + sqrt(s2_g[j]) * xn_unscaled[j + (size_t)s * N];
  for (int s = 0; s < S; ++s)
    for (int d = 0; d < D; ++d) {
      double acc = 0.0;
      for (int j = 0; j < N; ++j) acc += log(1.0 + exp(xn[s + (size_t)j * S] - shift)) * W[d + (size_t)j * D];
      a[s + (size_t)d * S] = (kind == 1) ? sqrt(acc) : acc;            /* PreCalcwn.m:44 */
    }
  const double pep = (kind == 1) ? pow(2.0 * PI * sn2, 0.5 * (1.0 - alpha)) * pow(alpha, -0.5) : 1.0;   /* :48 / :49 */
  double Zsum = 0.0;
  for (int i = 0; i < M; ++i) { s1[i] = 0.0; s2acc[i] = 0.0; }
  for (int s = 0; s < S; ++s) {
    double v = sn2 / alpha, m = 0.0;                                    /* :45-46 */
    for (int d = 0; d < D; ++d) {
      const double ad = a[s + (size_t)d * S];
      v += ad * ad * s2_z[d];
      m += ad * mu_z[d];
    }
    const double sd = sqrt(v), t = (y - m) / sd;
    pdf[s] = exp(-0.5 * t * t) / (sqrt(2.0 * PI) * sd);                 /* normpdf (:51-53) */
    Zsum += wn[s] * pdf[s];
    const double q = (y - m) / v;
    for (int d = 0; d < D; ++d) {
      const double ad = a[s + (size_t)d * S];
      s1[d] += wn[s] * (ad * (q * pdf[s]));                             /* :59-61 */
      s2acc[d] += wn[s] * (ad * ad * ((q * q - 1.0 / v) * pdf[s]));     /* :71-74 */
    }
    for (int j = 0; j < N; ++j) {
      const double e = (xn[s + (size_t)j * S] - mu_g[j]) / s2_g[j];     /* :47 */
      s1[D + j] += wn[s] * (e * pdf[s]);                                /* :65-66 */
      s2acc[D + j] += wn[s] * ((e * e - 1.0 / s2_g[j]) * pdf[s]);       /* :77-79 */
    }
  }
  const double Z = pep * max_nan(JITTER, Zsum);                         /* :55 */
  const double Zinv = 1.0 / Z;
  *lZ = log(Z);
  for (int i = 0; i < M; ++i) {
    dlZ[i] = Zinv * pep * s1[i];                                        /* :63,:68 */
    d2lZ[i] = -dlZ[i] * dlZ[i] + Zinv * pep * s2acc[i];                 /* :75,:80 */
  }
  free(buf);
  return 0;
}

static int mom_p(const oracle_problem *p, double alpha, double y, const double *mu, const double *s2,
                 double *lZ, double *dlZ, double *d2lZ) {
  return oracle_mom(p->kind, p->lik_param, p->shift, p->D, p->N, p->W, p->S, p->wn, p->xn, alpha, y, mu, s2, lZ, dlZ, d2lZ);
}

/* [~,ind] = min(abs(r-R))  (:239): first index on ties; all distances Inf/NaN -> first index. */
static int lookup_min(const double *r, int nr, double R) {
  int best = 0;
  double bd = fabs(r[0] - R);
  for (int i = 1; i < nr; ++i) {
    const double d = fabs(r[i] - R);
    if (d < bd) { bd = d; best = i; }
  }
  return best;
}

static size_t block_table_offset(const oracle_problem *p, int blk, int per) {   /* per = 1 (PP) or 2 (PG) */
  size_t off = 0;
  for (int i = 0; i < blk; ++i) {
    const int b = p->ilist[i + 1] - p->ilist[i];
    off += (size_t)p->nr * per * b * b;
  }
  return off;
}

/* Forward pass (:233-310).  first_pass: moments at every step, else only at the last.
 * running: _constraints nlZ variant, site vectors carried from step to step.
 * Returns the scalar lZ accumulated by the pass; lZk (may be NULL) gets the per-step terms. */
static double filter_pass(const oracle_problem *p, const double *y, long T, double ep_damp, int first_pass, int running,
                          double *ttau, double *tnu, double *R, double *MS, double *m, double *lZk) {
  const int D = p->D, N = p->N, M = D + N, n = p->n, nr = p->nr;
  double *HA = (double *)calloc((size_t)M * n, sizeof(double));
  double *PPd = (double *)calloc((size_t)n * n, sizeof(double));
  double *Wd = (double *)calloc((size_t)n * M, sizeof(double));
  double *fmu = (double *)calloc(8 * (size_t)M + 2 * (size_t)n, sizeof(double));
  double *HPH = fmu + M, *dl = HPH + M, *d2 = dl + M, *ttk = d2 + M, *tnk = ttk + M, *Rk = tnk + M, *trun = Rk + M;
  double *mnew = trun + M, *tmp = mnew + n;
  double *tt_run = (double *)calloc(3 * (size_t)M, sizeof(double)), *tn_run = tt_run + M, *R_run = tn_run + M;
  double lZ = 0.0;
  (void)trun;
  for (int i = 0; i < M; ++i)                                           /* H*A */
    for (int c = 0; c < n; ++c) {
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc += p->H[i + (size_t)l * M] * p->A[l + (size_t)c * n];
      HA[i + (size_t)c * M] = acc;
    }
  for (long k = 0; k < T; ++k) {
    if (k > 0) {                                                        /* :236-245 */
      memset(PPd, 0, sizeof(double) * n * n);
      for (int b_ = 0; b_ < M; ++b_) {
        const int i0 = p->ilist[b_], b = p->ilist[b_ + 1] - i0;
        const double Rprev = running ? R_run[b_] : R[b_ + (size_t)(k - 1) * M];
        const int ind = lookup_min(p->r, nr, Rprev);
        const double *row = p->PP + block_table_offset(p, b_, 1) + (size_t)ind * b * b;
        for (int c = 0; c < b; ++c)
          for (int r_ = 0; r_ < b; ++r_) PPd[(i0 + r_) + (size_t)(i0 + c) * n] = row[r_ + c * b];
      }
    } else {
      memcpy(PPd, p->Pinf, sizeof(double) * n * n);                     /* :246 */
    }
    for (int i = 0; i < M; ++i) {                                       /* fmu = H*A*m; W = PP*H'; HPH = diag(H*W) (:250) */
      double acc = 0.0;
      for (int c = 0; c < n; ++c) acc += HA[i + (size_t)c * M] * m[c];
      fmu[i] = acc;
      for (int r_ = 0; r_ < n; ++r_) {
        double w = 0.0;
        for (int c = 0; c < n; ++c) w += PPd[r_ + (size_t)c * n] * p->H[i + (size_t)c * M];
        Wd[r_ + (size_t)i * n] = w;
      }
      double hph = 0.0;
      for (int c = 0; c < n; ++c) hph += p->H[i + (size_t)c * M] * Wd[c + (size_t)i * n];
      HPH[i] = hph;
    }
    for (int i = 0; i < M; ++i) {
      ttk[i] = running ? tt_run[i] : ttau[i + (size_t)k * M];
      tnk[i] = running ? tn_run[i] : tnu[i + (size_t)k * M];
    }
    if (first_pass || k == T - 1) {                                     /* :253-271 */
      double lz;
      mom_p(p, 1.0, y[k], fmu, HPH, &lz, dl, d2);
      lZ += lz;
      if (lZk) lZk[k] = lz;
      for (int i = 0; i < M; ++i) {
        const double den = 1.0 + d2[i] * HPH[i];
        ttk[i] = (1.0 - ep_damp) * ttk[i] + ep_damp * (-d2[i] / den);                 /* :265 */
        tnk[i] = (1.0 - ep_damp) * tnk[i] + ep_damp * ((dl[i] - fmu[i] * d2[i]) / den);   /* :266 */
        Rk[i] = 1.0 / ttk[i];                                           /* :269, before the clamp */
      }
    } else {
      for (int i = 0; i < M; ++i) Rk[i] = R[i + (size_t)k * M];
    }
    for (int i = 0; i < M; ++i) ttk[i] = max_nan(0.0, ttk[i]);          /* :274 */
    for (int b_ = 0; b_ < M; ++b_) {                                    /* :280-304 */
      const int i0 = p->ilist[b_], b = p->ilist[b_ + 1] - i0;
      for (int r_ = 0; r_ < b; ++r_) {
        double acc = 0.0;
        for (int c = 0; c < b; ++c) acc += p->A[(i0 + r_) + (size_t)(i0 + c) * n] * m[i0 + c];
        tmp[r_] = acc;                                                  /* A_b m_b */
      }
      if (ttk[b_] == 0.0) {
        Rk[b_] = INFINITY;                                              /* :287 */
        for (int r_ = 0; r_ < b; ++r_) mnew[i0 + r_] = tmp[r_];
      } else {
        const double ys = tnk[b_] / ttk[b_];                            /* :277 */
        double hAm = 0.0;                                               /* H_b A_b m_b */
        for (int c = 0; c < b; ++c) hAm += p->H[b_ + (size_t)(i0 + c) * M] * tmp[c];
        for (int r_ = 0; r_ < b; ++r_) {
          const double K = Wd[(i0 + r_) + (size_t)b_ * n] / (HPH[b_] + Rk[b_]);       /* :293 */
          mnew[i0 + r_] = (tmp[r_] - K * hAm) + K * ys;                 /* (A - K H A) m + K ys (:296-299) */
        }
      }
    }
    memcpy(m, mnew, sizeof(double) * n);
    if (running) {
      for (int i = 0; i < M; ++i) { tt_run[i] = ttk[i]; tn_run[i] = tnk[i]; R_run[i] = Rk[i]; }
    } else {
      for (int i = 0; i < M; ++i) {
        ttau[i + (size_t)k * M] = ttk[i]; tnu[i + (size_t)k * M] = tnk[i]; R[i + (size_t)k * M] = Rk[i];
      }
    }
    if (MS) memcpy(MS + (size_t)k * n, m, sizeof(double) * n);          /* :307 */
  }
  free(HA); free(PPd); free(Wd); free(fmu); free(tt_run);
  return lZ;
}

/* Predict mode (:195-526).  Outputs: MS n-by-T, ttau/tnu/R M-by-T, nlZ[ep_itts], Eft M-by-T,
 * Varft[M] (one vector, replicated over time by the reference, :492), maxDiffM[ep_itts]. */
int oracle_ihgp_predict(const oracle_problem *p, const double *y, long T, double *MS, double *ttau, double *tnu,
                        double *R, double *nlZ, double *Eft, double *Varft, long *n_negcav, double *maxDiffM) {
  const int D = p->D, N = p->N, M = D + N, n = p->n, nr = p->nr;
  double *m = (double *)calloc(4 * (size_t)n + 6 * (size_t)M + 2 * (size_t)n * n, sizeof(double));
  double *t1 = m + n, *t2 = t1 + n, *mk = t2 + n;
  double *mm = mk + n, *vm = mm + M, *vc = vm + M, *mc = vc + M, *dl = mc + M, *d2 = dl + M;
  double *Pd = d2 + M, *Gd = Pd + (size_t)n * n;
  double *MSP = (double *)malloc(sizeof(double) * (size_t)n * T);
  if (!m || !MSP) return -1;
  memset(MS, 0, sizeof(double) * (size_t)n * T);
  memset(ttau, 0, sizeof(double) * (size_t)M * T);
  memset(tnu, 0, sizeof(double) * (size_t)M * T);
  for (size_t i = 0; i < (size_t)M * T; ++i) R[i] = exp(p->lik_param);  /* :209 */
  for (int i = 0; i < p->ep_itts; ++i) { nlZ[i] = 0.0; maxDiffM[i] = 0.0; }
  *n_negcav = 0;
  double ep_damp = p->damping[0];
  for (int itt = 1; itt <= p->ep_itts; ++itt) {
    double md = 0.0;
    memcpy(MSP, MS, sizeof(double) * (size_t)n * T);                    /* :226 */
    double lZ = filter_pass(p, y, T, ep_damp, itt == 1, 0, ttau, tnu, R, MS, m, NULL);
    if (itt == 1) nlZ[0] = -lZ;
    if (itt < p->ep_itts) ep_damp = p->damping[itt];                    /* ep_damping(itt+1) (:369-371) */
    memset(Pd, 0, sizeof(double) * n * n);
    memset(Gd, 0, sizeof(double) * n * n);
    for (long k = T - 2; k >= 0; --k) {                                 /* :373 */
      for (int b_ = 0; b_ < M; ++b_) {                                  /* :379-388 */
        const int i0 = p->ilist[b_], b = p->ilist[b_ + 1] - i0;
        const double Rn = R[b_ + (size_t)k * M];
        const int ind = isinf(Rn) ? nr - 1 : lookup_min(p->r, nr, Rn);
        const double *row = p->PG + block_table_offset(p, b_, 2) + (size_t)ind * 2 * b * b;
        for (int c = 0; c < b; ++c)
          for (int r_ = 0; r_ < b; ++r_) {
            Pd[(i0 + r_) + (size_t)(i0 + c) * n] = row[r_ + c * b];
            Gd[(i0 + r_) + (size_t)(i0 + c) * n] = row[b * b + r_ + c * b];
          }
      }
      const double *msk = MS + (size_t)k * n;
      for (int r_ = 0; r_ < n; ++r_) {                                  /* m - A*MS(:,k) */
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc += p->A[r_ + (size_t)c * n] * msk[c];
        t1[r_] = m[r_] - acc;
      }
      for (int r_ = 0; r_ < n; ++r_) {                                  /* m = MS(:,k) + G*(...) (:391) */
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc += Gd[r_ + (size_t)c * n] * t1[c];
        mk[r_] = msk[r_] + acc;
      }
      memcpy(m, mk, sizeof(double) * n);
      memcpy(MS + (size_t)k * n, m, sizeof(double) * n);
      for (int i = 0; i < M; ++i) {                                     /* H*m, diag(H*P*H') */
        double e = 0.0, v = 0.0;
        for (int c = 0; c < n; ++c) e += p->H[i + (size_t)c * M] * m[c];
        for (int c = 0; c < n; ++c) {
          double hp = 0.0;
          for (int l = 0; l < n; ++l) hp += p->H[i + (size_t)l * M] * Pd[l + (size_t)c * n];
          v += hp * p->H[i + (size_t)c * M];
        }
        mm[i] = e; vm[i] = v;
      }
      if (itt < p->ep_itts && !isnan(y[k])) {                           /* :397-437 */
        for (int i = 0; i < M; ++i) {
          vc[i] = 1.0 / (1.0 / vm[i] - p->alpha * ttau[i + (size_t)k * M]);             /* :407 */
          mc[i] = vc[i] * (mm[i] / vm[i] - p->alpha * tnu[i + (size_t)k * M]);          /* :408 */
        }
        double lz;
        mom_p(p, p->alpha, y[k], mc, vc, &lz, dl, d2);
        if (itt > 1) lZ += lz;                                          /* :420 */
        for (int i = 0; i < M; ++i) {
          if (vc[i] > 0.0) {                                            /* :411 */
            const double den = 1.0 + d2[i] * vc[i];
            const size_t o = i + (size_t)k * M;
            ttau[o] = (1.0 - ep_damp * p->alpha) * ttau[o] + ep_damp * (-d2[i] / den);             /* :428 */
            tnu[o] = (1.0 - ep_damp * p->alpha) * tnu[o] + ep_damp * ((dl[i] - mc[i] * d2[i]) / den);   /* :430 */
            R[o] = 1.0 / ttau[o];                                       /* :434 */
          } else {
            ++*n_negcav;
          }
        }
      }
      for (int i = 0; i < M; ++i) {                                     /* maxDiffM (:440) */
        double e = 0.0;
        for (int c = 0; c < n; ++c) e += p->H[i + (size_t)c * M] * MSP[c + (size_t)k * n];
        const double d = fabs(e - mm[i]);
        if (d > md) md = d;
      }
    }
    if (itt < p->ep_itts) nlZ[itt] = -lZ;
    maxDiffM[itt - 1] = md;
  }
  for (long k = 0; k < T; ++k)
    for (int i = 0; i < M; ++i) {
      double e = 0.0;
      for (int c = 0; c < n; ++c) e += p->H[i + (size_t)c * M] * MS[c + (size_t)k * n];
      Eft[i + (size_t)k * M] = e;                                       /* :488 */
    }
  for (int i = 0; i < M; ++i) Varft[i] = fabs(vm[i]);                   /* :492-496: P of the last look-up (k = 1) */
  (void)D; (void)N; (void)t2;
  free(m); free(MSP);
  return 0;
}

/* nlZ mode: a single ADF sweep (:533-624); running != 0: _constraints variant (:568-651). */
int oracle_ihgp_nlz(const oracle_problem *p, const double *y, long T, int running, double *edata,
                    double *ttau, double *tnu, double *R, double *lZk) {
  const int M = p->D + p->N;
  double *m = (double *)calloc((size_t)p->n, sizeof(double));
  if (!m) return -1;
  memset(ttau, 0, sizeof(double) * (size_t)M * T);
  memset(tnu, 0, sizeof(double) * (size_t)M * T);
  memset(R, 0, sizeof(double) * (size_t)M * T);
  memset(lZk, 0, sizeof(double) * (size_t)T);
  const double lZ = filter_pass(p, y, T, p->damping[0], 1, running, ttau, tnu, R, NULL, m, lZk);
  if (running) {
    *edata = -lZ;
  } else {
    double s = 0.0;
    for (long k = 0; k < T; ++k) s += lZk[k];                           /* -sum(lZ) (:619) */
    *edata = -s;
  }
  free(m);
  return 0;
}
