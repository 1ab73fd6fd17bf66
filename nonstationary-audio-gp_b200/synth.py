"""Synthetic signals of the reference's generative model (for benchmarks, tests, demos).

Restates the recipe of matlab/demo_toy_modulators_nmf.m:27-53 (and the "sqrt"
variant of matlab/experiments/synthetic_data_experiment.m:95-122) with a NumPy
generator: MATLAB's ``rng(100,'twister')`` stream cannot be reproduced without
MATLAB, so seeds are ours and recorded in every fixture.
"""
import dataclasses

import numpy as np

from . import ssmodel


@dataclasses.dataclass
class Hypers:
    """Natural-scale hyper-parameters in the reference's packing order
    ``w = log([w_lik; var_fast; len_fast; omega; var_slow; len_slow; W(:)])``
    (demo_toy_modulators_nmf.m:89)."""
    w_lik: float
    var_fast: np.ndarray
    len_fast: np.ndarray
    omega: np.ndarray
    var_slow: np.ndarray
    len_slow: np.ndarray
    W: np.ndarray            # (D, N)

    @property
    def D(self):
        return self.W.shape[0]

    @property
    def N(self):
        return self.W.shape[1]

    def pack_log(self):
        return np.log(np.concatenate([[self.w_lik], self.var_fast, self.len_fast, self.omega,
                                      self.var_slow, self.len_slow, self.W.reshape(-1, order="F")]))

    def w_sub(self):
        return np.concatenate([self.var_fast, self.len_fast, self.omega])

    def w_mod(self):
        return np.concatenate([self.var_slow, self.len_slow])


def demo_hypers(D, N, rng, w_lik=1e-4):
    """demo_toy_modulators_nmf.m:28-33."""
    return Hypers(w_lik,
                  var_fast=0.01 * np.ones(D),
                  len_fast=150 + 400 * rng.random(D),
                  omega=np.linspace(np.pi / 3, np.pi / 50, D),
                  var_slow=5 + 5 * rng.random(N),
                  len_slow=np.linspace(200, 1500, N),
                  W=0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2))


def speech_hypers(D, N, rng, w_lik=1e-4, fs=16000.0):
    """"Speech-shaped" configuration of BASELINE.md C2/C3: centre
    frequencies log-spaced over 80 Hz .. 4 kHz at fs = 16 kHz, subband length
    scales in [100, 2000], modulator length scales in [200, 5000] samples (ranges
    of experiments/train_model.m:158-163)."""
    return Hypers(w_lik,
                  var_fast=0.01 * np.ones(D),
                  len_fast=np.exp(rng.uniform(np.log(100.0), np.log(2000.0), D)),
                  omega=2 * np.pi * np.geomspace(4000.0, 80.0, D) / fs,
                  var_slow=5 + 5 * rng.random(N),
                  len_slow=np.geomspace(200.0, 5000.0, N),
                  W=0.1 * np.abs((2.0 * rng.random((D, N))) ** 2 - 0.2))


def _var1_paths(A, C, T, rng, chunk=1024):
    """x_0 ~ given below by caller; here: responses of x_k = A x_{k-1} + C e_k,
    e_k ~ N(0,I), x_{-1} = 0, for k = 0..T-1, computed chunk-parallel."""
    n = A.shape[0]
    nch = -(-T // chunk)
    E = rng.standard_normal((nch, chunk, n)) @ C.T
    X = np.empty((nch, chunk, n))
    x = np.zeros((nch, n))
    for j in range(chunk):                       # all chunks advance together
        x = x @ A.T + E[:, j]
        X[:, j] = x
    pw = np.empty((chunk, n, n))
    P = np.eye(n)
    for j in range(chunk):
        P = A @ P
        pw[j] = P                                 # A^(j+1)
    carry = np.zeros(n)
    for c in range(nch):                          # chunk-to-chunk carry
        if c > 0:
            X[c] += np.einsum("jab,b->ja", pw, carry)
        carry = X[c, -1]
    return X.reshape(nch * chunk, n)[:T]


def sample_signal(hyp, kernel1, kernel2, T, rng, link_shift=0.0, sqrt_model=False):
    """Draw the latent state from the discretised prior and form
    ``y_k = (H_z x_k)' a(H_g x_k)`` with a(g) = W softplus(g - shift), or its
    elementwise sqrt for the spectrogram model.  No observation noise is added
    (demo_toy_modulators_nmf.m:47-52).  Returns (y (T,), subbands (T,D), modulators (T,N))."""
    F, L, Qc, H, Pinf = ssmodel.ss_modulators_nmf(hyp.w_sub(), hyp.w_mod(), kernel1, kernel2)[:5]
    A, Q = ssmodel.lti_disc(F, L, Qc, 1.0)
    Q = 0.5 * (Q + Q.T)
    n = A.shape[0]
    ev, V = np.linalg.eigh(Q)
    C = V * np.sqrt(np.clip(ev, 0.0, None))       # C C' = Q (Q is tiny and ill-conditioned: no chol)
    X = _var1_paths(A, C, T, rng)
    # stationary start: x_0 ~ N(0, Pinf) propagated as A^k x_0
    evp, Vp = np.linalg.eigh(Pinf)
    x0 = (Vp * np.sqrt(np.clip(evp, 0.0, None))) @ rng.standard_normal(n)
    # remove the k=0 innovation's double counting: x_k = A^k x0 + sum_{j<=k, j>=1} A^(k-j) C e_j
    X = X - _impulse(A, X[0], T) + _impulse(A, x0, T)
    D, N = hyp.D, hyp.N
    zf = X @ H[:D].T
    g = X @ H[D:].T
    a = np.log1p(np.exp(g - link_shift)) @ hyp.W.T
    if sqrt_model:
        a = np.sqrt(a)
    y = np.sum(zf * a, axis=1)
    return y, zf, g


def _impulse(A, x0, T, chunk=1024):
    """A^k x0 for k = 0..T-1."""
    n = A.shape[0]
    out = np.empty((T, n))
    pw = np.empty((chunk, n, n))
    P = np.eye(n)
    for j in range(chunk):
        pw[j] = P
        P = A @ P
    AL = P                                         # A^chunk
    x = np.asarray(x0, float)
    for s in range(0, T, chunk):
        e = min(chunk, T - s)
        out[s:s + e] = np.einsum("jab,b->ja", pw[:e], x)
        x = AL @ x
    return out


def add_gaps(y, rng, n_gaps_per_20k=6, min_len=10, max_len=320):
    """Missing-data gaps as in experiments/missing_data_music.m:51,57: NaN runs of
    10..320 samples, six per 20 000 samples."""
    y = np.array(y, float)
    T = y.size
    n_gaps = max(1, int(round(n_gaps_per_20k * T / 20000.0)))
    lens = np.exp(rng.uniform(np.log(min_len), np.log(max_len), n_gaps)).astype(int)
    for ln in lens:
        s = int(rng.integers(1, max(2, T - ln - 1)))
        y[s:s + ln] = np.nan
    return y
