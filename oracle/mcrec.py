"""Oracle: Monte-Carlo reconstruction of the signal and the NMF components from the posterior
marginals (test infrastructure only).

Restates matlab/demo_toy_modulators_nmf.m:119-165 and, for the "sqrt" model,
matlab/experiments/missing_data_music.m:138-176, with the normal draws passed in explicitly
(``Z[:, :, i]`` is the ``randn(T, s)`` matrix the reference draws for latent i; it draws them in the
order i = 1, D+1, 2, D+2, ... inside its plotting loop).
"""
import numpy as np


def reconstruct(Eft, Varft, W, Z, link_shift=0.0, sqrt_model=False):
    """Eft, Varft: (M, T); W: (D, N); Z: (T, s, M).  Returns Esig, Vsig (T,), Eft_mod, Varft_mod (N, T)."""
    Eft = np.asarray(Eft, float); Varft = np.asarray(Varft, float); W = np.asarray(W, float)
    D, N = W.shape
    T, s = Z.shape[0], Z.shape[1]
    link = lambda g: np.log(1 + np.exp(g - link_shift))
    sub_samp = np.zeros((D, T, s)); mod_samp = np.zeros((N, T, s))
    Eft_mod = np.zeros((N, T)); Varft_mod = np.zeros((N, T))
    for i in range(D):
        sub_samp[i] = Z[:, :, i] * np.sqrt(Varft[i])[:, None] + Eft[i][:, None]                  # :135
        if i < N:
            mod_samp[i] = Z[:, :, D + i] * np.sqrt(Varft[D + i])[:, None] + Eft[D + i][:, None]  # :145
    for i in range(N):          # the reference only fills i <= min(D, N); D >= N in every caller
        Eft_mod[i] = np.mean(link(mod_samp[i]), axis=1)                                          # :146
        Varft_mod[i] = np.var(link(mod_samp[i]), axis=1, ddof=1) if s > 1 else 0.0               # :147
    sig_samp = np.zeros((T, s))
    for v in range(s):
        a = W @ link(mod_samp[:, :, v])                                                          # :160
        if sqrt_model:
            a = np.sqrt(a)                                                                       # missing_data_music.m:168
        sig_samp[:, v] = np.sum(a * sub_samp[:, :, v], axis=0)
    Esig = np.mean(sig_samp, axis=1)                                                             # :162
    Vsig = np.var(sig_samp, axis=1, ddof=1) if s > 1 else np.zeros(T)                            # :163
    return Esig, Vsig, Eft_mod, Varft_mod
