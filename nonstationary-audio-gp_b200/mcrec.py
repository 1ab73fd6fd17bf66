"""Monte-Carlo reconstruction of the signal and the NMF components from the posterior marginals --
what every demo of the reference does with ``Eft, Varft`` (matlab/demo_toy_modulators_nmf.m:119-165;
sqrt model: matlab/experiments/missing_data_music.m:138-176).  Device kernel: csrc/mcrec.cuh."""
import ctypes as C

import numpy as np

from . import _lib


def reconstruct_signal(Eft, Varft, W, s=250, link_shift=0.0, sqrt_model=False, Z=None, seed=0):
    """Eft, Varft: (M, T) marginals of the M = D + N latents; W: (D, N).
    ``Z`` (T, s, M): explicit standard-normal draws (page i for latent i), else the draws are
    generated on the device from ``seed``.  Returns dict(Esig, Vsig (T,), Eft_mod, Varft_mod (N, T))."""
    Eft = np.asarray(Eft, float); Varft = np.asarray(Varft, float); W = np.asarray(W, float)
    D, N = W.shape
    M, T = Eft.shape
    if M != D + N or Varft.shape != Eft.shape:
        raise ValueError("Eft/Varft must be (D+N, T)")
    Ef = _lib.as_f64(Eft.T); Vf = _lib.as_f64(Varft.T)          # device layout [T][M] = MATLAB column-major
    Wf = np.asfortranarray(W)
    zp = None
    if Z is not None:
        Z = np.asarray(Z, float)
        if Z.shape != (T, s, M):
            raise ValueError("Z must be (T, s, M)")
        zb = np.ascontiguousarray(np.transpose(Z, (2, 1, 0)))   # [M][s][T]
        zp = _lib.dptr(zb)
    Esig = np.empty(T); Vsig = np.empty(T); Em = np.empty((T, N)); Vm = np.empty((T, N))
    _lib.check(_lib.lib().nsagp_mc_reconstruct(D, N, T, int(s), _lib.dptr(Ef), _lib.dptr(Vf), Wf.ctypes.data_as(_lib.c_double_p),
                                               float(link_shift), int(bool(sqrt_model)), zp, C.c_uint64(int(seed)),
                                               _lib.dptr(Esig), _lib.dptr(Vsig), _lib.dptr(Em), _lib.dptr(Vm)))
    return dict(Esig=Esig, Vsig=Vsig, Eft_mod=Em.T.copy(), Varft_mod=Vm.T.copy())
