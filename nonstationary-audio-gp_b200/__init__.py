"""nsagp-b200: the EP-in-Kalman inference hot path of AaltoML/nonstationary-audio-gp
on B200 (sm_100a), behind the reference's own entry points.

The directory name carries a hyphen (it mirrors the reference repository's name),
so import it with importlib:

    import importlib
    nsagp = importlib.import_module("nonstationary-audio-gp_b200")

All time loops run in csrc/libnsagp.so (C ABI: include/nsagp.h).  Nothing here
falls back to the CPU.
"""
from . import _lib, batch, chunked, cubature, mcrec, ssmodel, synth, tables       # noqa: F401
from ._lib import NsagpError, build                                       # noqa: F401
from .cubature import gauher, mvhermgauss_unit, utp_ws                    # noqa: F401
from .entry import (Plan, gf_ep_modulator, gf_ep_modulator_nmf, gf_ep_modulator_nmf_constraints,      # noqa: F401
                    gf_giekf_modulator_nmf, gf_giekf_modulator_nmf_constraints,
                    ihgp_ep_modulator_nmf, ihgp_ep_modulator_nmf_constraints,
                    inv_sigmoid, lambda_map, merge_inputs, sigmoid)
from .mcrec import reconstruct_signal                                     # noqa: F401
from .lik import Moments, Softplus, likModulatorNMFPower, likModulatorPower, likModulatorPreCalcwn      # noqa: F401
from .ssmodel import BlockModel, lti_disc, ss_modulators, ss_modulators_nmf, to_block_model         # noqa: F401

__version__ = "0.1.0"
