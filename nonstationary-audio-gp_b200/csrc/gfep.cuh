// Full-state Power-EP kernels (gf_ep_modulator_nmf) -- see gfep.cuh body below.
#pragma once
#include "common.cuh"
