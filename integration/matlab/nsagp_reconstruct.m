function [Esig, Vsig, Eft_mod, Varft_mod] = nsagp_reconstruct(Eft, Varft, W, link_shift, sqrt_model, s, Z_or_seed)
% NSAGP_RECONSTRUCT - Monte-Carlo reconstruction of the signal and of the NMF components from the posterior
% marginals on the GPU: what matlab/demo_toy_modulators_nmf.m:119-165 (sqrt_model = 0) and
% experiments/missing_data_music.m:138-176 (sqrt_model = 1) compute with s samples per time step.
%   Z_or_seed : T-by-s-by-(D+N) standard-normal draws, page i belonging to latent i.  Filling the pages with
%               randn(T,s) in the order i = 1, D+1, 2, D+2, ... (the reference's loop) reproduces its random stream;
%               or a scalar seed: the draws are generated on the device.
  if nargin < 7, Z_or_seed = 0; end
  [Esig, Vsig, Eft_mod, Varft_mod] = nsagp_mex('mc_reconstruct', Eft, Varft, W, link_shift, sqrt_model, s, Z_or_seed);
end
