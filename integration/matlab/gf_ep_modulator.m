function [varargout] = gf_ep_modulator(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,ep_fraction,ep_damping,ep_itts)
% GF_EP_MODULATOR - drop-in for matlab/gf_ep_modulator.m, the model WITHOUT NMF weights (demo_toy_modulators.m:99,109):
% D carrier x modulator pairs, y = sum_d z_d link(g_d).  That file is gf_ep_modulator_nmf.m line by line with W = I
% and N = D (they differ in the `mom` signature, :138,:227) on the BALANCED model (:75-81), so this wrapper packs the
% pairs as an NMF model with identity weights and runs the same GPU path.  `ss` is the reference's
% @(x,p,kern1,kern2) ss_modulators(p,kern1,kern2); `mom` its likModulatorPower closure (demo_toy_modulators.m:88) or an
% nsagp_mom descriptor.  The sigma points live in D dimensions: D <= 4 pairs.  One deviation: the floor under the
% tilted normaliser is the NMF files' 1e-10, not likModulatorPower.m:29's 1e-8 (differs only where Z < 1e-8).
  D = (numel(w) - num_lik_params) / 5;
  assert(D == round(D) && D >= 1 && D <= 4, 'gf_ep_modulator on the GPU: 1..4 carrier x modulator pairs');
  mom = nsagp_resolve_mom(mom, D);
  [yall, return_ind] = nsagp_merge(x, y, xt);
  lik_param = w(1:num_lik_params);
  param = exp(w(num_lik_params+1:end));
  ss_nmf = @(x_,p1,p2,k1,k2) ss(x_, [p1(:); p2(:)], k1, k2);
  out = nsagp_run('full', lik_param, param(1:3*D), param(3*D+1:end), eye(D), x, yall, ss_nmf, mom, xt, kernel1, kernel2, D, D, ...
                  ep_fraction, ep_damping, ep_itts, true, 1);
  varargout = nsagp_outputs(out, return_ind, numel(w), isempty(xt), nargout);
end
