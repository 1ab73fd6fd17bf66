function [varargout] = gf_ep_modulator_nmf_constraints(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,ep_fraction,ep_damping,ep_itts,constraints,w_fixed,tune_hypers)
% Drop-in for matlab/gf_ep_modulator_nmf_constraints.m (box-constrained parameters,
% balanced model, :75-117; the loops are those of gf_ep_modulator_nmf).
  mom = nsagp_resolve_mom(mom, N);      % the reference's own closure (or an nsagp_mom descriptor)
  [yall, return_ind] = nsagp_merge(x, y, xt);
  [lik_param, param1, param2, Wnmf] = nsagp_unpack_constraints(w, w_fixed, tune_hypers, constraints, num_lik_params, D, N);
  out = nsagp_run('full', lik_param, param1, param2, Wnmf, x, yall, ss, mom, xt, kernel1, kernel2, D, N, ...
                  ep_fraction, ep_damping, ep_itts, true, 1);
  varargout = nsagp_outputs(out, return_ind, numel(w), isempty(xt), nargout);
end
