function [varargout] = ihgp_ep_modulator_nmf_constraints(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,ep_fraction,ep_damping,ep_itts,constraints,w_fixed,tune_hypers)
% Drop-in for matlab/ihgp_ep_modulator_nmf_constraints.m (box-constrained parameters;
% its nlZ mode carries the site vectors from step to step, :568-615 -> mode 2).
  mom = nsagp_resolve_mom(mom, N);      % the reference's own closure (or an nsagp_mom descriptor)
  [yall, return_ind] = nsagp_merge(x, y, xt);
  [lik_param, param1, param2, Wnmf] = nsagp_unpack_constraints(w, w_fixed, tune_hypers, constraints, num_lik_params, D, N);
  out = nsagp_run('ihgp', lik_param, param1, param2, Wnmf, x, yall, ss, mom, xt, kernel1, kernel2, D, N, ...
                  ep_fraction, ep_damping, ep_itts, true, 2);
  varargout = nsagp_outputs(out, return_ind, numel(w), isempty(xt), nargout);
end
