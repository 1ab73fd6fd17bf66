function [varargout] = ihgp_ep_modulator_nmf(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,ep_fraction,ep_damping,ep_itts)
% IHGP_EP_MODULATOR_NMF - drop-in for matlab/ihgp_ep_modulator_nmf.m of
% AaltoML/nonstationary-audio-gp with the time loops on a B200 (libnsagp.so via nsagp_mex).
%
% Same arguments and outputs as the reference ({nlZ, grad} when xt is empty, otherwise
% {Eft, Varft, Covft, lb, ub, out}); `mom` may be the reference's own closure
% (resolved by nsagp_resolve_mom from the variables it captured) or an nsagp_mom descriptor.
% What stays in MATLAB is what the reference also does once per call with built-ins:
% merging inputs, unpacking w, ss(...), balance, lti_disc and the DARE tables.
  mom = nsagp_resolve_mom(mom, N);      % the reference's own closure (or an nsagp_mom descriptor)
  [yall, return_ind] = nsagp_merge(x, y, xt);
  lik_param = w(1:num_lik_params);
  param1 = exp(w(num_lik_params+1:num_lik_params+3*D));
  param2 = exp(w(num_lik_params+3*D+1:num_lik_params+3*D+2*N));
  Wnmf = reshape(exp(w(num_lik_params+3*D+2*N+1:end)), [D,N]);
  out = nsagp_run('ihgp', lik_param, param1, param2, Wnmf, x, yall, ss, mom, xt, kernel1, kernel2, D, N, ...
                  ep_fraction, ep_damping, ep_itts, true, 1);
  varargout = nsagp_outputs(out, return_ind, numel(w), isempty(xt), nargout);
end
