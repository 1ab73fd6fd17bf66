function [varargout] = gf_ep_modulator_nmf(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,ep_fraction,ep_damping,ep_itts)
% GF_EP_MODULATOR_NMF - drop-in for matlab/gf_ep_modulator_nmf.m (full-state Power EP:
% Kalman filter, RTS smoother, site updates) with the time loops on a B200.
% `mom`: the reference's closure or an nsagp_mom descriptor (nsagp_resolve_mom).  The model is NOT balanced here, as in the reference (:80).
  mom = nsagp_resolve_mom(mom, N);      % the reference's own closure (or an nsagp_mom descriptor)
  [yall, return_ind] = nsagp_merge(x, y, xt);
  lik_param = w(1:num_lik_params);
  param1 = exp(w(num_lik_params+1:num_lik_params+3*D));
  param2 = exp(w(num_lik_params+3*D+1:num_lik_params+3*D+2*N));
  Wnmf = reshape(exp(w(num_lik_params+3*D+2*N+1:end)), [D,N]);
  out = nsagp_run('full', lik_param, param1, param2, Wnmf, x, yall, ss, mom, xt, kernel1, kernel2, D, N, ...
                  ep_fraction, ep_damping, ep_itts, false, 1);
  varargout = nsagp_outputs(out, return_ind, numel(w), isempty(xt), nargout);
end
