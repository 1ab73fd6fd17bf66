function [varargout] = gf_giekf_modulator_nmf(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,g_iter,l_iter,GradObj)
% Drop-in for matlab/gf_giekf_modulator_nmf.m (globally iterated EKF + RTS smoother; callers:
% experiments/missing_data_music.m:128, noise_reduction_speech.m:97, synthetic_data_experiment.m:176)
% with the time loops on a B200.  Differences to the _constraints file: log-scale parameter vector
% (:70-73) and the state (m, P) is initialised on the first global iteration only (:127-131).
% `mom` is ignored, as in the reference.  Only GradObj = 'off' is supported: the second output is zeros.
  if nargin > 13 && ~isempty(GradObj) && ~strcmpi(GradObj, 'off')
    error('nsagp:grad', 'analytic EKF gradients are not provided; use GradObj = ''off''');
  end
  [yall, return_ind] = nsagp_merge(x, y, xt);
  lik_param = w(1:num_lik_params);
  param1 = exp(w(num_lik_params+1:num_lik_params+3*D));
  param2 = exp(w(num_lik_params+3*D+1:num_lik_params+3*D+2*N));
  Wnmf = reshape(exp(w(num_lik_params+3*D+2*N+1:end)),[D,N]);
  [F,L,Qc,H,Pinf] = ss(x, param1, param2, kernel1, kernel2);
  [T,F] = balance(F); L = T\L; H = H*T;                   % :78-85
  LL = T\chol(Pinf,'lower'); Pinf = LL*LL';
  sigma2 = exp(lik_param(1));
  if ~isempty(xt)
    [A,Q] = lti_disc(F, L, Qc, 1);
    out = nsagp_mex('giekf_carry', nsagp_blocks(A,Q,H,Pinf,D,N), Wnmf, sigma2, g_iter, l_iter, yall, 0);
    out.R = zeros(D+N, numel(yall));
    c = {out.Eft(:,return_ind), out.Varft(:,return_ind), [], out.lb(:,return_ind), out.ub(:,return_ind), out};
    varargout = c(1:max(nargout,1));
  else
    A = expm(F); Q = Pinf - A*Pinf*A';                      % :311-316, :342-344
    out = nsagp_mex('giekf_carry', nsagp_blocks(A,Q,H,Pinf,D,N), Wnmf, sigma2, 1, 1, yall, 1);
    varargout = {out.edata, zeros(1, numel(w))};
  end
end
