function [varargout] = gf_giekf_modulator_nmf_constraints(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,g_iter,l_iter,constraints,w_fixed,tune_hypers,GradObj)
% Drop-in for matlab/gf_giekf_modulator_nmf_constraints.m (globally iterated EKF + RTS smoother,
% hard-wired measurement y = (H_z x)' W softplus(H_g x)) with the time loops on a B200.
% `mom` is ignored, as in the reference.  Only GradObj = 'off' is supported (how
% experiments/train_model.m:226,239-240 runs it): the second output is zeros.
  if nargin > 16 && ~isempty(GradObj) && ~strcmpi(GradObj, 'off')
    error('nsagp:grad', 'analytic EKF gradients are not provided; use GradObj = ''off''');
  end
  [yall, return_ind] = nsagp_merge(x, y, xt);
  [lik_param, param1, param2, Wnmf] = nsagp_unpack_constraints(w, w_fixed, tune_hypers, constraints, num_lik_params, D, N);
  [F,L,Qc,H,Pinf] = ss(x, param1, param2, kernel1, kernel2);
  [T,F] = balance(F); L = T\L; H = H*T;                   % :113-120
  LL = T\chol(Pinf,'lower'); Pinf = LL*LL';
  sigma2 = exp(lik_param(1));
  if ~isempty(xt)
    [A,Q] = lti_disc(F, L, Qc, 1);
    out = nsagp_mex('giekf', nsagp_blocks(A,Q,H,Pinf,D,N), Wnmf, sigma2, g_iter, l_iter, yall, 0);
    out.R = zeros(D+N, numel(yall));
    c = {out.Eft(:,return_ind), out.Varft(:,return_ind), [], out.lb(:,return_ind), out.ub(:,return_ind), out};
    varargout = c(1:max(nargout,1));
  else
    A = expm(F); Q = Pinf - A*Pinf*A';                      % :376-380
    out = nsagp_mex('giekf', nsagp_blocks(A,Q,H,Pinf,D,N), Wnmf, sigma2, 1, 1, yall, 1);
    varargout = {out.edata, zeros(1, numel(w))};
  end
end
