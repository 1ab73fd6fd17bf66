function [varargout] = gf_giekf_modulator_nmf(w,x,y,ss,mom,xt,kernel1,kernel2,num_lik_params,D,N,g_iter,l_iter,GradObj)
% Drop-in for matlab/gf_giekf_modulator_nmf.m (globally iterated EKF + RTS smoother; callers:
% experiments/missing_data_music.m:128, noise_reduction_speech.m:97, synthetic_data_experiment.m:176)
% with the time loops on a B200.  Differences to the _constraints file: log-scale parameter vector
% (:70-73) and the state (m, P) is initialised on the first global iteration only (:127-131).
% `mom` is ignored, as in the reference.  GradObj = 'on' with xt empty returns the analytic gradient of
% :296-437 (sensitivity equations, one CUDA block per hyper-parameter: csrc/ekfgrad.cuh); as in the reference,
% dF and dPinf are NOT balanced (:82-84 is commented out there).
  if nargin < 14 || isempty(GradObj), GradObj = 'on'; end   % the reference runs the derivative loops unless 'off'
  [yall, return_ind] = nsagp_merge(x, y, xt);
  lik_param = w(1:num_lik_params);
  param1 = exp(w(num_lik_params+1:num_lik_params+3*D));
  param2 = exp(w(num_lik_params+3*D+1:num_lik_params+3*D+2*N));
  Wnmf = reshape(exp(w(num_lik_params+3*D+2*N+1:end)),[D,N]);
  if isempty(xt) && ~strcmpi(GradObj, 'off')
    [F,L,Qc,H,Pinf,dF,dQc,dPinf] = ss(x, param1, param2, kernel1, kernel2); %#ok<ASGLU>
  else
    [F,L,Qc,H,Pinf] = ss(x, param1, param2, kernel1, kernel2);
  end
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
    model = nsagp_blocks(A,Q,H,Pinf,D,N);
    if strcmpi(GradObj, 'off')
      out = nsagp_mex('giekf_carry', model, Wnmf, sigma2, 1, 1, yall, 1);
      varargout = {out.edata, zeros(1, numel(w))};
      return;
    end
    % per-parameter blocks for nsagp_giekf_grad (include/nsagp.h): every slice of dF / dPinf lives in one latent
    starts = [find(sum(abs(H),1) > 0), size(H,2)+1];
    nparam = 1 + size(dF,3); bmax = max(model.bz, model.bg);
    latent = -ones(nparam,1); dA = zeros(bmax,bmax,nparam); dQ = dA; dP0 = dA;
    dR = zeros(nparam,1); dR(1) = 1;                        % :93-96
    for j = 2:nparam
      [r,c] = find(dF(:,:,j-1) | dPinf(:,:,j-1));
      if isempty(r), continue; end
      lat = find(starts <= min([r;c]), 1, 'last'); ii = starts(lat):starts(lat+1)-1; b = numel(ii);
      assert(max([r;c]) <= ii(end), 'derivative slice %d is not confined to one latent', j-1);
      AA = expm([F(ii,ii) zeros(b); dF(ii,ii,j-1) F(ii,ii)]);   % :328-338 for that latent's block
      Al = AA(1:b,1:b); dAl = AA(b+1:end,1:b); dPl = dPinf(ii,ii,j-1);
      dAPAt = dAl*Pinf(ii,ii)*Al';
      latent(j) = lat-1; dA(1:b,1:b,j) = dAl; dP0(1:b,1:b,j) = dPl;
      dQ(1:b,1:b,j) = dPl - dAPAt - Al*dPl*Al' - dAPAt';       % :362-364
    end
    [edata, gdata] = nsagp_mex('giekf_grad', model, Wnmf, sigma2, latent, dA, dQ, dP0, dR, yall);
    ww = w(1:end-D*N);
    varargout = {edata, gdata(:)'.*exp(ww(:)')};            % :431-436
  end
end
