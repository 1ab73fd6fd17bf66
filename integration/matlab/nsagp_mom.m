function mom = nsagp_mom(likname, link_shift, varargin)
% NSAGP_MOM - descriptor that replaces the `mom` function handle of the reference
%
%   mom = nsagp_mom('likModulatorNMFPower', c, p_cubature, N)
%   mom = nsagp_mom('likModulatorPreCalcwn', c, wn, xn_unscaled)
%
% The reference builds  mom = @(hyp,mu,s2,nmfW,ep_frac,yall,k) feval(likfunc,link,...)
% (demo_toy_modulators_nmf.m:78-81, experiments/train_model.m:186-190).  A handle cannot
% cross into CUDA, so the B200 entry points take this struct instead: which likelihood
% file, the shift c of the link log(1+exp(g-c)), and the unit sigma points.  The sigma
% points come from the reference's own utp_ws / mvhermgauss, so quirks such as the
% ut9_ws weight sum (SURVEY.md F7) carry over unchanged.
  mom.link_shift = link_shift;
  switch likname
    case 'likModulatorNMFPower'
      p = varargin{1}; N = varargin{2};
      mom.kind = 0;
      if any(p == [3 5 7 9])
        [wn, xn] = utp_ws(p, N);                 % likModulatorNMFPower.m:32-35
      else
        [xn, wn] = mvhermgauss(zeros(N,1), ones(N,1), p);   % :41, unit points
        xn = xn'; wn = wn(:)';
      end
      mom.wn = wn(:)'; mom.xn = xn;              % xn is N-by-S
    case 'likModulatorPreCalcwn'
      mom.kind = 1;
      mom.wn = varargin{1}(:)'; mom.xn = varargin{2};
    otherwise
      error('nsagp:mom', 'unknown likelihood %s', likname);
  end
end
