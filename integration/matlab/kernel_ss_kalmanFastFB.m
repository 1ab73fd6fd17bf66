function  [lik,Xfin,Pfin,varargout] = kernel_ss_kalmanFastFB(A,Q,C,P0,K,vary,y,varargin)
% KERNEL_SS_KALMANFASTFB - drop-in for matlab/unifying_prob_tf/kernel_ss_kalmanFastFB.m of
% AaltoML/nonstationary-audio-gp with the two time loops (:86-112 filter, :134-148 smoother) on a B200
% (libnsagp.so via nsagp_mex 'fastfb').  Same arguments and outputs.  What stays in MATLAB is what the reference
% does once per call with built-ins: the two dare solves and a few n-by-n products.
  T = length(y);
  if nargin<=8, KF = 0; else, KF = varargin{2}; end
  H = C; R = vary;
  try
    PP = dare(A',H',Q,R);                       % :50
    S  = H*PP*H'+R;                             % :53
  catch
    error('Unstable DARE solution!')
  end
  Kg = PP*H'/S;                                 % :60
  AKHA = A-Kg*H*A;                              % :63
  PF2 = PP - Kg*H*PP;                           % :76
  HA = H*A;
  G = [];
  if KF~=1
    G = PF2*A'/PP;                              % :126
    QQ = PF2-G*PP*G'; QQ = (QQ+QQ')/2;
    P = dare(G',zeros(size(QQ)),QQ);            % :131
  end
  [MS, quad] = nsagp_mex('fastfb', A, AKHA, Kg(:), HA(:), S, G, y(:));
  lik = -(.5*log(2*pi)*T + .5*log(S)*T + quad); % :80, :99, :152
  PS = repmat(PF2,[1 1 T]);
  if KF~=1, PS(:,:,1:T-1) = repmat(P,[1 1 T-1]); end
  Xfin = reshape(MS,[1 size(MS)]);
  Pfin = PS;
  varargout = cell(1,max(nargout-3,0));
end
