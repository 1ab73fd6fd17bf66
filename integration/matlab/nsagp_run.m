function out = nsagp_run(kind, lik_param, param1, param2, Wnmf, x, yall, ss, mom, xt, kernel1, kernel2, D, N, ep_fraction, ep_damping, ep_itts, do_balance, nlz_mode)
% Host-side setup with MATLAB built-ins, then one call into the CUDA library.
  [F,L,Qc,H,Pinf] = ss(x, param1, param2, kernel1, kernel2);
  if do_balance                                   % ihgp_ep_modulator_nmf.m:81-87
    [T,F] = balance(F); L = T\L; H = H*T;
    LL = T\chol(Pinf,'lower'); Pinf = LL*LL';
  end
  [A,Q] = lti_disc(F, L, Qc, 1);                  % dt = 1 is hard-coded in the reference
  predict = ~isempty(xt);
  lik = struct('kind', mom.kind, 'sn2', exp(lik_param(1)), 'link_shift', mom.link_shift, 'W', Wnmf, ...
               'wn', mom.wn, 'xn', mom.xn);
  ep = struct('ep_fraction', ep_fraction, 'ep_damping', ep_damping(:)', 'ep_itts', ep_itts);
  if predict, mode = 0; else, mode = nlz_mode; end
  if strcmp(kind, 'ihgp')
    Q = (Q+Q')/2;                                 % :97
    tables = nsagp_ihgp_tables(A, Q, H, predict);
    out = nsagp_mex('ep_ihgp', nsagp_blocks(A,Q,H,Pinf,D,N), lik, ep, tables, yall, mode);
    out.r = tables.r;
  else
    out = nsagp_mex('ep_full', nsagp_blocks(A,Q,H,Pinf,D,N), lik, ep, [], yall, min(mode,1));
  end
end
