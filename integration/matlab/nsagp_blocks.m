function model = nsagp_blocks(A, Q, H, Pinf, D, N)
% NSAGP_BLOCKS - pack the block-diagonal discrete model for the C ABI (nsagp_model)
%
% ss_modulators_nmf builds F, L, Qc, H, Pinf with blkdiag, one block and one row of H
% per latent (ss_modulators_nmf.m:128-132), so A, Q, Pinf are block diagonal too.  The
% library takes the D subband blocks (size bz) and the N modulator blocks (size bg)
% packed column-major one after the other, and the matching pieces of the rows of H.
  starts = [find(sum(abs(H),1) > 0), size(H,2)+1];   % as ihgp_ep_modulator_nmf.m:104
  sizes = diff(starts);
  model.D = D; model.N = N; model.bz = sizes(1); model.bg = sizes(end);
  assert(all(sizes(1:D) == model.bz) && all(sizes(D+1:end) == model.bg), 'unexpected block structure');
  model.A = []; model.Q = []; model.Pinf = []; model.h = [];
  for i = 1:D+N
    ii = starts(i):starts(i+1)-1;
    Ab = A(ii,ii); Qb = Q(ii,ii); Pb = Pinf(ii,ii);
    model.A = [model.A; Ab(:)]; model.Q = [model.Q; Qb(:)]; model.Pinf = [model.Pinf; Pb(:)];
    model.h = [model.h; H(i,ii)'];
  end
  offblock = A; for i = 1:D+N, ii = starts(i):starts(i+1)-1; offblock(ii,ii) = 0; end
  assert(~any(offblock(:)), 'A is not block diagonal: not an ss_modulators_nmf model');
end
