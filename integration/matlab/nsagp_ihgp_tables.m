function tables = nsagp_ihgp_tables(A, Q, H, want_smoother)
% NSAGP_IHGP_TABLES - steady-state tables of the infinite-horizon path, built with the
% same MATLAB built-ins the reference uses (dare, apxGrid), once per call on the host:
% forward tables as ihgp_ep_modulator_nmf.m:106-134, smoother tables as :150-191.
% Output layout = nsagp_tables: per block nr rows (PP(:)' resp. [PS2(:)' G(:)']), blocks
% concatenated, each block stored row after row.
  starts = [find(sum(abs(H),1) > 0), size(H,2)+1];
  M = numel(starts) - 1;
  r = logspace(-2,4,200)';
  tables.r = r'; tables.PP = []; tables.PG = [];
  for n = 1:M
    ii = starts(n):starts(n+1)-1; b = numel(ii);
    ro = logspace(-2,4,32)';
    PPo = nan(numel(ro), b*b);
    for j = 1:numel(ro)
      try
        PP = dare(A(ii,ii)', H(n,ii)', Q(ii,ii), ro(j));
        PPo(j,:) = PP(:)';
      catch
        ro(j) = nan;                              % the reference drops the node (:118-127)
      end
    end
    PPo(isnan(ro),:) = []; ro(isnan(ro)) = [];
    U = apxGrid('interp', {ro}, r, 3);            % piecewise linear on this grid (SURVEY.md F8)
    PPn = U*PPo;
    tables.PP = [tables.PP; reshape(PPn', [], 1)];
    if want_smoother
      PGo = nan(numel(ro), 2*b*b);
      for j = 1:numel(ro)
        PP = reshape(PPo(j,:), b, b);
        h = H(n,ii); S = h*PP*h' + ro(j); K = PP*h'/S; P = PP - K*ro(j)*K';
        L = chol(A(ii,ii)*P*A(ii,ii)' + Q(ii,ii), 'lower');
        G = P*A(ii,ii)'/L'/L;
        QQ = P - G*PP*G'; QQ = (QQ+QQ')/2;
        [V,DD] = eig(QQ); keep = diag(DD) > 0;
        QQ = V(:,keep)*DD(keep,keep)*V(:,keep)';
        PS2 = dare(G', 0*G, QQ);
        PGo(j,:) = [PS2(:)' G(:)'];
      end
      PGn = U*PGo;
      tables.PG = [tables.PG; reshape(PGn', [], 1)];
    end
  end
end
