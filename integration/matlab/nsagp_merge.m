function [yall, return_ind] = nsagp_merge(x, y, xt)
% Combine observations and test points, sort, keep the first of duplicates; test-only
% points carry NaN (what gf_ep_modulator_nmf.m:58-66 does).
  xall = [x(:); xt(:)];
  yall = [y(:); nan(numel(xt),1)];
  [~, sort_ind, return_ind] = unique(xall, 'first');
  yall = yall(sort_ind);
  return_ind = return_ind(end-numel(xt)+1:end);
end
