function [lik_param, param1, param2, Wnmf] = nsagp_unpack_constraints(w, w_fixed, tune_hypers, constraints, nlik, D, N)
% Split tuned / fixed parameters by tune_hypers(1:7) and squash them into their boxes with
% the reference's sigmoid (gf_ep_modulator_nmf_constraints.m:75-110; sigmoid.m:17-19).
  src = {w_fixed(:), w(:)}; pos = [0 0];
  function v = take(flag, count)
    k = flag + 1; v = src{k}(pos(k)+1:pos(k)+count); pos(k) = pos(k) + count;
  end
  counts = [nlik D D D N N D*N];
  parts = cell(1,7);
  for i = 1:7, parts{i} = take(logical(tune_hypers(i)), counts(i)); end
  lik_param = parts{1};
  for i = 2:7, parts{i} = sigmoid(parts{i}, constraints(i-1,:)); end
  param1 = [parts{2}; parts{3}; parts{4}];
  param2 = [parts{5}; parts{6}];
  Wnmf = reshape(parts{7}, [D,N]);
end
