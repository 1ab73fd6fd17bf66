function mom = nsagp_resolve_mom(mom, N)
% NSAGP_RESOLVE_MOM - accept the reference's own `mom` closure.
%
% The callers of the reference build
%   mom = @(hyp,mu,s2,nmfW,ep_frac,yall,k) feval(likfunc,link,hyp,yall(k),mu,s2,nmfW,p_cubature,ep_frac,'infEP');
%       (demo_toy_modulators_nmf.m:78-81, experiments/synthetic_data_experiment.m:148)
%   mom = @(hyp,mu,s2,nmfW,ep_frac,yall,k) feval(likfunc,link,hyp,yall(k),mu,s2,nmfW,ep_frac,wn,xn_unscaled,'infEP');
%       (experiments/missing_data_music.m:98, noise_reduction_speech.m:75, source_sep_piano.m:94)
% A function handle cannot cross into CUDA, but an anonymous function carries the variables it captured:
% functions(mom).workspace{1} holds likfunc, link and p_cubature (or wn, xn_unscaled).  This function recognises
% exactly those closures and returns the descriptor nsagp_mom builds; anything else is an error (there is no CPU
% fallback that could evaluate an arbitrary handle).  A struct (already a descriptor) passes through unchanged.
  if isstruct(mom), return; end
  if ~isa(mom, 'function_handle')
    error('nsagp:mom', '`mom` must be the reference''s moment closure or an nsagp_mom descriptor');
  end
  info = functions(mom);
  if ~strcmp(info.type, 'anonymous') || isempty(info.workspace)
    error('nsagp:mom', 'cannot resolve `mom`: not an anonymous function with a captured workspace (use nsagp_mom)');
  end
  ws = info.workspace{1};
  if ~isfield(ws, 'likfunc') || ~isfield(ws, 'link')
    error('nsagp:mom', 'cannot resolve `mom`: the closure does not capture `likfunc` and `link` (use nsagp_mom)');
  end
  likname = ws.likfunc;
  if isa(likname, 'function_handle'), likname = func2str(likname); end
  likname = regexprep(likname, '^@', '');
  % link: log(1+exp(g)) or log(1+exp(g-c)) with a literal or captured shift c
  ls = regexprep(func2str(ws.link), '\s+', '');
  tok = regexp(ls, '^@\(g\)log\(1\+exp\(g(?:-([^\)]+))?\)\)$', 'tokens', 'once');
  if isempty(tok) && ~strcmp(ls, '@(g)log(1+exp(g))')
    error('nsagp:mom', 'unsupported link function %s: only log(1+exp(g-c)) runs on the GPU', ls);
  end
  shift = 0;
  if ~isempty(tok) && ~isempty(tok{1})
    shift = str2double(tok{1});
    if isnan(shift)                               % the shift is itself a captured variable of the link closure
      lw = functions(ws.link);
      if ~isempty(lw.workspace) && isfield(lw.workspace{1}, tok{1})
        shift = lw.workspace{1}.(tok{1});
      else
        error('nsagp:mom', 'cannot resolve the link shift `%s`', tok{1});
      end
    end
  end
  switch likname
    case 'likModulatorPower'       % the model without NMF weights (demo_toy_modulators.m:88): NMFPower with W = I
      if ~isfield(ws, 'p_cubature'), error('nsagp:mom', 'the closure does not capture `p_cubature`'); end
      mom = nsagp_mom('likModulatorNMFPower', shift, ws.p_cubature, N);
    case 'likModulatorNMFPower'
      if ~isfield(ws, 'p_cubature'), error('nsagp:mom', 'the closure does not capture `p_cubature`'); end
      mom = nsagp_mom('likModulatorNMFPower', shift, ws.p_cubature, N);
    case 'likModulatorPreCalcwn'
      if ~isfield(ws, 'wn') || ~isfield(ws, 'xn_unscaled')
        error('nsagp:mom', 'the closure does not capture `wn` and `xn_unscaled`');
      end
      mom = nsagp_mom('likModulatorPreCalcwn', shift, ws.wn, ws.xn_unscaled);
    otherwise
      error('nsagp:mom', 'unsupported likelihood %s (likModulatorNMFPower / likModulatorPower / likModulatorPreCalcwn run on the GPU)', likname);
  end
end
