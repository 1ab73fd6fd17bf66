function c = nsagp_outputs(out, return_ind, nw, nlz, nout)
% varargout packing of the reference (gf_ep_modulator_nmf.m:332-348, :525-531).
  if nlz
    c = {out.edata, zeros(1, nw)};                % the reference's gradient is identically zero
    return
  end
  Eft = out.Eft(:,return_ind); Varft = out.Varft(:,return_ind);
  c = {Eft, Varft, [], out.lb(:,return_ind), out.ub(:,return_ind), out};
  c = c(1:max(nout,1));
end
