import sys, numpy as np
a = np.load(sys.argv[1]); b = np.load(sys.argv[2])
def rel(x, y):
    m = np.isfinite(x) & np.isfinite(y)
    return float(np.max(np.abs(x[m] - y[m])) / max(np.max(np.abs(y[m])), 1e-300))
for k in a.files:
    if k.endswith("0") or k.endswith("1"):
        print(k, "new-vs-old", rel(a[k], b[k]) if a[k].ndim else (a[k], b[k]))
for nm in ("ihgp", "gfep"):
    for src, lab in ((a, "new"), (b, "old")):
        print(nm, lab, "form0-vs-form1 E", rel(src[nm + "_E0"], src[nm + "_E1"]), "V", rel(src[nm + "_V0"], src[nm + "_V1"]), "nlZ", rel(src[nm + "_nlZ0"], src[nm + "_nlZ1"]), "neg", src[nm + "_neg0"], src[nm + "_neg1"])
