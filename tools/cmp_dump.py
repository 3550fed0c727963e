import sys, numpy as np
sys.path.insert(0, "/root/repo/tests")
from helpers import load_golden, golden_names
def rowrel(a, b):
    return np.abs(a - b).max(axis=1) / np.maximum(1, np.abs(b).max(axis=1))
for name in golden_names():
    meta, params, d = load_golden(name)
    c = np.load("/root/repo/gpurun_out/cuda_%s.npz" % name)
    ok = np.isfinite(d["samp_logp"]) & np.isfinite(d["logp"])
    e_base = rowrel(c["base"], d["base"])[ok]
    e_logp = (np.abs(c["logp"] - d["logp"]) / np.maximum(1, np.abs(d["logp"])))[ok]
    e_sx = rowrel(c["samp_x"], d["samp_x"])[ok]
    e_sl = (np.abs(c["samp_logp"] - d["samp_logp"]) / np.maximum(1, np.abs(d["samp_logp"])))[ok]
    print("%-26s base %.1e (n>1e-10: %d) logp %.1e | samp_x %.1e (n>1e-10: %d) samp_logp %.1e  [ref rt %.1e/%.1e]" % (name, e_base.max(), (e_base > 1e-10).sum(), e_logp.max(), e_sx.max(), (e_sx>1e-10).sum(), e_sl.max(), d["ref_roundtrip_base_err"], d["ref_roundtrip_logp_err"]))
