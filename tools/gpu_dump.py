"""Dump CUDA-path outputs for the golden cases to gpurun_out/cuda_<name>.npz (analysis helper)."""
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import build_pdf, golden_names, load_golden
os.makedirs("gpurun_out", exist_ok=True)
names = sys.argv[1:] or golden_names()
for name in names:
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    dt = getattr(torch, meta["dtype"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    cond = t(data["cond"]) if "cond" in data else None
    with torch.no_grad():
        logp, lb, base = p(t(data["x"]), conditional_input=cond)
        x, _, slogp, slb = p._obtain_sample(conditional_input=cond, predefined_target_input=t(data["z"]))
    np.savez("gpurun_out/cuda_%s.npz" % name, logp=logp.cpu().numpy(), logp_base=lb.cpu().numpy(), base=base.cpu().numpy(),
             samp_x=x.cpu().numpy(), samp_logp=slogp.cpu().numpy())
