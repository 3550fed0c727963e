"""cfg4 in fp32 at full size (4M rows): where do the non-finite / unconverged status counts come from?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_cuda_manifolds import _model
n = 4_000_000
p64 = _model("e6+s2", "gggggg+v", 0.02, cond=64)
g = torch.Generator(device="cuda").manual_seed(31)
cond64 = torch.randn(n, 64, generator=g, dtype=torch.float64, device="cuda")
z64 = torch.randn(n, 8, generator=g, dtype=torch.float64, device="cuda")
import copy
p32 = copy.deepcopy(p64).float().cuda()
p64 = p64.cuda()
cond, z = cond64.float(), z64.float()
for q in (p32, p64):
    q.chunk_rows = 1 << 18
with torch.no_grad():
    x, _, logp, _ = p32._obtain_sample(conditional_input=cond, predefined_target_input=z)
    st_s = p32.kernel_status()
    rt_logp, _, rt_z = p32(x, conditional_input=cond)
    st_l = p32.kernel_status()
    x64, _, logp64, _ = p64._obtain_sample(conditional_input=cond.double(), predefined_target_input=z.double())
    st64 = p64.kernel_status()
print("status sample", st_s, "logpdf", st_l, "fp64 sample", st64)
bad_x = ~torch.isfinite(x).all(dim=1)
bad_lp = ~torch.isfinite(logp)
bad_rt = ~torch.isfinite(rt_z).all(dim=1) | ~torch.isfinite(rt_logp)
r_s2 = z[:, 6:].norm(dim=1)
calm = (z[:, :6].abs().max(dim=1)[0] < 5.2) & (r_s2 < 5.2) & (r_s2 > 2e-3)
print("rows: nonfinite x %d, logp %d, round trip %d; not calm %d; nonfinite & calm: x %d logp %d rt %d" %
      (int(bad_x.sum()), int(bad_lp.sum()), int(bad_rt.sum()), int((~calm).sum()), int((bad_x & calm).sum()),
       int((bad_lp & calm).sum()), int((bad_rt & calm).sum())))
idx = torch.nonzero(bad_x | bad_lp | bad_rt)[:10, 0]
for i in idx.tolist():
    print(i, "z", z[i].tolist(), "x", x[i].tolist(), "logp", float(logp[i]), "x64", x64[i].tolist())
ok = calm & ~bad_x & ~bad_rt
err = ((rt_z - z).abs().max(dim=1)[0] / z.abs().max(dim=1)[0].clamp(min=1))[ok]
print("round trip (calm rows): median %.2e p999 %.2e max %.2e" % (float(err.median()), float(err.quantile(0.999)), float(err.max())))
e64 = ((x.double() - x64).abs().max(dim=1)[0] / x64.abs().max(dim=1)[0].clamp(min=1))[ok]
print("fp32 vs fp64 samples (calm rows): median %.2e p999 %.2e max %.2e" % (float(e64.median()), float(e64.quantile(0.999)), float(e64.max())))
el = ((logp.double() - logp64).abs() / logp64.abs().clamp(min=1))[ok]
print("fp32 vs fp64 logp: median %.2e p999 %.2e max %.2e" % (float(el.median()), float(el.quantile(0.999)), float(el.max())))
# which sub-pdf do unconverged elements belong to? run the e6 part alone is not possible here; report per-coordinate error
print("per-coordinate max |x32-x64| on calm rows:", (x.double() - x64).abs()[ok].max(dim=0)[0].tolist())
