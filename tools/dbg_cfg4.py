import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_cuda_manifolds import _model
p = _model("e6+s2", "gggggg+v", 0.02, cond=64).cuda()
n = 1_000_000
g = torch.Generator(device="cuda").manual_seed(31)
cond = torch.randn(n, 64, generator=g, dtype=torch.float64, device="cuda")
z = torch.randn(n, 8, generator=g, dtype=torch.float64, device="cuda")
with torch.no_grad():
    x, _, logp, _ = p._obtain_sample(conditional_input=cond, predefined_target_input=z)
    rt_logp, _, rt_z = p(x, conditional_input=cond)
err = (rt_z - z).abs()
rowerr = err.max(dim=1)[0]
idx = torch.argsort(rowerr, descending=True)[:8]
torch.set_printoptions(precision=6, linewidth=200)
for i in idx.tolist():
    print(i, "err per col", err[i].cpu().numpy().round(12), "\n   z", z[i].cpu().numpy().round(4), "\n   x", x[i].cpu().numpy().round(6), " znorm s2 %.3f" % float(z[i, 6:].norm()))
print(p.kernel_status())
print("quantiles", [float(rowerr.quantile(q)) for q in (0.5, 0.99, 0.999, 0.9999, 0.99999)])
