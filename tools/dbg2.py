import sys, numpy as np, torch, copy
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from test_cuda_properties import _perturbed, _inputs
p = _perturbed("e4+s2+e4", "gggg+n+gggg", scale=0.1).cuda()
n = 2_000_000
z = torch.randn(n, 10, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(11))
with torch.no_grad():
    x, _, logp, logp_base = p._obtain_sample(predefined_target_input=z)
    rt_logp, rt_logp_base, rt_z = p(x)
err = (rt_z - z).abs()
rows = torch.nonzero(err.max(dim=1)[0] > 1e-7)[:, 0]
print("bad rows", rows.numel(), p.kernel_status())
for r in rows[:8].tolist():
    print(r, "z", z[r].cpu().numpy().round(4), "\n   rt", rt_z[r].cpu().numpy().round(4), "\n   x", x[r].cpu().numpy().round(4), "logp", float(logp[r]), float(rt_logp[r]))
# fp32 check
p64 = _perturbed("e6", "gggggg", scale=0.05, cond=8)
xx, zz, c = _inputs(p64, 4000)
p32 = copy.deepcopy(p64).float().cuda(); p64 = p64.cuda()
with torch.no_grad():
    lp64, _, b64 = p64(xx.cuda(), conditional_input=c.cuda())
    lp32, _, b32 = p32(xx.float().cuda(), conditional_input=c.float().cuda())
bad = torch.nonzero(~torch.isfinite(lp32))[:, 0]
print("fp32 nonfinite rows:", bad.numel(), "fp64 nonfinite:", int((~torch.isfinite(lp64)).sum()), p32.kernel_status())
for r in bad[:4].tolist():
    print(r, "x", xx[r].numpy().round(3), "b64", b64[r].cpu().numpy().round(3), "b32", b32[r].cpu().numpy(), float(lp64[r]), float(lp32[r]))
err = ((lp32.double() - lp64).abs() / lp64.abs().clamp(min=1))
print("fp32 vs fp64 logp rel err quantiles", np.nanquantile(err.cpu().numpy(), [0.5, 0.9, 0.99, 1.0]))
