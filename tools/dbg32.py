import sys, numpy as np, torch, copy
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import build_pdf, load_golden, rel_err
meta, params, data = load_golden("g_e5_k7_cond_f32")
p32 = build_pdf(meta, params).cuda()
p64 = copy.deepcopy(p32).double()
t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
with torch.no_grad():
    lp32, lb32, b32 = p32(t32(data["x"]), conditional_input=t32(data["cond"]))
    lp64, lb64, b64 = p64(t32(data["x"]).double(), conditional_input=t32(data["cond"]).double())
lp32, lp64 = lp32.cpu().numpy().astype(np.float64), lp64.cpu().numpy()
ref = data["logp"].astype(np.float64)
print("cuda32 vs cuda64: median %.2e max %.2e" % (np.median(np.abs(lp32-lp64)), np.abs(lp32-lp64).max()))
print("ref32  vs cuda64: median %.2e max %.2e" % (np.median(np.abs(ref-lp64)), np.abs(ref-lp64).max()))
print("cuda32 vs ref32 : median %.2e max %.2e" % (np.median(np.abs(lp32-ref)), np.abs(lp32-ref).max()))
bd32 = np.abs(b32.cpu().numpy().astype(np.float64) - b64.cpu().numpy()).max(axis=1)
bdr = np.abs(data["base"].astype(np.float64) - b64.cpu().numpy()).max(axis=1)
print("base cuda32 vs cuda64: median %.2e max %.2e ; ref32 vs cuda64: median %.2e max %.2e" % (np.median(bd32), bd32.max(), np.median(bdr), bdr.max()))
i = np.argsort(-np.abs(lp32-lp64))[:5]
print(i, lp32[i], lp64[i], ref[i])
print(np.abs(b64.cpu().numpy()[i]).max(axis=1))
