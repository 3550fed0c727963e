import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import build_pdf, load_golden
for name in ["train_e2e2_cond", "train_e3_ggg_cond", "train_e10_gg_cond"]:
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    p.zero_grad()
    lp, _, _ = p(t(data["x"]), conditional_input=t(data["cond"]))
    lp.mean().backward()
    print(name, "logp err", np.abs(lp.detach().cpu().numpy() - data["logp"]).max())
    for k, q in p.named_parameters():
        ref = data["grad/" + k]; g = q.grad.cpu().numpy()
        print("   %-28s rel err %.3e   |ref| %.3e" % (k, np.abs(g - ref).max() / np.abs(ref).max(), np.abs(ref).max()))
    # per-output-column error of the last bias (= sum over rows of dP): which raw parameter index is wrong?
    k = [n for n, _ in p.named_parameters() if n.endswith(".2.bias")][0]
    ref = data["grad/" + k]; g = dict(p.named_parameters())[k].grad.cpu().numpy()
    bad = np.nonzero(np.abs(g - ref) > 1e-8 * np.abs(ref).max())[0]
    print("   bad bias columns:", bad[:40], "of", len(ref))
    print(p.kernel_status())
