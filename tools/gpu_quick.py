import sys, numpy as np, torch, time
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import build_pdf, golden_names, load_golden, rel_err
for name in golden_names():
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    dt = getattr(torch, meta["dtype"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    cond = t(data["cond"]) if "cond" in data else None
    with torch.no_grad():
        logp, lb, base = p(t(data["x"]), conditional_input=cond)
        x, _, slogp, _ = p._obtain_sample(conditional_input=cond, predefined_target_input=t(data["z"]))
        rt_logp, _, rt_base = p(x, conditional_input=cond)
    st = p.kernel_status()
    print("%-26s logp %.1e base %.1e | samp_x %.1e samp_logp %.1e | rt_base %.1e (ref %.1e) evals/elem %.2f st=%s" % (
        name, np.nanmax(rel_err(logp.cpu().numpy(), data["logp"])), np.nanmax(rel_err(base.cpu().numpy(), data["base"])),
        np.nanmax(rel_err(x.cpu().numpy(), data["samp_x"])), np.nanmax(rel_err(slogp.cpu().numpy(), data["samp_logp"])),
        np.nanmax(rel_err(rt_base.cpu().numpy(), data["z"])), data["ref_roundtrip_base_err"],
        st["evaluations"] / max(1, data["z"].size), {k: v for k, v in st.items() if k != "evaluations"}))
