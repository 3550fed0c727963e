import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import numpy as np, torch
from test_cuda_properties import _inputs, _perturbed
from oracle.jf_oracle import OraclePdf
from helpers import rel_err, row_rel_err
cases = [({"rotation_mode": "triangular_combination"}, 3), ({"add_skewness": 1}, 3),
         ({"center_mean": 1, "inverse_function_type": "inormal_partly_precise"}, None),
         ({"center_mean": 1, "inverse_function_type": "inormal_partly_precise"}, 3)]
for opts, cd in cases:
    p = _perturbed("e3+e2", "gg+ggt", cond=cd, scale=0.2, options_overwrite={"g": opts})
    x, z, c = _inputs(p, 2000)
    o = OraclePdf(p.export_program(), {k: v.numpy() for k, v in p.state_dict().items()})
    lp_o, _, b_o = o.log_pdf(x, c)
    xs_o, slp_o, _ = o.sample(z, c)
    pc = p.cuda()
    cc = c.cuda() if c is not None else None
    with torch.no_grad():
        lp, _, b = pc(x.cuda(), conditional_input=cc)
        st1 = pc.kernel_status()
        xs, _, slp, _ = pc._obtain_sample(conditional_input=cc, predefined_target_input=z.cuda())
        st2 = pc.kernel_status()
    xs = xs.cpu()
    e_lp = rel_err(lp.cpu().numpy(), lp_o.numpy()); e_b = row_rel_err(b.cpu().numpy(), b_o.numpy())
    e_x = row_rel_err(xs.numpy(), xs_o.numpy()); e_slp = rel_err(slp.cpu().numpy(), slp_o.numpy())
    zmax = np.abs(z.numpy()).max(axis=1); bmax = np.abs(b_o.numpy()).max(axis=1)
    print("\n==", opts, cd, "status logpdf", st1, "sample", st2)
    for name, e, key in (("logp", e_lp, bmax), ("base", e_b, bmax), ("x", e_x, zmax), ("slogp", e_slp, zmax)):
        idx = np.argsort(-e)[:4]
        print("  %-6s worst:" % name, ["row %d err %.1e |z|max %.2f" % (i, e[i], key[i]) for i in idx])
    bad = np.nonzero(e_x > 1e-7)[0][:3]
    if len(bad):
        _, _, bc = o.log_pdf(xs[bad], None if c is None else c[bad])
        _, _, bo = o.log_pdf(xs_o[bad], None if c is None else c[bad])
        for i, r in enumerate(bad):
            print("   row", r, "z", z[r].numpy().round(3), "x_cuda", xs[r].numpy().round(3), "x_orac", xs_o[r].numpy().round(3),
                  "back-err cuda %.1e orac %.1e" % ((bc[i] - z[r]).abs().max(), (bo[i] - z[r]).abs().max()))
