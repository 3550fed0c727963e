"""Throughput of the other BASELINE.json configs (cfg1, cfg3, cfg4) on one B200: log_pdf evaluations/s and samples/s
with device-resident inputs, CUDA-event timed, 3 warm-up + 5 timed calls each.  (cfg2 = bench.py, cfg5 = train_bench.py)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jammy_flows_b200 as jfb

CFG3_F = {"add_vertical_rq_spline_flow": 1, "spline_num_basis_functions": -1, "vertical_smooth": 1,
          "vertical_flow_defs": "rr", "circular_flow_defs": "oo", "vertical_fix_boundary_derivative": 1,
          "add_circular_rq_spline_flow": 1, "circular_add_rotation": 0, "vertical_fix_first_width_n_height_to_zero": 1,
          "vertical_also_fix_second_width_to_zero": 1, "vertical_independent_width_height_parametrization": 1}
CFGS = [
    ("cfg1 e2 'gg' fp64 100k", dict(pdf_defs="e2", flow_defs="gg"), 100_000, torch.float64, None),
    ("cfg1 e2 'gg' fp64 10M", dict(pdf_defs="e2", flow_defs="gg"), 10_000_000, torch.float64, None),
    ("cfg3 s2+i1 'f+r' smooth splines fp64 5M", dict(pdf_defs="s2+i1", flow_defs="f+r", options_overwrite={"f": CFG3_F}), 5_000_000, torch.float64, None),
    ("cfg4 e6+s2 'gggggg+v' cond 64 fp32 4M", dict(pdf_defs="e6+s2", flow_defs="gggggg+v", conditional_input_dim=64), 4_000_000, torch.float32, 64),
]


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = []
for name, kw, n, dt, cdim in CFGS:
    torch.manual_seed(1); np.random.seed(1)
    p = jfb.pdf(**kw).to(dt)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for q in p.parameters():
            q.add_((0.02 if cdim else 0.1) * torch.randn(q.shape, generator=g, dtype=torch.float64).to(q.dtype))
    p = p.cuda()
    gen = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randn(n, p.total_base_dim, dtype=dt, device="cuda", generator=gen)
    c = torch.randn(n, cdim, dtype=dt, device="cuda", generator=gen) if cdim else None
    with torch.no_grad():
        x = p._obtain_sample(conditional_input=c, predefined_target_input=z)[0]      # valid target points
        t_lp = timeit(lambda: p(x, conditional_input=c))
        t_s = timeit(lambda: p._obtain_sample(conditional_input=c, predefined_target_input=z))
    st = p.kernel_status()
    row = dict(config=name, rows=n, dtype=str(dt)[6:], logpdf_ms=round(t_lp, 3), sample_ms=round(t_s, 3),
               logpdf_evals_per_s=n / t_lp * 1e3, samples_per_s=n / t_s * 1e3, status=st)
    out.append(row)
    print(json.dumps(row))
