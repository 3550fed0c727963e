"""fully_amortized_pdf throughput on one B200 (device-resident inputs, CUDA events) and the achieved HBM rate of the
per-row Linear kernel `jf_rowwise_linear` (HBM-bound: every per-row weight is read once).
    python tools/fa_bench.py [rows]  -> one JSON line per measurement on stdout"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jammy_flows_b200 as jfb  # noqa: E402
from jammy_flows_b200 import _cabi  # noqa: E402


def timed(fn, reps=5, warm=2):
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


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    lib = _cabi.load()
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = float(peak.get("hbm_gbs", peak.get("hbm_GBps", 6458.1))) if isinstance(peak, dict) else 6458.1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    micro = "fa_only" not in sys.argv
    for dtype, code, es in ((torch.float64, _cabi.JF_F64, 8), (torch.float32, _cabi.JF_F32, 4)) if micro else ():
        for n_in, n_out in ((7, 128), (128, 64), (3, 128), (128, 3), (16, 130)):
            R = rows
            ld = n_in * n_out + n_out
            params = torch.randn(R, ld, dtype=dtype, device="cuda")
            inp = torch.randn(R, n_in, dtype=dtype, device="cuda")
            out = torch.empty(R, n_out, dtype=dtype, device="cuda")
            f = lambda: lib.jf_rowwise_linear(code, C.c_void_p(params.data_ptr()), ld, 0, n_in * n_out,
                                              C.c_void_p(inp.data_ptr()), n_in, n_in, n_out, 1, 0,
                                              C.c_void_p(out.data_ptr()), 1, n_out, R, st)
            ms = timed(f)
            nbytes = R * (ld + n_in + n_out) * es
            print(json.dumps(dict(kernel="rowwise_linear", dtype=str(dtype), n_in=n_in, n_out=n_out, rows=R, ms=round(ms, 4),
                                  alg_bytes=nbytes, achieved_gbs=round(nbytes / ms / 1e6, 1), hbm_peak_gbs=hbm,
                                  frac=round(nbytes / ms / 1e6 / hbm, 3))), flush=True)
            del params, inp, out
    torch.manual_seed(1)
    np.random.seed(1)
    fa = jfb.fully_amortized_pdf("e4+s2+e4", "gggg+f+gggg", conditional_input_dim=16, inner_mlp_dims_sub_pdfs="32",
                                 inner_mlp_ranks=4, inner_mlp_highway_mode=1, amortization_mlp_dims="128",
                                 amortization_mlp_ranks=8).cuda()
    t = fa.pdf_to_amortize.total_number_amortizable_params
    n = rows
    gen = torch.Generator(device="cuda").manual_seed(3)
    cond = torch.randn(n, 16, dtype=torch.float64, device="cuda", generator=gen)
    x = 1.5 * torch.randn(n, 10, dtype=torch.float64, device="cuda", generator=gen)
    x[:, 4] = torch.acos(1 - 2 * torch.rand(n, dtype=torch.float64, device="cuda", generator=gen))
    x[:, 5] = 2 * np.pi * torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    l0 = lib.jf_launch_count()
    ms_f = timed(lambda: fa(x, conditional_input=cond), reps=3, warm=1)
    launches = (lib.jf_launch_count() - l0) // 4
    ms_s_host_rng = timed(lambda: fa.sample(conditional_input=cond, seed=1), reps=3, warm=1)
    fa.rng_mode = "philox"          # counter-based normals drawn on the device (jf_normal_rows) instead of host numpy + H2D
    ms_s = timed(lambda: fa.sample(conditional_input=cond, seed=1), reps=3, warm=1)
    print(json.dumps(dict(what="fully_amortized_pdf e4+s2+e4 'gggg+f+gggg', cond 16, inner MLPs 32 rank 4 mode 1, outer 128 rank 8",
                          rows=n, amortizable_params_per_row=t, outer_params=fa.total_param_num,
                          logpdf_ms=round(ms_f, 3), logpdf_evals_per_s=round(n / ms_f * 1e3),
                          sample_ms=round(ms_s, 3), samples_per_s=round(n / ms_s * 1e3),
                          sample_ms_numpy_rng=round(ms_s_host_rng, 3), launches_per_forward=launches,
                          status=fa.kernel_status())), flush=True)


if __name__ == "__main__":
    main()
