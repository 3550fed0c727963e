"""Write-only bandwidth on this B200 (torch fill_ of the same [rows, T] block) next to `mlp_expand_kernel`
(rank 8 -> T = 6202 outputs per row): is the expand kernel at the write roofline?"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jammy_flows_b200 import engine, _cabi  # noqa: E402


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


R, K = 262144, 8
lib = _cabi.load()
for T in (6202, 6208):
    out = torch.empty(R, T, dtype=torch.float64, device="cuda")
    mid = torch.randn(R, K, dtype=torch.float64, device="cuda")
    w = torch.randn(T, K, dtype=torch.float64, device="cuda")
    b = torch.randn(T, dtype=torch.float64, device="cuda")
    nbytes = R * T * 8
    ms_fill = timed(lambda: out.fill_(1.0))
    ms_exp = timed(lambda: engine._run_chain(lib, torch.float64, out.device, [(w, b)], [mid], out, 1, T, R, False))
    ref = mid[:64] @ w.t() + b
    err = float((out[:64] - ref).abs().max())
    print(json.dumps(dict(T=T, rows=R, fill_ms=round(ms_fill, 3), fill_gbs=round(nbytes / ms_fill / 1e6, 1),
                          expand_ms=round(ms_exp, 3), expand_gbs=round(nbytes / ms_exp / 1e6, 1), err=err)), flush=True)
    del out
