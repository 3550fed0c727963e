import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from jammy_flows_b200 import engine
n = 4_000_000
pdf = bench.make_model().cuda()
x, z = bench.make_inputs(n, torch.device("cuda"), 100)
xh = x.cpu().pin_memory()
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
with torch.no_grad():
    print("device logpdf      %.1f ms" % t(lambda: engine.pdf_logpdf(pdf, x)))
    print("host   logpdf      %.1f ms" % t(lambda: engine.pdf_logpdf_host(pdf, xh)))
    for ch in (1 << 16, 1 << 18, 1 << 20):
        print("host   logpdf chunk %d  %.1f ms" % (ch, t(lambda: engine.pdf_logpdf_host(pdf, xh, chunk_rows=ch))))
    print("pinned alloc 4Mx12 doubles %.1f ms" % t(lambda: torch.empty(n, 12, dtype=torch.float64, pin_memory=True)))
    xd = torch.empty_like(x)
    print("H2D copy %.1f ms" % t(lambda: xd.copy_(xh, non_blocking=True)))
    bh = torch.empty(n, 10, dtype=torch.float64, pin_memory=True)
    print("D2H copy %.1f ms" % t(lambda: bh.copy_(x, non_blocking=True)))
