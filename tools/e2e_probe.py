"""End-to-end (host buffers in / out) timing of the README flow against the device-resident path, per chunk size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from jammy_flows_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pdf = bench.make_model().cuda()
x, z = bench.make_inputs(n, torch.device("cuda"), 100)
xh, zh = x.cpu().pin_memory(), z.cpu().pin_memory()
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
with torch.no_grad():
    print("device logpdf %.1f ms   sample %.1f ms" % (t(lambda: engine.pdf_logpdf(pdf, x)), t(lambda: engine.pdf_sample(pdf, z))))
    for ch in (1 << 17, 1 << 18, 1 << 19, 1 << 20):
        print("host chunk %8d: logpdf %.1f ms   sample %.1f ms" % (
            ch, t(lambda: engine.pdf_logpdf_host(pdf, xh, chunk_rows=ch)), t(lambda: engine.pdf_sample_host(pdf, zh, chunk_rows=ch))))
