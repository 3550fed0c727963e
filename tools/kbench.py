"""Per-kernel timing of one step (bench.kernel_breakdown) on N rows -- quick optimisation loop helper."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jammy_flows_b200 import _cabi, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
lib = _cabi.load()
pdf = bench.make_model().cuda()
x, z = bench.make_inputs(n, torch.device("cuda"), 100)
with torch.no_grad():
    for _ in range(2):
        engine.pdf_logpdf(pdf, x); engine.pdf_sample(pdf, z)
    pdf.kernel_status()
    engine.pdf_sample(pdf, z)
st = pdf.kernel_status()
ev = st["evaluations"] / (n * 32.0)
rows, total = bench.kernel_breakdown(pdf, x, z, lib, ev)
print("rows %d  evals/elem %.2f  total %.1f ms  (%.2f ns/row)" % (n, ev, total, total * 1e6 / n))
for r in rows:
    print("  %-30s %8.2f ms  %5.1f%%  %6.2f TF-eq" % (r["kernel"], r["ms"], 100 * r["share"], r["achieved_tflops"]))
