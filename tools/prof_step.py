"""One log_pdf + one sampling pass of the README flow on N rows (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_model, make_inputs
from jammy_flows_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pdf = make_model().cuda()
x, z = make_inputs(n, torch.device("cuda"), 100)
with torch.no_grad():
    for _ in range(2):
        engine.pdf_logpdf(pdf, x)
        engine.pdf_sample(pdf, z)
torch.cuda.synchronize()
print(pdf.kernel_status())
