"""One log_pdf + one sampling pass of BASELINE cfg4 (e6+s2 'gggggg+v', 64 conditional inputs, fp32) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jammy_flows_b200 as jfb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
torch.manual_seed(1); np.random.seed(1)
p = jfb.pdf("e6+s2", "gggggg+v", conditional_input_dim=64).float()
g = torch.Generator().manual_seed(2)
with torch.no_grad():
    for q in p.parameters():
        q.add_(0.02 * torch.randn(q.shape, generator=g, dtype=torch.float64).float())
p = p.cuda()
z = torch.randn(n, 8, device="cuda"); c = torch.randn(n, 64, device="cuda")
with torch.no_grad():
    for _ in range(2):
        x = p._obtain_sample(conditional_input=c, predefined_target_input=z)[0]
        p(x, conditional_input=c)
torch.cuda.synchronize()
print(p.kernel_status())
