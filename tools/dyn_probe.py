import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch._dynamo
import jammy_flows_b200 as jfb
p = jfb.pdf("e4+s2+e4", "gggg+n+gggg").double().cuda()
x = torch.randn(256, 10, dtype=torch.float64, device="cuda"); x[:,4]=1.0; x[:,5]=2.0
with torch.no_grad():
    ex = torch._dynamo.explain(p)(x)
    print("breaks", ex.graph_break_count)
    for r in ex.break_reasons: print("REASON:", r.reason[:300]); print("   at", [str(f)[:150] for f in r.user_stack][-2:])
    try:
        torch.compile(p, backend="eager", fullgraph=True)(x)
    except Exception as e:
        print("FULLGRAPH ERR:", str(e)[:1500])
