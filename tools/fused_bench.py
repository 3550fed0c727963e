"""Fused generator + "g"-chain kernel (jf_subpdf_apply_generated) of the README flow's last sub-pdf, timed alone on N rows
(CUDA events on the launching stream) -- optimisation loop helper and ncu target.

    python tools/fused_bench.py [rows] [reps] [logpdf|sample|both]
"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jammy_flows_b200 import _cabi, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "both"
lib = _cabi.load()
pdf = bench.make_model().cuda()
dev = torch.device("cuda")
x, z = bench.make_inputs(n, dev, 100)
desc = pdf._desc(torch.float64)
pack = engine.ParamPack(pdf, torch.float64, dev)
st = engine._stream_ptr(dev)
k = 2
emb = torch.randn(n, 3, dtype=torch.float64, device=dev)
emb /= emb.norm(dim=1, keepdim=True)
out = torch.empty(n, 10, dtype=torch.float64, device=dev)
ld = torch.zeros(n, dtype=torch.float64, device=dev)
lb = torch.zeros(n, dtype=torch.float64, device=dev)
status = torch.zeros(4, dtype=torch.int64, device=dev)
vp = lambda t, off=0: C.c_void_p(t.data_ptr() + off * 8)
md = _cabi.JfMlpDesc()
C.memmove(C.byref(md), C.byref(desc.mlp[k]), C.sizeof(md))
segs = [(vp(x), 10, 4), (vp(emb), 3, 3)]
md.n_segments = len(segs)
ptrs = (C.c_void_p * 2)(*[s[0] for s in segs])
lds = (C.c_int64 * 2)(*[s[1] for s in segs])
for i, s in enumerate(segs):
    md.seg_cols[i] = s[2]
nws = lib.jf_subpdf_generated_workspace_bytes(C.byref(desc.sub[k]), C.byref(md), _cabi.JF_F64)
assert nws > 0, nws
ws = torch.zeros(int(nws), dtype=torch.uint8, device=dev)


def run(direction, src, prepared):
    return lib.jf_subpdf_apply_generated(C.byref(desc.sub[k]), C.byref(md), _cabi.JF_F64, direction, ptrs, lds,
                                         pack.c.weights[k], pack.c.biases[k], vp(src, 6), 10, vp(ld), vp(ld), vp(lb), vp(lb),
                                         vp(out, 6), 10, n, C.c_void_p(ws.data_ptr()), nws, prepared, vp(status), st)


for direction, src, name in ((0, x, "logpdf"), (1, z, "sample")):
    if which not in ("both", name):
        continue
    assert run(direction, src, 0) == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert run(direction, src, 1) == 0
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("fused %s: %d rows  %.3f ms (best %.3f)  %.2f ns/row  -> %.1f ms per 10M rows" %
          (name, n, ts[len(ts) // 2], ts[0], ts[len(ts) // 2] * 1e6 / n, ts[len(ts) // 2] * 1e7 / n))
print("status", status.tolist())
