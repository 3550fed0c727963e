"""GPU: the tcgen05 int8-sliced fp64 MLP (jf_mlp_forward_ws) against torch fp64 and against the DMMA kernel; timing."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from jammy_flows_b200 import _cabi

lib = _cabi.load()
dev = torch.device("cuda")
torch.manual_seed(0)


def run(B, kin, N, scale_w2=0.1, dt=torch.float64):
    x = torch.randn(B, kin, dtype=dt, device=dev) * 1.5
    W1 = torch.randn(128, kin, dtype=dt, device=dev) * 0.5
    b1 = torch.randn(128, dtype=dt, device=dev) * 0.3
    W2 = torch.randn(N, 128, dtype=dt, device=dev) * scale_w2
    W2[: N // 3] *= 1e-3
    b2 = torch.randn(N, dtype=dt, device=dev)
    ref = (torch.tanh(x.double() @ W1.double().T + b1.double()) @ W2.double().T + b2.double())
    code = _cabi.JF_F64 if dt == torch.float64 else _cabi.JF_F32
    md = _cabi.JfMlpDesc()
    md.n_linear = 2
    md.dims[0], md.dims[1], md.dims[2] = kin, 128, N
    md.n_segments = 1
    md.seg_cols[0] = kin
    segs = (C.c_void_p * 1)(x.data_ptr())
    lds = (C.c_int64 * 1)(kin)
    ws_ = (C.c_void_p * 2)(W1.data_ptr(), W2.data_ptr())
    bs_ = (C.c_void_p * 2)(b1.data_ptr(), b2.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nws = lib.jf_mlp_workspace_bytes(C.byref(md), code)
    ws = torch.zeros(max(nws, 16), dtype=torch.uint8, device=dev)
    out_i8 = torch.full((N, B), float("nan"), dtype=dt, device=dev)
    out_dm = torch.full((N, B), float("nan"), dtype=dt, device=dev)
    rc = lib.jf_mlp_forward_ws(C.byref(md), code, segs, lds, ws_, bs_, C.c_void_p(out_i8.data_ptr()), B, 1, B,
                               C.c_void_p(ws.data_ptr()), nws, 0, st)
    assert rc == 0, rc
    rc = lib.jf_mlp_forward(C.byref(md), code, segs, lds, ws_, bs_, C.c_void_p(out_dm.data_ptr()), B, 1, B, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    sc = (ref.abs().max()).item()
    e_i8 = (out_i8.T - ref).abs().max().item()
    e_dm = (out_dm.T - ref).abs().max().item()

    def timeit(fn, n=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_i8 = timeit(lambda: lib.jf_mlp_forward_ws(C.byref(md), code, segs, lds, ws_, bs_, C.c_void_p(out_i8.data_ptr()),
                                                 B, 1, B, C.c_void_p(ws.data_ptr()), nws, 1, st))
    t_dm = timeit(lambda: lib.jf_mlp_forward(C.byref(md), code, segs, lds, ws_, bs_, C.c_void_p(out_dm.data_ptr()),
                                              B, 1, B, st))
    print(str(dt)[6:], "B=%d kin=%d N=%d: |i8-ref| %.2e  |dmma-ref| %.2e (scale %.1f)   i8 %.3f ms  dmma %.3f ms   nan(i8)=%d"
          % (B, kin, N, e_i8, e_dm, sc, t_i8, t_dm, int(torch.isnan(out_i8).sum())))


if len(sys.argv) > 1 and sys.argv[1] == "f32":
    run(1000, 7, 548, dt=torch.float32)
    run(401, 70, 50, dt=torch.float32)
    run(1 << 18, 64, 1302, dt=torch.float32)
    run(1 << 17, 64, 3210, dt=torch.float32)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    run(1 << 19, 7, 548)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    run(1 << 18, 7, 548)
    run(1 << 18, 4, 10)
    sys.exit(0)
run(1000, 7, 548)
run(128 * 3 + 17, 4, 10)
run(1 << 19, 7, 548)
run(1 << 19, 4, 10)
run(1 << 17, 16, 1302)
