"""BASELINE.json configs[4]: training step on the conditional e10 "gggggggg" flow (64 conditional inputs, 422 410
parameters), 1 M rows per GPU, fp32: forward + backward through the fused layer kernels, gradient all-reduce (NCCL over
NVLink when launched under torchrun), Adam.  Rows are processed in chunks (gradient accumulation) because the per-row
parameter buffer is 12.8 KB/row.

    python tools/train_bench.py [--rows 1000000] [--chunk 65536] [--steps 3] [--dtype float32]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG", "WARN")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import jammy_flows_b200 as jfb  # noqa: E402
from jammy_flows_b200 import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--dtype", default="float32")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    dt = getattr(torch, args.dtype)
    torch.manual_seed(1)
    np.random.seed(1)
    pdf = jfb.pdf("e10", "gggggggg", conditional_input_dim=64).to(dt).to(dev)
    n = args.rows
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    cond = torch.randn(n, 64, generator=g, dtype=dt, device=dev)
    y = (0.5 * cond[:, :10] + 0.8 * torch.randn(n, 10, generator=g, dtype=dt, device=dev))
    opt = torch.optim.Adam(pdf.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        tot = torch.zeros((), dtype=torch.float64, device=dev)
        for r0 in range(0, n, args.chunk):
            lp, _, _ = pdf(y[r0:r0 + args.chunk], conditional_input=cond[r0:r0 + args.chunk])
            loss = -lp.sum() / n
            loss.backward()                       # gradients accumulate over the chunks
            tot += loss.detach().double()
        if dist is not None:
            sharding.allreduce_gradients(pdf)     # one flat NCCL all-reduce (1.7 MB)
        opt.step()
        return tot

    losses = []
    for _ in range(args.warmup):
        losses.append(float(step()))
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses.append(float(step()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = pdf.kernel_status()
    if rank == 0:
        print(json.dumps(dict(metric="training rows/s, conditional e10 'gggggggg' (cond dim 64), fwd+bwd+allreduce+Adam",
                              value=world * n / (ms * 1e-3), unit="rows/s", n_gpus=world, ms_per_step=ms, dtype=args.dtype,
                              rows_per_gpu=n, chunk_rows=args.chunk, params=pdf.count_parameters(), losses=losses,
                              kernel_status=st)))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
