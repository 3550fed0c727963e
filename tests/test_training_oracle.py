"""CPU: (1) the oracle's autograd gradients against the gradients the unmodified reference computed (golden `grad/*`
entries) -- this pins the oracle as the checker of the CUDA backward kernels; (2) the data-parallel gradient
all-reduce helper over a 2-rank gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build_pdf, golden_names, load_golden
from jammy_flows_b200 import sharding
from oracle.jf_oracle import OraclePdf

TRAIN = [n for n in golden_names() if n.startswith("train_")]


def oracle_grads(pdf, params, x, cond):
    """d mean(log_pdf)/d(parameter tensors) by autograd through the oracle (torch CPU ops)."""
    o = OraclePdf(pdf.export_program("float64"), params)
    for t in o.params.values():
        t.requires_grad_(True)
    lp, _, _ = o.log_pdf(x, cond)
    lp.mean().backward()
    return {k: t.grad.detach().numpy() for k, t in o.params.items() if t.grad is not None}


@pytest.mark.parametrize("name", TRAIN)
def test_oracle_autograd_reproduces_reference_gradients(name):
    meta, params, data = load_golden(name)
    pdf = build_pdf(meta)
    g = oracle_grads(pdf, params, data["x"], data.get("cond"))
    n = 0
    for k in data:
        if not k.startswith("grad/"):
            continue
        ref = data[k]
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.abs(g[k[5:]] - ref).max() / scale < 1e-9, k
        n += 1
    if meta["conditional_input_dim"] is not None:
        assert n == 4 * len(meta["pdf_defs"].split("+"))       # weight + bias of both Linear layers of every MLP
    else:
        assert n == len(params)                                 # every permanent tensor and every MLP tensor


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
        g = torch.Generator().manual_seed(11)
        xs = torch.randn(10, 4, generator=g, dtype=torch.float64)
        lo, hi = sharding.shard_range(10, rank, world)
        m(xs[lo:hi]).pow(2).sum().backward()                # sum over this rank's rows
        n = sharding.allreduce_gradients(m, average=False)
        assert n == sum(p.numel() for p in m.parameters())
        ref = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
        ref.load_state_dict(m.state_dict())
        ref(xs).pow(2).sum().backward()                      # the same sum over all rows on one rank
        for a, b in zip(m.parameters(), ref.parameters()):
            assert torch.allclose(a.grad, b.grad, rtol=1e-12, atol=1e-14)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
