"""GPU: the backward kernels (jf_subpdf_backward through the autograd path of `pdf.forward`) against the reference's
own gradients (golden `grad/*`), against the pinned oracle's autograd on the BASELINE configs[4] structure
(e10 "gggggggg", 64 conditional inputs), fp32 against fp64, and a short Adam run."""
import numpy as np
import pytest
import torch

import jammy_flows_b200 as jfb
from helpers import build_pdf, golden_names, load_golden
from test_training_oracle import oracle_grads

pytestmark = pytest.mark.gpu
TRAIN = [n for n in golden_names() if n.startswith("train_")]


def cuda_grads(p, x, cond):
    p.zero_grad()
    lp, _, _ = p(x, conditional_input=cond)
    lp.mean().backward()
    return {k: q.grad.detach().cpu().numpy() for k, q in p.named_parameters()}, lp.detach()


@pytest.mark.parametrize("name", TRAIN)
def test_gradients_match_reference_golden(name, lib_built):
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g, lp = cuda_grads(p, t(data["x"]), t(data["cond"]) if "cond" in data else None)
    assert np.abs(lp.cpu().numpy() - data["logp"]).max() < 1e-9
    for k in data:
        if k.startswith("grad/"):
            ref = data[k]
            err = np.abs(g[k[5:]] - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert err < 1e-8, (k, err)
    assert p.kernel_status()["out_of_range"] == 0


def _cfg5(n, seed=3, scale=0.02):
    torch.manual_seed(seed)
    np.random.seed(seed)
    p = jfb.pdf("e10", "gggggggg", conditional_input_dim=64).double()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for q in p.parameters():
            q.add_(scale * torch.randn(q.shape, generator=g, dtype=torch.float64))
    y = 1.5 * torch.randn(n, 10, generator=g, dtype=torch.float64)
    c = torch.randn(n, 64, generator=g, dtype=torch.float64)
    return p, y, c


def test_cfg5_structure_gradients_match_oracle_autograd(lib_built):
    p, y, c = _cfg5(48)
    ref = oracle_grads(p, {k: v.numpy() for k, v in p.state_dict().items()}, y.numpy(), c.numpy())
    g, _ = cuda_grads(p.cuda(), y.cuda(), c.cuda())
    for k, r in ref.items():
        err = np.abs(g[k] - r).max() / max(np.abs(r).max(), 1e-30)
        assert err < 1e-8, (k, err)


def test_fp32_gradients_close_to_fp64(lib_built):
    p, y, c = _cfg5(4096)
    g64, _ = cuda_grads(p.cuda(), y.cuda(), c.cuda())
    p32 = p.float()
    g32, _ = cuda_grads(p32.cuda(), y.float().cuda(), c.float().cuda())
    for k in g64:
        err = np.abs(g32[k] - g64[k]).max() / max(np.abs(g64[k]).max(), 1e-30)
        assert err < 2e-3, (k, err)


def test_weighted_loss_gradients_match_oracle_autograd(lib_built):
    """a per-row upstream gradient (loss = sum_r w_r log p_r): the forward pass keeps the per-row Jacobian and the backward
    scales it by w_r (elementwise in fp64, as `row_scale` of the tensor-core generator backward in fp32)"""
    from oracle.jf_oracle import OraclePdf
    p, y, c = _cfg5(64)
    w = torch.linspace(-0.5, 2.0, 64, dtype=torch.float64)
    o = OraclePdf(p.export_program("float64"), {k: v.numpy() for k, v in p.state_dict().items()})
    for t in o.params.values():
        t.requires_grad_(True)
    lp, _, _ = o.log_pdf(y.numpy(), c.numpy())
    (lp * w).sum().backward()
    ref = {k: t.grad.detach().numpy() for k, t in o.params.items() if t.grad is not None}
    pc = p.cuda()
    pc.zero_grad()
    lpc, _, _ = pc(y.cuda(), conditional_input=c.cuda())
    (lpc * w.cuda()).sum().backward()
    for k, r in ref.items():
        g = dict(pc.named_parameters())[k].grad.cpu().numpy()
        err = np.abs(g - r).max() / max(np.abs(r).max(), 1e-30)
        assert err < 1e-8, (k, err)
    # fp32: the same loss through jf_mlp_backward's row_scale (tf32 products: 2e-3 of the tensor maximum)
    p32 = p.float().cuda()
    p32.zero_grad()
    lp32, _, _ = p32(y.float().cuda(), conditional_input=c.float().cuda())
    (lp32 * w.float().cuda()).sum().backward()
    for k, r in ref.items():
        g = dict(p32.named_parameters())[k].grad.double().cpu().numpy()
        err = np.abs(g - r).max() / max(np.abs(r).max(), 1e-30)
        assert err < 2e-3, (k, err)


def test_gradients_wrt_x_and_conditional_input_by_finite_differences(lib_built):
    """the reference differentiates log_pdf through the evaluation points and the conditional input as well
    (autograd); here d log_pdf / d x comes out of the forward + backward kernel (grad_x) and flows on through the
    generators of the later sub-pdfs, d / d cond through jf_mlp_backward / the fp64 GEMMs"""
    torch.manual_seed(2)
    np.random.seed(2)
    p = jfb.pdf("e3+e2", "ggg+gg", conditional_input_dim=5).double()
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for q in p.parameters():
            q.add_(0.05 * torch.randn(q.shape, generator=g, dtype=torch.float64))
    p = p.cuda()
    n = 16
    x = torch.randn(n, 5, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    c = torch.randn(n, 5, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    w = torch.linspace(0.5, 1.5, n, dtype=torch.float64).cuda()
    lp, _, _ = p(x, conditional_input=c)
    (lp * w).sum().backward()
    h = 1e-6
    with torch.no_grad():
        for t, gt in ((x, x.grad), (c, c.grad)):
            for col in range(t.shape[1]):
                e = torch.zeros_like(t)
                e[:, col] = h
                args = lambda s: (x + s * e, c) if t is x else (x, c + s * e)
                xa, ca = args(1.0)
                xb, cb = args(-1.0)
                fd = (p(xa, conditional_input=ca)[0] - p(xb, conditional_input=cb)[0]) / (2 * h) * w
                assert (fd - gt[:, col]).abs().max() < 1e-6 * max(1.0, float(gt.abs().max())), (t is x, col)


@pytest.mark.parametrize("n", [1, 3, 31, 33, 130])
def test_fp32_tiny_and_ragged_batches(lib_built, n):
    """row counts below / across the 32- and 64-row tiles of the tensor-core generator backward (TMA boxes reach past the
    tensor: zero fill) and of the chain kernel's 32-row warps"""
    p, y, c = _cfg5(max(n, 4))
    y, c = y[:n], c[:n]
    g64, _ = cuda_grads(p.cuda(), y.cuda(), c.cuda())
    g32, _ = cuda_grads(p.float().cuda(), y.float().cuda(), c.float().cuda())
    for k in g64:
        err = np.abs(g32[k] - g64[k]).max() / max(np.abs(g64[k]).max(), 1e-30)
        assert err < 3e-3, (n, k, err)


def test_adam_steps_decrease_the_loss(lib_built):
    """a few optimiser steps on synthetic conditional data: the negative log-likelihood goes down, nothing goes non-finite"""
    p, _, c = _cfg5(8192, scale=0.0)
    p = p.float().cuda()
    c = c.float().cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    y = (0.5 * c[:, :10] + 0.8 * torch.randn(8192, 10, generator=g, device="cuda")).float()
    opt = torch.optim.Adam(p.parameters(), lr=2e-3)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        lp, _, _ = p(y, conditional_input=c)
        loss = -lp.mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 0.05, losses
    assert p.kernel_status()["nonfinite"] == 0


STRAIN = [n for n in golden_names() if n.startswith("strain_")]


@pytest.mark.parametrize("name", STRAIN)
def test_sample_gradients_match_reference_golden(name, lib_built):
    """gradients through samples: the reference differentiates through its bisection / Newton iterations
    (sample(allow_gradients=True), main/default.py:1342); here the implicit-function reverse pass
    jf_subpdf_sample_backward.  Stated tolerance 1e-6 of each tensor's maximum (the reference's own iteration stops at
    1e-7 of the target)."""
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cond = t(data["cond"]) if "cond" in data else None
    p.zero_grad()
    x, _, lp, _ = p._obtain_sample(conditional_input=cond, predefined_target_input=t(data["z"]), _trainable=True)
    assert np.abs(x.detach().cpu().numpy() - data["samp_x"]).max() < 1e-9
    ((x * t(data["sgrad_w"])).sum() + 0.3 * lp.sum()).backward()
    g = {k: q.grad.detach().cpu().numpy() for k, q in p.named_parameters() if q.grad is not None}
    n = 0
    for k in data:
        if k.startswith("sgrad/"):
            ref = data[k]
            err = np.abs(g[k[6:]] - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert err < 1e-6, (k, err)
            n += 1
    assert n == len(g) and n > 0
    assert p.kernel_status()["out_of_range"] == 0


def test_sample_allow_gradients_api(lib_built):
    """pdf.sample(allow_gradients=True): reparametrised samples carry gradients to the parameters and the conditional input"""
    p = jfb.pdf("e3", "gg", conditional_input_dim=2).double().cuda()
    c = torch.randn(64, 2, dtype=torch.float64, device="cuda", requires_grad=True)
    x, z, lp, lb = p.sample(conditional_input=c, seed=3, allow_gradients=True)
    (x.pow(2).sum() - lp.sum()).backward()
    assert c.grad is not None and torch.isfinite(c.grad).all() and float(c.grad.abs().max()) > 0
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in p.parameters())
    x2, _, lp2, _ = p.sample(conditional_input=c.detach(), seed=3)
    assert torch.equal(x.detach(), x2) and torch.allclose(lp.detach(), lp2, atol=1e-12)
    with pytest.raises(NotImplementedError):
        jfb.pdf("e2", "gg", options_overwrite={"g": {"rotation_mode": "angles"}}).double().cuda().sample(samplesize=4, allow_gradients=True)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_readme_flow_trains(lib_built, dtype):
    """the README flow (Euclidean + S2 + Euclidean, autoregressively conditioned) through pdf.forward + autograd: closed-form
    reverse pass for the "g" chains, dual-number sweep for the "f" layer, generator gradients on the MLP kernels"""
    torch.manual_seed(0)
    np.random.seed(0)
    p = jfb.pdf("e4+s2+e4", "gggg+f+gggg").to(dtype).cuda()
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 4096
    a = 0.7 * torch.randn(n, 4, generator=g, device="cuda", dtype=torch.float64) + 0.3
    th = torch.acos(1 - 2 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)).clamp(0.05, 3.0)
    ph = (2 * np.pi * torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 0.5 * a[:, 0]) % (2 * np.pi)
    b = 0.5 * a + 0.5 * torch.randn(n, 4, generator=g, device="cuda", dtype=torch.float64)
    x = torch.cat([a, th[:, None], ph[:, None], b], dim=1).to(dtype)
    opt = torch.optim.Adam(p.parameters(), lr=5e-3)
    losses = []
    for _ in range(15):
        opt.zero_grad()
        lp, _, _ = p(x)
        loss = -lp.mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 0.1, losses
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in p.parameters())


def test_backward_of_unsupported_pdfs_fails_loudly(lib_built):
    p = jfb.pdf("e2", "gg", options_overwrite={"g": {"rotation_mode": "angles"}}).double().cuda()   # non-default "g" options: no backward kernel
    x = torch.tensor([[1.0, 2.0], [0.5, 4.0]], dtype=torch.float64, device="cuda")
    lp, _, _ = p(x)
    with pytest.raises(NotImplementedError):
        lp.sum().backward()


def test_generator_backward_kernel_matches_fp64_reference(lib_built):
    """jf_mlp_backward (tcgen05 kind::tf32 products, csrc/mlp_bwd.cuh) against a plain torch fp64 evaluation of the same
    gradient.  Stated tolerance of the tf32 path: 2e-3 of each tensor's maximum (10-bit operand mantissas)."""
    import torch
    from jammy_flows_b200 import engine
    torch.manual_seed(3)
    dev = torch.device("cuda")
    for B, n_in, P in ((1000, 7, 300), (4099, 64, 3210), (130, 96, 129)):
        inp = torch.randn(B, n_in, device=dev)
        w1 = 0.3 * torch.randn(128, n_in, device=dev)
        b1 = 0.1 * torch.randn(128, device=dev)
        w2 = 0.2 * torch.randn(P, 128, device=dev)
        g = torch.randn(P, B, device=dev) * torch.rand(P, 1, device=dev)
        got = engine._mlp_backward_tc(inp, w1, b1, w2, g, True)
        i64, w164, b164, w264, g64 = (t.double() for t in (inp, w1, b1, w2, g))
        h = torch.tanh(torch.addmm(b164, i64, w164.t()))
        g_pre = (g64.t() @ w264) * (1.0 - h * h)
        ref = (g_pre @ w164, g_pre.t() @ i64, g_pre.sum(0), g64 @ h, g64.sum(1))
        for name, a, r in zip(("inp", "w1", "b1", "w2", "b2"), got, ref):
            err = float((a.double() - r).abs().max() / r.abs().max())
            assert err < 2e-3, (B, n_in, P, name, err)
