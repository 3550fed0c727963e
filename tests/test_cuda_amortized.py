"""GPU: `fully_amortized_pdf` / `pdf(..., amortize_everything=True)` (reference
main/fully_amortized.py:22-278, amortizable_mlp.py:470-682 with per-row weights) against the reference goldens, the
oracle, and the per-row Linear kernel `jf_rowwise_linear` against torch on every shape class it distinguishes."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import REL_TOL, base_tolerance, build_fa, fa_golden_names, fa_outer_spec, load_golden, rel_err
from oracle.jf_oracle import OraclePdf, fully_amortized_parameters

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", fa_golden_names())
def test_fully_amortized_matches_reference(name, lib_built):
    meta, params, data = load_golden(name)
    fa = build_fa(meta, params).cuda()
    x, cond, z = (torch.from_numpy(data[k]).cuda() for k in ("x", "cond", "z"))
    tol = REL_TOL["float64"]
    with torch.no_grad():
        am = fa.amortization_parameters(cond)
        logp, logp_base, base = fa(x, conditional_input=cond)
        xs, _, slp, slb = fa.pdf_to_amortize._obtain_sample(amortization_parameters=am, predefined_target_input=z)
        rt_lp, _, rt_base = fa(xs, conditional_input=cond)
    assert am.shape == (x.shape[0], meta["total_number_amortizable_params"])
    assert rel_err(am[:8].cpu().numpy(), data["amort_head"]).max() < 1e-12
    assert rel_err(logp.cpu().numpy(), data["logp"]).max() < tol
    # base coordinates behind an inverse-normal-CDF stage carry the REFERENCE's erfinv(2 cdf - 1) rounding
    # (helpers.icdf_conditioning); log N(base) alone inherits z*dz of it (it cancels in log_pdf)
    base_err = np.abs(base.cpu().numpy() - data["base"]).max(axis=1)
    assert (base_err <= base_tolerance(fa.pdf_to_amortize, "float64", data["base"])).all(), base_err.max()
    cond_term = (np.abs(data["base"]) * base_err[:, None]).sum(axis=1)
    assert (np.abs(logp_base.cpu().numpy() - data["logp_base"])
            <= tol * np.maximum(1, np.abs(data["logp_base"])) + cond_term).all()
    stol = max(1e-9, 10 * float(data["ref_roundtrip_base_err"]))
    assert rel_err(xs.cpu().numpy(), data["samp_x"]).max() < stol
    assert rel_err(slp.cpu().numpy(), data["samp_logp"]).max() < stol
    assert rel_err(slb.cpu().numpy(), data["samp_logp_base"]).max() < 1e-12
    # own round trip
    assert (rt_base - z).abs().max() < 1e-9
    assert (rt_lp - slp).abs().max() < 1e-9
    st = fa.kernel_status()
    assert st["nonfinite"] == 0 and st["unconverged"] == 0


def test_fully_amortized_fresh_inputs_chunking_and_sample(lib_built):
    """fresh parameters and 20 000 rows against the oracle; results do not depend on the row chunking; sample() draws
    the same base normals for any chunking (numpy RNG, seed) and round-trips through forward()."""
    meta, _, _ = load_golden("fa_e2s2e2_lowrank_mode1")
    fa = build_fa(meta, seed=11)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p_ in fa.parameters():
            p_.add_(0.03 * torch.randn(p_.shape, generator=gen, dtype=torch.float64))
    n = 20000
    cond = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    x = 1.5 * torch.randn(n, 6, generator=gen, dtype=torch.float64)
    x[:, 2] = torch.acos(1 - 2 * torch.rand(n, generator=gen, dtype=torch.float64))
    x[:, 3] = 2 * np.pi * torch.rand(n, generator=gen, dtype=torch.float64)
    params = {k: v.numpy() for k, v in fa.state_dict().items()}
    m = 1500                                                     # oracle rows (per-row einsums on the CPU)
    am_o = fully_amortized_parameters(fa_outer_spec(fa), params, cond[:m])
    o = OraclePdf(fa.pdf_to_amortize.export_program("float64"), {})
    lp_o, _, base_o = o.log_pdf(x[:m], None, amort=am_o)
    fa = fa.cuda()
    xc, cc = x.cuda(), cond.cuda()
    with torch.no_grad():
        lp, lb, base = fa(xc, conditional_input=cc)
        fa.chunk_rows = 3331
        lp2, lb2, base2 = fa(xc, conditional_input=cc)
        xs2, z2, slp2, _ = fa.sample(conditional_input=cc, seed=9)
        fa.chunk_rows = None
        xs, z, slp, _ = fa.sample(conditional_input=cc, seed=9)
        rt_lp, _, rt_base = fa(xs, conditional_input=cc)
    assert rel_err(lp[:m].cpu().numpy(), lp_o.numpy()).max() < 1e-10
    assert rel_err(base[:m].cpu().numpy(), base_o.numpy()).max() < 1e-9
    assert torch.equal(lp, lp2) and torch.equal(base, base2) and torch.equal(lb, lb2)
    assert torch.equal(z, z2) and torch.equal(xs, xs2) and torch.equal(slp, slp2)
    assert (rt_base - z).abs().max() < 1e-8 and (rt_lp - slp).abs().max() < 1e-8
    st = fa.kernel_status()
    assert st["nonfinite"] == 0 and st["unconverged"] == 0


def test_amortize_everything_argument_checks(lib_built):
    import jammy_flows_b200 as jfb
    p = jfb.pdf("e2+e1", "gg+g", amortization_mlp_use_custom_mode=True, amortization_mlp_dims="8",
                amortize_everything=True)
    t = p.total_number_amortizable_params
    x = torch.randn(10, 3, dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):
        p(x)                                                              # needs amortization_parameters
    with pytest.raises(AssertionError):
        p(x, amortization_parameters=torch.zeros(10, t + 1, dtype=torch.float64, device="cuda"))
    with pytest.raises(AssertionError):
        p(x, amortization_parameters=torch.zeros(9, t, dtype=torch.float64, device="cuda"))
    q = jfb.pdf("e2", "gg").cuda()
    with pytest.raises(AssertionError):
        q(x[:, :2], amortization_parameters=torch.zeros(10, 5, dtype=torch.float64, device="cuda"))
    init = p.init_params()
    assert init.shape == (t,)
    with torch.no_grad():
        lp, _, _ = p(x, amortization_parameters=init.double().cuda().unsqueeze(0).repeat(10, 1))
    assert torch.isfinite(lp).all()


SHAPES = [(1, 1), (5, 7), (7, 128), (31, 33), (32, 5), (33, 64), (128, 10), (128, 548), (300, 3), (3, 700)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n_in,n_out", SHAPES)
def test_rowwise_linear_matches_torch(n_in, n_out, dtype, lib_built):
    """out[r] (+)= act(W_r in[r] + b_r) with W_r, b_r at arbitrary column offsets of a wider per-row block: narrow
    (n_in < 32, staged tiles) and wide (n_in >= 32, shuffle reduction) paths, bias / tanh / accumulate, strided output."""
    from jammy_flows_b200 import _cabi
    lib = _cabi.load()
    code = _cabi.JF_F64 if dtype == torch.float64 else _cabi.JF_F32
    gen = torch.Generator().manual_seed(n_in * 1000 + n_out)
    R = 777
    off_w, off_b = 3, 3 + n_in * n_out + 2
    ld = off_b + n_out + 5
    params = (torch.randn(R, ld, generator=gen, dtype=torch.float64) / np.sqrt(n_in)).to(dtype).cuda()
    inp_full = torch.randn(R, n_in + 2, generator=gen, dtype=torch.float64).to(dtype).cuda()
    inp = inp_full[:, 1:1 + n_in]                                    # leading dimension > n_in, offset start
    w = params[:, off_w:off_w + n_in * n_out].reshape(R, n_out, n_in).double()
    b = params[:, off_b:off_b + n_out].double()
    lin = torch.einsum("roi,ri->ro", w, inp.double())
    tol = 1e-13 if dtype == torch.float64 else 3e-6
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for bias, act, acc, param_major in [(1, 1, 0, 0), (0, 0, 0, 0), (1, 0, 1, 0), (1, 1, 1, 1), (0, 0, 0, 1)]:
        ref = lin + (b if bias else 0.0)
        if act:
            ref = torch.tanh(ref)
        prev = torch.randn(R, n_out, generator=gen, dtype=torch.float64).to(dtype).cuda()
        if acc:
            ref = ref + prev.double()
        out = prev.t().contiguous() if param_major else prev.clone()
        so_p, so_r = (R, 1) if param_major else (1, n_out)
        rc = lib.jf_rowwise_linear(code, C.c_void_p(params.data_ptr()), ld, off_w, off_b if bias else -1,
                                   C.c_void_p(inp.data_ptr()), inp.stride(0), n_in, n_out, act, acc,
                                   C.c_void_p(out.data_ptr()), so_p, so_r, R, st)
        assert rc == 0
        torch.cuda.synchronize()
        got = (out.t() if param_major else out).double()
        assert ((got - ref).abs() / ref.abs().clamp(min=1)).max() < tol, (bias, act, acc, param_major)
    # argument checks: a weight block that does not fit the row, no rows
    assert lib.jf_rowwise_linear(code, C.c_void_p(params.data_ptr()), n_in * n_out - 1, 0, -1, C.c_void_p(inp.data_ptr()),
                                 inp.stride(0), n_in, n_out, 0, 0, C.c_void_p(out.data_ptr()), 1, n_out, R, st) < 0
    assert lib.jf_rowwise_linear(code, None, ld, off_w, -1, None, n_in, n_in, n_out, 0, 0, None, 1, n_out, 0, st) == 0


def test_fully_amortized_embedding_coordinates_and_fp32(lib_built):
    """force_embedding_coordinates through the amortized path (charts before / after the chain, main/default.py:906-909,
    :1522-1524) and the same pdf evaluated in float32 (dtype follows the inputs: the inner pdf owns no parameters)."""
    meta, params, data = load_golden("fa_e2s2e2_lowrank_mode1")
    fa = build_fa(meta, params).cuda()
    inner = fa.pdf_to_amortize
    x, cond, z = (torch.from_numpy(data[k]).cuda() for k in ("x", "cond", "z"))
    with torch.no_grad():
        lp, _, base = fa(x, conditional_input=cond)
        x_emb, ld = inner.transform_target_space(x, 0.0, transform_from="default", transform_to="embedding")
        lp_e, _, base_e = fa(x_emb, conditional_input=cond, force_embedding_coordinates=True)
        am = fa.amortization_parameters(cond)
        xs_e, _, slp_e, _ = inner._obtain_sample(amortization_parameters=am, predefined_target_input=z,
                                                 force_embedding_coordinates=True)
        xs, _, slp, _ = inner._obtain_sample(amortization_parameters=am, predefined_target_input=z)
        back, ld_s = inner.transform_target_space(xs, 0.0, transform_from="default", transform_to="embedding")
    assert x_emb.shape[1] == 7 and xs_e.shape[1] == 7
    assert (lp_e - (lp - ld)).abs().max() < 1e-9          # the round trip through acos near the poles costs ~1e-10
    assert (base_e - base).abs().max() < 1e-7
    assert torch.equal(xs_e, back) and (slp_e - (slp - ld_s)).abs().max() < 1e-12
    assert ((xs_e[:, 2:5] ** 2).sum(dim=1) - 1).abs().max() < 1e-14
    # float32: amortization parameters, inputs and outputs in single precision against the fp64 result
    with torch.no_grad():
        lp32, _, base32 = inner(x.float(), amortization_parameters=am.float())
    assert lp32.dtype == torch.float32
    err = (lp32.double() - lp).abs() / lp.abs().clamp(min=1)
    assert err.median() < 1e-5 and err.max() < 1e-3       # rows in the far tails amplify the fp32 rounding of the weights
