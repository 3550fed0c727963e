"""GPU: properties and edge cases of the widened rows (non-default "g" options on the general chain kernel, charts,
marginal entropies, AmortizableMLP generators) at sizes and on inputs the goldens do not cover."""
import numpy as np
import pytest
import torch

import jammy_flows_b200 as jfb
from helpers import rel_err, row_rel_err
from oracle.jf_oracle import OraclePdf
from test_cuda_properties import _inputs, _perturbed

pytestmark = pytest.mark.gpu

G_OPTS = [
    {"rotation_mode": "angles"},
    {"rotation_mode": "triangular_combination"},
    {"rotation_mode": "none", "softplus_for_width": 1},
    {"clamp_widths": 1, "upper_bound_for_widths": -1, "width_smooth_saturation": 0},
    {"add_skewness": 1},
    {"nonlinear_stretch_type": "rq_splines"},
    {"center_mean": 1, "inverse_function_type": "inormal_partly_precise"},
]


@pytest.mark.parametrize("opts", G_OPTS, ids=lambda o: "-".join("%s=%s" % kv for kv in o.items()))
@pytest.mark.parametrize("cond_dim", [None, 3])
def test_g_options_fresh_inputs_match_oracle(opts, cond_dim, lib_built):
    """fresh seeded parameters and inputs, shared and per-row parameters, every option of the reference's sweep"""
    p = _perturbed("e3+e2", "gg+ggt", cond=cond_dim, scale=0.2, options_overwrite={"g": opts})
    x, z, c = _inputs(p, 2000)
    o = OraclePdf(p.export_program(), {k: v.numpy() for k, v in p.state_dict().items()})
    lp_o, _, b_o = o.log_pdf(x, c)
    xs_o, slp_o, _ = o.sample(z, c)
    pc = p.cuda()
    cc = c.cuda() if c is not None else None
    with torch.no_grad():
        lp, _, b = pc(x.cuda(), conditional_input=cc)
        xs, _, slp, _ = pc._obtain_sample(conditional_input=cc, predefined_target_input=z.cuda())
        rt_lp, _, rt_b = pc(xs, conditional_input=cc)
    # with an inverse-normal stage in EVERY layer the oracle's own Phi^-1 noise (eps/phi(z), helpers.icdf_conditioning)
    # enters the later layers and the log-det, not just the base coordinates
    all_inormal = opts.get("inverse_function_type", "").startswith("inormal")
    assert rel_err(lp.cpu().numpy(), lp_o.numpy()).max() < (2e-9 if all_inormal else 1e-10)
    assert np.median(rel_err(lp.cpu().numpy(), lp_o.numpy())) < 1e-13
    # base coordinates of layer 0 come out of an inverse normal CDF: conditioning of the oracle's own erfinv(2cdf-1)
    calm = np.abs(b_o.numpy()).max(axis=1) < 4.5
    assert row_rel_err(b.cpu().numpy(), b_o.numpy())[calm].max() < 1e-9
    # sampling inverts the same stages.  Reference-side effects that bound the comparison:
    #  * the ORACLE's Phi^-1 (erfinv(2cdf-1)) loses eps/phi(y) near |y| ~ 5 (helpers.icdf_conditioning); with an inverse
    #    normal stage in every layer that also hits intermediate values, and targets inside the bulk/Pade jump of an
    #    intermediate stage have no unique pre-image (two x that both map back to z within 1e-10 differ by 1e-3);
    #  * the skewed mixture's log(exp(sp)-1) cancellation (extra_functions.py:56) leaves ~1e-9 in log p.
    # The own round trip below is the sharp check in those cases.
    zc = z.numpy()
    calm_z = np.abs(zc).max(axis=1) < 4.5
    ex = row_rel_err(xs.cpu().numpy(), xs_o.numpy())[calm_z]
    el = rel_err(slp.cpu().numpy(), slp_o.numpy())[calm_z]
    assert np.median(ex) < 1e-12 and np.median(el) < 1e-12
    if all_inormal:
        assert np.quantile(ex, 0.99) < 1e-8 and np.quantile(el, 0.99) < 1e-8
    else:
        tol_s = 5e-9 if opts.get("add_skewness") else 1e-9
        assert ex.max() < tol_s and el.max() < tol_s
    # own round trip
    # (a few rows of these deliberately wild parameter sets pass through logit targets ~1e3 and scales ~1e6 between the
    #  sub-pdfs, where a coordinate of size 1 is recovered from one of size 1e6: bounded, not tight)
    ok = np.abs(zc).max(axis=1) < 5.0
    rt = row_rel_err(rt_b.cpu().numpy(), zc)[ok]
    assert np.quantile(rt, 0.99) < 1e-8 and np.median(rt) < 1e-12 and rt.max() < 1e-4
    assert np.quantile(rel_err(rt_lp.cpu().numpy(), slp.cpu().numpy())[ok], 0.99) < 1e-8
    st = pc.kernel_status()
    assert st["nonfinite"] == 0 and st["unconverged"] == 0


def test_general_kernel_round_trip_chunking_and_edges(lib_built):
    """1 M rows through a chain that mixes options (angles + skewness, rq_splines, default "g", "t"): round trip,
    chunk/shard invariance, empty batch, single row, far tails"""
    opts = {(0, 0): {"g": {"rotation_mode": "angles", "add_skewness": 1}},
            (0, 1): {"g": {"nonlinear_stretch_type": "rq_splines", "rotation_mode": "triangular_combination"}}}
    p = _perturbed("e4", "gggt", scale=0.15, options_overwrite=opts).cuda()
    n = 1_000_000
    z = torch.randn(n, 4, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    with torch.no_grad():
        x, _, logp, _ = p._obtain_sample(predefined_target_input=z)
        rt_logp, _, rt_z = p(x)
        calm = z.abs().max(dim=1)[0] < 5.2
        err = (rt_z - z).abs().max(dim=1)[0]
        assert float(err[calm].max()) < 1e-7 and float(err[calm].quantile(0.999)) < 1e-10
        assert float(((rt_logp - logp).abs() / logp.abs().clamp(min=1))[calm].max()) < 1e-8
        st = p.kernel_status()
        assert st["nonfinite"] == 0 and st["unconverged"] <= int((~calm).sum())
        # chunk / shard invariance (the multi-GPU contract)
        ref = p(x[:20011])
        p.chunk_rows = 777
        a = p(x[:20011])
        p.chunk_rows = None
        b = p(x[5000:9000])
        for r, q in zip(ref, a):
            assert torch.equal(r, q)
        for r, q in zip(ref, b):
            assert torch.equal(r[5000:9000], q)
        # empty batch, single row, far tails
        out = p(torch.zeros(0, 4, dtype=torch.float64, device="cuda"))
        assert out[0].shape == (0,) and out[2].shape == (0, 4)
        one = p(x[:1].contiguous())
        assert torch.equal(one[0], ref[0][:1])
        far = torch.tensor([[1e4, -1e4, 3e3, -50.0], [-30.0, 25.0, 40.0, -1e3]], dtype=torch.float64, device="cuda")
        lp, _, base = p(far)
        assert torch.isfinite(lp).all() and torch.isfinite(base).all()


def test_general_kernel_fp32_matches_fp64(lib_built):
    """fp32 instantiation of the general kernel against its own fp64 run on the same (fp32-rounded) inputs"""
    opts = {"g": {"rotation_mode": "angles", "softplus_for_width": 1}}
    p64 = _perturbed("e3", "gg", cond=2, scale=0.1, options_overwrite=opts)
    x, z, c = _inputs(p64, 4000)
    x32, z32, c32 = x.float(), z.float(), c.float()
    p32 = jfb.pdf("e3", "gg", conditional_input_dim=2, options_overwrite=opts).float()
    p32.load_state_dict({k: v.float() for k, v in p64.state_dict().items()})
    p64.load_state_dict({k: v.float().double() for k, v in p64.state_dict().items()})
    p64, p32 = p64.cuda(), p32.cuda()
    with torch.no_grad():
        lp64, _, b64 = p64(x32.double().cuda(), conditional_input=c32.double().cuda())
        lp32, _, b32 = p32(x32.cuda(), conditional_input=c32.cuda())
    calm = (b64.abs().max(dim=1)[0] < 4.0).cpu().numpy()
    assert rel_err(lp32.cpu().numpy(), lp64.cpu().numpy())[calm].max() < 2e-4
    assert np.median(rel_err(lp32.cpu().numpy(), lp64.cpu().numpy())[calm]) < 1e-5
    assert row_rel_err(b32.cpu().numpy(), b64.cpu().numpy())[calm].max() < 2e-3


def test_chart_and_entropy_edges(lib_built):
    """charts: round trip on random points incl. near the poles, empty batch; entropy: shapes, determinism, the total
    entropy of a standard normal-like flow is finite and equals -mean(log p) of its own samples"""
    p = _perturbed("e2+s2+s1", "gg+f+m", scale=0.2, cond=2).cuda()
    x, _, c = _inputs_mixed(p, 5000)
    with torch.no_grad():
        xe, ld = p.transform_target_space(x, 0.0, transform_from="default", transform_to="embedding")
        assert xe.shape == (5000, 2 + 3 + 2)
        xb, ld2 = p.transform_target_space(xe, 0.0, transform_from="embedding", transform_to="default")
        assert float((xb - x).abs().max()) < 1e-7
        assert torch.allclose(ld, -ld2, atol=1e-9)
        assert torch.allclose(ld, torch.log(torch.sin(x[:, 2])), atol=1e-12)
        e0, _ = p.transform_target_space(x[:0], 0.0, transform_from="default", transform_to="embedding")
        assert e0.shape == (0, 7)
        # log_pdf in embedding coordinates = intrinsic log_pdf - log sin(theta)
        lp_i, _, _ = p(x, conditional_input=c)
        lp_e, _, _ = p(xe, conditional_input=c, force_embedding_coordinates=True)
        assert torch.allclose(lp_e, lp_i - torch.log(torch.sin(x[:, 2])), rtol=1e-9, atol=1e-9)
        torch.manual_seed(9)
        ent = p.entropy(sub_manifolds=[-1, 0, 1, 2], conditional_input=c[:4], samplesize=32)
        torch.manual_seed(9)
        ent2 = p.entropy(sub_manifolds=[-1, 0, 1, 2], conditional_input=c[:4], samplesize=32)
        for k in ("total", 0, 1, 2):
            assert ent[k].shape == (4,) and torch.isfinite(ent[k]).all()
            assert torch.equal(ent[k], ent2[k])
        # chain rule sanity: sum of marginal entropies >= total entropy up to Monte-Carlo noise
        assert float((ent[0] + ent[1] + ent[2] - ent["total"]).min()) > -1.0


def _inputs_mixed(p, n, seed=7):
    g = torch.Generator().manual_seed(seed)
    cols = []
    for k, d in enumerate(p.pdf_defs_list):
        if d[0] == "e":
            cols.append(1.5 * torch.randn(n, p.target_dims[k], generator=g, dtype=torch.float64))
        elif d == "s2":
            u = torch.rand(n, generator=g, dtype=torch.float64)
            th = torch.acos(1 - 2 * u)
            th[:10] = torch.tensor([1e-6, 1e-4, np.pi - 1e-6, np.pi - 1e-4, 0.5, 1.0, 2.0, 3.0, 1e-3, 3.14], dtype=torch.float64)
            cols.append(torch.stack([th, 2 * np.pi * torch.rand(n, generator=g, dtype=torch.float64)], 1))
        else:
            cols.append(2 * np.pi * torch.rand(n, 1, generator=g, dtype=torch.float64))
    c = torch.randn(n, p.conditional_input_dim, generator=g, dtype=torch.float64) if p.conditional_input_dim else None
    return torch.cat(cols, 1).cuda(), None, (c.cuda() if c is not None else None)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_amortizable_mlp_module_matches_oracle(mode, lib_built):
    """AmortizableMLP.forward on its own (reference amortizable_mlp.py:586-682) and inside a pdf at 50k rows"""
    from jammy_flows_b200.amortizable_mlp import AmortizableMLP
    torch.manual_seed(mode)
    ranks = {0: "3-0-2", 1: "3-0-2-4", 2: "3-0-2-0-5", 3: 2, 4: 0}[mode]
    m = AmortizableMLP(6, "20-12", 9, highway_mode=mode, low_rank_approximations=ranks)
    x = torch.randn(1000, 6, dtype=torch.float64)
    ref = OraclePdf._custom_mlp(m.structure(), m.u_v_b_pars.detach(), x)
    out = m.cuda()(x.cuda())
    assert out.shape == (1000, 9)
    assert float((out.cpu() - ref).abs().max()) < 1e-12 * max(1.0, float(ref.abs().max()))
    p = _perturbed("e2+e2", "gg+gg", cond=3, scale=0.05, amortization_mlp_use_custom_mode=True,
                   amortization_mlp_dims="32", amortization_mlp_ranks=4, amortization_mlp_highway_mode=mode).cuda()
    n = 50_000
    z = torch.randn(n, 4, dtype=torch.float64, device="cuda")
    c = torch.randn(n, 3, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        xs, _, lp, _ = p._obtain_sample(conditional_input=c, predefined_target_input=z)
        rt_lp, _, rt_z = p(xs, conditional_input=c)
    calm = z.abs().max(dim=1)[0] < 5.0
    assert float((rt_z - z).abs().max(dim=1)[0][calm].max()) < 1e-8
    assert float(((rt_lp - lp).abs() / lp.abs().clamp(min=1))[calm].max()) < 1e-8
    # the staged path is row-chunked: any chunk size gives bit-identical results
    p.chunk_rows = 7001
    with torch.no_grad():
        xs2, _, lp2, _ = p._obtain_sample(conditional_input=c, predefined_target_input=z)
    assert torch.equal(xs, xs2) and torch.equal(lp, lp2)


def test_failsafe_crosscheck_resamples_deviating_rows(lib_built):
    """`failsafe_crosscheck_tolerance` (reference extra_functions.py:413-533): a generous tolerance changes nothing, a
    tolerance inside the round-trip noise re-draws exactly the rows that exceed it"""
    p = _perturbed("e2+s2", "gg+v", scale=0.0).cuda()
    p.rng_mode = "device"
    with torch.no_grad():
        a = p.sample(samplesize=4000, seed=3)
        b = p.sample(samplesize=4000, seed=3, failsafe_crosscheck_tolerance=1e-6)
        for u, w in zip(a, b):
            assert torch.equal(u, w)
        tol = 2e-13
        c = p.sample(samplesize=4000, seed=3, failsafe_crosscheck_tolerance=tol)
        lp, _, base = p(c[0])
        assert float((base - c[1]).abs().max()) <= 10 * tol and float((lp - c[2]).abs().max()) <= 10 * tol
        changed = int((c[1] != a[1]).any(dim=1).sum())
        print("\nfailsafe at %.0e: %d of 4000 rows re-drawn" % (tol, changed))
        assert 0 < changed < 4000
        ent = p.entropy(samplesize=500, failsafe_crosscheck_tolerance=1e-6)
        assert torch.isfinite(ent["total"]).all()


def test_device_rng_matches_oracle_and_is_shard_invariant(lib_built):
    """`jf_normal_rows` (Philox4x32-10 + Box-Muller) against its numpy restatement, which is pinned to the published
    known-answer vectors; the union of per-rank draws equals the single-process draw bit for bit."""
    from jammy_flows_b200 import engine, sharding
    from oracle.philox import normal_rows as oracle_rows
    n, d, seed = 100_003, 5, 1234567890123
    z = engine.normal_rows(n, d, seed)
    ref = oracle_rows(seed, 0, n, d)
    assert np.abs(z.cpu().numpy() - ref).max() < 1e-12
    for world in (2, 3, 8):
        parts = [sharding.shard_base_normals(n, d, seed, rank=r, world=world) for r in range(world)]
        assert torch.equal(torch.cat(parts), z)
    z32 = engine.normal_rows(1000, 3, 9, first_row=2 ** 33, dtype=torch.float32)
    assert np.abs(z32.cpu().numpy() - oracle_rows(9, 2 ** 33, 1000, 3)).max() < 1e-6
    assert engine.normal_rows(0, 3, 1).shape == (0, 3)
    # pdf.sample with the counter-based generator: same seed -> same samples, shards of the batch agree with the whole
    p = _perturbed("e2+s2", "gg+f", scale=0.2).cuda()
    p.rng_mode = "philox"
    with torch.no_grad():
        a = p.sample(samplesize=5000, seed=11)
        b = p.sample(samplesize=5000, seed=11)
        p.rng_first_row = 3000
        c = p.sample(samplesize=2000, seed=11)
    assert torch.equal(a[0], b[0]) and torch.equal(a[0][3000:], c[0]) and torch.equal(a[2][3000:], c[2])


@pytest.mark.parametrize("name", ["last_e3s2e2", "last_e2s1s2_cond"])
def test_only_last_matches_reference(name, lib_built):
    """forward(..., only_last=True) / _obtain_sample(..., only_last=True) (main/default.py:1015-1024, :1490-1502): a
    one-layer program per sub-pdf (its last layer, sphere layers with the base chart forced on) on the same kernels."""
    from helpers import build_pdf, load_golden
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    t = lambda a: torch.from_numpy(a).cuda()
    cond = t(data["cond"]) if "cond" in data else None
    with torch.no_grad():
        lp, lb, base = p(t(data["x"]), conditional_input=cond, only_last=True)
        xs, _, slp, _ = p._obtain_sample(conditional_input=cond, predefined_target_input=t(data["z"]), only_last=True)
        rt_lp, _, rt_base = p(xs, conditional_input=cond, only_last=True)
        full_lp, _, _ = p(t(data["x"]), conditional_input=cond)
    assert rel_err(lp.cpu().numpy(), data["last_logp"]).max() < 1e-10
    assert rel_err(lb.cpu().numpy(), data["last_logp_base"]).max() < 1e-10
    assert rel_err(base.cpu().numpy(), data["last_base"]).max() < 1e-10
    stol = max(1e-9, 10 * float(data["ref_roundtrip_base_err"]))
    assert rel_err(xs.cpu().numpy(), data["last_samp_x"]).max() < stol
    assert rel_err(slp.cpu().numpy(), data["last_samp_logp"]).max() < stol
    assert (rt_base - t(data["z"])).abs().max() < 1e-8 and (rt_lp - slp).abs().max() < 1e-8
    assert (full_lp - lp).abs().max() > 1e-3          # it really is a different (shorter) flow
    st = p.kernel_status()
    assert st["nonfinite"] == 0 and st["unconverged"] == 0


@pytest.mark.parametrize("name", ["poisson_e2_joint_cond", "poisson_s2_joint_cond", "poisson_e2_uncond"])
def test_poisson_log_lambda_matches_reference(name, lib_built):
    """pdf.log_mean_poisson (main/default.py:832-877): the last output of the first generator (joint prediction) or
    the free parameter; the flow evaluated next to it is covered by the golden parity tests on the same files."""
    from helpers import build_pdf, load_golden
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    cond = torch.from_numpy(data["cond"]).cuda() if "cond" in data else None
    ll = p.log_mean_poisson(conditional_input=cond)
    assert tuple(ll.shape) == data["log_lambda"].shape
    assert rel_err(ll.detach().cpu().numpy(), data["log_lambda"]).max() < 1e-12
    with torch.no_grad():
        lp, _, _ = p(torch.from_numpy(data["x"]).cuda(), conditional_input=cond)
    assert rel_err(lp.cpu().numpy(), data["logp"]).max() < 1e-10
