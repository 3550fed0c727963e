"""GPU parity tests proper: the CUDA path (through the C-ABI) against the committed golden vectors of the unmodified
reference and against the oracle on fresh seeded inputs.

Tolerances (BASELINE.json north_star): fp64 relative 1e-10, fp32 relative 1e-5, measured per row as
max-norm(error)/max(1, max-norm(reference row)) for coordinates and |error|/max(1,|reference|) for log-densities.
Two documented, reference-side effects widen the bound for specific entries (helpers.icdf_conditioning):
  * base coordinates that leave an inverse-normal-CDF stage carry the reference's own eps/phi(z) rounding noise;
  * `inormal_full_pade` has a numerically unstable log-derivative next to cdf=0.5 (the reference's own
    sample->forward log_pdf round trip is only 2e-8 there), so its log_pdf is compared at the reference's own
    round-trip error.
Samples: within max(tolerance, 10 x the reference's own sample->forward round-trip error); the new path's own
round-trip error is printed and bounded."""
import numpy as np
import pytest
import torch

from helpers import (REL_TOL, base_tolerance, build_pdf, golden_names, icdf_conditioning, load_golden, rel_err,
                     row_rel_err)

pytestmark = pytest.mark.gpu


def _cuda_model(name):
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    dt = getattr(torch, meta["dtype"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    cond = None
    if "cond" in data:
        cond = [t(c) for c in data["cond"]] if isinstance(data["cond"], list) else t(data["cond"])
    return meta, data, p, t, cond


def _base_tolerance(p, meta, base_ref):
    return base_tolerance(p, meta["dtype"], base_ref)


@pytest.mark.parametrize("name", golden_names())
def test_logpdf_matches_reference_golden(name, lib_built):
    meta, data, p, t, cond = _cuda_model(name)
    tol = REL_TOL[meta["dtype"]]
    with torch.no_grad():
        logp, logp_base, base = p(t(data["x"]), conditional_input=cond)
    logp, logp_base, base = logp.cpu().numpy(), logp_base.cpu().numpy(), base.cpu().numpy()
    ok = np.isfinite(data["logp"])
    base_err = np.abs(base - data["base"]).max(axis=1)
    if meta["dtype"] == "float32":
        # fp32 (SURVEY.md F3 decision): the yardstick is the fp64 oracle on the same fp32-rounded inputs/parameters.
        # The fp32 reference itself is noisy here: measured against that yardstick it is off by up to 2.6e-2 in log_pdf
        # and 1.6 in a base coordinate (rows next to the |z|=5.33 bulk/Pade switch, where rounding flips the branch);
        # the fp32 kernel is compared at 1e-5 relative + the fp32 inverse-normal conditioning, away from that switch.
        from oracle.jf_oracle import OraclePdf
        o64 = OraclePdf(p.export_program("float64"), {k: v.cpu().numpy() for k, v in p.state_dict().items()})
        lp64, _, b64 = o64.log_pdf(data["x"].astype(np.float64), data["cond"].astype(np.float64) if "cond" in data else None)
        lp64, b64 = lp64.numpy(), b64.numpy()
        calm = ok & (np.abs(b64).max(axis=1) < 5.0)
        berr = np.abs(base - b64).max(axis=1)
        btol = tol * np.maximum(1.0, np.abs(b64).max(axis=1)) + icdf_conditioning(b64, "float32").max(axis=1)
        assert (berr[calm] <= btol[calm]).all(), (berr[calm] / btol[calm]).max()
        assert (np.abs(logp - lp64)[calm] <= 10 * tol * np.maximum(1, np.abs(lp64))[calm]).all()
        assert np.median(np.abs(logp - lp64)[calm] / np.maximum(1, np.abs(lp64))[calm]) < tol
        # the fp32 golden of the reference agrees at the same level on the rows where the reference itself is sane
        sane = calm & (np.abs(data["logp"] - lp64) <= 10 * tol * np.maximum(1, np.abs(lp64)))
        assert sane.sum() > 0.9 * ok.sum()
        assert (np.abs(logp - data["logp"])[sane] <= 20 * tol * np.maximum(1, np.abs(lp64))[sane]).all()
        assert np.abs(logp - lp64)[ok].max() < 0.1        # rows at the branch switch: bounded, not wild
    else:
        extra = 0.0
        if "natural" in name:
            # natural_direction=1: the log_pdf direction is itself an iterative inverse; the reference's descent stops on
            # 1 - y.target (cancellation floor ~1e-8), so it is only as accurate as its own round trip (6.7e-8 here)
            extra = 10 * float(data["ref_roundtrip_base_err"])
        assert (base_err[ok] <= (_base_tolerance(p, meta, data["base"]) + extra)[ok]).all(), base_err[ok].max()
        logp_tol = tol + extra
        if "full_pade" in str(meta["options_overwrite"]):
            logp_tol = max(tol, float(data["ref_roundtrip_logp_err"]))
        assert rel_err(logp, data["logp"])[ok].max() < logp_tol
        # log N(base) on its own inherits z*dz of the ill-conditioned entries (it cancels in log_pdf)
        cond_term = (np.abs(data["base"]) * base_err[:, None]).sum(axis=1)
        assert (np.abs(logp_base - data["logp_base"])[ok] <= ((tol + extra) * np.maximum(1, np.abs(data["logp_base"])) + cond_term)[ok]).all()
    st = p.kernel_status()
    assert st["nonfinite"] == int((~ok).sum())


@pytest.mark.parametrize("name", golden_names())
def test_sample_matches_reference_golden(name, lib_built):
    meta, data, p, t, cond = _cuda_model(name)
    tol = REL_TOL[meta["dtype"]]
    z = t(data["z"])
    with torch.no_grad():
        x, _, logp, logp_base = p._obtain_sample(conditional_input=cond, predefined_target_input=z)
        rt_logp, _, rt_base = p(x, conditional_input=cond)      # the new path's own round trip
    ok = np.isfinite(data["samp_x"]).all(axis=1) & np.isfinite(data["samp_logp"])
    ref_rt = float(np.nan_to_num(data["ref_roundtrip_base_err"], nan=0.0))
    ref_rt_lp = float(np.nan_to_num(data["ref_roundtrip_logp_err"], nan=0.0))
    ex = row_rel_err(x.cpu().numpy(), data["samp_x"])[ok].max()
    el = rel_err(logp.cpu().numpy(), data["samp_logp"])[ok].max()
    rt = row_rel_err(rt_base.cpu().numpy(), data["z"])[ok].max()
    rtl = rel_err(rt_logp.cpu().numpy(), logp.cpu().numpy())[ok].max()
    print("\n%s: |x-x_ref| %.2e  |logp-logp_ref| %.2e  own round trip: base %.2e logp %.2e  (reference's own: %.2e / %.2e)"
          % (name, ex, el, rt, rtl, ref_rt, ref_rt_lp))
    assert ex < max(tol, 10 * ref_rt), (ex, ref_rt)
    # log p is compared at two slightly different points when the reference's sample is off by its own inverse error
    # ("v": descent stops on 1 - y.target, cancellation floor ~1e-7): allow |grad log p| ~ O(1) times that displacement
    assert el < max(tol, 10 * ref_rt_lp, 10 * ref_rt), (el, ref_rt_lp, ref_rt)
    assert rel_err(logp_base.cpu().numpy(), data["samp_logp_base"])[ok].max() < tol
    assert rt < max(100 * tol, 10 * ref_rt)
    st = p.kernel_status()
    if meta["dtype"] == "float64":
        assert st["unconverged"] == 0


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("emb_")])
def test_embedding_coordinates_match_reference_golden(name, lib_built):
    """force_embedding_coordinates in forward / _obtain_sample / entropy and pdf.transform_target_space
    (reference main/default.py:906-913, :1522-1529, :1737-1813, :2263-2369) through the chart kernel."""
    meta, data, p, t, cond = _cuda_model(name)
    tol = REL_TOL[meta["dtype"]]
    ref_rt = float(np.nan_to_num(data["ref_roundtrip_base_err"], nan=0.0))
    with torch.no_grad():
        xe, ld = p.transform_target_space(t(data["x"]), 0.0, transform_from="default", transform_to="embedding")
        assert np.abs(xe.cpu().numpy() - data["x_emb"]).max() < 1e-14
        back, ld2 = p.transform_target_space(t(data["x_emb"]), 0.0, transform_from="embedding", transform_to="intrinsic")
        assert np.abs(back.cpu().numpy() - data["x"]).max() < 1e-7          # acos near the poles
        assert torch.allclose(ld, -ld2, atol=1e-9)
        logp, _, base = p(t(data["x_emb"]), conditional_input=cond, force_embedding_coordinates=True)
        ok = np.isfinite(data["logp_emb"])
        extra = 10 * ref_rt if "v" in meta["flow_defs"] else 0.0
        assert rel_err(logp.cpu().numpy(), data["logp_emb"])[ok].max() < tol + extra
        assert (np.abs(base.cpu().numpy() - data["base_emb"]).max(axis=1)[ok]
                <= (base_tolerance(p, meta["dtype"], data["base_emb"]) + extra)[ok]).all()
        x, _, slogp, _ = p._obtain_sample(conditional_input=cond, predefined_target_input=t(data["z"]),
                                          force_embedding_coordinates=True)
        assert x.shape[1] == p.total_target_dim_embedded
        assert row_rel_err(x.cpu().numpy(), data["samp_x_emb"]).max() < max(tol, 10 * ref_rt)
        assert rel_err(slogp.cpu().numpy(), data["samp_logp_emb"]).max() < max(tol, 10 * ref_rt)
        # entropy (total): -mean log p over device-RNG samples, in embedding coordinates by default
        torch.manual_seed(5)
        n = 64
        c_small = cond[:3] if cond is not None else None
        ent = p.entropy(conditional_input=c_small, samplesize=n)["total"]
        torch.manual_seed(5)
        rows = n * (3 if cond is not None else 1)
        z = torch.randn(rows, p.total_base_dim, dtype=base.dtype, device=base.device)
        cr = c_small.repeat_interleave(n, dim=0) if cond is not None else None
        _, _, lp, _ = p._obtain_sample(conditional_input=cr, predefined_target_input=z, force_embedding_coordinates=True)
        assert torch.allclose(ent, -lp.reshape(-1, n).mean(dim=1), rtol=1e-12, atol=1e-12)
        # marginal entropies against the reference, on the reference's own base normals
        nb = 3 if cond is not None else 1
        S = int(data["ent_S"])
        subs = [-1] + list(range(len(meta["pdf_defs"].split("+"))))
        for flag, tag in ((True, "emb"), (False, "intr")):
            ent = p.entropy(sub_manifolds=subs, conditional_input=None if cond is None else cond[:nb], samplesize=S,
                            force_embedding_coordinates=flag, _base_samples=t(data["ent_z"]))
            for k_, v_ in ent.items():
                ref = data["ent_%s_%s" % (tag, k_)]
                assert rel_err(v_.cpu().numpy(), ref).max() < max(tol, 10 * ref_rt), (tag, k_)
