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

from helpers import REL_TOL, build_pdf, golden_names, icdf_conditioning, load_golden, rel_err, row_rel_err

pytestmark = pytest.mark.gpu


def _cuda_model(name):
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params).cuda()
    dt = getattr(torch, meta["dtype"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    cond = t(data["cond"]) if "cond" in data else None
    return meta, data, p, t, cond


def _base_tolerance(p, meta, base_ref):
    """per-row tolerance for base coordinates: relative tolerance + inverse-normal-CDF conditioning of the reference"""
    tol = REL_TOL[meta["dtype"]] * np.maximum(1.0, np.abs(base_ref).max(axis=1))
    extra = np.zeros(base_ref.shape[0])
    for k, layers in enumerate(p.layer_list):
        l0 = layers[0]
        if getattr(l0, "inverse_function_type", "isigmoid") in ("inormal_partly_precise", "inormal_partly_crude"):
            b0, b1 = p.base_dim_indices[k]
            extra = np.maximum(extra, icdf_conditioning(base_ref[:, b0:b1], meta["dtype"]).max(axis=1))
    return tol + extra


@pytest.mark.parametrize("name", golden_names())
def test_logpdf_matches_reference_golden(name, lib_built):
    meta, data, p, t, cond = _cuda_model(name)
    tol = REL_TOL[meta["dtype"]]
    with torch.no_grad():
        logp, logp_base, base = p(t(data["x"]), conditional_input=cond)
    logp, logp_base, base = logp.cpu().numpy(), logp_base.cpu().numpy(), base.cpu().numpy()
    ok = np.isfinite(data["logp"])
    base_err = np.abs(base - data["base"]).max(axis=1)
    assert (base_err[ok] <= _base_tolerance(p, meta, data["base"])[ok]).all(), base_err[ok].max()
    logp_tol = tol
    if "full_pade" in str(meta["options_overwrite"]):
        logp_tol = max(tol, float(data["ref_roundtrip_logp_err"]))
    if meta["dtype"] == "float32":
        # fp32: log N(z) inherits z*dz from the conditioning term above
        cond_term = (np.abs(data["base"]) * icdf_conditioning(data["base"], "float32")).sum(axis=1)
        assert (np.abs(logp - data["logp"])[ok] <= (tol * np.maximum(1, np.abs(data["logp"])) + cond_term)[ok]).all()
    else:
        assert rel_err(logp, data["logp"])[ok].max() < logp_tol
        # log N(base) on its own inherits z*dz of the ill-conditioned entries
        cond_term = (np.abs(data["base"]) * (base_err[:, None] + 0 * data["base"])).sum(axis=1)
        assert (np.abs(logp_base - data["logp_base"])[ok] <= (tol * np.maximum(1, np.abs(data["logp_base"])) + cond_term)[ok]).all()
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
    assert el < max(tol, 10 * ref_rt_lp), (el, ref_rt_lp)
    assert rel_err(logp_base.cpu().numpy(), data["samp_logp_base"])[ok].max() < tol
    assert rt < max(100 * tol, 10 * ref_rt)
    st = p.kernel_status()
    if meta["dtype"] == "float64":
        assert st["unconverged"] == 0
