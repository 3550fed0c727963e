"""Shared helpers of the test-suite: golden-vector loading, model construction, and the parity tolerances."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json north_star: "fp64 relative 1e-10, fp32 relative 1e-5"
REL_TOL = {"float64": 1e-10, "float32": 1e-5}
EPS = {"float64": 2.220446049250313e-16, "float32": 1.1920929e-07}


def golden_names():
    """goldens of `pdf` (the `fa_*` files hold `fully_amortized_pdf` cases: fa_golden_names)"""
    return sorted(n for n in (os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith("fa_"))


def fa_golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "fa_*.npz")))


def build_fa(meta, params=None, seed=None):
    """Construct jammy_flows_b200.fully_amortized_pdf as the golden's reference model was constructed."""
    import jammy_flows_b200 as jfb
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    fa = jfb.fully_amortized_pdf(meta["pdf_defs"], meta["flow_defs"], **meta["fa_kw"])
    if params is not None:
        fa.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    return fa


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    params = {k[6:]: g[k] for k in g.files if k.startswith("param/")}
    data = {k: g[k] for k in g.files if not k.startswith("param/") and k != "meta"}   # includes "grad/<name>" entries
    if "cond0" in data:     # one conditional input per sub-pdf
        data["cond"] = [data["cond%d" % i] for i in range(len(meta["conditional_input_dim"]))]
    return meta, params, data


def _opts(meta):
    """options_overwrite as stored in the golden's json: keys are layer codes, or str() of an int / (int, int) tuple"""
    import ast
    out = {}
    for k, v in meta["options_overwrite"].items():
        key = ast.literal_eval(k) if (k.startswith("(") or k.lstrip("-").isdigit()) else k
        out[key] = v
    return out


def build_pdf(meta, params=None, seed=None):
    """Construct jammy_flows_b200.pdf as the golden's reference model was constructed; optionally load its params."""
    import jammy_flows_b200 as jfb
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    p = jfb.pdf(meta["pdf_defs"], meta["flow_defs"], options_overwrite=_opts(meta),
                conditional_input_dim=meta["conditional_input_dim"], **meta.get("pdf_kw", {}))
    p = p.to(getattr(torch, meta["dtype"]))
    if params is not None:
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
        p.load_state_dict(sd)
    return p


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def row_rel_err(a, b):
    """max-norm error of a row relative to the row's max-norm (rotations mix the coordinates of a row)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))


def icdf_conditioning(base, dtype):
    """Irreducible fp noise of the REFERENCE in base coordinates that come out of an inverse-normal-CDF stage.

    The reference evaluates Phi^-1(cdf) as sqrt2*erfinv(2*cdf-1) (gaussianization_flow.py:499-500): the argument is
    rounded to eps, which maps to eps/phi(z) in z.  Measured against 40-digit mpmath evaluations of the reference's
    formulas (tools/mp_truth.py, DESIGN.md section 'Parity'): the reference is off by up to 1e-9 at |z|~5.3 in fp64
    while the CUDA path (erfcinv on the small tail) is within 1e-14.  Outside the bulk (cdf within 0.5e-7 of 0/1,
    |z| > 5.33) the Pade tail takes over, which is well conditioned, so the term is capped there."""
    z = np.minimum(np.abs(np.asarray(base, dtype=np.float64)), 5.4)
    return 4.0 * EPS[dtype] * np.sqrt(2.0 * np.pi) * np.exp(0.5 * z * z)


def base_tolerance(p, dtype, base_ref):
    """per-row tolerance for base coordinates: relative tolerance + inverse-normal-CDF conditioning of the reference
    for the sub-pdfs whose last-applied stage (layer 0) is an inverse normal CDF"""
    tol = REL_TOL[dtype] * np.maximum(1.0, np.abs(base_ref).max(axis=1))
    extra = np.zeros(base_ref.shape[0])
    for k, layers in enumerate(p.layer_list):
        l0 = layers[0]
        # the base charts of intervals and of S1 are erfinv-based too (interval_base.py:52, sphere_base.py:474)
        if getattr(l0, "inverse_function_type", "isigmoid") in ("inormal_partly_precise", "inormal_partly_crude") \
                or getattr(l0, "code", "") in ("r", "o", "m"):
            b0, b1 = p.base_dim_indices[k]
            extra = np.maximum(extra, icdf_conditioning(base_ref[:, b0:b1], dtype).max(axis=1))
    return tol + extra


def fa_outer_spec(fa):
    """plain-dict description of the outer generator of a fully_amortized_pdf for the oracle"""
    if fa.use_amortizable_mlp:
        return fa.amortization_mlp.structure()
    return dict(linear_indices=[i for i, m in enumerate(fa.amortization_mlp) if isinstance(m, torch.nn.Linear)])
