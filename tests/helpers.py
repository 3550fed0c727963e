"""Shared helpers of the test-suite: golden-vector loading and model construction from a golden file."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    params = {k[6:]: g[k] for k in g.files if k.startswith("param/")}
    data = {k: g[k] for k in g.files if not k.startswith("param/") and k != "meta"}
    return meta, params, data


def _opts(meta):
    out = {}
    for k, v in meta["options_overwrite"].items():
        out[k] = v
    return out


def build_pdf(meta, params=None, seed=None):
    """Construct jammy_flows_b200.pdf as the golden's reference model was constructed; optionally load its params."""
    import jammy_flows_b200 as jfb
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    p = jfb.pdf(meta["pdf_defs"], meta["flow_defs"], options_overwrite=_opts(meta),
                conditional_input_dim=meta["conditional_input_dim"])
    p = p.to(getattr(torch, meta["dtype"]))
    if params is not None:
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
        p.load_state_dict(sd)
    return p


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))
