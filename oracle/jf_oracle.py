"""ORACLE -- test infrastructure only.  CPU restatement (torch CPU tensors, fp64/fp32) of the reference algorithm
for the jammy_flows hot path: log_pdf (target->base, with log-det) and sampling (base->target).

This module is NOT part of the product.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline /
`--impl reference` legs may import it, and only as the checker / the CPU arm -- never as a fallback of the CUDA path.

Parity status: PINNED.  Every function below is checked against golden vectors produced by executing the unmodified
reference (thoglu/jammy_flows v1.1.0) in the build container -- see tests/golden/make_golden.py and
tests/test_oracle_golden.py.  The reference holds no stored numeric fixtures of its own.

The oracle consumes a plain-dict "program" (produced by jammy_flows_b200.pdf.export_program(), no CUDA involved) plus
a dict of parameter tensors named as in the reference state_dict.  It is written in the reference's own arithmetic
style (log-space logsumexp / softplus, masks) so that branch structure and cancellation behaviour match; it does not
share any code with the CUDA kernels, which use a different (linear-space, rescaled) formulation.

Reference line citations are relative to /root/reference/jammy_flows/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# ---------------------------------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------------------------------
def _bounded_log_fn(x, lo, hi, center):
    """log(lo + hi/(1+exp(-x+c))), c = log(hi) if center else 0.
    Reference: layers/euclidean/gaussianization_flow.py:23-47 (generate_log_function_bounded_in_logspace)."""
    ln_max, ln_min = math.log(hi), math.log(lo)
    c = ln_max if center else 0.0
    first = ln_max - torch.logsumexp(torch.stack([torch.zeros_like(x), -x + c], dim=-1), dim=-1)
    return torch.logsumexp(torch.stack([first, torch.full_like(first, ln_min)], dim=-1), dim=-1)


def householder_matrix(vs):
    """Q = prod_i (I - 2 v_i v_i^T / |v_i|^2), accumulated left to right.  vs: [B, n_iter, d] -> [B, d, d].
    Reference: gaussianization_flow.py:457-471, layers/spheres/sphere_base.py:222-240."""
    b, n_iter, d = vs.shape
    eye = torch.eye(d, dtype=vs.dtype).unsqueeze(0)
    q = eye.repeat(b, 1, 1)
    for i in range(n_iter):
        v = vs[:, i, :]
        v = v / v.norm(dim=1, keepdim=True)
        q = torch.bmm(q, eye - 2.0 * v.unsqueeze(2) * v.unsqueeze(1))
    return q


def _safe_angle(x, margin=1e-7):
    """Reference: layers/spheres/sphere_base.py:8-19."""
    return torch.clamp(x, min=margin, max=math.pi - margin)


def _safe_costheta(x, margin=None):
    """Reference: layers/spheres/sphere_base.py:21-38."""
    if margin is None:
        margin = 1e-7 if x.dtype == torch.float32 else 1e-10
    return torch.clamp(x, min=-1.0 + margin, max=1.0 - margin)


# ---------------------------------------------------------------------------------------------------------------------
# Gaussianization-flow layer "g"
# ---------------------------------------------------------------------------------------------------------------------
class GfLayer:
    """One gf_block.  `spec` keys: dim, num_kde, hh_iter, inverse_function_type, fit_normalization,
    regulate_normalization, model_offset, w_min, w_max, n_min, n_max."""

    def __init__(self, spec):
        self.s = spec
        self.d = spec["dim"]
        self.k = spec["num_kde"]

    # -- parameter unpacking: [offset d][hh iter*d][means K*d][log_w K*d][log_n K*d]
    #    Reference: layers/euclidean/euclidean_base.py:34-50, gaussianization_flow.py:699-861
    def unpack(self, p):
        s, d, k = self.s, self.d, self.k
        i = 0
        offset = None
        if s["model_offset"]:
            offset = p[:, :d]
            i = d
        q = None
        if s["hh_iter"] > 0:
            n = s["hh_iter"] * d
            q = householder_matrix(p[:, i:i + n].reshape(-1, s["hh_iter"], d))
            i += n
        means = p[:, i:i + k * d].reshape(-1, k, d)
        i += k * d
        log_w = _bounded_log_fn(p[:, i:i + k * d].reshape(-1, k, d), s["w_min"], s["w_max"], center=True)
        i += k * d
        if s["fit_normalization"]:
            log_n = p[:, i:i + k * d].reshape(-1, k, d)
            i += k * d
            if s["regulate_normalization"]:
                log_n = _bounded_log_fn(log_n, s["n_min"], s["n_max"], center=False)
        else:
            log_n = torch.zeros_like(log_w)
        assert i == p.shape[1], (i, p.shape)
        return offset, q, means, log_w, log_n

    # -- K-logistic mixture in log space.  Reference: gaussianization_flow.py:389-454 (add_skewness=0 branch)
    @staticmethod
    def mixture(x, means, log_w, log_n):
        a = (x.unsqueeze(1) - means) / torch.exp(log_w)
        nrm = log_n - torch.logsumexp(log_n, dim=1, keepdim=True)
        sp = F.softplus(-a)
        log_cdf = torch.logsumexp(-sp + nrm, dim=1)
        log_sf = torch.logsumexp(-a - sp + nrm, dim=1)
        log_pdf = torch.logsumexp(-a - log_w - 2.0 * sp + nrm, dim=1)
        return log_cdf, log_sf, log_pdf

    # -- inverse-CDF stage.  Reference: gaussianization_flow.py:480-560
    def value(self, log_cdf, log_sf):
        t = self.s["inverse_function_type"]
        if t == "isigmoid":
            return log_cdf - log_sf
        eps, a = 0.5e-7, 0.147
        c = 2.0 / (math.pi * a)
        cdf = torch.exp(log_cdf)
        ln_fac = log_cdf + log_sf + math.log(4.0)
        comb = c + ln_fac / 2.0
        pos = 2.0 * (torch.sqrt(comb ** 2 - ln_fac / a) - comb)
        pade = torch.sqrt(torch.clamp(pos, min=0.0))
        if t == "inormal_full_pade":
            return torch.where(cdf <= 0.5, -pade, pade)
        bulk = (cdf > eps) & (cdf < 1.0 - eps)
        good = torch.where(bulk, cdf, torch.full_like(cdf, 0.5))
        ret = math.sqrt(2.0) * torch.erfinv(2.0 * good - 1.0)       # == Normal(0,1).icdf
        if t == "inormal_partly_crude":
            tail = torch.sqrt(-2.0 * (log_sf + log_cdf)) - 0.4717
        else:
            tail = pade
        zero = torch.zeros_like(ret)
        ret = ret + torch.where(cdf >= 1.0 - eps, tail, zero) - torch.where(cdf <= eps, tail, zero)
        return ret

    # -- log d(value)/dx.  Reference: gaussianization_flow.py:568-671
    def log_deriv(self, log_cdf, log_sf, log_pdf):
        t = self.s["inverse_function_type"]
        if t == "isigmoid":
            return torch.logaddexp(-log_sf, -log_cdf) + log_pdf
        eps, a = 0.5e-7, 0.147
        c = 2.0 / (math.pi * a)
        cdf = torch.exp(log_cdf)

        def pade_total():
            ln_fac = log_cdf + log_sf + math.log(4.0)
            f1 = ln_fac / 2.0 + c
            f2 = torch.sqrt(f1 ** 2 - ln_fac / a)
            log_num = torch.log(-(f1 - 1.0 / a - f2))
            log_den = 0.5 * math.log(8.0) + 0.5 * torch.log(f2 - f1) + torch.log(f2)
            sign_fac = torch.log(torch.where(cdf <= 0.5, 1.0 - 2.0 * cdf, -1.0 + 2.0 * cdf))
            tot = log_num - log_den - log_sf - log_cdf + sign_fac
            bad = (cdf > 0.49999) & (cdf < 0.50001)
            return torch.where(bad, torch.full_like(tot, math.log(2.506628)), tot)

        if t == "inormal_full_pade":
            return pade_total() + log_pdf
        bulk = (cdf > eps) & (cdf < 1.0 - eps)
        good = torch.where(bulk, cdf, torch.full_like(cdf, 0.5))
        bulk_val = LOG_SQRT_2PI + torch.erfinv(2.0 * good - 1.0) ** 2 + log_pdf
        if t == "inormal_partly_crude":
            tail = -0.5 * torch.log(-2.0 * (log_cdf + log_sf)) - log_sf - log_cdf
        else:
            tail = pade_total()
        return torch.where(bulk, bulk_val, tail + log_pdf)

    # -- log_pdf direction.  Reference: euclidean_base.py:34-50 + gaussianization_flow.py:995-1057
    def inverse(self, x, log_det, p):
        offset, q, means, log_w, log_n = self.unpack(p)
        if offset is not None:
            x = x - offset
        if q is not None:
            x = torch.einsum("bji,bj->bi", q.expand(x.shape[0], -1, -1), x)       # Q^T x
        lc, ls, lp = self.mixture(x, means, log_w, log_n)
        return self.value(lc, ls), log_det + self.log_deriv(lc, ls, lp).sum(dim=-1)

    # -- sampling direction: 25 bisections on [-1e5,1e5] then <=20 Newton steps, tolerance 1e-14 on the row sum.
    #    Reference: gaussianization_flow.py:911-989 + layers/bisection_n_newton.py:11-135
    def forward(self, z, log_det, p):
        offset, q, means, log_w, log_n = self.unpack(p)
        b = z.shape[0]
        means, log_w, log_n = (t.expand(b, -1, -1) for t in (means, log_w, log_n))

        def f(xx, sel=None):
            m, w, n = (means, log_w, log_n) if sel is None else (means[sel], log_w[sel], log_n[sel])
            lc, ls, lp = self.mixture(xx, m, w, n)
            return self.value(lc, ls), lc, ls, lp

        lo = torch.full_like(z, -1e5)
        hi = torch.full_like(z, 1e5)
        mid = None
        for _ in range(25):
            mid = (hi + lo) / 2.0
            val = f(mid)[0]
            right = val < z
            ok = torch.abs(val - z) <= 1e-6 * torch.abs(z)
            lo = torch.where(ok, mid, torch.where(right, mid, lo))
            hi = torch.where(ok, mid, torch.where(right, hi, mid))
        x = mid
        active = torch.ones(b, dtype=torch.bool)
        for _ in range(20):
            if not bool(active.any()):
                break
            val, lc, ls, lp = f(x[active], active)
            upd = (val - z[active]) / torch.exp(self.log_deriv(lc, ls, lp))
            new = x[active] - upd
            new = torch.where(torch.isfinite(new), new, x[active])
            x = x.clone()
            x[active] = new
            still = torch.abs(upd).sum(dim=1) >= 1e-14
            idx = active.nonzero(as_tuple=True)[0]
            active = active.clone()
            active[idx] = still
        _, lc, ls, lp = f(x)
        log_det = log_det - self.log_deriv(lc, ls, lp).sum(dim=-1)
        if q is not None:
            x = torch.einsum("bij,bj->bi", q.expand(b, -1, -1), x)              # Q x
        if offset is not None:
            x = x + offset
        return x, log_det

    def embedding(self, x):
        return x


# ---------------------------------------------------------------------------------------------------------------------
# S2: charts + "f" layer (Householder rotation in R^3 + von-Mises-Fisher z-scaling), default options
# ---------------------------------------------------------------------------------------------------------------------
def s2_to_embedding(x, log_det):
    """(theta,phi) -> (x,y,z), log_det += log sin(theta).  Reference: sphere_base.py:305-332."""
    theta = _safe_angle(x[:, 0:1])
    phi = x[:, 1:2]
    e = torch.cat([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)], dim=1)
    return e, log_det + torch.log(torch.sin(theta))[:, 0]


def s2_from_embedding(x, log_det):
    """(x,y,z) -> (theta,phi), log_det -= log sin(theta).  Reference: sphere_base.py:266-282."""
    theta = _safe_angle(torch.acos(x[:, 2:3] / torch.sqrt((x ** 2).sum(dim=-1, keepdim=True))))
    log_det = log_det - torch.log(torch.sin(theta))[:, 0]
    arg = torch.clamp(x[:, 0:1] / torch.sqrt((x[:, :2] ** 2).sum(dim=-1, keepdim=True)), min=-1.0, max=1.0)
    phi = torch.acos(arg)
    phi = torch.where(x[:, 1:2] < 0, 2 * math.pi - phi, phi)
    return torch.cat([theta, phi], dim=1), log_det


def s2_sphere_to_plane(x, log_det):
    """Reference: sphere_base.py:496-513 and :416-430."""
    theta = _safe_angle(x[:, 0:1])
    c = _safe_costheta(torch.cos(theta), margin=1e-6)
    r = torch.sqrt(-torch.log((1.0 - c) / 2.0) * 2.0)
    log_det = log_det - torch.log(1.0 - c[:, 0]) + torch.log(torch.sin(theta[:, 0]))
    return torch.cat([r * torch.cos(x[:, 1:2]), r * torch.sin(x[:, 1:2])], dim=1), log_det


def s2_plane_to_sphere(x, log_det):
    """Reference: sphere_base.py:364-408 (in-plane polar) and :569-592."""
    r = torch.sqrt((x ** 2).sum(dim=1, keepdim=True))
    arg = torch.where(r == 0, torch.ones_like(r), x[:, :1] / r)
    ang = torch.acos(arg)
    ang = torch.where(x[:, 1:2] < 0, 2 * math.pi - ang, ang)
    theta = _safe_angle(torch.acos(1.0 - 2.0 * torch.exp(-(r ** 2) / 2.0)))
    log_det = log_det + torch.log(1.0 - torch.cos(theta[:, 0])) - torch.log(torch.sin(theta[:, 0]))
    return torch.cat([theta, ang], dim=1), log_det


class FvmLayer:
    """fisher_von_mises_2d with the sub-flow options off (reference defaults, flow_options.py:154-180).
    `spec` keys: add_rotation, hh_iter, z_sign, min_kappa, first (chart to the plane as first layer of the sub-pdf)."""

    def __init__(self, spec):
        self.s = spec

    def _split(self, p):
        n_hh = self.s["hh_iter"] * 3 if self.s["add_rotation"] else 0
        q = householder_matrix(p[:, :n_hh].reshape(-1, self.s["hh_iter"], 3)) if n_hh > 0 else None
        kappa = torch.exp(p[:, n_hh:n_hh + 1]) + self.s["min_kappa"]          # fvm_2d.py:123
        return q, kappa

    # log_pdf direction.  Reference: sphere_base.py:601-650 + fvm_2d.py:273-500
    def inverse(self, x, log_det, p):
        q, kappa = self._split(p)
        s = self.s["z_sign"]
        if q is not None:
            e, log_det = s2_to_embedding(x, log_det)
            e = torch.einsum("bji,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x, log_det = s2_from_embedding(e, log_det)
        ct = torch.cos(x[:, :1])
        log_det = log_det + torch.log(torch.sin(_safe_angle(x[:, 0])))
        kappa = kappa.expand(x.shape[0], -1)
        safe = torch.where(kappa < 100, torch.log(torch.exp(2 * torch.clamp(kappa, max=100.0)) - 1.0), 2 * kappa)
        log_det = log_det + (torch.log(2 * kappa) + kappa * (s * ct + 1) - safe)[:, 0]
        ret = s * ((1.0 + torch.exp(-2 * kappa) - 2 * torch.exp(kappa * (s * ct - 1))) / (-1 + torch.exp(-2 * kappa)))
        ret = torch.where(kappa < (1e-4 if x.dtype == torch.float32 else 1e-8), ct, ret)
        ret = _safe_costheta(_safe_costheta(ret))
        theta = torch.acos(ret)
        log_det = log_det - torch.log(torch.sin(_safe_angle(theta[:, 0])))
        x = torch.cat([theta, x[:, 1:]], dim=1)
        if self.s["first"]:
            x, log_det = s2_sphere_to_plane(x, log_det)
        return x, log_det

    # sampling direction.  Reference: sphere_base.py:653-695 + fvm_2d.py:502-726
    def forward(self, x, log_det, p):
        q, kappa = self._split(p)
        s = self.s["z_sign"]
        if self.s["first"]:
            x, log_det = s2_plane_to_sphere(x, log_det)
        ct = torch.cos(x[:, :1])
        log_det = log_det + torch.log(torch.sin(_safe_angle(x[:, 0])))
        kappa = kappa.expand(x.shape[0], -1)
        log_det = log_det - torch.log(kappa * s * ct + kappa / torch.tanh(kappa))[:, 0]
        ret = s * (1.0 + (1.0 / kappa) * torch.log(0.5 * (1.0 + s * ct) + (0.5 - 0.5 * s * ct) * torch.exp(-2.0 * kappa)))
        ret = torch.where(kappa < (1e-4 if x.dtype == torch.float32 else 1e-8), ct, ret)
        theta = torch.acos(_safe_costheta(ret))
        log_det = log_det - torch.log(torch.sin(_safe_angle(theta[:, 0])))
        x = torch.cat([theta, x[:, 1:]], dim=1)
        if q is not None:
            e, log_det = s2_to_embedding(x, log_det)
            e = torch.einsum("bij,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x, log_det = s2_from_embedding(e, log_det)
        return x, log_det

    def embedding(self, x):
        """Reference: sphere_base.py:779-794 (the conditioning vector handed to later MLPs is the embedding)."""
        return s2_to_embedding(x, torch.zeros(x.shape[0], dtype=x.dtype))[0]


LAYER_TYPES = {"g": GfLayer, "f": FvmLayer}


# ---------------------------------------------------------------------------------------------------------------------
# pdf-level: autoregressive wiring.  Reference: main/default.py:879-1057 (log_pdf), :1373-1531 + :1533-1707 (sample)
# ---------------------------------------------------------------------------------------------------------------------
class OraclePdf:
    def __init__(self, program, params):
        """program: dict from jammy_flows_b200.pdf.export_program(); params: {state_dict name: tensor/ndarray}."""
        self.prog = program
        self.dtype = getattr(torch, program["dtype"])
        self.params = {k: torch.as_tensor(np.asarray(v)).to(self.dtype) for k, v in params.items()}
        self.subs = []
        for sp in program["subpdfs"]:
            self.subs.append([LAYER_TYPES[ls["code"]](ls) for ls in sp["layers"]])

    def _mlp(self, k, inp):
        """nn.Sequential(Linear, Tanh, ..., Linear).  Reference: main/default.py:654-670."""
        m = self.prog["subpdfs"][k]["mlp"]
        h = inp
        for li, idx in enumerate(m["linear_indices"]):
            w = self.params["mlp_predictors.%d.%d.weight" % (k, idx)]
            b = self.params["mlp_predictors.%d.%d.bias" % (k, idx)]
            h = F.linear(h, w, b)
            if li < len(m["linear_indices"]) - 1:
                h = torch.tanh(h)
        return h

    def _sub_params(self, k, cond, prev_emb, batch):
        sp = self.prog["subpdfs"][k]
        if sp["mlp"] is not None:
            pieces = ([cond] if cond is not None else []) + prev_emb
            return self._mlp(k, torch.cat(pieces, dim=1))
        # permanent parameters: concatenate the layers' tensors in extra_inputs order, broadcast over the batch
        vec = torch.cat([self.params[name].reshape(-1) for name in sp["permanent_param_names"]]) \
            if sp["permanent_param_names"] else torch.zeros(0, dtype=self.dtype)
        return vec.unsqueeze(0)

    def log_pdf(self, x, cond=None):
        x = torch.as_tensor(np.asarray(x)).to(self.dtype)
        cond = None if cond is None else torch.as_tensor(np.asarray(cond)).to(self.dtype)
        b = x.shape[0]
        log_det = torch.zeros(b, dtype=self.dtype)
        prev_emb, base = [], []
        for k, layers in enumerate(self.subs):
            sp = self.prog["subpdfs"][k]
            p = self._sub_params(k, cond, prev_emb, b)
            t0, t1 = sp["target_cols"]
            cur = x[:, t0:t1]
            for li in reversed(range(len(layers))):
                o0, o1 = sp["layer_param_ranges"][li]
                cur, log_det = layers[li].inverse(cur, log_det, p[:, o0:o1])
            base.append(cur)
            prev_emb.append(layers[-1].embedding(x[:, t0:t1]))
        base = torch.cat(base, dim=1)
        logp_base = (-0.5 * base ** 2 - LOG_SQRT_2PI).sum(dim=-1)
        return logp_base + log_det, logp_base, base

    def sample(self, z, cond=None):
        z = torch.as_tensor(np.asarray(z)).to(self.dtype)
        cond = None if cond is None else torch.as_tensor(np.asarray(cond)).to(self.dtype)
        b = z.shape[0]
        log_det = torch.zeros(b, dtype=self.dtype)
        prev_emb, out = [], []
        for k, layers in enumerate(self.subs):
            sp = self.prog["subpdfs"][k]
            p = self._sub_params(k, cond, prev_emb, b)
            b0, b1 = sp["base_cols"]
            cur = z[:, b0:b1]
            for li in range(len(layers)):
                o0, o1 = sp["layer_param_ranges"][li]
                cur, log_det = layers[li].forward(cur, log_det, p[:, o0:o1])
            out.append(cur)
            prev_emb.append(layers[-1].embedding(cur))
        x = torch.cat(out, dim=1)
        logp_base = (-0.5 * z ** 2 - LOG_SQRT_2PI).sum(dim=-1)
        return x, logp_base - log_det, logp_base
