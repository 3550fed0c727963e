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


def s2_rotation_matrix(mode, p):
    """Rotation matrix of the non-Householder modes of the sphere layers.  p: [B, n] rotation parameters.
    Reference: layers/spheres/sphere_base.py:127-216 ("angles": Givens rotations over itertools.combinations(range(3), 2)
    multiplied from the left; "xyz": z axis onto the normalised vector; "quaternion": unnormalised quaternion a,i,j,k)."""
    b = p.shape[0]
    if mode == "angles":
        m = torch.eye(3, dtype=p.dtype).unsqueeze(0).repeat(b, 1, 1)
        for ind, (a_, b_) in enumerate(((0, 1), (0, 2), (1, 2))):
            g = torch.eye(3, dtype=p.dtype).unsqueeze(0).repeat(b, 1, 1)
            g[:, a_, a_] = torch.cos(p[:, ind])
            g[:, b_, b_] = g[:, a_, a_]
            g[:, a_, b_] = torch.sin(p[:, ind])
            g[:, b_, a_] = -g[:, a_, b_]
            m = torch.bmm(g, m)
        return m
    if mode == "xyz":
        n = p[:, :3] / (p[:, :3] ** 2).sum(dim=-1, keepdim=True).sqrt()
        nx, ny, nz = n[:, 0], n[:, 1], n[:, 2]
        m = torch.zeros(b, 3, 3, dtype=p.dtype)
        m[:, 0, 0] = 1.0 - nx ** 2 / (1 + nz); m[:, 1, 1] = 1.0 - ny ** 2 / (1 + nz); m[:, 2, 2] = nz
        m[:, 0, 1] = -nx * ny / (1 + nz); m[:, 1, 2] = ny
        m[:, 1, 0] = -nx * ny / (1 + nz); m[:, 2, 1] = -ny
        m[:, 0, 2] = nx; m[:, 2, 0] = -nx
        return m
    if mode == "quaternion":
        n2 = (p[:, :4] ** 2).sum(dim=-1)
        a, i, j, k = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
        m = torch.zeros(b, 3, 3, dtype=p.dtype)
        m[:, 0, 0] = 1 - 2 * (j ** 2 + k ** 2) / n2; m[:, 1, 1] = 1 - 2 * (i ** 2 + k ** 2) / n2; m[:, 2, 2] = 1 - 2 * (i ** 2 + j ** 2) / n2
        m[:, 0, 1] = 2 * (i * j - a * k) / n2; m[:, 1, 2] = 2 * (j * k - i * a) / n2
        m[:, 1, 0] = 2 * (i * j + a * k) / n2; m[:, 2, 1] = 2 * (j * k + i * a) / n2
        m[:, 0, 2] = 2 * (i * k + j * a) / n2; m[:, 2, 0] = 2 * (i * k - j * a) / n2
        return m
    raise ValueError(mode)


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
def _lower_unit(d, low, upper=False):
    """Unit-diagonal triangular matrix from `low` (sub-diagonals from the bottom-left corner, matrix_fns.py:27-52);
    upper=True transposes it."""
    m = torch.eye(d, dtype=low.dtype).unsqueeze(0).repeat(low.shape[0], 1, 1)
    pos = 0
    for ind in range(d - 1):
        n = ind + 1
        m = m + torch.diag_embed(low[:, pos:pos + n], offset=-(d - 1 - ind))
        pos += n
    return m.permute(0, 2, 1) if upper else m


def _log_one_plus_exp_x_to_a_minus_1(x, a):
    """log(((1+e^x)^a - 1)/(1+e^x)^a), branch by branch as the reference: extra_functions.py:14-61."""
    sp = a * F.softplus(x)
    small = x <= -20
    res = torch.where(small, torch.log(a) + x, torch.zeros_like(x))
    large = sp > 20
    res = torch.where((~small) & large, sp, res)
    tiny = sp < 1e-8
    res = torch.where((~small) & tiny, torch.log(sp), res)
    mid = (~small) & (~large) & (~tiny)
    res = torch.where(mid, torch.log(torch.exp(sp) - 1.0), res)
    return res - sp


class GfLayer:
    """One gf_block.  `spec` keys: dim, num_kde, hh_iter, inverse_function_type, fit_normalization,
    regulate_normalization, model_offset, w_min, w_max, n_min, n_max, rotation_mode, width_mode, width_clamp,
    add_skewness, center_mean, stretch."""

    def __init__(self, spec):
        self.s = spec
        self.d = spec["dim"]
        self.k = spec["num_kde"]

    # -- width regulator variants.  Reference: gaussianization_flow.py:264-317
    def _log_width(self, raw):
        s = self.s
        mode = s.get("width_mode", "smooth")
        clamp = s.get("width_clamp")
        if clamp is not None:
            raw = torch.clamp(raw, min=clamp[0], max=(None if math.isinf(clamp[1]) else clamp[1]))
        if mode == "smooth":
            return _bounded_log_fn(raw, s["w_min"], s["w_max"], center=True)
        if mode == "softplus":
            return torch.log(F.softplus(raw) + s["w_min"])
        return torch.log(torch.exp(raw) + s["w_min"])

    # -- rotation parameters -> (mode, payload).  Reference: gaussianization_flow.py:711-800
    def _rotation(self, p, i):
        s, d = self.s, self.d
        mode = s.get("rotation_mode", "householder")
        if mode == "householder":
            if s["hh_iter"] > 0:
                n = s["hh_iter"] * d
                return ("matrix", householder_matrix(p[:, i:i + n].reshape(-1, s["hh_iter"], d))), i + n
            return None, i
        if mode == "none" or d == 1:
            return None, i
        if mode == "angles":
            n = d * (d - 1) // 2
            ang = p[:, i:i + n]
            eye = torch.eye(d, dtype=p.dtype).unsqueeze(0).repeat(p.shape[0], 1, 1)
            q = eye
            import itertools
            for ind, (a, b) in enumerate(itertools.combinations(range(d), 2)):
                g = eye.clone()
                g[:, a, a] = torch.cos(ang[:, ind])
                g[:, b, b] = g[:, a, a]
                g[:, a, b] = torch.sin(ang[:, ind])
                g[:, b, a] = -g[:, a, b]
                q = torch.bmm(g, q)
            return ("matrix", q), i + n
        if mode == "cayley":
            c = p[:, i:i + 1]
            f = 1.0 / (1.0 + c ** 2)
            q = torch.diag_embed(((1.0 - c ** 2) * f).repeat(1, 2))
            q[:, 0:1, 1:2] = (-2.0 * c * f).unsqueeze(-1)
            q[:, 1:2, 0:1] = (2.0 * c * f).unsqueeze(-1)
            return ("matrix", q), i + 1
        if mode == "triangular_combination":
            npm = d * (d - 1) // 2
            left = p[:, i:i + npm]
            mid = p[:, i + npm:i + npm + d - 1]
            right = p[:, i + npm + d - 1:i + 2 * npm + d - 1]
            return ("tri", (left, mid, right)), i + 2 * npm + d - 1
        raise ValueError(mode)

    def _rotate(self, rot, x, inverse):
        """inverse=True: log_pdf direction (gaussianization_flow.py:1004-1055); False: sampling (:942-987)."""
        if rot is None:
            return x
        kind, pay = rot
        b = x.shape[0]
        if kind == "matrix":
            q = pay.expand(b, -1, -1)
            return torch.einsum("bji,bj->bi", q, x) if inverse else torch.einsum("bij,bj->bi", q, x)
        left, mid, right = pay
        d = self.d
        lm = _lower_unit(d, left).expand(b, -1, -1)
        um = _lower_unit(d, right, upper=True).expand(b, -1, -1)
        diag = torch.cat([mid, -mid.sum(dim=1, keepdim=True)], dim=1)
        if inverse:
            x = torch.linalg.solve_triangular(lm, x.unsqueeze(-1), upper=False, unitriangular=True).squeeze(-1)
            x = x / torch.exp(diag)
            return torch.linalg.solve_triangular(um, x.unsqueeze(-1), upper=True, unitriangular=True).squeeze(-1)
        x = torch.einsum("bij,bj->bi", um, x)
        x = x * torch.exp(diag)
        return torch.einsum("bij,bj->bi", lm, x)

    # -- parameter unpacking: [offset d][rotation][means][log_w][log_n][log_skew]  /  rq_splines layout
    #    Reference: layers/euclidean/euclidean_base.py:34-50, gaussianization_flow.py:699-909
    def unpack(self, p):
        s, d, k = self.s, self.d, self.k
        i = 0
        offset = None
        if s["model_offset"]:
            offset = p[:, :d]
            i = d
        rot, i = self._rotation(p, i)
        if s.get("stretch", "classic") == "rq_splines":
            lw = p[:, i:i + d * k].reshape(-1, d, k)
            i += d * k
            lh = p[:, i:i + d * k].reshape(-1, d, k)
            i += d * k
            ld = p[:, i:i + d * (k + 1)].reshape(-1, d, k + 1)
            i += d * (k + 1)
            bp = p[:, i:i + 4 * d].reshape(-1, d, 4)
            i += 4 * d
            assert i == p.shape[1], (i, p.shape)
            left = bp[:, :, 0:1]
            right = left + torch.exp(bp[:, :, 1:2]) + 0.5
            bottom = bp[:, :, 2:3]
            top = bottom + torch.exp(bp[:, :, 3:4]) + 0.5
            return offset, rot, ("rqs", (lw, lh, ld, left, right, bottom, top))
        cm = int(s.get("center_mean", 0))
        means = p[:, i:i + (k - cm) * d].reshape(-1, k - cm, d)
        i += (k - cm) * d
        log_w = self._log_width(p[:, i:i + k * d].reshape(-1, k, d))
        i += k * d
        if s["fit_normalization"]:
            log_n = p[:, i:i + k * d].reshape(-1, k, d)
            i += k * d
            if s["regulate_normalization"]:
                log_n = _bounded_log_fn(log_n, s["n_min"], s["n_max"], center=False)
        else:
            log_n = torch.zeros_like(log_w)
        log_s = None
        if s.get("add_skewness", 0):
            log_s = _bounded_log_fn(p[:, i:i + k * d].reshape(-1, k, d), 0.1, 9.0, center=True)
            i += k * d
        if cm:
            nw = log_n.exp()
            new_mean = -(means * nw[:, :-1, :]).sum(dim=1, keepdim=True) / nw[:, -1:, :]
            means = torch.cat([means, new_mean], dim=1)
        assert i == p.shape[1], (i, p.shape)
        return offset, rot, ("classic", (means, log_w, log_n, log_s))

    # -- K-logistic mixture in log space.  Reference: gaussianization_flow.py:389-454
    @staticmethod
    def mixture(x, means, log_w, log_n, log_s=None):
        a = (x.unsqueeze(1) - means) / torch.exp(log_w)
        nrm = log_n - torch.logsumexp(log_n, dim=1, keepdim=True)
        if log_s is None:
            sp = F.softplus(-a)
            log_cdf = torch.logsumexp(-sp + nrm, dim=1)
            log_sf = torch.logsumexp(-a - sp + nrm, dim=1)
            log_pdf = torch.logsumexp(-a - log_w - 2.0 * sp + nrm, dim=1)
            return log_cdf, log_sf, log_pdf
        # skewed kernels: the first K//2 are sigmoid(a)^s, the rest mirrored (gaussianization_flow.py:363-372, 417-442)
        k = means.shape[1]
        sgn = torch.ones(1, k, 1, dtype=x.dtype)
        sgn[:, k // 2:, :] = -1.0
        se = torch.exp(log_s)
        log_pdf = torch.logsumexp(-sgn * a - log_w + log_s - (se + 1.0) * F.softplus(-sgn * a) + nrm, dim=1)
        pos = (sgn > 0).expand_as(a)
        lc = torch.where(pos, -se * F.softplus(-a), _log_one_plus_exp_x_to_a_minus_1(a, se.expand_as(a)))
        ls = torch.where(pos, _log_one_plus_exp_x_to_a_minus_1(-a, se.expand_as(a)), -se * F.softplus(a))
        return torch.logsumexp(lc + nrm, dim=1), torch.logsumexp(ls + nrm, dim=1), log_pdf

    # -- rational-quadratic spline with linear tails.  Reference: layers/spline_fns.py:188-358
    @staticmethod
    def rqs_linear_ext(x, pars, inverse):
        lw, lh, ld, left, right, bottom, top = pars
        k = lw.shape[-1]
        x = x.unsqueeze(-1)

        def knots(raw, lo, hi):
            w = 1e-3 + (1.0 - 1e-3 * k) * F.softmax(raw, dim=-1)
            c = F.pad(torch.cumsum(w, dim=-1), pad=(1, 0), mode="constant", value=0.0)
            c = (hi - lo) * c + lo
            return c, c[..., 1:] - c[..., :-1]

        cw, w = knots(lw, left, right)
        ch, h = knots(lh, bottom, top)
        der = 1e-3 + F.softplus(ld)
        b = x.shape[0]
        cw, w, ch, h, der = (t.expand(b, -1, -1) for t in (cw, w, ch, h, der))
        left, right, bottom, top = (t.expand(b, -1, -1) for t in (left, right, bottom, top))
        idx = torch.sum(x >= (ch if inverse else cw), dim=-1, keepdim=True) - 1
        idx = torch.clamp(idx, 0, k - 1)
        g = lambda t: t.gather(-1, idx)
        x0, wk, y0, hk = g(cw), g(w), g(ch), g(h)
        delta = g(h / w)
        d0, d1 = g(der), g(der[..., 1:])
        if inverse:
            dy = x - y0
            qa = dy * (d0 + d1 - 2 * delta) + hk * (delta - d0)
            qb = hk * d0 - dy * (d0 + d1 - 2 * delta)
            qc = -delta * dy
            root = (2 * qc) / (-qb - torch.sqrt(qb.pow(2) - 4 * qa * qc))
            out = root * wk + x0
            tt = root * (1 - root)
            den = delta + (d0 + d1 - 2 * delta) * tt
            num = delta.pow(2) * (d1 * root.pow(2) + 2 * delta * tt + d0 * (1 - root).pow(2))
            lad = -(torch.log(num) - 2 * torch.log(den))
            lo_off = cw[..., 0:1] - ch[..., 0:1] / der[..., 0:1]
            out = torch.where(x <= bottom, x / der[..., 0:1] + lo_off, out)
            hi_off = cw[..., -1:] - ch[..., -1:] / der[..., -1:]
            out = torch.where(x >= top, x / der[..., -1:] + hi_off, out)
            lad = torch.where(x <= bottom, -torch.log(der[..., 0:1]), lad)
            lad = torch.where(x >= top, -torch.log(der[..., -1:]), lad)
        else:
            th = (x - x0) / wk
            tt = th * (1 - th)
            den = delta + (d0 + d1 - 2 * delta) * tt
            out = y0 + hk * (delta * th.pow(2) + d0 * tt) / den
            num = delta.pow(2) * (d1 * th.pow(2) + 2 * delta * tt + d0 * (1 - th).pow(2))
            lad = torch.log(num) - 2 * torch.log(den)
            lo_off = ch[..., 0:1] - cw[..., 0:1] * der[..., 0:1]
            out = torch.where(x <= left, x * der[..., 0:1] + lo_off, out)
            hi_off = ch[..., -1:] - cw[..., -1:] * der[..., -1:]
            out = torch.where(x >= right, x * der[..., -1:] + hi_off, out)
            lad = torch.where(x <= left, torch.log(der[..., 0:1]), lad)
            lad = torch.where(x >= right, torch.log(der[..., -1:]), lad)
        return out.squeeze(-1), lad.squeeze(-1)

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
        offset, rot, (kind, pars) = self.unpack(p)
        if offset is not None:
            x = x - offset
        x = self._rotate(rot, x, inverse=True)
        if kind == "rqs":
            y, lad = self.rqs_linear_ext(x, pars, inverse=False)
            return y, log_det + lad.sum(dim=-1)
        lc, ls, lp = self.mixture(x, *pars)
        return self.value(lc, ls), log_det + self.log_deriv(lc, ls, lp).sum(dim=-1)

    # -- sampling direction: 25 bisections on [-1e5,1e5] then <=20 Newton steps, tolerance 1e-14 on the row sum.
    #    Reference: gaussianization_flow.py:911-989 + layers/bisection_n_newton.py:11-135
    def forward(self, z, log_det, p):
        offset, rot, (kind, pars) = self.unpack(p)
        b = z.shape[0]
        if kind == "rqs":
            x, lad = self.rqs_linear_ext(z, pars, inverse=True)
            log_det = log_det + lad.sum(dim=-1)
            x = self._rotate(rot, x, inverse=False)
            if offset is not None:
                x = x + offset
            return x, log_det
        means, log_w, log_n, log_s = pars
        means, log_w, log_n = (t.expand(b, -1, -1) for t in (means, log_w, log_n))
        if log_s is not None:
            log_s = log_s.expand(b, -1, -1)

        def f(xx, sel=None):
            if sel is None:
                m, w, n, sk = means, log_w, log_n, log_s
            else:
                m, w, n, sk = means[sel], log_w[sel], log_n[sel], (None if log_s is None else log_s[sel])
            lc, ls, lp = self.mixture(xx, m, w, n, sk)
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
        x = self._rotate(rot, x, inverse=False)
        if offset is not None:
            x = x + offset
        return x, log_det

    def embedding(self, x):
        return x


# ---------------------------------------------------------------------------------------------------------------------
# Affine layer "t"
# ---------------------------------------------------------------------------------------------------------------------
class MvnLayer:
    """mvn_block.  `spec` keys: dim, cov_type, model_offset, w_min, w_max.
    Reference: layers/euclidean/multivariate_normal.py:191-272, layers/matrix_fns.py:4-145, euclidean_base.py:34-75."""

    def __init__(self, spec):
        self.s = spec
        self.d = spec["dim"]

    def _lower(self, log_diag, low):
        """Lower-triangular matrix: exp(log_diag) on the diagonal, `low` filled sub-diagonal by sub-diagonal starting at
        the bottom-left corner (matrix_fns.py:27-44)."""
        d = self.d
        m = torch.diag_embed(torch.exp(log_diag))
        pos = 0
        for ind in range(d - 1):
            n = ind + 1
            m = m + torch.diag_embed(low[:, pos:pos + n], offset=-(d - 1 - ind))
            pos += n
        return m

    def _split(self, p):
        s, d = self.s, self.d
        off = None
        if s["model_offset"]:
            off, p = p[:, :d], p[:, d:]
        return off, p

    def inverse(self, x, log_det, p):
        off, p = self._split(p)
        if off is not None:
            x = x - off
        ct = self.s["cov_type"]
        if ct == "identity":
            return x, log_det
        if ct == "diagonal_symmetric":
            lw = _bounded_log_fn(p, self.s["w_min"], self.s["w_max"], center=True)
            return torch.exp(-lw) * x, log_det - self.d * lw.sum(dim=-1)
        lw = _bounded_log_fn(p[:, :self.d], self.s["w_min"], self.s["w_max"], center=True)
        if ct == "diagonal":
            return torch.exp(-lw) * x, log_det - lw.sum(dim=-1)
        m = self._lower(lw, p[:, self.d:]).expand(x.shape[0], -1, -1)
        z = torch.linalg.solve_triangular(m, x.unsqueeze(-1), upper=False).squeeze(-1)
        return z, log_det - lw.sum(dim=-1)

    def forward(self, z, log_det, p):
        off, p = self._split(p)
        ct = self.s["cov_type"]
        if ct == "identity":
            x = z
        elif ct == "diagonal_symmetric":
            lw = _bounded_log_fn(p, self.s["w_min"], self.s["w_max"], center=True)
            x, log_det = torch.exp(lw) * z, log_det + self.d * lw.sum(dim=-1)
        else:
            lw = _bounded_log_fn(p[:, :self.d], self.s["w_min"], self.s["w_max"], center=True)
            if ct == "diagonal":
                x = torch.exp(lw) * z
            else:
                x = torch.einsum("bij,bj->bi", self._lower(lw, p[:, self.d:]).expand(z.shape[0], -1, -1), z)
            log_det = log_det + lw.sum(dim=-1)
        if off is not None:
            x = x + off
        return x, log_det

    def embedding(self, x):
        return x


# ---------------------------------------------------------------------------------------------------------------------
# Rational-quadratic splines (shared by "r", "o" and the sub-flows of "f")
# ---------------------------------------------------------------------------------------------------------------------
def _softplus(t):
    return F.softplus(t)


def _knots(raw, lo, hi, floor, max_ratio):
    """raw [B,n] -> (knot positions [B,n+1], bin sizes [B,n]): softmax with a floor, cumulated, ends pinned to lo/hi and
    the sizes recomputed as knot differences.  Reference: layers/spline_fns.py:77-111 (same in :400-424, :595-616)."""
    n = raw.shape[-1]
    if max_ratio > 0.0:
        ln_max = (math.log(max_ratio) - math.log(n - 1)) / 2.0                       # spline_fns.py:79-84
        raw = 2.0 * torch.sigmoid(raw) * ln_max - ln_max
    frac = floor + (1.0 - floor * n) * torch.softmax(raw, dim=-1)
    cum = F.pad(torch.cumsum(frac, dim=-1), pad=(1, 0), value=0.0)
    cum = (hi - lo) * cum + lo
    cum[..., 0] = lo
    cum[..., -1] = hi
    return cum, cum[..., 1:] - cum[..., :-1]


def _rq_apply(x, kx, w, ky, h, d, inverse):
    """Evaluate the monotone rational-quadratic spline with knots (kx, ky), bin sizes (w, h) and knot derivatives d at
    x [B,1].  Returns (y, log|dy/dx|) -- for inverse=True the value returned is MINUS the forward log-derivative at the
    pre-image, exactly as the reference does.  Reference: spline_fns.py:113-186 (bin search :13-19)."""
    search = (ky if inverse else kx).clone()
    search[..., -1] += 1e-6
    idx = (x >= search).sum(dim=-1, keepdim=True) - 1
    b = x.shape[0]
    ex = lambda t: t.expand(b, -1) if t.shape[0] == 1 else t
    kx, w, ky, h, d = ex(kx), ex(w), ex(ky), ex(h), ex(d)
    g = lambda t: t.gather(-1, idx)
    x0, wk, y0, hk = g(kx), g(w), g(ky), g(h)
    s = g(h / w)
    d0, d1 = g(d), g(d[..., 1:])
    if inverse:
        dy = x - y0
        t = d0 + d1 - 2 * s
        qa = dy * t + hk * (s - d0)
        qb = hk * d0 - dy * t
        qc = -s * dy
        disc = qb.pow(2) - 4 * qa * qc
        assert (disc >= 0).all()
        xi = (2 * qc) / (-qb - torch.sqrt(disc))
        out = xi * wk + x0
    else:
        xi = (x - x0) / wk
        out = None
    xx = xi * (1 - xi)
    den = s + (d0 + d1 - 2 * s) * xx
    if not inverse:
        out = y0 + hk * (s * xi.pow(2) + d0 * xx) / den
    num = s.pow(2) * (d1 * xi.pow(2) + 2 * s * xx + d0 * (1 - xi).pow(2))
    lad = torch.log(num) - 2 * torch.log(den)
    return out, (-lad if inverse else lad)


class Spline:
    """One 1-d spline transformation as configured by an "r" or "o" layer.  `spec` keys: kind (plain | smooth |
    circular), n_bins, n_w, n_h, n_d, fix_first, fix_second, indep, bd_mode (0 derivatives are parameters, 1 boundary
    derivatives fixed to softplus^-1 value bd_fixed, 2 periodic: first copied to the end), lo, hi, min_w, min_h, min_d,
    max_ratio."""

    def __init__(self, spec):
        self.s = spec

    def unpack(self, p):
        """Raw layer parameters -> (raw widths [B,n], raw heights [B,n], raw derivatives or None).
        Reference: layers/intervals/rational_quadratic_spline.py:180-246, layers/spheres/splines_1d.py:111-170."""
        s = self.s
        uw = p[:, :s["n_w"]]
        uh = p[:, s["n_w"]:s["n_w"] + s["n_h"]]
        ud = p[:, s["n_w"] + s["n_h"]:s["n_w"] + s["n_h"] + s["n_d"]] if s["n_d"] > 0 else None
        zero = torch.zeros(p.shape[0], 1, dtype=p.dtype)
        if s["fix_first"]:
            uh = torch.cat([zero, uh], dim=1)
            uw = torch.cat([zero, zero, uw] if s["fix_second"] else [zero, uw], dim=1)
        if s["indep"]:
            uh = uw + uh
        if s["kind"] == "smooth" and s["n_bins"] == 3:
            uw = torch.cat([uw, uw[:, 0:1]], dim=1)
            uh = torch.cat([uh, uh[:, 0:1]], dim=1)
        assert uw.shape[1] == s["n_bins"] and uh.shape[1] == s["n_bins"], (uw.shape, uh.shape, s)
        fixed = torch.full((p.shape[0], 1), float(s["bd_fixed"]), dtype=p.dtype)
        if s["kind"] == "plain":
            if s["bd_mode"] == 1:
                ud = torch.cat([fixed, ud, fixed], dim=-1) if ud is not None else torch.cat([fixed, fixed], dim=-1)
            elif s["bd_mode"] == 2:
                ud = torch.cat([ud, ud[:, 0:1]], dim=-1)
        elif s["kind"] == "smooth":
            if s["bd_mode"] == 1:
                ud = torch.cat([fixed, fixed], dim=-1)
        return uw, uh, ud

    def apply(self, x, p, inverse):
        s = self.s
        uw, uh, ud = self.unpack(p)
        lo, hi = s["lo"], s["hi"]
        kx, w = _knots(uw, lo, hi, s["min_w"], s["max_ratio"])
        ky, h = _knots(uh, lo, hi, s["min_h"], s["max_ratio"])
        if s["kind"] == "plain":
            return _rq_apply(x, kx, w, ky, h, s["min_d"] + _softplus(ud), inverse)
        if s["kind"] == "smooth":
            return _rq_apply(x, kx, w, ky, h, self._smooth_derivs(w, h, s["min_d"] + _softplus(ud)), inverse)
        return self._circular(x, kx, w, ky, h, inverse)

    @staticmethod
    def _smooth_derivs(w, h, bd):
        """Interior knot derivatives that make the second derivative continuous (2 bins, or 3 symmetric bins).
        Reference: spline_fns.py:426-484."""
        n = w.shape[-1]
        if n == 1:
            return bd
        if n == 2:
            h1, h2, w1, w2 = h[..., :1], h[..., 1:], w[..., :1], w[..., 1:]
            hs = h1 + h2
            half = 0.5 * ((h1 / hs) * (h2 / w2 - bd[..., 1:]) + (h2 / hs) * (h1 / w1 - bd[..., :1]))
            q = -(h1 * h2) * ((h1 / hs) * (1.0 / w1 ** 2) + (h2 / hs) * (1.0 / w2 ** 2))
            mid = half + (half ** 2 - q).sqrt()
            return torch.cat([bd[..., :1], mid, bd[..., 1:]], dim=-1)
        assert n == 3
        w1, w2, h1, h2 = w[..., 0:1], w[..., 1:2], h[..., 0:1], h[..., 1:2]
        cd = w1 * w2 * (2 * h1 + h2)
        pp = h2 * (bd[..., :1] * w1 * w2 - h1 * (w1 + w2)) / cd
        q = -h1 * h2 * (h1 * w2 ** 2 + h2 * w1 ** 2) / (cd * w1 * w2)
        mid = -pp / 2.0 + torch.sqrt((-pp / 2.0) ** 2 - q)
        return torch.cat([bd[..., :1], mid, mid, bd[..., 1:]], dim=-1)

    @staticmethod
    def _circular(x, kx, w, ky, h, inverse):
        """Two-bin periodic spline with one common knot derivative and a shift that keeps the map centred.
        Reference: spline_fns.py:618-668 and :727-759."""
        two_pi = 2 * math.pi
        w1, w2, h1, h2 = w[..., :1], w[..., 1:], h[..., :1], h[..., 1:]
        hp, wp = h1 * h2, w1 * w2
        root = torch.sqrt(hp * (8 * ((h2 * w1) ** 2 + (h1 * w2) ** 2) + (9 * (w1 + w2) ** 2 - 16 * wp) * hp))
        res = (hp * (w1 + w2) + root) / (4 * (h1 + h2) * wp)
        d = torch.cat([res, res, res], dim=-1)
        a = -math.pi + w1 / 2.0
        ab = a + w2
        corr = two_pi - (h1 + h2 * a * (a * h1 - res * w1 * ab) / (h1 * w2 ** 2 + 2 * (h1 - res * w1) * a * ab))
        shift_in = corr if inverse else (math.pi - w1 / 2.0)
        u = x - shift_in
        u = torch.where(u < 0.0, u + two_pi, u)
        out, lad = _rq_apply(u, kx, w, ky, h, d, inverse)
        out = out + ((math.pi - w1 / 2.0) if inverse else corr)
        out = torch.where(out > two_pi, out - two_pi, out)
        out = torch.where(x == 0.0, torch.zeros_like(out), out)
        out = torch.where(x == two_pi, torch.full_like(out, two_pi), out)
        return out, lad


# ---------------------------------------------------------------------------------------------------------------------
# interval sub-pdfs: erf chart + "r"
# ---------------------------------------------------------------------------------------------------------------------
class RLayer:
    """rational_quadratic_spline on [lo,hi].  `spec` keys: spline, first, lo, hi.
    Reference: layers/intervals/rational_quadratic_spline.py:180-400, layers/intervals/interval_base.py:33-79."""

    def __init__(self, spec):
        self.s = spec
        self.sp = Spline(spec["spline"])

    def inverse(self, x, log_det, p):
        x = torch.clamp(x, min=-1.0, max=1.0)                       # rational_quadratic_spline.py:297-298 (sic)
        x, lad = self.sp.apply(x, p, True)
        log_det = log_det + lad.sum(dim=-1)
        x = torch.clamp(x, min=-1.0, max=1.0)
        if self.s["first"]:                                         # interval_base.py:47-59
            width = self.s["hi"] - self.s["lo"]
            z = torch.erfinv(2.0 * ((x - self.s["lo"]) / width) - 1.0) * math.sqrt(2.0)
            log_det = log_det - (-(z[:, 0] ** 2) / 2.0 - 0.5 * math.log(2 * math.pi) + math.log(width))
            x = z
        return x, log_det

    def forward(self, x, log_det, p):
        if self.s["first"]:                                         # interval_base.py:33-45
            width = self.s["hi"] - self.s["lo"]
            log_det = log_det - (x[:, 0] ** 2) / 2.0 - 0.5 * math.log(2 * math.pi) + math.log(width)
            x = (0.5 + 0.5 * torch.erf(x / math.sqrt(2.0))) * width + self.s["lo"]
        x = torch.clamp(x, min=-1.0, max=1.0)
        x, lad = self.sp.apply(x, p, False)
        log_det = log_det + lad.sum(dim=-1)
        return torch.clamp(x, min=-1.0, max=1.0), log_det

    def embedding(self, x):
        return x


# ---------------------------------------------------------------------------------------------------------------------
# S1 sub-pdfs: chart, rotation in R^2, "o" and "m"
# ---------------------------------------------------------------------------------------------------------------------
def _safe_angle_2pi(x, margin=1e-7):
    """Reference: spline_fns.py:22-43."""
    return torch.clamp(x, min=margin, max=2 * math.pi - margin)


def s1_from_embedding(e):
    """(x,y) -> angle in [0,2pi].  Reference: sphere_base.py:248-266."""
    a = torch.acos(e[:, 0:1] / torch.sqrt((e ** 2).sum(dim=1, keepdim=True)))
    return torch.where(e[:, 1:2] < 0, 2 * math.pi - a, a)


class S1Layer:
    """Common part of the S1 layers.  Reference: sphere_base.py:601-695 (rotation wrapper), :460-480 / :529-539 chart."""

    def __init__(self, spec):
        self.s = spec

    def _rot(self, p):
        n = self.s["hh_iter"] * 2 if self.s["add_rotation"] else 0
        return (householder_matrix(p[:, :n].reshape(-1, self.s["hh_iter"], 2)) if n > 0 else None), p[:, n:]

    def inverse(self, x, log_det, p):
        q, rest = self._rot(p)
        if q is not None:
            e = torch.cat([torch.cos(x), torch.sin(x)], dim=1)
            e = torch.einsum("bji,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x = s1_from_embedding(e)
        x, log_det = self._inner(x, log_det, rest, logpdf=True)
        if self.s["first"]:
            sign = torch.where(x > math.pi, -1.0, 1.0).to(x.dtype)
            y = torch.where(sign > 0, x, 2 * math.pi - x)
            eps = 1e-5 if x.dtype == torch.float32 else 1e-8
            y = torch.where(y <= 0.0, torch.full_like(y, eps), y)
            y = torch.where(y >= 2 * math.pi, torch.full_like(y, 2 * math.pi - eps), y)
            z = math.sqrt(2.0) * torch.erfinv(1.0 - y / math.pi)
            log_det = log_det - math.log(math.sqrt(2.0 * math.pi)) + (z[:, 0] ** 2) / 2.0
            x = z * sign
        return x, log_det

    def forward(self, x, log_det, p):
        q, rest = self._rot(p)
        if self.s["first"]:
            r = torch.sqrt((x ** 2).sum(dim=1, keepdim=True))
            keep = (x >= 0).to(x.dtype)
            log_det = log_det + math.log(math.sqrt(2.0 * math.pi)) - (r[:, 0] ** 2) / 2.0
            a = math.pi * (1.0 - torch.erf(r / math.sqrt(2.0)))
            x = keep * a + (1.0 - keep) * (2 * math.pi - a)
        x, log_det = self._inner(x, log_det, rest, logpdf=False)
        if q is not None:
            e = torch.cat([torch.cos(x), torch.sin(x)], dim=1)
            e = torch.einsum("bij,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x = s1_from_embedding(e)
        return x, log_det

    def embedding(self, x):
        return torch.cat([torch.cos(x), torch.sin(x)], dim=1)


class OLayer(S1Layer):
    """spline_1d.  Reference: layers/spheres/splines_1d.py:111-306."""

    def __init__(self, spec):
        super().__init__(spec)
        self.sp = Spline(spec["spline"])

    def _inner(self, x, log_det, p, logpdf):
        if logpdf:
            x = _safe_angle_2pi(x)
            x, lad = self.sp.apply(x, p, self.s["natural_direction"] != 0)
            return _safe_angle_2pi(x), log_det + lad.sum(dim=-1)
        x = torch.clamp(x, min=0.0, max=2 * math.pi)
        x, lad = self.sp.apply(x, p, self.s["natural_direction"] == 0)
        return torch.clamp(x, min=0.0, max=2 * math.pi), log_det + lad.sum(dim=-1)


class MLayer(S1Layer):
    """moebius.  `spec` keys: K (num_basis_functions), natural_direction.  Reference: moebius_1d.py:57-259."""

    def _omega(self, pars):
        ll = pars[:, :, 2:3]
        length = 0.001 + torch.exp(math.log(0.999 - 0.001) - torch.logsumexp(
            torch.cat([torch.zeros_like(ll), -ll], dim=2), dim=2, keepdim=True))
        vec = pars[:, :, :2] / (pars[:, :, :2] ** 2).sum(dim=2, keepdim=True).sqrt() * length
        return length, vec

    def trafo(self, x, pars):
        cx, sx = torch.cos(x)[:, None, :], torch.sin(x)[:, None, :]
        length, vec = self._omega(pars)
        ox, oy = vec[:, :, 0:1], vec[:, :, 1:2]
        om = 1.0 - length ** 2
        cmp, smp = np.cos(-np.pi), np.sin(-np.pi)
        opo = 1.0 + length ** 2 - 2 * (cx * ox + sx * oy)
        opo_mp = 1.0 + length ** 2 - 2 * (cmp * ox + smp * oy)
        rot = -math.pi - torch.atan2(om * (smp - oy) - oy * opo_mp, om * (cmp - ox) - ox * opo_mp)
        yv = om * (sx - oy) - oy * opo
        xv = om * (cx - ox) - ox * opo
        xp = torch.cos(rot) * xv - torch.sin(rot) * yv
        yp = torch.sin(rot) * xv + torch.cos(rot) * yv
        at = torch.atan2(yp, xp)[:, :, -1:] + math.pi
        ln = pars[:, :, 3:4]
        return torch.sum(at * torch.exp(ln - torch.logsumexp(ln, dim=1, keepdim=True)), dim=1) - math.pi

    def deriv(self, x, pars):
        cx, sx = torch.cos(x)[:, None, :], torch.sin(x)[:, None, :]
        length, vec = self._omega(pars)
        om = 1.0 - length ** 2
        opo = 1.0 + length ** 2 - 2 * (cx * vec[:, :, 0:1] + sx * vec[:, :, 1:2])
        ln = pars[:, :, 3:4]
        return torch.exp(torch.logsumexp(torch.log(om / opo) + ln - torch.logsumexp(ln, dim=1, keepdim=True), dim=1))

    def _solve(self, target, pars):
        """20 bisections on [-pi,pi] + <=20 Newton steps.  Reference: bisection_n_newton.py:137-256."""
        b = target.shape[0]
        pars = pars.expand(b, -1, -1)
        lo = torch.full_like(target, -math.pi)
        hi = torch.full_like(target, math.pi)
        mid = None
        for _ in range(20):
            mid = (hi + lo) / 2.0
            val = self.trafo(mid, pars)
            right = val < target
            ok = torch.abs(val - target) <= 1e-6 * torch.abs(target)
            lo = torch.where(ok, mid, torch.where(right, mid, lo))
            hi = torch.where(ok, mid, torch.where(right, hi, mid))
        x = mid
        active = torch.ones(b, dtype=torch.bool)
        for _ in range(20):
            if not bool(active.any()):
                break
            upd = (self.trafo(x[active], pars[active]) - target[active]) / self.deriv(x[active], pars[active])
            x = x.clone()
            x[active] = x[active] - upd
            idx = active.nonzero(as_tuple=True)[0]
            active = active.clone()
            active[idx] = torch.abs(upd).sum(dim=1) >= 1e-14
        return x

    def _inner(self, x, log_det, p, logpdf):
        pars = p.reshape(p.shape[0], self.s["K"], 4)
        x = torch.where(x > math.pi, x - 2 * math.pi, x)
        direct = (self.s["natural_direction"] == 0) if logpdf else (self.s["natural_direction"] != 0)
        if direct:
            ld = torch.log(self.deriv(x, pars)).sum(dim=-1)
            x = self.trafo(x, pars)
        else:
            x = self._solve(x, pars)
            ld = -torch.log(self.deriv(x, pars)).sum(dim=-1)
        x = torch.where(x < 0, 2 * math.pi + x, x)
        return x, log_det + ld


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
    """fisher_von_mises_2d (reference flow_options.py:154-180), with the optional vertical ("r...") and circular
    ("o...") spline sub-flows.  `spec` keys: add_rotation, hh_iter, z_sign, min_kappa, first (chart to the plane as
    first layer of the sub-pdf), vertical / circular (lists of spline specs; parameters follow kappa in that order)."""

    def __init__(self, spec):
        self.s = spec
        self.vertical = [Spline(v) for v in spec.get("vertical", [])]
        self.circular = [Spline(v) for v in spec.get("circular", [])]

    @staticmethod
    def _window(c):
        """Reference: fvm_2d.py:267-271."""
        return torch.where(c <= 0, 6 * c ** 5 + 15 * c ** 4 + 10 * c ** 3 + 1.0, -6 * c ** 5 + 15 * c ** 4 - 10 * c ** 3 + 1.0)

    @staticmethod
    def _extra_rotation(ret, angle, log_det, transpose):
        """add_extra_rotation_inbetween: cylinder -> angles -> embedding -> fixed matrix -> angles -> cylinder."""
        th = torch.acos(ret)
        log_det = log_det - torch.log(torch.sin(_safe_angle(th[:, 0])))
        e, log_det = s2_to_embedding(torch.cat([th, angle], dim=1), log_det)
        m = torch.tensor([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]], dtype=ret.dtype)
        e = torch.einsum("ij,bj->bi", m.t() if transpose else m, e)
        comb, log_det = s2_from_embedding(e, log_det)
        log_det = log_det + torch.log(torch.sin(_safe_angle(comb[:, 0])))
        return torch.cos(comb[:, :1]), comb[:, 1:], log_det

    def _n_rot(self):
        if not self.s["add_rotation"]:
            return 0
        mode = self.s.get("rotation_mode", "householder")
        return {"householder": self.s["hh_iter"] * 3, "angles": 3, "xyz": 3, "quaternion": 4}[mode]

    def _sub_params(self, p):
        n_hh = self._n_rot()
        o = n_hh + (1 if self.s.get("kappa_mode", 0) <= 2 else 0)
        nv = sum(v.s["n_w"] + v.s["n_h"] + v.s["n_d"] for v in self.vertical)
        nc = sum(v.s["n_w"] + v.s["n_h"] + v.s["n_d"] for v in self.circular)
        return p[:, o:o + nv], p[:, o + nv:o + nv + nc]

    @staticmethod
    def _chain(splines, x, log_det, p, logpdf, angle):
        """Pass-through chain of nested layers (main/default.py:998-1031 reversed / :1482-1506 in order)."""
        offs, o = [], 0
        for sp in splines:
            n = sp.s["n_w"] + sp.s["n_h"] + sp.s["n_d"]
            offs.append((o, o + n))
            o += n
        order = reversed(range(len(splines))) if logpdf else range(len(splines))
        for i in order:
            sp, (a, b) = splines[i], offs[i]
            if angle:       # "o": splines_1d.py:111-306
                if logpdf:
                    x = _safe_angle_2pi(x)
                    x, lad = sp.apply(x, p[:, a:b], sp.s["natural_direction"] != 0)
                    x = _safe_angle_2pi(x)
                else:
                    x = torch.clamp(x, min=0.0, max=2 * math.pi)
                    x, lad = sp.apply(x, p[:, a:b], sp.s["natural_direction"] == 0)
                    x = torch.clamp(x, min=0.0, max=2 * math.pi)
            else:           # "r": rational_quadratic_spline.py:180-400
                x = torch.clamp(x, min=-1.0, max=1.0)
                x, lad = sp.apply(x, p[:, a:b], logpdf)
                x = torch.clamp(x, min=-1.0, max=1.0)
            log_det = log_det + lad.sum(dim=-1)
        return x, log_det

    def _split(self, p):
        n_hh = self._n_rot()
        mode = self.s.get("rotation_mode", "householder")
        q = None
        if n_hh > 0:
            q = (householder_matrix(p[:, :n_hh].reshape(-1, self.s["hh_iter"], 3)) if mode == "householder"
                 else s2_rotation_matrix(mode, p[:, :n_hh]))
        # concentration: fvm_2d.py:108-138 (link functions) and :289-330 (norm of the rotation parameters)
        km, clamp = self.s.get("kappa_mode", 0), self.s.get("kappa_clamping", 0)
        raw = p[:, n_hh:n_hh + 1]
        if km == 0:
            kappa = torch.exp(torch.clamp(raw, min=-5.0) if clamp else raw) + self.s["min_kappa"]
        elif km == 1:
            kappa = F.softplus(torch.clamp(raw, min=-5.0) if clamp else raw) + self.s["min_kappa"]
        elif km == 2:
            sp = F.softplus(raw)
            kappa = ((torch.clamp(sp, min=-5.0) if clamp else sp) + math.log(self.s["min_kappa"])).exp()
        elif km == 3:
            kappa = (p[:, :3] ** 2).sum(dim=-1, keepdim=True).sqrt()
        elif km == 4:
            kappa = (p[:, :3] ** 2).sum(dim=-1, keepdim=True)
        elif km == 5:
            kappa = (p[:, 1:4] ** 2).sum(dim=-1, keepdim=True).sqrt()
        else:
            kappa = (p[:, 1:4] ** 2).sum(dim=-1, keepdim=True)
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
        ret = _safe_costheta(ret)
        angle = x[:, 1:]
        if self.s.get("extra_rotation", 0):                                     # fvm_2d.py:381-402
            ret, angle, log_det = self._extra_rotation(ret, angle, log_det, True)
        pv, pc = self._sub_params(p)
        region = self.s.get("identity_region", 0.0)
        # fvm_2d.py:404-470: with an identity region only the rows inside (-1 + region, 1 - region) pass the sub-flows
        mask = (torch.ones_like(ret[:, 0], dtype=torch.bool) if region == 0.0
                else ((ret > -1.0 + region) & (ret < 1.0 - region))[:, 0])
        if len(self.circular) > 0 and bool(mask.any()):                         # fvm_2d.py:413-427
            a2, l2 = self._chain(self.circular, angle[mask], log_det[mask],
                                 (pc.expand(x.shape[0], -1) * self._window(ret))[mask], True, True)
            angle, log_det = angle.clone(), log_det.clone()
            angle[mask], log_det[mask] = a2, l2
        if len(self.vertical) > 0 and bool(mask.any()):                         # fvm_2d.py:430-432
            r2, l2 = self._chain(self.vertical, ret[mask], log_det[mask], pv.expand(x.shape[0], -1)[mask], True, False)
            ret, log_det = ret.clone(), log_det.clone()
            ret[mask], log_det[mask] = r2, l2
        ret = _safe_costheta(ret)
        theta = torch.acos(ret)
        log_det = log_det - torch.log(torch.sin(_safe_angle(theta[:, 0])))
        x = torch.cat([theta, angle], dim=1)
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
        angle = x[:, 1:]
        pv, pc = self._sub_params(p)
        region = self.s.get("identity_region", 0.0)
        mask = (torch.ones_like(ct[:, 0], dtype=torch.bool) if region == 0.0
                else ((ct > -1.0 + region) & (ct < 1.0 - region))[:, 0])               # fvm_2d.py:612 (before the vertical flow)
        if len(self.vertical) > 0 and bool(mask.any()):                         # fvm_2d.py:591-592
            c2, l2 = self._chain(self.vertical, ct[mask], log_det[mask], pv.expand(x.shape[0], -1)[mask], False, False)
            ct, log_det = ct.clone(), log_det.clone()
            ct[mask], log_det[mask] = c2, l2
        if len(self.circular) > 0 and bool(mask.any()):                         # fvm_2d.py:595-607
            a2, l2 = self._chain(self.circular, angle[mask], log_det[mask],
                                 (pc.expand(x.shape[0], -1) * self._window(ct))[mask], False, True)
            angle, log_det = angle.clone(), log_det.clone()
            angle[mask], log_det[mask] = a2, l2
        if self.s.get("extra_rotation", 0):                                     # fvm_2d.py:664-688
            ct, angle, log_det = self._extra_rotation(ct, angle, log_det, False)
        kappa = kappa.expand(x.shape[0], -1)
        log_det = log_det - torch.log(kappa * s * ct + kappa / torch.tanh(kappa))[:, 0]
        ret = s * (1.0 + (1.0 / kappa) * torch.log(0.5 * (1.0 + s * ct) + (0.5 - 0.5 * s * ct) * torch.exp(-2.0 * kappa)))
        ret = torch.where(kappa < (1e-4 if x.dtype == torch.float32 else 1e-8), ct, ret)
        theta = torch.acos(_safe_costheta(ret))
        log_det = log_det - torch.log(torch.sin(_safe_angle(theta[:, 0])))
        x = torch.cat([theta, angle], dim=1)
        if q is not None:
            e, log_det = s2_to_embedding(x, log_det)
            e = torch.einsum("bij,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x, log_det = s2_from_embedding(e, log_det)
        return x, log_det

    def embedding(self, x):
        """Reference: sphere_base.py:779-794 (the conditioning vector handed to later MLPs is the embedding)."""
        return s2_to_embedding(x, torch.zeros(x.shape[0], dtype=x.dtype))[0]


# ---------------------------------------------------------------------------------------------------------------------
# S2: exponential-map flow "v" (exponential potential)
# ---------------------------------------------------------------------------------------------------------------------
class VLayer:
    """exponential_map_s2 with exp_map_type="exponential", mean_parametrization="old".
    `spec` keys: add_rotation, hh_iter, first, natural_direction, K, max_iter.
    Reference: layers/spheres/exponential_map_s2.py:248-442 (map and Jacobian), :446-528 (directions),
    layers/bisection_n_newton.py:394-465 (inverse), rotation/chart wrapper sphere_base.py:601-695."""

    def __init__(self, spec):
        self.s = spec

    def _split(self, p):
        n_hh = self.s["hh_iter"] * 3 if self.s["add_rotation"] else 0
        q = householder_matrix(p[:, :n_hh].reshape(-1, self.s["hh_iter"], 3)) if n_hh > 0 else None
        kind = self.s.get("exp_map_type", "exponential")
        n_pot = 5 if kind == "exponential" else (35 if kind == "splines" else 4)
        return q, p[:, n_hh:].reshape(p.shape[0], n_pot, self.s["K"])

    @staticmethod
    def _log_map(base, g, jac_g):
        """Unit tangent at `base` towards g, the projection of g on it, and both Jacobians w.r.t. base (:153-214)."""
        tn = (g ** 2).sum(dim=1, keepdim=True).sqrt()
        nh = g / tn
        ca = (nh * base).sum(dim=1, keepdim=True)
        alpha = torch.arccos(ca)
        sa = torch.sin(alpha)
        th = (nh - base * ca) / sa
        proj = (g * th).sum(dim=1, keepdim=True)
        d_t_base = torch.diag_embed((-ca / sa).repeat(1, 3))
        d_t_theta = ((base - nh * ca) / sa ** 2).unsqueeze(-1)
        inv_sq = -1.0 / torch.sqrt(1.0 - ca ** 2)
        jt = d_t_base + d_t_theta @ (inv_sq * nh).unsqueeze(1)
        jp = (jt * g.unsqueeze(-1)).sum(dim=1, keepdim=True)
        d_theta_norm = (inv_sq * base).unsqueeze(1)
        d_norm_unnorm = (-g / tn ** 2).unsqueeze(-1) @ nh.unsqueeze(1) + torch.diag_embed(1.0 / tn.repeat(1, 3))
        d_t_norm = torch.diag_embed(1.0 / sa.repeat(1, 3))
        jt = jt + d_t_theta @ d_theta_norm @ d_norm_unnorm @ jac_g
        jt = jt + d_t_norm @ d_norm_unnorm @ jac_g
        jp = jp + (th.unsqueeze(-1) * jac_g).sum(dim=1, keepdim=True)
        return th, proj, jt, jp

    def map(self, x, pars):
        """-> (y, (J B)^T (J B) [B,2,2], J [B,3,3]).  Reference :248-442."""
        norm = (pars[:, :3, :] ** 2).sum(dim=1, keepdim=True).sqrt()
        mu = pars[:, :3, :] / norm
        fake = -torch.log(1.0 + (math.e - 1.0) * torch.exp(-norm / 10.0)) + 1.0
        lw = pars[:, 3:4, :] - torch.logsumexp(pars[:, 3:4, :], dim=2, keepdim=True) + fake.log()
        w = lw.exp()
        xm = (x[:, :, None] * mu).sum(dim=1, keepdim=True)
        kind = self.s.get("exp_map_type", "exponential")
        if kind == "exponential":                                 # :285-314
            beta = pars[:, 4:5, :].exp()
            e = torch.exp(beta * (xm - 1.0))
            g = (w * mu * e).sum(dim=-1)
            jac_g = torch.einsum("biu,bju->bij", beta * w * mu * e, mu)
        elif kind == "splines":                                   # :346-388: the potential is the integral of a spline
            nb, K = 10, self.s["K"]
            b = max(x.shape[0], pars.shape[0])
            sp = Spline(dict(kind="plain", n_bins=nb, n_w=nb, n_h=nb, n_d=nb + 1, fix_first=0, fix_second=0, indep=0,
                             bd_mode=0, bd_fixed=0.0, lo=-1.0, hi=1.0, min_w=1e-3, min_h=1e-3, min_d=1e-3, max_ratio=-1.0))
            res, der = [], []
            for k in range(K):
                pk = pars[:, 4:4 + 3 * nb + 1, k].expand(b, -1)
                r_k, lad_k = sp.apply(xm[:, 0, k:k + 1].expand(b, -1), pk, False)
                res.append(r_k)
                der.append(lad_k.exp())
            res, der = torch.stack(res, dim=2), torch.stack(der, dim=2)          # [B, 1, K]
            g = (w * mu * res).sum(dim=-1)
            jac_g = torch.einsum("biu,bju->bij", w * mu * der, mu.expand(b, -1, -1))
        elif kind == "linear":                                    # :315-324 (no Jacobian of the gradient)
            g = (w * mu).sum(dim=-1).expand(x.shape[0], -1)
            jac_g = torch.zeros(x.shape[0], 3, 3, dtype=x.dtype)
        else:                                                     # quadratic, :325-343
            g = (w * mu * xm).sum(dim=-1)
            jac_g = torch.einsum("biu,bju->bij", (w * mu).expand(x.shape[0], -1, -1), mu.expand(x.shape[0], -1, -1))
        th, proj, jt, jp = self._log_map(x, g, jac_g)
        y = x * torch.cos(proj) + th * torch.sin(proj)
        first = torch.diag_embed(torch.cos(proj).repeat(1, 3)) + (-x * torch.sin(proj)).unsqueeze(-1) @ jp
        second = jt * torch.sin(proj.unsqueeze(-1)) + (th * torch.cos(proj)).unsqueeze(-1) @ jp
        jac = first + second
        basis = torch.cat([th.unsqueeze(2), torch.cross(x, th, dim=1).unsqueeze(2)], dim=2)
        pj = torch.bmm(jac, basis)
        return y, torch.bmm(pj.permute(0, 2, 1), pj), jac

    def _solve(self, target, pars):
        """Damped descent on the sphere: start (0,0,-1), step 0.4 x Newton length, stop at 1e-12.  Reference
        bisection_n_newton.py:394-465 with exponential_map_s2.py:216-246 as the tangent finder."""
        b = target.shape[0]
        pars = pars.expand(b, -1, -1)
        prev = torch.zeros_like(target)
        prev[:, 2] = -1.0
        active = torch.ones(b, dtype=torch.bool)
        for _ in range(self.s["max_iter"]):
            y, _, jac = self.map(prev[active], pars[active])
            tg = target[active]
            fn = -(y * tg).sum(dim=-1, keepdim=True) + 1.0
            res = -torch.bmm(jac.permute(0, 2, 1), tg.unsqueeze(2)).squeeze(-1)
            gn = (res ** 2).sum(dim=1, keepdim=True).sqrt()
            base, tdir = prev[active], -(res / gn)
            ca = (tdir * base).sum(dim=1, keepdim=True)
            conv = ca >= 1
            alt = torch.zeros_like(base)
            alt[:, 0] = 1.0
            ca = torch.where(conv, (tdir * alt).sum(dim=1, keepdim=True), ca)
            alpha = torch.arccos(ca)
            ub = torch.where(conv, alt, base)
            vs = (tdir - ub * ca) / torch.sin(alpha)
            alpha = torch.where(conv, torch.zeros_like(alpha), alpha)
            step = -(fn / (vs * res).sum(dim=1, keepdim=True))
            step = torch.where(alpha == 0, torch.zeros_like(step), step)
            prev = prev.clone()
            prev[active] = base * torch.cos(0.4 * step) + vs * torch.sin(0.4 * step)
            idx = active.nonzero(as_tuple=True)[0]
            active = active.clone()
            active[idx] = torch.abs(step[:, 0]) >= 1e-12
            if not bool(active.any()):
                break
        return prev

    def _inner(self, x, log_det, pars, logpdf):
        e, log_det = s2_to_embedding(x, log_det)
        direct = (self.s["natural_direction"] == 0) if logpdf else (self.s["natural_direction"] != 0)
        if direct:
            y, m2, _ = self.map(e, pars.expand(e.shape[0], -1, -1))
            log_det = log_det + 0.5 * torch.slogdet(m2)[1]
        else:
            y = self._solve(e, pars)
            _, m2, _ = self.map(y, pars.expand(e.shape[0], -1, -1))
            log_det = log_det - 0.5 * torch.slogdet(m2)[1]
        return s2_from_embedding(y, log_det)

    def inverse(self, x, log_det, p):
        q, pars = self._split(p)
        if q is not None:
            e, log_det = s2_to_embedding(x, log_det)
            e = torch.einsum("bji,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x, log_det = s2_from_embedding(e, log_det)
        x, log_det = self._inner(x, log_det, pars, True)
        if self.s["first"]:
            x, log_det = s2_sphere_to_plane(x, log_det)
        return x, log_det

    def forward(self, x, log_det, p):
        q, pars = self._split(p)
        if self.s["first"]:
            x, log_det = s2_plane_to_sphere(x, log_det)
        x, log_det = self._inner(x, log_det, pars, False)
        if q is not None:
            e, log_det = s2_to_embedding(x, log_det)
            e = torch.einsum("bij,bj->bi", q.expand(e.shape[0], -1, -1), e)
            x, log_det = s2_from_embedding(e, log_det)
        return x, log_det

    def embedding(self, x):
        return s2_to_embedding(x, torch.zeros(x.shape[0], dtype=x.dtype))[0]


LAYER_TYPES = {"t": MvnLayer, "v": VLayer, "g": GfLayer, "f": FvmLayer, "r": RLayer, "o": OLayer, "m": MLayer}


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
        if m.get("custom", False):
            return self._custom_mlp(m, self.params["mlp_predictors.%d.u_v_b_pars" % k], inp)
        h = inp
        for li, idx in enumerate(m["linear_indices"]):
            w = self.params["mlp_predictors.%d.%d.weight" % (k, idx)]
            b = self.params["mlp_predictors.%d.%d.bias" % (k, idx)]
            h = F.linear(h, w, b)
            if li < len(m["linear_indices"]) - 1:
                h = torch.tanh(h)
        return h

    @staticmethod
    def _custom_mlp(m, flat, x):
        """AmortizableMLP with permanent parameters.  Reference: amortizable_mlp.py:508-682 (U (V^T x) for factorised
        layers, connectivity modes 0-4)."""
        if flat.dim() == 2 and m.get("amortised", False):
            return OraclePdf._custom_mlp_rows(m, flat, x)
        flat = flat.reshape(-1)

        def chain(spec, pars, inp):
            h, pos = inp, 0
            n = len(spec["layers"])
            for i, l in enumerate(spec["layers"]):
                u = pars[pos:pos + l["n_u"]]
                pos += l["n_u"]
                v = pars[pos:pos + l["n_v"]]
                pos += l["n_v"]
                b = pars[pos:pos + l["n_b"]]
                pos += l["n_b"]
                if l["full"]:
                    h = torch.einsum("ij,bj->bi", u.view(l["n_out"], l["n_in"]), h)
                else:
                    mid = torch.einsum("ij,bj->bi", v.view(l["rank"], l["n_in"]), h)
                    h = torch.einsum("ij,bj->bi", u.view(l["n_out"], l["rank"]), mid)
                if l["n_b"] > 0:
                    h = h + b
                if i < n - 1:
                    h = torch.tanh(h)
            return h

        mode = m["highway_mode"]
        prev = 0.0
        if mode > 0:
            hw = m["highway"]
            prev = chain(hw, flat[-hw["num_params"]:], x)
        pos = 0
        nxt = x
        for ci, ch in enumerate(m["chains"]):
            nonlinear = chain(ch, flat[pos:pos + ch["num_params"]], x if ci == 0 else nxt)
            pos += ch["num_params"]
            if mode == 3:
                nxt = prev + nonlinear
            elif mode == 4:
                nxt = torch.cat([x, prev + nonlinear], dim=1)
            else:
                nxt = x
            prev = prev + nonlinear
        return prev

    @staticmethod
    def _custom_mlp_rows(m, rows, x):
        """AmortizableMLP "being amortised": row b applies the network encoded by rows[b] (per-row weights).
        Reference: amortizable_mlp.py:470-585 (`_adaptive_matmul`: bmm of the [B, out, in] view with the input) and
        :586-682; same layout and connectivity as the permanent-parameter case above."""
        def chain(spec, pars, inp):
            h, pos = inp, 0
            n = len(spec["layers"])
            for i, l in enumerate(spec["layers"]):
                u = pars[:, pos:pos + l["n_u"]]
                pos += l["n_u"]
                v = pars[:, pos:pos + l["n_v"]]
                pos += l["n_v"]
                b = pars[:, pos:pos + l["n_b"]]
                pos += l["n_b"]
                if l["full"]:
                    h = torch.einsum("bij,bj->bi", u.reshape(-1, l["n_out"], l["n_in"]), h)
                else:
                    mid = torch.einsum("bij,bj->bi", v.reshape(-1, l["rank"], l["n_in"]), h)
                    h = torch.einsum("bij,bj->bi", u.reshape(-1, l["n_out"], l["rank"]), mid)
                if l["n_b"] > 0:
                    h = h + b
                if i < n - 1:
                    h = torch.tanh(h)
            return h

        mode = m["highway_mode"]
        prev = 0.0
        if mode > 0:
            hw = m["highway"]
            prev = chain(hw, rows[:, -hw["num_params"]:], x)
        pos = 0
        nxt = x
        for ci, ch in enumerate(m["chains"]):
            nonlinear = chain(ch, rows[:, pos:pos + ch["num_params"]], x if ci == 0 else nxt)
            pos += ch["num_params"]
            if mode == 3:
                nxt = prev + nonlinear
            elif mode == 4:
                nxt = torch.cat([x, prev + nonlinear], dim=1)
            else:
                nxt = x
            prev = prev + nonlinear
        return prev

    def _split_amortization(self, amort):
        """amortization_parameters [B, T] -> per sub-pdf column block (flow parameters of a first sub-pdf without
        generator, else the flat vector of its generator), in sub-pdf order.  Reference: main/default.py:925-991."""
        self._amort = None
        if amort is None:
            return
        amort = torch.as_tensor(np.asarray(amort)).to(self.dtype)
        out, pos = [], 0
        for sp in self.prog["subpdfs"]:
            m = sp["mlp"]
            if m is not None:
                n = sum(c["num_params"] for c in m["chains"]) + (m["highway"]["num_params"] if m["highway"] else 0)
            else:
                n = sp["layer_param_ranges"][-1][1] if len(sp["layer_param_ranges"]) else 0
            out.append(amort[:, pos:pos + n])
            pos += n
        assert pos == amort.shape[1], (pos, amort.shape)
        self._amort = out

    def _sub_params(self, k, cond, prev_emb, batch):
        sp = self.prog["subpdfs"][k]
        am = getattr(self, "_amort", None)
        if am is not None:
            if sp["mlp"] is None:
                return am[k]
            pieces = ([cond] if cond is not None else []) + prev_emb
            return self._custom_mlp(sp["mlp"], am[k], torch.cat(pieces, dim=1))
        if sp["mlp"] is not None:
            if isinstance(cond, (list, tuple)):          # one conditional input per sub-pdf (main/default.py:944-949)
                cond = cond[k]
            pieces = ([cond] if cond is not None else []) + prev_emb
            return self._mlp(k, torch.cat(pieces, dim=1))
        # permanent parameters: concatenate the layers' tensors in extra_inputs order, broadcast over the batch
        vec = torch.cat([self.params[name].reshape(-1) for name in sp["permanent_param_names"]]) \
            if sp["permanent_param_names"] else torch.zeros(0, dtype=self.dtype)
        return vec.unsqueeze(0)

    def transform_target(self, x, to_embedding):
        """intrinsic <-> embedding charts of all sub-pdfs; returns (new x, chart log_det).
        Reference: main/default.py:1737-1813, sphere_base.py:242-332, :796-841 (e / i sub-pdfs: identity)."""
        x = torch.as_tensor(np.asarray(x)).to(self.dtype)
        log_det = torch.zeros(x.shape[0], dtype=self.dtype)
        out, col = [], 0
        for sp in self.prog["subpdfs"]:
            d = sp["target_cols"][1] - sp["target_cols"][0]
            sphere = sp["manifold"] == "s"
            n_in = d + (1 if (sphere and not to_embedding) else 0)
            cur = x[:, col:col + n_in]
            col += n_in
            if sphere and d == 2:
                cur, log_det = s2_to_embedding(cur, log_det) if to_embedding else s2_from_embedding(cur, log_det)
            elif sphere and d == 1:
                cur = torch.cat([torch.cos(cur), torch.sin(cur)], dim=1) if to_embedding else s1_from_embedding(cur)
            out.append(cur)
        return torch.cat(out, dim=1), log_det

    def log_pdf_embedding(self, x_emb, cond=None):
        """pdf.forward(..., force_embedding_coordinates=True).  Reference: main/default.py:906-909."""
        x, ld = self.transform_target(x_emb, to_embedding=False)
        logp, logp_base, base = self.log_pdf(x, cond)
        return logp + ld, logp_base, base

    def sample_embedding(self, z, cond=None):
        """pdf._obtain_sample(..., force_embedding_coordinates=True).  Reference: main/default.py:1522-1524."""
        x, logp, logp_base = self.sample(z, cond)
        xe, ld = self.transform_target(x, to_embedding=True)
        return xe, logp - ld, logp_base

    def subpdf_log_pdf(self, k, x_k, cond, prev_emb):
        """log p_k(x_k | cond, embedded earlier targets) of one sub-pdf (default coordinates).
        Reference: one iteration of main/default.py:2783-2870 (all_layer_inverse_individual_subdims)."""
        sp, layers = self.prog["subpdfs"][k], self.subs[k]
        b = x_k.shape[0]
        p = self._sub_params(k, cond, prev_emb, b)
        log_det = torch.zeros(b, dtype=self.dtype)
        cur = x_k
        for li in reversed(range(len(layers))):
            o0, o1 = sp["layer_param_ranges"][li]
            cur, log_det = layers[li].inverse(cur, log_det, p[:, o0:o1])
        return (-0.5 * cur ** 2 - LOG_SQRT_2PI).sum(dim=-1) + log_det

    def entropy(self, z, cond, samplesize, sub_manifolds, embedding=True):
        """pdf.entropy on given base normals z [batch*S, D] (cond: [batch, C] or None).
        Reference: main/default.py:2263-2454 (total and marginal entropies)."""
        S = samplesize
        z = torch.as_tensor(np.asarray(z)).to(self.dtype)
        cond = None if cond is None else torch.as_tensor(np.asarray(cond)).to(self.dtype)
        cr = None if cond is None else cond.repeat_interleave(S, dim=0)
        x, logp, _ = self.sample(z, cr)
        subs = self.prog["subpdfs"]
        out = {}

        def chart(k, xk):
            """-log sin(theta) of an S2 sub-pdf in embedding coordinates"""
            if embedding and subs[k]["manifold"] == "s" and xk.shape[1] == 2:
                return -torch.log(torch.sin(_safe_angle(xk[:, 0])))
            return torch.zeros(xk.shape[0], dtype=self.dtype)

        embs = [self.subs[k][-1].embedding(x[:, sp["target_cols"][0]:sp["target_cols"][1]]) for k, sp in enumerate(subs)]
        for sm in sub_manifolds:
            if sm == -1:
                tot = logp + sum(chart(k, x[:, sp["target_cols"][0]:sp["target_cols"][1]]) for k, sp in enumerate(subs))
                out["total"] = -tot.reshape(-1, S).mean(dim=1)
                continue
            t0, t1 = subs[sm]["target_cols"]
            xk = x[:, t0:t1]
            if sm == 0:
                lp = self.subpdf_log_pdf(0, xk, cr, []) + chart(0, xk)
                out[0] = -lp.reshape(-1, S).mean(dim=1)
                continue
            nb = x.shape[0] // S
            rep_first = lambda t: t.reshape(nb, S, -1).repeat(1, S, 1).reshape(nb * S * S, -1)
            final = xk.reshape(nb, S, -1).repeat_interleave(S, dim=1).reshape(nb * S * S, -1)
            prev = [rep_first(e) for e in embs[:sm]]
            crr = None if cr is None else cr.repeat_interleave(S, dim=0)
            lp = (self.subpdf_log_pdf(sm, final, crr, prev) + chart(sm, final)).reshape(-1, S, S)
            out[sm] = -(torch.logsumexp(lp, dim=-1) - math.log(float(S))).mean(dim=1)
        return out

    def _as_cond(self, cond):
        if cond is None:
            return None
        if isinstance(cond, (list, tuple)):
            return [torch.as_tensor(np.asarray(c)).to(self.dtype) for c in cond]
        return torch.as_tensor(np.asarray(cond)).to(self.dtype)

    def _last_layer(self, k):
        """the layer `only_last` applies for sub-pdf k: the last one, sphere layers with the base chart forced on
        (`fix_euclidean_to_sphere_first=True`, main/default.py:1018-1021, :1497-1498)"""
        sp = self.prog["subpdfs"][k]
        ls = dict(sp["layers"][-1])
        if sp["manifold"] == "s":
            ls["first"] = 1
        return LAYER_TYPES[ls["code"]](ls)

    def log_pdf(self, x, cond=None, amort=None, only_last=False):
        """amort: amortization_parameters [B, T] of a pdf built with amortize_everything (main/default.py:1062-1076).
        only_last: apply only the last layer of every sub-pdf (main/default.py:1015-1024)."""
        x = torch.as_tensor(np.asarray(x)).to(self.dtype)
        cond = self._as_cond(cond)
        self._split_amortization(amort)
        b = x.shape[0]
        log_det = torch.zeros(b, dtype=self.dtype)
        prev_emb, base = [], []
        for k, layers in enumerate(self.subs):
            sp = self.prog["subpdfs"][k]
            p = self._sub_params(k, cond, prev_emb, b)
            t0, t1 = sp["target_cols"]
            cur = x[:, t0:t1]
            if only_last:
                o0, o1 = sp["layer_param_ranges"][-1]
                cur, log_det = self._last_layer(k).inverse(cur, log_det, p[:, o0:o1])
            else:
                for li in reversed(range(len(layers))):
                    o0, o1 = sp["layer_param_ranges"][li]
                    cur, log_det = layers[li].inverse(cur, log_det, p[:, o0:o1])
            base.append(cur)
            prev_emb.append(layers[-1].embedding(x[:, t0:t1]))
        base = torch.cat(base, dim=1)
        self._amort = None
        logp_base = (-0.5 * base ** 2 - LOG_SQRT_2PI).sum(dim=-1)
        return logp_base + log_det, logp_base, base

    def sample(self, z, cond=None, amort=None, only_last=False):
        z = torch.as_tensor(np.asarray(z)).to(self.dtype)
        cond = self._as_cond(cond)
        self._split_amortization(amort)
        b = z.shape[0]
        log_det = torch.zeros(b, dtype=self.dtype)
        prev_emb, out = [], []
        for k, layers in enumerate(self.subs):
            sp = self.prog["subpdfs"][k]
            p = self._sub_params(k, cond, prev_emb, b)
            b0, b1 = sp["base_cols"]
            cur = z[:, b0:b1]
            if only_last:
                if sp["manifold"] not in ("s", "e"):
                    raise Exception("Flow type ", sp["manifold"], " does not supported *only_last*!")   # :1499-1502
                o0, o1 = sp["layer_param_ranges"][-1]
                cur, log_det = self._last_layer(k).forward(cur, log_det, p[:, o0:o1])
            else:
                for li in range(len(layers)):
                    o0, o1 = sp["layer_param_ranges"][li]
                    cur, log_det = layers[li].forward(cur, log_det, p[:, o0:o1])
            out.append(cur)
            prev_emb.append(layers[-1].embedding(cur))
        x = torch.cat(out, dim=1)
        self._amort = None
        logp_base = (-0.5 * z ** 2 - LOG_SQRT_2PI).sum(dim=-1)
        return x, logp_base - log_det, logp_base


def fully_amortized_parameters(outer_spec, params, cond, dtype=torch.float64):
    """Outer generator of `fully_amortized_pdf`: conditional input [B, C] -> amortization parameters [B, T], which
    OraclePdf.log_pdf / sample take as `amort`.  Reference: main/fully_amortized.py:96-131 (construction), :175 / :218
    (`all_flow_params=self.amortization_mlp(conditional_input)`).
    outer_spec: AmortizableMLP structure dict (custom mode, parameters `amortization_mlp.u_v_b_pars`) or
    dict(linear_indices=[...]) for the plain nn.Sequential(Linear, Tanh, ..., Linear)."""
    cond = torch.as_tensor(np.asarray(cond)).to(dtype)
    if outer_spec.get("custom", False):
        flat = torch.as_tensor(np.asarray(params["amortization_mlp.u_v_b_pars"])).to(dtype)
        return OraclePdf._custom_mlp(dict(outer_spec, amortised=False), flat, cond)
    h = cond
    idx = outer_spec["linear_indices"]
    for li, i in enumerate(idx):
        w = torch.as_tensor(np.asarray(params["amortization_mlp.%d.weight" % i])).to(dtype)
        b = torch.as_tensor(np.asarray(params["amortization_mlp.%d.bias" % i])).to(dtype)
        h = F.linear(h, w, b)
        if li < len(idx) - 1:
            h = torch.tanh(h)
    return h
