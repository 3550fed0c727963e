"""High-precision (mpmath, 40 digits) evaluation of a shared-parameter 'g' chain in the log_pdf direction.
Used to decide, for rows where CUDA and the reference disagree by more than 1e-10, which one is closer to the
exact value of the reference's own formulas (conditioning analysis for DESIGN.md)."""
import sys, json, numpy as np
import mpmath as mp
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from helpers import load_golden, build_pdf
mp.mp.dps = 40

def regulate(x, lo, hi, center):
    c = mp.log(hi) if center else mp.mpf(0)
    first = mp.log(hi) - mp.log(1 + mp.exp(-x + c))
    return mp.log(mp.exp(first) + lo)

def g_layer_inverse(x, desc, p):
    d, K = desc["dim"], desc["num_kde"]
    i = 0
    if desc["model_offset"]:
        x = [x[j] - p[j] for j in range(d)]; i = d
    for it in range(desc["hh_iter"]):
        v = p[i + it * d: i + (it + 1) * d]
        nrm = sum(a * a for a in v); dot = sum(a * b for a, b in zip(v, x))
        x = [x[j] - 2 * v[j] * dot / nrm for j in range(d)]
    i += desc["hh_iter"] * d
    m = p[i:i + K * d]; i += K * d
    lw = [regulate(a, desc["w_min"], desc["w_max"], True) for a in p[i:i + K * d]]; i += K * d
    if desc["fit_normalization"]:
        ln = p[i:i + K * d]
        if desc["regulate_normalization"]:
            ln = [regulate(a, desc["n_min"], desc["n_max"], False) for a in ln]
    else:
        ln = [mp.mpf(0)] * (K * d)
    ys, ld = [], mp.mpf(0)
    for j in range(d):
        nsum = sum(mp.exp(ln[k * d + j]) for k in range(K))
        cdf = sf = pdf = mp.mpf(0)
        for k in range(K):
            w = mp.exp(lw[k * d + j]); a = (x[j] - m[k * d + j]) / w; n = mp.exp(ln[k * d + j]) / nsum
            s = 1 / (1 + mp.exp(-a)); cdf += n * s; sf += n * (1 - s) if a < 30 else n * mp.exp(-a) / (1 + mp.exp(-a)); pdf += n * s * (1 / (1 + mp.exp(a))) / w
        lc, ls, lp = mp.log(cdf), mp.log(sf), mp.log(pdf)
        t = desc["inverse_function_type"]
        if t == "isigmoid":
            y = lc - ls; logd = lp - lc - ls
        else:
            eps = mp.mpf("0.5e-7"); a_ = mp.mpf("0.147"); c_ = 2 / (mp.pi * a_)
            lnf = lc + ls + mp.log(4); F = c_ + lnf / 2; F2 = mp.sqrt(F * F - lnf / a_)
            bulk = cdf > eps and cdf < 1 - eps
            if t != "inormal_full_pade" and bulk:
                e = mp.erfinv(2 * cdf - 1); y = mp.sqrt(2) * e; logd = mp.log(mp.sqrt(2 * mp.pi)) + e * e + lp
            elif t == "inormal_partly_crude":
                s_ = -2 * (ls + lc); tail = mp.sqrt(s_) - mp.mpf("0.4717"); y = tail if cdf >= 1 - eps else -tail
                logd = -mp.log(s_) / 2 - ls - lc + lp
            else:
                pade = mp.sqrt(max(0, 2 * (F2 - F)))
                if cdf > mp.mpf("0.49999") and cdf < mp.mpf("0.50001"):
                    tot = mp.log(mp.mpf("2.506628"))
                else:
                    tot = mp.log(-(F - 1 / a_ - F2)) - (mp.log(8) / 2 + mp.log(F2 - F) / 2 + mp.log(F2)) - ls - lc + mp.log(abs(1 - 2 * cdf))
                y = (-pade if cdf <= mp.mpf("0.5") else pade) if t == "inormal_full_pade" else (pade if cdf >= 1 - eps else -pade)
                logd = tot + lp
        ys.append(y); ld += logd
    return ys, ld

def truth_rows(name, rows):
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params)
    prog = p.export_program()
    sp = prog["subpdfs"][0]
    vec = np.concatenate([params[n].reshape(-1) for n in sp["permanent_param_names"]])
    out = []
    for r in rows:
        x = [mp.mpf(float(v)) for v in data["x"][r]]
        ld = mp.mpf(0)
        for li in reversed(range(len(sp["layers"]))):
            o0, o1 = sp["layer_param_ranges"][li]
            x, l = g_layer_inverse(x, sp["layers"][li], [mp.mpf(float(v)) for v in vec[o0:o1]])
            ld += l
        logp = ld + sum(-v * v / 2 - mp.log(mp.sqrt(2 * mp.pi)) for v in x)
        out.append(([float(v) for v in x], float(logp)))
    return out

if __name__ == "__main__":
    name = sys.argv[1]
    meta, params, data = load_golden(name)
    c = np.load("/root/repo/gpurun_out/cuda_%s.npz" % name)
    err = np.abs(c["base"] - data["base"]).max(axis=1) / np.maximum(1, np.abs(data["base"]).max(axis=1))
    rows = np.argsort(-err)[:6]
    tr = truth_rows(name, rows)
    for r, (zb, lp) in zip(rows, tr):
        print("row %4d x=%s  base_ref=%s" % (r, data["x"][r], data["base"][r]))
        print("    |cuda-ref| %.2e   |ref-truth| %.2e   |cuda-truth| %.2e   logp: |ref-truth| %.2e |cuda-truth| %.2e" % (
            np.abs(c["base"][r] - data["base"][r]).max(), np.abs(np.array(zb) - data["base"][r]).max(),
            np.abs(np.array(zb) - c["base"][r]).max(), abs(lp - data["logp"][r]), abs(lp - c["logp"][r])))


def truth_rows_cond(name, rows):
    """single conditional e sub-pdf: params from the MLP evaluated in mp arithmetic"""
    meta, params, data = load_golden(name)
    p = build_pdf(meta, params)
    prog = p.export_program()
    sp = prog["subpdfs"][0]
    W = [params["mlp_predictors.0.%d.weight" % i] for i in sp["mlp"]["linear_indices"]]
    b = [params["mlp_predictors.0.%d.bias" % i] for i in sp["mlp"]["linear_indices"]]
    out = []
    for r in rows:
        h = [mp.mpf(float(v)) for v in data["cond"][r]]
        for li in range(len(W)):
            h = [sum(mp.mpf(float(W[li][o, i])) * h[i] for i in range(len(h))) + mp.mpf(float(b[li][o])) for o in range(W[li].shape[0])]
            if li < len(W) - 1:
                h = [mp.tanh(v) for v in h]
        x = [mp.mpf(float(v)) for v in data["x"][r]]
        ld = mp.mpf(0)
        for li in reversed(range(len(sp["layers"]))):
            o0, o1 = sp["layer_param_ranges"][li]
            x, l = g_layer_inverse(x, sp["layers"][li], h[o0:o1])
            ld += l
        logp = ld + sum(-v * v / 2 - mp.log(mp.sqrt(2 * mp.pi)) for v in x)
        out.append(([float(v) for v in x], float(logp)))
    return out
