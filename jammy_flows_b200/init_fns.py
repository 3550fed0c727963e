"""Data-driven initialisation of a chain of Euclidean layers (`pdf.init_params(data=...)`).

What it computes follows the reference's `find_init_pars_of_chained_blocks` (extra_functions.py:179-409, helpers
:100-176): walking the chain from the target side to the base side, every layer gets initial parameters that roughly
Gaussianise the data it sees, and the data are pushed through that layer before the next one is initialised:

  * offset           = column means (euclidean_base `model_offset`)
  * "t" (mvn_block)  = lower-triangular factor fitted to the (eigenvalue-floored) second-moment matrix by minimising the
                       reverse KL with `scipy.optimize.minimize`, data whitened with the matching matrix square root
  * "g" (gf_block)   = Householder vectors fitted (first layer, d < 30) so that Q(v) maps the unit diagonal like the PCA
                       basis of X^T X; kernel means = K percentiles per dimension, log-widths = log(1.5 x the smallest
                       percentile gap); data passed through the mixture-CDF / inverse-CDF stage with exactly those
                       widths (no width regulator -- the reference calls `sigmoid_inv_error_pass_w_params` on the raw
                       values)

This is host logic executed once (numpy / scipy, like the reference).  The one data-parallel step -- pushing the data
through a "g" layer -- runs on the sm_100a layer kernel through the layer plugin API (a temporary `gf_block` whose
width regulator is the plain exponential), so the data must be (movable to) a CUDA device; there is no CPU
implementation of the layer math in this package.  The random draws (Householder start vectors, torch and numpy) are
made in the reference's order, so equal seeds start the optimiser from the same point.
"""
import numpy
import scipy.linalg
import torch
from scipy.optimize import minimize

from . import layers


def _householder_matrix_np(vs):
    """Q = H_0 H_1 ... for vs [n_iter, d] (reference gaussianization_flow.py:457-471), float64 numpy."""
    n_iter, d = vs.shape
    q = numpy.eye(d)
    for i in range(n_iter):
        v = vs[i] / numpy.sqrt((vs[i] ** 2).sum())
        q = q @ (numpy.eye(d) - 2.0 * numpy.outer(v, v))
    return q


def _bounded_exp(raw, lo, hi):
    """w = lo + 1/(1/hi + exp(-raw)): the smooth width regulator of "t" / "g" (gaussianization_flow.py:23-47, center=True)."""
    return lo + 1.0 / (1.0 / hi + numpy.exp(-raw))


def _mvn_lower(layer, a):
    """Lower-triangular factor of a "t" layer from its raw covariance parameters (matrix_fns.py:4-52)."""
    d, ct = layer.dimension, layer.cov_type
    if ct == "diagonal_symmetric":
        return numpy.eye(d) * _bounded_exp(a[0], layer.width_min, layer.width_max)
    diag = _bounded_exp(a[:d], layer.width_min, layer.width_max)
    m = numpy.diag(diag)
    if ct == "full":
        low, pos = a[d:], 0
        for ind in range(d - 1):                       # sub-diagonals, starting with the bottom-left corner
            n = ind + 1
            m = m + numpy.diag(low[pos:pos + n], k=-(d - 1 - ind))
            pos += n
    return m


def _fit_mvn(layer, second_moment):
    """raw covariance parameters of `layer` minimising KL(N(0, L L^T) || N(0, M)) (extra_functions.py:123-152)."""
    d = layer.dimension
    inv_target = scipy.linalg.pinv(second_moment)
    logdet_target = numpy.linalg.slogdet(second_moment)[1]

    def loss(a):
        lo = _mvn_lower(layer, a)
        pred = lo @ lo.T
        return 0.5 * (numpy.trace(inv_target @ pred) - numpy.linalg.slogdet(pred)[1] + logdet_target - d)

    n_par = layer.total_param_num - (d if layer.model_offset else 0)
    start = numpy.random.normal(size=n_par)
    return minimize(loss, start)["x"]


def _mvn_whitening(layer, a):
    """sqrt of the inverse of the fitted covariance (extra_functions.py:154-176)."""
    lo = _mvn_lower(layer, a)
    inv_pred = scipy.linalg.pinv(lo @ lo.T)
    _, sigma, r = scipy.linalg.svd(inv_pred)
    return numpy.sqrt(sigma) * r


def _fit_householder(target, n_iter):
    """Householder vectors whose Q maps the normalised all-ones vector like `target` does (extra_functions.py:100-121)."""
    d = target.shape[0]
    test = numpy.ones(d) / numpy.sqrt(float(d))
    want = target @ test

    def loss(a):
        return -((_householder_matrix_np(a.reshape(n_iter, d)) @ test) * want).sum()

    start = numpy.random.normal(size=d * d)
    return minimize(loss, start)["x"]


def _push_through_g(layer, data, means, log_widths):
    """data -> inverse-CDF(mixture CDF(data)) with widths exp(log_widths), on the layer kernel (C-ABI)."""
    if not torch.cuda.is_available():
        raise RuntimeError("jammy_flows_b200: data-driven initialisation pushes the data through the sm_100a layer "
                           "kernel and needs a CUDA device (there is no CPU fallback)")
    d, k = layer.dimension, layer.num_kde
    tmp = layers.gf_block(d, num_kde=k, num_householder_iter=0, use_permanent_parameters=False, fit_normalization=0,
                          inverse_function_type=layer.inverse_function_type, model_offset=0,
                          width_smooth_saturation=0, lower_bound_for_widths=1e-300, upper_bound_for_widths=-1,
                          add_skewness=layer.add_skewness, rotation_mode="none")
    pieces = [means.reshape(-1), log_widths.reshape(-1)]
    if layer.add_skewness:
        pieces.append(torch.zeros(k * d, dtype=data.dtype))
    p = torch.cat([t.to(data.dtype).cpu() for t in pieces]).unsqueeze(0)
    dev = data.device if data.is_cuda else torch.device("cuda")
    out, _ = tmp.inv_flow_mapping([data.to(dev).contiguous(), 0.0], extra_inputs=p.to(dev))
    return out.to(data.device)


def find_init_pars_of_chained_blocks(layer_list, data, mvn_min_max_sv_ratio=1e-4):
    """Initial parameter vector (in `extra_inputs` order, first layer first) of one Euclidean sub-pdf."""
    if data is None:
        # reference order: the chain is traversed from the last layer to the first (matters for the RNG stream)
        rev = [l.get_desired_init_parameters() for l in list(layer_list)[::-1]]
        return torch.cat(rev[::-1])
    cur = data
    dim = data.shape[1]
    per_layer = []
    with torch.no_grad():
        for layer_ind, layer in enumerate(list(layer_list)[::-1]):
            pars = []
            if layer.model_offset:
                mean = cur.mean(axis=0, keepdim=True)
                pars.append(mean.squeeze(0))
                cur = cur - mean
            if isinstance(layer, layers.mvn_block):
                if layer.cov_type == "identity":
                    if len(pars) > 0:
                        per_layer.append(torch.cat(pars))
                    continue
                moment = (torch.matmul(cur.T, cur) / float(data.shape[0])).cpu().numpy().astype(numpy.float64)
                l_, sigma, r_ = scipy.linalg.svd(moment)
                sigma = numpy.where(sigma < mvn_min_max_sv_ratio * max(sigma), mvn_min_max_sv_ratio * max(sigma), sigma)
                a = _fit_mvn(layer, (l_ * sigma) @ r_)
                pars.append(torch.from_numpy(a).to(data))
                white = torch.from_numpy(_mvn_whitening(layer, a)).to(cur)
                cur = torch.matmul(cur, white.T)
            elif isinstance(layer, layers.gf_block):
                vs = None
                if layer.rotation_mode == "householder":
                    if layer.use_householder:
                        if layer.dimension < 30 and layer_ind == 0:
                            moment = torch.matmul(cur.T, cur).cpu().numpy().astype(numpy.float64)
                            _, _, r_ = scipy.linalg.svd(moment)
                            # the reference builds a throw-away gf_block with permanent parameters here: same RNG draws
                            layers.gf_block(dim, num_householder_iter=layer.householder_iter, use_permanent_parameters=True)
                            vs = torch.from_numpy(_fit_householder(r_, layer.householder_iter)).to(data)
                        else:
                            vs = torch.randn(layer.dimension * layer.householder_iter).to(data)
                        pars.append(vs)
                        layers.gf_block(dim, num_householder_iter=layer.householder_iter, use_permanent_parameters=True)
                        q = _householder_matrix_np(vs.cpu().numpy().astype(numpy.float64).reshape(layer.householder_iter, dim))
                        cur = torch.matmul(cur, torch.from_numpy(q).to(cur))          # rows: (Q^T x)^T = x^T Q
                elif layer.rotation_mode != "none":
                    pars.append(torch.zeros(layer.num_rotation_params))
                k = layer.num_kde
                assert (k < 100)
                if layer.nonlinear_stretch_type != "classic":
                    raise Exception("Data initilaization only implemented (and probably only makes sense) for classic "
                                    "Gaussianization Flow structure")
                # kernel means = percentiles of the (rotated) data, one common log-width per dimension
                perc = torch.from_numpy(numpy.percentile(cur.cpu().numpy(), numpy.linspace(0, 100, k), axis=0)).to(data)
                pars.append(perc.flatten() if layer.center_mean == 0 else perc[:-1].flatten())
                gaps = perc[1:, :] - perc[:-1, :]
                bw = torch.ones_like(perc) * torch.log(gaps.min(axis=0, keepdim=True)[0] * 1.5)
                pars.append(bw.flatten())
                if layer.fit_normalization:
                    pars.append(torch.ones_like(bw.flatten()))
                if layer.add_skewness:
                    pars.append(torch.zeros_like(bw.flatten()))
                cur = _push_through_g(layer, cur, perc, bw)
            else:
                raise NotImplementedError("data-driven initialisation of layer %r" % type(layer).__name__)
            per_layer.append(torch.cat([p.to(data) for p in pars]))
    out = torch.cat(per_layer[::-1])
    assert len(out) == sum(l.total_param_num for l in layer_list), (len(out), sum(l.total_param_num for l in layer_list))
    return out
