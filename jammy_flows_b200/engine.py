"""Host-side dispatch: compiles a `pdf` into the C-ABI descriptors and launches the sm_100a kernels.

PyTorch is used for device memory (torch.empty), the current CUDA stream and, later, torch.distributed -- plumbing
only.  Every number comes out of libjammy_b200.so; nothing here computes layer math and nothing falls back to
eager PyTorch or the CPU.
"""
import ctypes as C
import math

import torch

from . import _cabi

_DT = {torch.float32: _cabi.JF_F32, torch.float64: _cabi.JF_F64}
DEFAULT_CHUNK_ROWS = 1 << 19

_workspaces = {}


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "jammy_flows_b200: %s lives on '%s'. The hot path only exists as sm_100a CUDA kernels -- move the model and "
            "its inputs to a CUDA device (there is no CPU fallback)." % (what, t.device))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _workspace(device, nbytes):
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ---------------------------------------------------------------------------------------------------------------------
# descriptors
# ---------------------------------------------------------------------------------------------------------------------
def fill_layer_desc(ld, desc, param_offset):
    """python layer descriptor dict (layers.*.descriptor()) -> JfLayerDesc"""
    ld.dim = desc["dim"]
    ld.n_params = desc["n_params"]
    ld.param_offset = param_offset
    if desc["code"] == "g":
        ld.kind = _cabi.JF_LAYER_GF
        ld.K = desc["num_kde"]
        ld.hh_iter = desc["hh_iter"]
        ld.inv_type = desc["inv_type"]
        ld.norm_mode = (_cabi.JF_NORM_NONE if not desc["fit_normalization"] else
                        (_cabi.JF_NORM_REGULATED if desc["regulate_normalization"] else _cabi.JF_NORM_RAW))
        ld.has_offset = desc["model_offset"]
        ld.w_min, ld.w_max, ld.n_min, ld.n_max = desc["w_min"], desc["w_max"], desc["n_min"], desc["n_max"]
        from .layers import ROT_MODES, WIDTH_MODES
        ld.rotation_mode = ROT_MODES[desc.get("rotation_mode", "householder")]
        ld.width_mode = WIDTH_MODES[desc.get("width_mode", "smooth")]
        clamp = desc.get("width_clamp")
        ld.width_clamp = 0 if clamp is None else 1
        ld.clamp_lo, ld.clamp_hi = (0.0, 0.0) if clamp is None else (clamp[0], clamp[1])
        ld.skew = desc.get("add_skewness", 0)
        ld.center_mean = desc.get("center_mean", 0)
        ld.stretch = (_cabi.JF_STRETCH_RQS if desc.get("stretch", "classic") == "rq_splines"
                      else _cabi.JF_STRETCH_CLASSIC)
    elif desc["code"] == "t":
        ld.kind = _cabi.JF_LAYER_MVN
        ld.inv_type = desc["cov"]
        ld.has_offset = desc["model_offset"]
        ld.w_min, ld.w_max = desc["w_min"], desc["w_max"]
    elif desc["code"] == "f":
        ld.kind = _cabi.JF_LAYER_FVM
        ld.hh_iter = desc["hh_iter"] if desc["add_rotation"] else 0
        ld.rotation_mode = {"householder": _cabi.JF_ROT_HOUSEHOLDER, "angles": _cabi.JF_ROT_ANGLES, "xyz": _cabi.JF_ROT_XYZ,
                            "quaternion": _cabi.JF_ROT_QUATERNION}[desc.get("rotation_mode", "householder")] \
            if desc["add_rotation"] else _cabi.JF_ROT_HOUSEHOLDER
        ld.width_mode = desc.get("kappa_mode", 0)
        ld.width_clamp = desc.get("kappa_clamping", 0)
        ld.skew = desc.get("extra_rotation", 0)
        ld.clamp_lo = desc.get("identity_region", 0.0)
        ld.first = desc["first"]
        ld.z_sign = desc["z_sign"]
        ld.min_kappa = desc["min_kappa"]
        nested = list(desc.get("vertical", [])) + list(desc.get("circular", []))
        if len(nested) > _cabi.JF_MAX_NESTED:
            raise NotImplementedError("more than %d nested spline sub-flows" % _cabi.JF_MAX_NESTED)
        ld.n_vertical = len(desc.get("vertical", []))
        ld.n_circular = len(desc.get("circular", []))
        for i, sp in enumerate(nested):
            fill_spline_desc(ld.spline[i], sp)
    elif desc["code"] == "r":
        ld.kind = _cabi.JF_LAYER_RQS
        ld.first = desc["first"]
        ld.lo, ld.hi = desc["lo"], desc["hi"]
        fill_spline_desc(ld.spline[0], desc["spline"])
    elif desc["code"] == "o":
        ld.kind = _cabi.JF_LAYER_S1SPLINE
        ld.hh_iter = desc["hh_iter"] if desc["add_rotation"] else 0
        ld.first = desc["first"]
        ld.natural_direction = desc["natural_direction"]
        fill_spline_desc(ld.spline[0], desc["spline"])
    elif desc["code"] == "m":
        ld.kind = _cabi.JF_LAYER_MOEBIUS
        ld.hh_iter = desc["hh_iter"] if desc["add_rotation"] else 0
        ld.first = desc["first"]
        ld.natural_direction = desc["natural_direction"]
        ld.K = desc["K"]
    elif desc["code"] == "v":
        ld.kind = _cabi.JF_LAYER_EXPMAP
        ld.hh_iter = desc["hh_iter"] if desc["add_rotation"] else 0
        ld.first = desc["first"]
        ld.natural_direction = desc["natural_direction"]
        ld.K = desc["K"]
        ld.max_iter = desc["max_iter"]
        ld.inv_type = desc.get("potential", 0)
    else:
        raise NotImplementedError("no sm_100a kernel for layer code %r" % desc["code"])


_SPLINE_KINDS = {"plain": _cabi.JF_SPLINE_PLAIN, "smooth": _cabi.JF_SPLINE_SMOOTH, "circular": _cabi.JF_SPLINE_CIRCULAR}


def fill_spline_desc(sd, spec):
    """python spline spec (layers._spline_options._spline_spec) -> JfSplineDesc"""
    sd.kind = _SPLINE_KINDS[spec["kind"]]
    for name in ("n_bins", "n_w", "n_h", "n_d", "fix_first", "fix_second", "indep", "bd_mode", "natural_direction",
                 "param_offset", "lo", "hi", "min_w", "min_h", "min_d", "bd_fixed", "max_ratio"):
        setattr(sd, name, spec[name])


def fill_subpdf_desc(sd, manifold, dim, layer_descs):
    if len(layer_descs) > _cabi.JF_MAX_LAYERS:
        raise NotImplementedError("more than %d layers in one sub-pdf" % _cabi.JF_MAX_LAYERS)
    sd.manifold = ord(manifold)
    sd.dim = dim
    sd.n_layers = len(layer_descs)
    off = 0
    for i, d in enumerate(layer_descs):
        fill_layer_desc(sd.layers[i], d, off)
        off += d["n_params"]
    sd.n_params = off


def compile_pdf(pdf, dtype):
    """`pdf` -> JfPdfDesc (static structure only; parameter pointers are gathered per call)."""
    if len(pdf.layer_list) > _cabi.JF_MAX_SUBPDFS:
        raise NotImplementedError("more than %d sub-pdfs" % _cabi.JF_MAX_SUBPDFS)
    d = _cabi.JfPdfDesc()
    d.abi_version = _cabi.JF_ABI_VERSION
    d.dtype = _DT[dtype]
    d.n_sub = len(pdf.layer_list)
    d.cond_dim = int(pdf.conditional_input_dim or 0) if type(pdf.conditional_input_dim) != list else 0
    d.total_target_dim = pdf.total_target_dim
    d.total_base_dim = pdf.total_base_dim
    for k, layers in enumerate(pdf.layer_list):
        manifold = pdf.pdf_defs_list[k][0]
        fill_subpdf_desc(d.sub[k], manifold, layers[0].dimension, [l.descriptor() for l in layers])
        d.target_col[k] = pdf.target_dim_indices[k][0]
        d.base_col[k] = pdf.base_dim_indices[k][0]
        d.emb_dim[k] = layers[-1]._embedding_conditional_return_num()
        mlp = pdf.mlp_predictors[k]
        d.has_mlp[k] = 0 if mlp is None else 1
        if mlp is not None and not hasattr(mlp, "u_v_b_pars"):      # AmortizableMLP: composed on the host (staged path)
            linears = [m for m in mlp if isinstance(m, torch.nn.Linear)]
            if len(linears) > _cabi.JF_MAX_MLP_LINEAR:
                raise NotImplementedError("MLP deeper than %d Linear layers" % _cabi.JF_MAX_MLP_LINEAR)
            d.mlp[k].n_linear = len(linears)
            d.mlp[k].dims[0] = linears[0].in_features
            for i, lin in enumerate(linears):
                d.mlp[k].dims[i + 1] = lin.out_features
    return d


class ParamPack:
    """Device pointers of one call (keeps the temporaries alive until the call has been enqueued)."""

    def __init__(self, pdf, dtype, device):
        self.keep = []
        self.c = _cabi.JfPdfParams()
        for k, layers in enumerate(pdf.layer_list):
            mlp = pdf.mlp_predictors[k]
            if mlp is None and pdf.amortize_everything:
                continue            # the flow parameters arrive per row in `amortization_parameters`
            if mlp is None:
                vecs = [l.packed_permanent_params() for l in layers]
                vecs = [v for v in vecs if v is not None]
                if len(vecs) > 0:
                    vec = torch.cat([v.detach().to(device=device, dtype=dtype) for v in vecs]).contiguous()
                    _require_cuda(vec, "parameters of sub-pdf %d" % k)
                    self.keep.append(vec)
                    self.c.shared[k] = vec.data_ptr()
            elif hasattr(mlp, "u_v_b_pars"):
                continue
            else:
                linears = [m for m in mlp if isinstance(m, torch.nn.Linear)]
                for i, lin in enumerate(linears):
                    wt = lin.weight.detach().to(device=device, dtype=dtype).contiguous()
                    b = lin.bias.detach().to(device=device, dtype=dtype).contiguous()
                    _require_cuda(wt, "MLP weights of sub-pdf %d" % k)
                    self.keep += [wt, b]
                    self.c.weights[k][i] = wt.data_ptr()
                    self.c.biases[k][i] = b.data_ptr()


# ---------------------------------------------------------------------------------------------------------------------
# whole-pdf calls
# ---------------------------------------------------------------------------------------------------------------------
def _prep_inputs(pdf, t, cond, what):
    _require_cuda(t, what)
    if t.dtype not in _DT:
        raise TypeError("jammy_flows_b200 supports float32/float64, got %s" % t.dtype)
    if t.dim() != 2:
        raise ValueError("%s must be 2-dimensional (B, D)" % what)
    if t.stride(1) != 1:
        t = t.contiguous()
    if isinstance(cond, (list, tuple)):
        out = []
        for ci in cond:
            _require_cuda(ci, "conditional_input")
            assert ci.dtype == t.dtype and ci.device == t.device
            out.append(ci if ci.stride(1) == 1 else ci.contiguous())
        return t, out
    if cond is not None:
        _require_cuda(cond, "conditional_input")
        assert cond.dtype == t.dtype and cond.device == t.device
        if cond.stride(1) != 1:
            cond = cond.contiguous()
    return t, cond


def uses_custom_mlp(pdf):
    """True when the per-sub-pdf orchestration on the host is needed: AmortizableMLP generators, or one conditional input
    per sub-pdf (`conditional_input_dim` list), neither of which the single-call C entries describe."""
    return type(pdf.conditional_input_dim) == list or \
        (pdf.predict_log_normalization and pdf.join_poisson_and_pdf_description) or \
        any(m is not None and hasattr(m, "u_v_b_pars") for m in pdf.mlp_predictors)


def _run_chain(lib, dt, dev, layers_wb, segs, out, so_p, so_r, R, accumulate):
    """one Linear/tanh chain with dense weights [(W, b)] on column blocks `segs` -> out (optionally +=)."""
    # slices of a flat parameter vector may start at an odd element: the tensor-pipe kernels stream weights in 16-byte
    # pieces (the library falls back to its generic kernel for unaligned pointers; aligned copies keep the fast one)
    layers_wb = [(w if w.data_ptr() % 16 == 0 else w.clone(), b if b.data_ptr() % 16 == 0 else b.clone())
                 for w, b in layers_wb]
    md = _cabi.JfMlpDesc()
    md.n_linear = len(layers_wb)
    md.dims[0] = layers_wb[0][0].shape[1]
    for i, (w, _) in enumerate(layers_wb):
        md.dims[i + 1] = w.shape[0]
    if md.n_linear > _cabi.JF_MAX_MLP_LINEAR or len(segs) > _cabi.JF_MAX_MLP_SEGMENTS:
        raise NotImplementedError("MLP chain deeper than %d layers / more than %d input blocks"
                                  % (_cabi.JF_MAX_MLP_LINEAR, _cabi.JF_MAX_MLP_SEGMENTS))
    md.n_segments = len(segs)
    for i, sg in enumerate(segs):
        md.seg_cols[i] = sg.shape[1]
    assert sum(sg.shape[1] for sg in segs) == md.dims[0], ([sg.shape for sg in segs], md.dims[0])
    ptrs = (C.c_void_p * len(segs))(*[sg.data_ptr() for sg in segs])
    lds = (C.c_int64 * len(segs))(*[sg.stride(0) for sg in segs])
    ws = (C.c_void_p * len(layers_wb))(*[w.data_ptr() for w, _ in layers_wb])
    bs = (C.c_void_p * len(layers_wb))(*[b.data_ptr() for _, b in layers_wb])
    with torch.cuda.device(dev):
        rc = lib.jf_mlp_forward_acc(C.byref(md), _DT[dt], ptrs, lds, ws, bs, _ptr(out), so_p, so_r, R,
                                    1 if accumulate else 0, _stream_ptr(dev))
    _cabi.check(rc, "jf_mlp_forward_acc")


def custom_mlp_forward(mlp, segs, R):
    """AmortizableMLP (permanent parameters) on the column blocks `segs` -> [R, output_dim] row-major.
    The connectivity modes (reference amortizable_mlp.py:586-682) are composed from accumulating chain launches:
    out = highway(x); out += chain_0(x); out += chain_i(x | out | [x, out]) for the later chains."""
    lib = _cabi.load()
    segs = [s if s.stride(1) == 1 else s.contiguous() for s in segs]
    for s_ in segs:
        _require_cuda(s_, "MLP input")
    dt, dev = segs[0].dtype, segs[0].device
    chains, hw = mlp.chain_segments(dt, dev)
    P = mlp.output_dim
    # wide outputs (the [rows, T] block of a fully amortized pdf): rows padded to a multiple of 128 bytes, so that the
    # write-bound expand kernel stores whole lines (measured 4.3 -> 5.7 TB/s); the result is the [R, P] view
    ld = (P + 15) // 16 * 16 if P >= 256 else P
    out = torch.empty(R, ld, dtype=dt, device=dev)[:, :P]

    def run_pieces(pieces, inp, accumulate):
        # consecutive Linear/tanh chains (cut behind the V^T of a factorised layer): narrow row-major intermediates
        for piece in pieces[:-1]:
            width = piece[-1][0].shape[0]
            tmp = torch.empty(R, width, dtype=dt, device=dev)
            _run_chain(lib, dt, dev, piece, inp, tmp, 1, width, R, False)
            inp = [tmp]
        _run_chain(lib, dt, dev, pieces[-1], inp, out, 1, out.stride(0), R, accumulate)

    started = False
    if hw is not None:
        run_pieces(hw, segs, False)
        started = True
    for ci, ch in enumerate(chains):
        if ci == 0 or mlp.highway_mode <= 2:
            inp = segs
        elif mlp.highway_mode == 3:
            inp = [out]
        else:
            inp = segs + [out]
        # a CTA gathers the input rows it owns before it writes them, so reading `out` while accumulating into it is safe
        run_pieces(ch, inp, started)
        started = True
    if not started:
        out.zero_()
    return out


def sequential_mlp_forward(mlp, x):
    """nn.Sequential(Linear, Tanh, ..., Linear) on x [R, in] -> [R, out] row-major (jf_mlp_forward_acc)."""
    lib = _cabi.load()
    _require_cuda(x, "MLP input")
    if x.stride(1) != 1:
        x = x.contiguous()
    dt, dev, R = x.dtype, x.device, x.shape[0]
    wb = [(l.weight.detach().to(device=dev, dtype=dt).contiguous(), l.bias.detach().to(device=dev, dtype=dt).contiguous())
          for l in mlp if isinstance(l, torch.nn.Linear)]
    out = torch.empty(R, wb[-1][0].shape[0], dtype=dt, device=dev)
    _run_chain(lib, dt, dev, wb, [x], out, 1, out.shape[1], R, False)
    return out


def _rowwise_linear(lib, dt, dev, extra, off_w, off_b, inp, n_in, n_out, act, accumulate, out, R):
    with torch.cuda.device(dev):
        rc = lib.jf_rowwise_linear(_DT[dt], _ptr(extra), extra.stride(0), off_w, off_b, _ptr(inp), inp.stride(0), n_in,
                                   n_out, 1 if act else 0, 1 if accumulate else 0, _ptr(out), 1, out.stride(0), R,
                                   _stream_ptr(dev))
    _cabi.check(rc, "jf_rowwise_linear")


def amortized_mlp_forward(mlp, segs, extra, R):
    """AmortizableMLP "being amortised": row r runs the network whose flat parameter vector is extra[r]
    (reference amortizable_mlp.py:508-682 with `extra_inputs`; used by pdf(amortize_everything=True) and
    fully_amortized_pdf).  Every layer is one `jf_rowwise_linear` launch that reads its weights straight out of `extra`
    (column offsets follow the reference's layout: per layer U, V, bias; chains in order; the linear map last);
    factorised layers are V^T x then U (.), as the reference multiplies them.  -> [R, output_dim] row-major."""
    lib = _cabi.load()
    _require_cuda(extra, "amortization parameters")
    if extra.stride(1) != 1:
        extra = extra.contiguous()
    segs = [s_ for s_ in segs]
    for s_ in segs:
        _require_cuda(s_, "MLP input")
    x = segs[0] if len(segs) == 1 else torch.cat(segs, dim=1)
    if x.stride(1) != 1:
        x = x.contiguous()
    dt, dev = x.dtype, x.device
    assert extra.dtype == dt and extra.device == dev and extra.shape[0] == R and x.shape[0] == R
    assert extra.shape[1] == mlp.num_amortization_params
    P = mlp.output_dim
    out = torch.empty(R, P, dtype=dt, device=dev)

    def run_chain(chain, pos, inp, accumulate):
        h = inp
        n = len(chain.layers)
        for i, l in enumerate(chain.layers):
            last = i == n - 1
            off_u, off_v, off_b = pos, pos + l["n_u"], pos + l["n_u"] + l["n_v"]
            pos = off_b + l["n_b"]
            if l["n_b"] == 0:
                off_b = -1
            dst = out if last else torch.empty(R, l["n_out"], dtype=dt, device=dev)
            assert h.shape[1] == l["n_in"], (h.shape, l)
            if l["full"]:
                _rowwise_linear(lib, dt, dev, extra, off_u, off_b, h, l["n_in"], l["n_out"], not last,
                                last and accumulate, dst, R)
            else:
                mid = torch.empty(R, l["rank"], dtype=dt, device=dev)
                _rowwise_linear(lib, dt, dev, extra, off_v, -1, h, l["n_in"], l["rank"], False, False, mid, R)
                _rowwise_linear(lib, dt, dev, extra, off_u, off_b, mid, l["rank"], l["n_out"], not last,
                                last and accumulate, dst, R)
            h = dst

    started = False
    if mlp.highway is not None:
        run_chain(mlp.highway, mlp.num_amortization_params - mlp.highway.num_params, x, False)
        started = True
    pos = 0
    for ci, ch in enumerate(mlp.chains):
        if ci == 0 or mlp.highway_mode <= 2:
            inp = x
        elif mlp.highway_mode == 3:
            inp = out       # read by the chain's first launch, accumulated into by its last
        else:
            inp = torch.cat([x, out], dim=1)
        run_chain(ch, pos, inp, started)
        pos += ch.num_params
        started = True
    if not started:
        out.zero_()
    return out


def _pdf_staged(pdf, src, cond, direction, amort=None, only_last=False):
    """Row-chunked wrapper of `_pdf_staged_chunk` (the per-row parameter buffers are [rows, P]: bounded by the chunk)."""
    chunk = int(pdf.chunk_rows or DEFAULT_CHUNK_ROWS)
    R = src.shape[0]
    if amort is not None:
        assert amort.shape[0] == R, "batch size of amortization_parameters must agree with the batch size of the input"
    if R <= chunk:
        return _pdf_staged_chunk(pdf, src, cond, direction, amort, only_last)
    parts = []
    for r0 in range(0, R, chunk):
        c = None
        if cond is not None:
            c = [ci[r0:r0 + chunk] for ci in cond] if isinstance(cond, (list, tuple)) else cond[r0:r0 + chunk]
        parts.append(_pdf_staged_chunk(pdf, src[r0:r0 + chunk], c, direction,
                                       None if amort is None else amort[r0:r0 + chunk], only_last))
    return tuple(torch.cat([p[i] for p in parts], dim=0) for i in range(3))


def _last_layer_desc(pdf, k, dt):
    """`only_last` (reference main/default.py:1015-1024, :1490-1502): a one-layer program made of the LAST layer of
    sub-pdf k, sphere layers with the base chart forced on (fix_euclidean_to_sphere_first=True); -> (desc, offset of
    that layer's parameters inside the sub-pdf's parameter vector)."""
    layers = pdf.layer_list[k]
    d = dict(layers[-1].descriptor())
    if pdf.pdf_defs_list[k][0] == "s":
        d["first"] = 1
    sd = _cabi.JfSubPdfDesc()
    fill_subpdf_desc(sd, pdf.pdf_defs_list[k][0], layers[0].dimension, [d])
    return sd, sum(l.total_param_num for l in layers[:-1])


def _pdf_staged_chunk(pdf, src, cond, direction, amort=None, only_last=False):
    """Per-sub-pdf orchestration on the host: parameter generator (nn.Sequential or AmortizableMLP) -> layer chain, the
    embedding of each sub-pdf's target feeding the later generators (reference main/default.py:931-1053 / :1413-1514).
    Used when a generator is an AmortizableMLP, which the single-call C entries do not describe."""
    lib = _cabi.load()
    src, cond = _prep_inputs(pdf, src, cond, "input")
    R = src.shape[0]
    dt, dev = src.dtype, src.device
    desc = pdf._desc(dt)
    pack = ParamPack(pdf, dt, dev)
    status = pdf._status(dev)
    logpdf = direction == _cabi.JF_DIR_LOGPDF
    dst = torch.empty(R, pdf.total_base_dim if logpdf else pdf.total_target_dim, dtype=dt, device=dev)
    logdet = torch.empty(R, dtype=dt, device=dev)
    logbase = torch.empty(R, dtype=dt, device=dev)
    prev, keep = [], []
    amort_pos = 0
    if amort is not None:
        # reference main/default.py:925-927 / :1404-1407: the whole pdf (flow parameters of a first sub-pdf without
        # generator, then every generator's flat vector, in sub-pdf order) comes per row from `amortization_parameters`
        _require_cuda(amort, "amortization_parameters")
        assert amort.dtype == dt and amort.device == dev
        assert amort.shape[1] == pdf.total_number_amortizable_params, (amort.shape[1], pdf.total_number_amortizable_params)
        if amort.stride(1) != 1:
            amort = amort.contiguous()
    for k, layers in enumerate(pdf.layer_list):
        mlp = pdf.mlp_predictors[k]
        segs = ([cond[k] if isinstance(cond, list) else cond] if cond is not None else []) + prev
        n_par = desc.sub[k].n_params
        if mlp is None and amort is not None and n_par > 0:
            # param-major copy [n_par, R]: the thread-per-row layer kernels read it coalesced (a strided view of the
            # [R, T] block costs them 4x; the transposing copy moves 2 x n_par x 8 B per row once)
            buf = amort[:, amort_pos:amort_pos + n_par].t().contiguous()
            amort_pos += n_par
            keep.append(buf)
            params, sp, sr = _ptr(buf), R, 1
        elif mlp is None:
            params, sp, sr = C.c_void_p(pack.c.shared[k]), 1, 0
        elif amort is not None:
            n_am = mlp.num_amortization_params
            buf = amortized_mlp_forward(mlp, segs, amort[:, amort_pos:amort_pos + n_am], R).t().contiguous()
            amort_pos += n_am
            keep.append(buf)
            params, sp, sr = _ptr(buf), R, 1
        elif hasattr(mlp, "u_v_b_pars"):
            buf = custom_mlp_forward(mlp, segs, R).t().contiguous()     # param-major for the layer kernels, as above
            keep.append(buf)
            params, sp, sr = _ptr(buf), R, 1
        else:
            linears = [m for m in mlp if isinstance(m, torch.nn.Linear)]
            wb = [(pack_w, pack_b) for pack_w, pack_b in
                  ((l.weight.detach().to(device=dev, dtype=dt).contiguous(), l.bias.detach().to(device=dev, dtype=dt).contiguous())
                   for l in linears)]
            # a generator may predict more than the flow parameters (the Poisson log-lambda as its last output,
            # predict_log_normalization + join_poisson_and_pdf_description): the layer kernels read the first n_par
            buf = torch.empty(max(n_par, wb[-1][0].shape[0], 1), R, dtype=dt, device=dev)
            keep += [buf, wb]
            _run_chain(lib, dt, dev, wb, segs, buf, R, 1, R, False)
            params, sp, sr = _ptr(buf), R, 1
        t0, t1 = pdf.target_dim_indices[k]
        b0, b1 = pdf.base_dim_indices[k]
        (i0, i1), (o0, o1) = ((t0, t1), (b0, b1)) if logpdf else ((b0, b1), (t0, t1))
        v_in, v_out = src[:, i0:i1], dst[:, o0:o1]
        emb = torch.empty(R, layers[-1]._embedding_conditional_return_num(), dtype=dt, device=dev)
        first = k == 0
        sub_desc = desc.sub[k]
        if only_last:
            if not logpdf and pdf.pdf_defs_list[k][0] not in ("s", "e"):
                raise Exception("Flow type ", pdf.pdf_defs_list[k][0], " does not supported *only_last*!")
            sub_desc, p_off = _last_layer_desc(pdf, k, dt)
            keep.append(sub_desc)
            if params.value is not None:
                params = C.c_void_p(params.value + p_off * sp * (8 if dt == torch.float64 else 4))
        with torch.cuda.device(dev):
            rc = lib.jf_subpdf_apply(C.byref(sub_desc), _DT[dt], direction, _ptr(v_in), src.stride(0), params, sp, sr,
                                     None if first else _ptr(logdet), _ptr(logdet), None if first else _ptr(logbase),
                                     _ptr(logbase), _ptr(v_out), dst.stride(0), _ptr(emb), emb.shape[1], R,
                                     _ptr(status), _stream_ptr(dev))
        _cabi.check(rc, "jf_subpdf_apply")
        prev.append(emb)
    return dst, logdet, logbase


def pdf_logpdf(pdf, x, cond=None, chunk_rows=None, want_base=True, amort=None, only_last=False):
    """-> (log_pdf [B], log_pdf_base [B], base [B, D_base]) on x's device.  Reference: main/default.py:1059-1117."""
    if pdf.amortize_everything and amort is None:
        raise AssertionError("a pdf built with amortize_everything needs amortization_parameters")
    if uses_custom_mlp(pdf) or amort is not None or only_last:
        base, logdet, logbase = _pdf_staged(pdf, x, cond, _cabi.JF_DIR_LOGPDF, amort, only_last)
        return logdet + logbase, logbase, base
    lib = _cabi.load()
    x, cond = _prep_inputs(pdf, x, cond, "x")
    B = x.shape[0]
    dev, dt = x.device, x.dtype
    desc = pdf._desc(dt)
    chunk = int(chunk_rows or min(max(B, 1), DEFAULT_CHUNK_ROWS))
    nbytes = lib.jf_pdf_workspace_bytes(C.byref(desc), chunk)
    ws = _workspace(dev, nbytes)
    logp = torch.empty(B, dtype=dt, device=dev)
    logp_base = torch.empty(B, dtype=dt, device=dev)
    base = torch.empty(B, pdf.total_base_dim, dtype=dt, device=dev) if want_base else None
    pack = ParamPack(pdf, dt, dev)
    status = pdf._status(dev)
    with torch.cuda.device(dev):
        rc = lib.jf_pdf_logpdf(C.byref(desc), C.byref(pack.c), _ptr(x), x.stride(0), _ptr(cond),
                               cond.stride(0) if cond is not None else 0, _ptr(logp), _ptr(logp_base), _ptr(base),
                               pdf.total_base_dim, B, _ptr(ws), ws.numel(), chunk, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_pdf_logpdf")
    return logp, logp_base, base


def pdf_sample(pdf, z, cond=None, chunk_rows=None, amort=None, only_last=False):
    """z [B, D_base] -> (x [B, D], log_pdf [B], log_pdf_base [B]).  Reference: main/default.py:1373-1531, :1533-1707."""
    if pdf.amortize_everything and amort is None:
        raise AssertionError("a pdf built with amortize_everything needs amortization_parameters")
    if uses_custom_mlp(pdf) or amort is not None or only_last:
        xs, logdet, logbase = _pdf_staged(pdf, z, cond, _cabi.JF_DIR_SAMPLE, amort, only_last)
        return xs, logbase - logdet, logbase
    lib = _cabi.load()
    z, cond = _prep_inputs(pdf, z, cond, "base sample")
    B = z.shape[0]
    dev, dt = z.device, z.dtype
    desc = pdf._desc(dt)
    chunk = int(chunk_rows or min(max(B, 1), DEFAULT_CHUNK_ROWS))
    nbytes = lib.jf_pdf_workspace_bytes(C.byref(desc), chunk)
    ws = _workspace(dev, nbytes)
    x = torch.empty(B, pdf.total_target_dim, dtype=dt, device=dev)
    logp = torch.empty(B, dtype=dt, device=dev)
    logp_base = torch.empty(B, dtype=dt, device=dev)
    pack = ParamPack(pdf, dt, dev)
    status = pdf._status(dev)
    with torch.cuda.device(dev):
        rc = lib.jf_pdf_sample(C.byref(desc), C.byref(pack.c), _ptr(z), z.stride(0), _ptr(cond),
                               cond.stride(0) if cond is not None else 0, _ptr(x), pdf.total_target_dim, _ptr(logp),
                               _ptr(logp_base), B, _ptr(ws), ws.numel(), chunk, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_pdf_sample")
    return x, logp, logp_base


def pdf_transform_target(pdf, target, log_det=None, to_embedding=True):
    """Charts of every sub-pdf in one launch: intrinsic -> embedding coordinates (to_embedding) or back.
    Returns (new target, log_det [B]).  Reference: main/default.py:1737-1813, sphere_base.py:242-332."""
    lib = _cabi.load()
    _require_cuda(target, "target")
    if target.dtype not in _DT:
        raise TypeError("jammy_flows_b200 supports float32/float64, got %s" % target.dtype)
    if target.stride(1) != 1:
        target = target.contiguous()
    B = target.shape[0]
    dev, dt = target.device, target.dtype
    n_in = pdf.total_target_dim_intrinsic if to_embedding else pdf.total_target_dim_embedded
    n_out = pdf.total_target_dim_embedded if to_embedding else pdf.total_target_dim_intrinsic
    assert target.shape[1] == n_in, (target.shape[1], n_in)
    desc = pdf._desc(dt)
    out = torch.empty(B, n_out, dtype=dt, device=dev)
    ld_in = None
    if torch.is_tensor(log_det):
        ld_in = log_det.to(device=dev, dtype=dt).expand(B).contiguous()
    ld_out = torch.empty(B, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_pdf_transform_target(C.byref(desc), 1 if to_embedding else 0, _ptr(target), target.stride(0),
                                         _ptr(out), n_out, _ptr(ld_in), _ptr(ld_out), B, _stream_ptr(dev))
    _cabi.check(rc, "jf_pdf_transform_target")
    if log_det is not None and not torch.is_tensor(log_det) and log_det != 0:
        ld_out = ld_out + log_det
    return out, ld_out


def subpdf_logpdf(pdf, k, x_k, cond_segments, embedding_coordinates=False):
    """log p_k(x_k | conditioning) [R] of ONE sub-pdf: its parameter generator (if any) on the column blocks
    `cond_segments` = [conditional input] + embedded earlier targets, then its layer chain in the log_pdf direction.
    x_k: [R, d_k] in default (intrinsic) coordinates, or in embedding coordinates when `embedding_coordinates` (the chart
    and its log-det are then applied first).  This is one step of the reference's
    `all_layer_inverse_individual_subdims` (main/default.py:2713-2901); used by the marginal entropies."""
    lib = _cabi.load()
    _require_cuda(x_k, "x")
    dt, dev = x_k.dtype, x_k.device
    R = x_k.shape[0]
    desc = pdf._desc(dt)
    st = _stream_ptr(dev)
    chart_ld = None
    if embedding_coordinates and pdf.pdf_defs_list[k][0] == "s":
        one = _cabi.JfPdfDesc()
        one.abi_version, one.dtype, one.n_sub = desc.abi_version, desc.dtype, 1
        C.memmove(C.byref(one.sub[0]), C.byref(desc.sub[k]), C.sizeof(_cabi.JfSubPdfDesc))
        d_k = desc.sub[k].dim
        x_in = x_k.contiguous()
        x_k = torch.empty(R, d_k, dtype=dt, device=dev)
        chart_ld = torch.empty(R, dtype=dt, device=dev)
        with torch.cuda.device(dev):
            rc = lib.jf_pdf_transform_target(C.byref(one), 0, _ptr(x_in), x_in.stride(0), _ptr(x_k), d_k, None,
                                             _ptr(chart_ld), R, st)
        _cabi.check(rc, "jf_pdf_transform_target")
    x_k = x_k.contiguous() if x_k.stride(1) != 1 else x_k
    pack = ParamPack(pdf, dt, dev)
    n_par = desc.sub[k].n_params
    if pdf.mlp_predictors[k] is None:
        params, sp, sr = C.c_void_p(pack.c.shared[k]), 1, 0
        keep = None
    else:
        segs = [s if s.stride(1) == 1 else s.contiguous() for s in cond_segments]
        md = _cabi.JfMlpDesc()
        C.memmove(C.byref(md), C.byref(desc.mlp[k]), C.sizeof(md))
        md.n_segments = len(segs)
        for i, sg in enumerate(segs):
            assert sg.shape[0] == R and sg.dtype == dt
            md.seg_cols[i] = sg.shape[1]
        ptrs = (C.c_void_p * len(segs))(*[sg.data_ptr() for sg in segs])
        lds = (C.c_int64 * len(segs))(*[sg.stride(0) for sg in segs])
        # param-major [P, R]; the generator may predict one more value than the layers use (joint log-lambda)
        keep = torch.empty(max(n_par, int(md.dims[md.n_linear]), 1), R, dtype=dt, device=dev)
        nws = lib.jf_mlp_workspace_bytes(C.byref(md), _DT[dt])
        ws = torch.zeros(max(int(nws), 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.jf_mlp_forward_ws(C.byref(md), _DT[dt], ptrs, lds, pack.c.weights[k], pack.c.biases[k], _ptr(keep),
                                       R, 1, R, _ptr(ws), nws, 0, st)
        _cabi.check(rc, "jf_mlp_forward_ws")
        params, sp, sr = _ptr(keep), R, 1
    d_base = pdf.base_dim_indices[k][1] - pdf.base_dim_indices[k][0]
    base = torch.empty(R, d_base, dtype=dt, device=dev)
    logdet = torch.empty(R, dtype=dt, device=dev)
    logbase = torch.empty(R, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_apply(C.byref(desc.sub[k]), _DT[dt], _cabi.JF_DIR_LOGPDF, _ptr(x_k), x_k.stride(0), params, sp,
                                 sr, _ptr(chart_ld), _ptr(logdet), None, _ptr(logbase), _ptr(base), d_base, None, 0, R,
                                 _ptr(pdf._status(dev)), st)
    _cabi.check(rc, "jf_subpdf_apply")
    return logdet + logbase


def normal_rows(n_rows, dim, seed, first_row=0, dtype=torch.float64, device="cuda"):
    """[n_rows, dim] standard normals generated on the device by the counter-based generator (`jf_normal_rows`): row i
    depends only on (seed, first_row + i), so shards of one global batch agree with the unsharded draw."""
    lib = _cabi.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("jammy_flows_b200: the device RNG needs a CUDA device (there is no CPU fallback)")
    out = torch.empty(n_rows, dim, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_normal_rows(_DT[dtype], int(seed) & (2 ** 64 - 1), int(first_row), n_rows, dim, _ptr(out), dim,
                                _stream_ptr(dev))
    _cabi.check(rc, "jf_normal_rows")
    return out


def row_logmeanexp(t):
    """[R, S] -> [R]: log(mean(exp(row))) on the device (jf_row_logmeanexp)."""
    lib = _cabi.load()
    _require_cuda(t, "log-probabilities")
    t = t.contiguous()
    out = torch.empty(t.shape[0], dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        rc = lib.jf_row_logmeanexp(_DT[t.dtype], _ptr(t), t.shape[0], t.shape[1], _ptr(out), _stream_ptr(t.device))
    _cabi.check(rc, "jf_row_logmeanexp")
    return out


_PINNED_OUT = {}


def _pinned_out(tag, shape, dtype, reuse):
    """Pinned host output buffer; with `reuse` the buffer of the previous call with the same shape is handed out again
    (page-locking 2 GB of fresh host memory per call costs more than the copies it receives)."""
    if not reuse:
        return torch.empty(*shape, dtype=dtype, pin_memory=True)
    key = (tag, tuple(shape), dtype)
    buf = _PINNED_OUT.get(key)
    if buf is None:
        for k in [k for k in _PINNED_OUT if k[0] == tag]:
            del _PINNED_OUT[k]                              # one live buffer per role
        buf = _PINNED_OUT[key] = torch.empty(*shape, dtype=dtype, pin_memory=True)
    return buf


def _host_call(pdf, direction, src, cond, chunk_rows, device, reuse_outputs=False):
    """HOST tensors in, HOST tensors out (pinned): the end-to-end path with copies inside the library call.
    reuse_outputs=True returns the SAME pinned output tensors on every call of that shape (they are overwritten by the
    next call): what a serving loop wants."""
    lib = _cabi.load()
    assert not src.is_cuda
    if uses_custom_mlp(pdf) or pdf.amortize_everything:
        raise NotImplementedError("the host-buffer entries describe nn.Sequential generators only; AmortizableMLP / "
                                  "per-sub-pdf conditional inputs / joint log-lambda prediction run through the "
                                  "device-tensor API (pdf.forward / pdf.sample)")
    dt = src.dtype
    dev = torch.device(device)
    B = src.shape[0]
    desc = pdf._desc(dt)
    chunk = int(chunk_rows or min(max(B, 1), DEFAULT_CHUNK_ROWS))   # measured: 2^19 beats 2^18 by 5 % (tools/e2e_probe.py)
    nbytes = lib.jf_pdf_host_workspace_bytes(C.byref(desc), chunk)
    ws = _workspace(dev, nbytes)
    pack = ParamPack(pdf, dt, dev)
    status = pdf._status(dev)
    tag = "lp" if direction == _cabi.JF_DIR_LOGPDF else "s"
    logp = _pinned_out(tag + "_logp", (B,), dt, reuse_outputs)
    logp_base = _pinned_out(tag + "_logp_base", (B,), dt, reuse_outputs)
    src = src.contiguous()
    cond = cond.contiguous() if cond is not None else None
    with torch.cuda.device(dev):
        # the library's copy/compute streams are non-blocking streams: the parameter vector (torch.cat / casts), the
        # zeroed status words and the workspace were prepared on torch's current stream and must be complete first
        torch.cuda.current_stream(dev).synchronize()
        if direction == _cabi.JF_DIR_LOGPDF:
            out = _pinned_out(tag + "_out", (B, pdf.total_base_dim), dt, reuse_outputs)
            rc = lib.jf_pdf_logpdf_host(C.byref(desc), C.byref(pack.c), _ptr(src), src.stride(0), _ptr(cond),
                                        cond.stride(0) if cond is not None else 0, _ptr(logp), _ptr(logp_base), _ptr(out),
                                        pdf.total_base_dim, B, _ptr(ws), ws.numel(), chunk, _ptr(status))
        else:
            out = _pinned_out(tag + "_out", (B, pdf.total_target_dim), dt, reuse_outputs)
            rc = lib.jf_pdf_sample_host(C.byref(desc), C.byref(pack.c), _ptr(src), src.stride(0), _ptr(cond),
                                        cond.stride(0) if cond is not None else 0, _ptr(out), pdf.total_target_dim,
                                        _ptr(logp), _ptr(logp_base), B, _ptr(ws), ws.numel(), chunk, _ptr(status))
    _cabi.check(rc, "jf_pdf_*_host")
    return out, logp, logp_base


def pdf_logpdf_host(pdf, x_host, cond_host=None, chunk_rows=None, device="cuda", reuse_outputs=False):
    base, logp, logp_base = _host_call(pdf, _cabi.JF_DIR_LOGPDF, x_host, cond_host, chunk_rows, device, reuse_outputs)
    return logp, logp_base, base


def pdf_sample_host(pdf, z_host, cond_host=None, chunk_rows=None, device="cuda", reuse_outputs=False):
    return _host_call(pdf, _cabi.JF_DIR_SAMPLE, z_host, cond_host, chunk_rows, device, reuse_outputs)


# ---------------------------------------------------------------------------------------------------------------------
# layer plugin API: a one-layer flow program (reference layers/layer_base.py:58-70)
# ---------------------------------------------------------------------------------------------------------------------
def run_single_layer(layer, direction, x, log_det, extra_inputs, **kw):
    lib = _cabi.load()
    if len(kw) > 0 and any(v for v in kw.values()):
        raise NotImplementedError("layer keyword arguments %s are not supported by the CUDA path yet" % list(kw))
    _require_cuda(x, "layer input")
    dt, dev = x.dtype, x.device
    x = x.contiguous()
    B = x.shape[0]
    sd = _cabi.JfSubPdfDesc()
    fill_subpdf_desc(sd, layer.manifold, layer.dimension, [layer.descriptor()])
    if extra_inputs is None:
        params = layer.packed_permanent_params().detach().to(device=dev, dtype=dt).contiguous()
        sj, sr = 1, 0
    else:
        params = extra_inputs.to(dt).contiguous()
        assert params.shape[1] == layer.total_param_num, (params.shape, layer.total_param_num)
        if params.shape[0] == 1:
            sj, sr = 1, 0
        else:
            assert params.shape[0] == B
            sj, sr = 1, params.stride(0)
    out = torch.empty_like(x)
    ld_in = log_det.to(dt).contiguous() if torch.is_tensor(log_det) else None
    ld_out = torch.empty(B, dtype=dt, device=dev)
    status = torch.zeros(_cabi.JF_STATUS_WORDS, dtype=torch.int64, device=dev)
    d = _cabi.JF_DIR_LOGPDF if direction == "logpdf" else _cabi.JF_DIR_SAMPLE
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_apply(C.byref(sd), _DT[dt], d, _ptr(x), x.stride(0), _ptr(params), sj, sr, _ptr(ld_in),
                                 _ptr(ld_out), None, None, _ptr(out), out.stride(0), None, 0, B, _ptr(status),
                                 _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_apply")
    if not torch.is_tensor(log_det):
        ld_out = ld_out + log_det
    return out, ld_out


def s2_embedding(x):
    """(theta, phi) -> (x, y, z) through the S2 kernel's embedding output (reference sphere_base.py:305-332)."""
    lib = _cabi.load()
    _require_cuda(x, "s2 coordinates")
    dt, dev = x.dtype, x.device
    x = x.contiguous()
    B = x.shape[0]
    sd = _cabi.JfSubPdfDesc()
    fill_subpdf_desc(sd, "s", 2, [dict(code="f", dim=2, add_rotation=0, hh_iter=0, z_sign=-1.0, min_kappa=1e-10,
                                       first=0, n_params=1)])
    params = torch.zeros(1, dtype=dt, device=dev)
    out = torch.empty_like(x)
    emb = torch.empty(B, 3, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_apply(C.byref(sd), _DT[dt], _cabi.JF_DIR_LOGPDF, _ptr(x), x.stride(0), _ptr(params), 1, 0,
                                 None, None, None, None, _ptr(out), out.stride(0), _ptr(emb), 3, B, None,
                                 _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_apply(embedding)")
    return emb


# ---------------------------------------------------------------------------------------------------------------------
# training path (BASELINE configs[4]): differentiable log_pdf for conditional pdfs made of Euclidean "g" sub-pdfs.
# The layer chain runs in the fused kernels in both directions (jf_subpdf_apply / jf_subpdf_backward); the parameter
# generator stays a torch module so that its backward is torch's own (plain library GEMMs, cuBLAS).  Its last Linear is
# evaluated transposed ([P, B] = W2 h^T + b2), which IS the param-major layout the layer kernels read coalesced.
# ---------------------------------------------------------------------------------------------------------------------
_JAC_CODES = ("f", "v", "r", "o", "m")       # non-Euclidean layers: Jacobian by the forward-mode sweep (csrc/jac_sweep.cuh)
_JAC_MAX_PARAMS = 160


def supports_backward(pdf):
    """True when every sub-pdf has a backward path: Euclidean sub-pdfs made of "g" layers with default options and a stage
    the closed-form reverse pass covers and / or "t" layers (`jf_subpdf_forward_backward`), non-Euclidean sub-pdfs ("f", "v", "r", "o", "m",
    any option; "v" / "m" in their closed-form direction) through the dual-number sweep `jf_subpdf_jacobian`.  Parameters
    may come from an MLP (per-row) or be permanent: a permanent vector is expanded to per-row form and autograd sums the
    per-row gradients back into it."""
    if uses_custom_mlp(pdf):
        return False
    for k, layers in enumerate(pdf.layer_list):
        if pdf.pdf_defs_list[k][0] != "e":
            for l in layers:
                if getattr(l, "code", "") not in _JAC_CODES:
                    return False
                if l.code in ("v", "m") and int(getattr(l, "natural_direction", 0)) != 0:
                    return False
            if sum(int(l.get_total_param_num()) for l in layers) > _JAC_MAX_PARAMS:
                return False
            continue
        for l in layers:
            if getattr(l, "code", "") == "t":          # affine layer: reverse pass inside the chain kernel (fb_mvn_backward)
                continue
            if getattr(l, "code", "") != "g" or l.inverse_function_type not in ("isigmoid", "inormal_partly_precise"):
                return False
            if not (l.is_default_kernel_config or _default_without_rotation(l)):
                return False
    return True


def _default_without_rotation(l):
    """default "g" options except rotation_mode="none": the chain kernel's reverse pass simply has no reflections to undo"""
    return (l.nonlinear_stretch_type == "classic" and l.rotation_mode == "none" and not l.add_skewness
            and not l.center_mean and l.width_mode == "smooth" and l.width_clamp is None)


def _embedding_torch(pdf, k, x_k):
    """embedding coordinates of the targets of sub-pdf k with torch ops (differentiable: the conditioning input of the
    later generators, reference sphere_base.py:779-794 / main/default.py:1050-1053)"""
    name = pdf.pdf_defs_list[k]
    if name[0] == "s" and x_k.shape[1] == 2:
        th = x_k[:, 0:1].clamp(1e-7, math.pi - 1e-7)
        ph = x_k[:, 1:2]
        return torch.cat([torch.sin(th) * torch.cos(ph), torch.sin(th) * torch.sin(ph), torch.cos(th)], dim=1)
    if name[0] == "s":
        return torch.cat([torch.cos(x_k), torch.sin(x_k)], dim=1)
    return x_k


def subpdf_logpdf_forward(pdf, k, params_t, x_k):
    """log p_k(x_k | params) = log N(base) + logdet of Euclidean sub-pdf k with per-row (param-major) parameters
    -> (log_pdf_k [B], log_base_k [B], base_k [B, d]).  Body of the op `jammy_b200::subpdf_logpdf` (ops.py)."""
    lib = _cabi.load()
    B, d = x_k.shape
    dt, dev = x_k.dtype, x_k.device
    sub_desc, status = pdf._desc(dt).sub[k], pdf._status(dev)
    base = torch.empty(B, d, dtype=dt, device=dev)
    logdet = torch.empty(B, dtype=dt, device=dev)
    logbase = torch.empty(B, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_apply(C.byref(sub_desc), _DT[dt], _cabi.JF_DIR_LOGPDF, _ptr(x_k), x_k.stride(0),
                                 _ptr(params_t), params_t.stride(0), 1, None, _ptr(logdet), None, _ptr(logbase),
                                 _ptr(base), d, None, 0, B, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_apply")
    return logdet + logbase, logbase, base


def subpdf_logpdf_fb(pdf, k, params_t, x_k, g_logp=None):
    """Forward AND backward of Euclidean sub-pdf k in one kernel (`jf_subpdf_forward_backward`, csrc/gf_fb.cuh)
    -> (log_pdf_k [B], log_base_k [B], base_k [B, d], jac [P, B], jx [B, d]).  With g_logp = None, jac / jx are the
    per-row Jacobians d log_pdf_k[row] / d params[:, row] and / d x_k[row]; the autograd formulas of ops.py scale them by
    the upstream gradient (the tensor-core generator backward takes it as `row_scale`, so the [P, B] block is never
    rescaled in HBM).  Body of the ops `jammy_b200::subpdf_logpdf` / `jammy_b200::generated_logpdf`."""
    lib = _cabi.load()
    B, d = x_k.shape
    dt, dev = x_k.dtype, x_k.device
    sub_desc, status = pdf._desc(dt).sub[k], pdf._status(dev)
    base = torch.empty(B, d, dtype=dt, device=dev)
    logdet = torch.empty(B, dtype=dt, device=dev)
    logbase = torch.empty(B, dtype=dt, device=dev)
    jac = torch.empty_like(params_t)
    jx = torch.empty(B, d, dtype=dt, device=dev)
    if pdf.pdf_defs_list[k][0] != "e":
        # non-Euclidean sub-pdf: values from the layer kernel, Jacobians from the dual-number sweep over the same device code
        assert g_logp is None
        with torch.cuda.device(dev):
            rc = lib.jf_subpdf_apply(C.byref(sub_desc), _DT[dt], _cabi.JF_DIR_LOGPDF, _ptr(x_k), x_k.stride(0),
                                     _ptr(params_t), params_t.stride(0), 1, None, _ptr(logdet), None, _ptr(logbase),
                                     _ptr(base), d, None, 0, B, _ptr(status), _stream_ptr(dev))
            _cabi.check(rc, "jf_subpdf_apply")
            rc = lib.jf_subpdf_jacobian(C.byref(sub_desc), _DT[dt], _ptr(x_k), x_k.stride(0), _ptr(params_t),
                                        params_t.stride(0), 1, _ptr(jac), jac.stride(0), _ptr(jx), d, None, B, _ptr(status),
                                        _stream_ptr(dev))
        if rc == -2:                      # JF_ERR_UNSUPPORTED
            raise NotImplementedError("jf_subpdf_jacobian: no backward for this sub-pdf (more than %d parameters, or an "
                                      "iterative direction)" % _JAC_MAX_PARAMS)
        _cabi.check(rc, "jf_subpdf_jacobian")
        return logdet + logbase, logbase, base, jac, jx
    g = None if g_logp is None else g_logp.contiguous()
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_forward_backward(C.byref(sub_desc), _DT[dt], _ptr(x_k), x_k.stride(0), _ptr(params_t),
                                            params_t.stride(0), 1, _ptr(g), _ptr(jac), _ptr(jx), d, _ptr(base), d,
                                            _ptr(logdet), _ptr(logbase), B, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_forward_backward")
    return logdet + logbase, logbase, base, jac, jx


def subpdf_logpdf_backward(pdf, k, params_t, x_k, g_logp):
    """Gradient of sum(g_logp * log p_k) with respect to the per-row parameters (`jf_subpdf_backward`: the same kernel
    with the forward outputs switched off)."""
    lib = _cabi.load()
    B = x_k.shape[0]
    dt, dev = x_k.dtype, x_k.device
    sub_desc, status = pdf._desc(dt).sub[k], pdf._status(dev)
    grad = torch.empty_like(params_t)
    g = g_logp.contiguous()
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_backward(C.byref(sub_desc), _DT[dt], _ptr(x_k), x_k.stride(0), _ptr(params_t),
                                    params_t.stride(0), 1, _ptr(g), _ptr(grad), B, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_backward")
    return grad


def subpdf_sample_forward(pdf, k, params_t, z_k):
    """x_k = T_k(z_k; params) of Euclidean sub-pdf k with per-row (param-major) parameters (`jf_subpdf_apply`, sampling
    direction) -> (x_k [B, d], log_pdf_k [B], log_base_k [B]).  Body of the op `jammy_b200::subpdf_sample`."""
    lib = _cabi.load()
    B, d = z_k.shape
    dt, dev = z_k.dtype, z_k.device
    sub_desc, status = pdf._desc(dt).sub[k], pdf._status(dev)
    x = torch.empty(B, d, dtype=dt, device=dev)
    logdet = torch.empty(B, dtype=dt, device=dev)
    logbase = torch.empty(B, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_apply(C.byref(sub_desc), _DT[dt], _cabi.JF_DIR_SAMPLE, _ptr(z_k), z_k.stride(0),
                                 _ptr(params_t), params_t.stride(0), 1, None, _ptr(logdet), None, _ptr(logbase),
                                 _ptr(x), d, None, 0, B, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_apply")
    return x, logbase - logdet, logbase


def subpdf_sample_backward(pdf, k, params_t, x_k, z_k, g_x, g_logp):
    """(gradient with respect to the per-row parameters [P, B], cotangent of z_k [B, d]) of a sample x_k given the
    cotangents of x_k and of log_pdf_k.  Euclidean "g" chains: `jf_subpdf_sample_backward` (implicit differentiation
    layer by layer, no root finder).  Non-Euclidean sub-pdfs: the Jacobian blocks of the log_pdf direction at x_k from
    the dual-number sweep (`jf_subpdf_jacobian` with jac_base), then dx/dtheta = -(dbase/dx)^-1 dbase/dtheta on the
    d x d blocks (d <= 2)."""
    lib = _cabi.load()
    B, d = x_k.shape
    dt, dev = x_k.dtype, x_k.device
    sub_desc, status = pdf._desc(dt).sub[k], pdf._status(dev)
    grad = torch.empty_like(params_t)
    g_z = torch.empty(B, d, dtype=dt, device=dev)
    gx = g_x.contiguous()
    gl = g_logp.contiguous()
    if pdf.pdf_defs_list[k][0] != "e":
        P = params_t.shape[0]
        jx = torch.empty(B, d, dtype=dt, device=dev)
        jb = torch.empty(P + d, B, d, dtype=dt, device=dev)
        with torch.cuda.device(dev):
            rc = lib.jf_subpdf_jacobian(C.byref(sub_desc), _DT[dt], _ptr(x_k), x_k.stride(0), _ptr(params_t),
                                        params_t.stride(0), 1, _ptr(grad), grad.stride(0), _ptr(jx), d, _ptr(jb), B,
                                        _ptr(status), _stream_ptr(dev))
        if rc == -2:
            raise NotImplementedError("jf_subpdf_jacobian: no backward for this sub-pdf")
        _cabi.check(rc, "jf_subpdf_jacobian")
        # the sweep differentiates log N(base) + logdet: add base . dbase back to get the logdet part alone
        dld_dp = grad + torch.einsum("pbd,bd->pb", jb[:P], z_k)
        dld_dx = jx + torch.einsum("sbd,bd->bs", jb[P:], z_k)
        Jx = jb[P:].permute(1, 2, 0)                                   # [B, d_out, d_seed] = dbase/dx
        u = gx + gl.unsqueeze(1) * dld_dx
        lam = torch.linalg.solve(Jx.transpose(1, 2), u.unsqueeze(2)).squeeze(2)
        g_params = gl.unsqueeze(0) * dld_dp - torch.einsum("pbd,bd->pb", jb[:P], lam)
        return g_params, lam
    with torch.cuda.device(dev):
        rc = lib.jf_subpdf_sample_backward(C.byref(sub_desc), _DT[dt], _ptr(x_k), x_k.stride(0), _ptr(params_t),
                                           params_t.stride(0), 1, _ptr(gx), gx.stride(0), _ptr(gl), _ptr(grad), _ptr(g_z), d,
                                           B, _ptr(status), _stream_ptr(dev))
    _cabi.check(rc, "jf_subpdf_sample_backward")
    return grad, g_z


def mlp_params_forward(inp, w1, b1, w2, b2):
    """Per-row parameters [P, B] (param-major) of a Linear-tanh-Linear generator with 128 hidden units on the tcgen05
    MLP kernel (`jf_mlp_forward_ws`, the same kernel as inference).  Body of the op `jammy_b200::mlp_params`."""
    lib = _cabi.load()
    dt, dev = inp.dtype, inp.device
    B, P = inp.shape[0], w2.shape[0]
    md = _cabi.JfMlpDesc()
    md.n_linear = 2
    md.dims[0], md.dims[1], md.dims[2] = w1.shape[1], w1.shape[0], P
    md.n_segments = 1
    md.seg_cols[0] = inp.shape[1]
    inp_c = inp if inp.stride(1) == 1 else inp.contiguous()
    ws_ = [w1.detach().contiguous(), w2.detach().contiguous()]
    bs_ = [b1.detach().contiguous(), b2.detach().contiguous()]
    ptrs = (C.c_void_p * 1)(inp_c.data_ptr())
    lds = (C.c_int64 * 1)(inp_c.stride(0))
    wp = (C.c_void_p * 2)(*[t.data_ptr() for t in ws_])
    bp = (C.c_void_p * 2)(*[t.data_ptr() for t in bs_])
    out = torch.empty(P, B, dtype=dt, device=dev)
    nws = lib.jf_mlp_workspace_bytes(C.byref(md), _DT[dt])
    ws = _workspace(dev, max(int(nws), 16))
    with torch.cuda.device(dev):
        rc = lib.jf_mlp_forward_ws(C.byref(md), _DT[dt], ptrs, lds, wp, bp, _ptr(out), B, 1, B, _ptr(ws), nws, 0,
                                   _stream_ptr(dev))
    _cabi.check(rc, "jf_mlp_forward_ws")
    return out


def mlp_params_backward(inp, w1, b1, w2, g, want_inp_grad, row_scale=None):
    """Gradient of the generator given d loss / d params [P, B] = g (* row_scale[None, :] if given).  fp32:
    `jf_mlp_backward` (tensor cores, tf32); fp64: library GEMMs (the fp64 training contract is 1e-8).  The hidden
    activations are recomputed."""
    if g.dtype == torch.float32:
        return _mlp_backward_tc(inp, w1, b1, w2, g, want_inp_grad, row_scale)
    if row_scale is not None:
        g = g * row_scale.unsqueeze(0)
    h = torch.tanh(torch.addmm(b1, inp, w1.t()))         # [B, 128]
    g_w2 = torch.mm(g, h)                                # [P, 128]
    g_b2 = g.sum(dim=1)
    g_h = torch.mm(g.t(), w2)                            # [B, 128]
    g_pre = g_h * (1.0 - h * h)
    g_w1 = torch.mm(g_pre.t(), inp)                      # [128, in]
    g_b1 = g_pre.sum(dim=0)
    # the generator's input may itself carry history (a conditional_input produced by an upstream encoder)
    g_inp = torch.mm(g_pre, w1) if want_inp_grad else None
    return g_inp, g_w1, g_b1, g_w2, g_b2


def _mlp_backward_tc(inp, w1, b1, w2, g, want_inp_grad, row_scale=None):
    """fp32 gradient of the generator on the tensor cores (`jf_mlp_backward`, csrc/mlp_bwd.cuh): h is recomputed, the two
    [P, B]-sized products run as tcgen05 kind::tf32 MMAs, the small ones as FFMA; g is read twice from HBM and nothing of
    its size is written."""
    lib = _cabi.load()
    dev = inp.device
    B, P, n_in = inp.shape[0], w2.shape[0], inp.shape[1]
    md = _cabi.JfMlpDesc()
    md.n_linear = 2
    md.dims[0], md.dims[1], md.dims[2] = n_in, w1.shape[0], P
    md.n_segments = 1
    md.seg_cols[0] = n_in
    g = g if (g.stride(1) == 1 and g.stride(0) % 4 == 0 and g.data_ptr() % 16 == 0) else g.contiguous()
    if g.stride(0) % 4 != 0:                                 # rows of a multiple of 4 floats (16-byte cp.async units)
        gp = torch.zeros(P, (B + 3) // 4 * 4, dtype=g.dtype, device=dev)
        gp[:, :B] = g
        g = gp
    rs = None if row_scale is None else row_scale.to(dtype=g.dtype).contiguous()
    inp_c = inp if inp.stride(1) == 1 else inp.contiguous()
    ws_ = [w1.detach().contiguous(), w2.detach().contiguous()]
    bs_ = [b1.detach().contiguous(), b1.detach().contiguous()]
    wp = (C.c_void_p * 2)(*[t.data_ptr() for t in ws_])
    bp = (C.c_void_p * 2)(*[t.data_ptr() for t in bs_])
    g_w1, g_b1 = torch.zeros_like(ws_[0]), torch.zeros_like(bs_[0])
    g_w2, g_b2 = torch.zeros_like(ws_[1]), torch.zeros(P, dtype=g.dtype, device=dev)
    g_inp = torch.empty(B, n_in, dtype=g.dtype, device=dev) if want_inp_grad else None
    nws = lib.jf_mlp_backward_workspace_bytes(C.byref(md), _cabi.JF_F32, B)
    if nws < 0:
        raise RuntimeError("jf_mlp_backward: shape not eligible (%d -> %d -> %d)" % (n_in, w1.shape[0], P))
    ws = _workspace(dev, max(int(nws), 16))
    with torch.cuda.device(dev):
        rc = lib.jf_mlp_backward(C.byref(md), _cabi.JF_F32, _ptr(inp_c), inp_c.stride(0), wp, bp, _ptr(g), g.stride(0), 1,
                                 _ptr(rs), _ptr(g_w1), _ptr(g_b1), _ptr(g_w2), _ptr(g_b2), _ptr(g_inp),
                                 g_inp.stride(0) if g_inp is not None else 0, B, _ptr(ws), ws.numel(), _stream_ptr(dev))
    _cabi.check(rc, "jf_mlp_backward")
    return g_inp, g_w1, g_b1, g_w2, g_b2


def _tc_mlp_eligible(mlp, dt, dev):
    """Linear-tanh-Linear with 128 hidden units and few enough inputs for the tcgen05 kernel (csrc/mlp_i8.cuh).  The
    kernel reads the weights through raw pointers with the element type of the input: weights of another dtype / on
    another device take the torch path, which raises the same dtype error as the reference."""
    mods = list(mlp)
    if len(mods) != 3 or not isinstance(mods[0], torch.nn.Linear) or not isinstance(mods[2], torch.nn.Linear):
        return False
    if mods[0].out_features != 128:
        return False
    if any(q.dtype != dt or q.device != dev for q in mlp.parameters()):
        return False
    return mods[0].in_features <= (16 if dt == torch.float64 else 96)


def sequential_mlp_forward_trainable(mlp, x):
    """nn.Sequential generator on x [R, in] -> [R, out] WITH autograd history (the tcgen05 forward + GEMM backward of
    the op `jammy_b200::mlp_params` when the shape fits, the torch module otherwise)."""
    _require_cuda(x, "MLP input")
    if _tc_mlp_eligible(mlp, x.dtype, x.device):
        from . import ops  # noqa: F401  (registers torch.ops.jammy_b200.*)
        mods = list(mlp)
        return torch.ops.jammy_b200.mlp_params(x, mods[0].weight, mods[0].bias, mods[2].weight, mods[2].bias).t()
    return mlp(x)


def pdf_logpdf_trainable(pdf, x, cond):
    """-> (log_pdf [B] with autograd history, log_pdf_base [B], base [B, D]).  Reference: main/default.py:1059-1117 with
    `torch.is_grad_enabled()`; the conditioning on earlier sub-pdfs uses the data x (no gradient flows through it)."""
    x, cond = _prep_inputs(pdf, x, cond, "x")
    from . import ops
    dt, dev = x.dtype, x.device
    handle = ops.handle_of(pdf)
    logp, logp_base, bases = None, None, []
    prev = []
    for k, layers in enumerate(pdf.layer_list):
        mlp = pdf.mlp_predictors[k]
        t0, t1 = pdf.target_dim_indices[k]
        x_k = x[:, t0:t1]
        if mlp is None:
            # permanent parameters (the reference's nn.Parameters broadcast over the batch): one vector in extra_inputs
            # order, expanded to the per-row layout of the backward kernel; autograd reduces the row gradients
            vecs = [v for v in (l.packed_permanent_params() for l in layers) if v is not None]
            vec = torch.cat(vecs).to(device=dev, dtype=dt) if len(vecs) > 0 else torch.zeros(0, dtype=dt, device=dev)
            params_t = vec.unsqueeze(1).expand(vec.shape[0], x.shape[0]).contiguous()
        else:
            pieces = ([cond] if cond is not None else []) + prev
            inp = torch.cat(pieces, dim=1) if len(pieces) > 1 else pieces[0]
            mods = list(mlp)
            if _tc_mlp_eligible(mlp, dt, dev):
                # generator + layer chain as ONE autograd node: the forward keeps the per-row Jacobian instead of the
                # parameter block, the backward is the tensor-core generator gradient with the upstream gradient as row scale
                lp_k, lb_k, base_k, _, _ = torch.ops.jammy_b200.generated_logpdf(
                    inp, mods[0].weight, mods[0].bias, mods[2].weight, mods[2].bias, x_k.contiguous(), handle, k)
                logp = lp_k if logp is None else logp + lp_k
                logp_base = lb_k if logp_base is None else logp_base + lb_k
                bases.append(base_k)
                prev.append(_embedding_torch(pdf, k, x_k))
                continue
            else:
                h = inp
                for m in mods[:-1]:
                    h = m(h)
                last = mods[-1]
                params_t = torch.addmm(last.bias.unsqueeze(1), last.weight, h.t())    # [P, B], param-major
        lp_k, lb_k, base_k, _, _ = torch.ops.jammy_b200.subpdf_logpdf(params_t, x_k.contiguous(), handle, k)
        logp = lp_k if logp is None else logp + lp_k
        logp_base = lb_k if logp_base is None else logp_base + lb_k
        bases.append(base_k)
        prev.append(_embedding_torch(pdf, k, x_k))
    return logp, logp_base, torch.cat(bases, dim=1)


def supports_sample_backward(pdf):
    """differentiable sampling: every pdf whose log_pdf has a backward path (Euclidean "g" chains: implicit-function
    reverse pass; non-Euclidean sub-pdfs: Jacobian blocks of the log_pdf direction at the sample)"""
    return supports_backward(pdf)


def pdf_sample_trainable(pdf, z, cond):
    """-> (x [B, D], log_pdf [B], log_pdf_base [B]) WITH autograd history: the reference's `sample(allow_gradients=True)`
    (main/default.py:1342), which differentiates through its bisection / Newton iterations; here the samples come from
    the same kernels as without gradients and the backward is the implicit-function pass `jf_subpdf_sample_backward`."""
    z, cond = _prep_inputs(pdf, z, cond, "z")
    from . import ops
    dt, dev = z.dtype, z.device
    handle = ops.handle_of(pdf)
    logp, logp_base, xs, prev = None, None, [], []
    for k, layers in enumerate(pdf.layer_list):
        mlp = pdf.mlp_predictors[k]
        b0, b1 = pdf.base_dim_indices[k]
        z_k = z[:, b0:b1].contiguous()
        if mlp is None:
            vecs = [v for v in (l.packed_permanent_params() for l in layers) if v is not None]
            vec = torch.cat(vecs).to(device=dev, dtype=dt) if len(vecs) > 0 else torch.zeros(0, dtype=dt, device=dev)
            params_t = vec.unsqueeze(1).expand(vec.shape[0], z.shape[0]).contiguous()
        else:
            pieces = ([cond] if cond is not None else []) + prev
            inp = torch.cat(pieces, dim=1) if len(pieces) > 1 else pieces[0]
            mods = list(mlp)
            if _tc_mlp_eligible(mlp, dt, dev):
                params_t = torch.ops.jammy_b200.mlp_params(inp, mods[0].weight, mods[0].bias, mods[2].weight, mods[2].bias)
            else:
                h = inp
                for m in mods[:-1]:
                    h = m(h)
                last = mods[-1]
                params_t = torch.addmm(last.bias.unsqueeze(1), last.weight, h.t())
        x_k, lp_k, lb_k = torch.ops.jammy_b200.subpdf_sample(params_t, z_k, handle, k)
        logp = lp_k if logp is None else logp + lp_k
        logp_base = lb_k if logp_base is None else logp_base + lb_k
        xs.append(x_k)
        prev.append(_embedding_torch(pdf, k, x_k))
    return torch.cat(xs, dim=1), logp, logp_base
