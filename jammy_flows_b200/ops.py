"""torch.library custom ops over the C-ABI (north_star: "thin C-ABI torch custom-op layer").

    torch.ops.jammy_b200.pdf_logpdf(x, cond, handle, chunk_rows)  -> (log_pdf, log_pdf_base, base)   jf_pdf_logpdf
    torch.ops.jammy_b200.pdf_sample(z, cond, handle, chunk_rows)  -> (x, log_pdf, log_pdf_base)      jf_pdf_sample
    torch.ops.jammy_b200.subpdf_logpdf(params_t, x_k, handle, k)  -> (log_pdf_k, log_base_k, base_k) jf_subpdf_apply
        + register_autograd: gradient with respect to the per-row parameters through jf_subpdf_backward
    torch.ops.jammy_b200.mlp_params(inp, w1, b1, w2, b2)          -> params [P, B]                   jf_mlp_forward_ws
        + register_autograd: jf_mlp_backward (fp32, tensor cores) / library GEMMs (fp64)

ctypes stays the loader of libjammy_b200.so; the ops are what `pdf.forward` / `pdf.sample` and the training path call,
so `torch.compile(pdf)` sees opaque, shape-annotated ops instead of ctypes calls (no graph break at the library
boundary).  A pdf is passed as an integer handle (custom ops take tensors and scalars): the descriptor and the
parameters are looked up from the live module, i.e. they are not graph inputs of the inference ops.
"""
import itertools
import weakref
from typing import Optional, Tuple

import torch

from . import engine

_PDFS = weakref.WeakValueDictionary()
_next_handle = itertools.count(1)


def register(pdf):
    """new integer handle for a live pdf module; `pdf.__init__` and `pdf.__setstate__` (deepcopy / unpickle) call it, the
    module keeps it as the plain int attribute `_op_handle` (a guarded constant for torch.compile)"""
    h = next(_next_handle)
    _PDFS[h] = pdf
    return h


def handle_of(pdf):
    return pdf._op_handle


def _pdf(handle):
    p = _PDFS.get(handle)
    if p is None:
        raise RuntimeError("jammy_b200 op: the pdf behind handle %d no longer exists" % handle)
    return p


# ---------------------------------------------------------------------------------------------------------------------
# whole-pdf inference ops
# ---------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("jammy_b200::pdf_logpdf", mutates_args=(), device_types="cuda")
def pdf_logpdf(x: torch.Tensor, cond: Optional[torch.Tensor], handle: int, chunk_rows: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    p = _pdf(handle)
    with torch.no_grad():
        return engine.pdf_logpdf(p, x, cond, chunk_rows=chunk_rows or None)


@pdf_logpdf.register_fake
def _(x, cond, handle, chunk_rows):
    p = _pdf(handle)
    B = x.shape[0]
    return x.new_empty(B), x.new_empty(B), x.new_empty(B, p.total_base_dim)


@torch.library.custom_op("jammy_b200::pdf_sample", mutates_args=(), device_types="cuda")
def pdf_sample(z: torch.Tensor, cond: Optional[torch.Tensor], handle: int, chunk_rows: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    p = _pdf(handle)
    with torch.no_grad():
        return engine.pdf_sample(p, z, cond, chunk_rows=chunk_rows or None)


@pdf_sample.register_fake
def _(z, cond, handle, chunk_rows):
    p = _pdf(handle)
    B = z.shape[0]
    return z.new_empty(B, p.total_target_dim), z.new_empty(B), z.new_empty(B)


# ---------------------------------------------------------------------------------------------------------------------
# training ops
# ---------------------------------------------------------------------------------------------------------------------
_Out5 = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


@torch.library.custom_op("jammy_b200::subpdf_logpdf", mutates_args=(), device_types="cuda")
def subpdf_logpdf(params_t: torch.Tensor, x_k: torch.Tensor, handle: int, k: int) -> _Out5:
    """-> (log_pdf, log_base, base, jac, jx): forward and per-row Jacobians from one kernel (jf_subpdf_forward_backward)"""
    p = _pdf(handle)
    return engine.subpdf_logpdf_fb(p, k, params_t, x_k)


@subpdf_logpdf.register_fake
def _(params_t, x_k, handle, k):
    B = x_k.shape[0]
    return x_k.new_empty(B), x_k.new_empty(B), x_k.new_empty(B, x_k.shape[1]), torch.empty_like(params_t), torch.empty_like(x_k)


@torch.library.custom_op("jammy_b200::subpdf_logpdf_backward", mutates_args=(), device_types="cuda")
def subpdf_logpdf_backward(params_t: torch.Tensor, x_k: torch.Tensor, g_logp: torch.Tensor, handle: int, k: int) -> torch.Tensor:
    p = _pdf(handle)
    return engine.subpdf_logpdf_backward(p, k, params_t, x_k, g_logp)


@subpdf_logpdf_backward.register_fake
def _(params_t, x_k, g_logp, handle, k):
    return torch.empty_like(params_t)


def _subpdf_setup(ctx, inputs, output):
    ctx.save_for_backward(output[3], output[4])          # the Jacobians; the parameter block itself is not kept
    ctx.set_materialize_grads(False)                     # (no [P, B] block of zeros for the unused Jacobian outputs)


def _subpdf_bwd(ctx, g_logp, g_logbase, g_base, g_jac, g_jx):
    jac, jx = ctx.saved_tensors
    if g_logp is None:                                    # log_pdf unused downstream (only log_base / base, which carry no history)
        return None, None, None, None
    g_params = jac * g_logp.unsqueeze(0) if ctx.needs_input_grad[0] else None
    g_x = jx * g_logp.unsqueeze(1) if ctx.needs_input_grad[1] else None
    return g_params, g_x, None, None


subpdf_logpdf.register_autograd(_subpdf_bwd, setup_context=_subpdf_setup)


@torch.library.custom_op("jammy_b200::generated_logpdf", mutates_args=(), device_types="cuda")
def generated_logpdf(inp: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
                     x_k: torch.Tensor, handle: int, k: int) -> _Out5:
    """Generator (tcgen05 MLP kernel) + layer chain forward/backward kernel of sub-pdf k; the [P, B] parameter block is a
    temporary of this call, what stays for the backward is the Jacobian block."""
    p = _pdf(handle)
    params_t = engine.mlp_params_forward(inp, w1, b1, w2, b2)
    return engine.subpdf_logpdf_fb(p, k, params_t, x_k)


@generated_logpdf.register_fake
def _(inp, w1, b1, w2, b2, x_k, handle, k):
    B = x_k.shape[0]
    return (x_k.new_empty(B), x_k.new_empty(B), x_k.new_empty(B, x_k.shape[1]), inp.new_empty(w2.shape[0], B),
            torch.empty_like(x_k))


def _generated_setup(ctx, inputs, output):
    inp, w1, b1, w2, b2, x_k, handle, k = inputs
    ctx.save_for_backward(inp, w1, b1, w2, output[3], output[4])
    ctx.set_materialize_grads(False)


def _generated_bwd(ctx, g_logp, g_logbase, g_base, g_jac, g_jx):
    inp, w1, b1, w2, jac, jx = ctx.saved_tensors
    if g_logp is None:
        return None, None, None, None, None, None, None, None
    want = ctx.needs_input_grad[0]
    g_inp, g_w1, g_b1, g_w2, g_b2 = torch.ops.jammy_b200.mlp_params_backward_scaled(inp, w1, b1, w2, jac, g_logp.contiguous(), want)
    g_x = jx * g_logp.unsqueeze(1) if ctx.needs_input_grad[5] else None
    return (g_inp if want else None), g_w1, g_b1, g_w2, g_b2, g_x, None, None


generated_logpdf.register_autograd(_generated_bwd, setup_context=_generated_setup)


@torch.library.custom_op("jammy_b200::mlp_params_backward_scaled", mutates_args=(), device_types="cuda")
def mlp_params_backward_scaled(inp: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, jac: torch.Tensor,
                               row_scale: torch.Tensor, want_inp_grad: bool) -> _Out5:
    g_inp, g_w1, g_b1, g_w2, g_b2 = engine.mlp_params_backward(inp, w1, b1, w2, jac, want_inp_grad, row_scale)
    if g_inp is None:
        g_inp = inp.new_zeros(0)
    return g_inp, g_w1, g_b1, g_w2, g_b2


@mlp_params_backward_scaled.register_fake
def _(inp, w1, b1, w2, jac, row_scale, want_inp_grad):
    return (torch.empty_like(inp) if want_inp_grad else inp.new_empty(0), torch.empty_like(w1), torch.empty_like(b1),
            torch.empty_like(w2), w2.new_empty(w2.shape[0]))


@torch.library.custom_op("jammy_b200::subpdf_sample", mutates_args=(), device_types="cuda")
def subpdf_sample(params_t: torch.Tensor, z_k: torch.Tensor, handle: int, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (x_k, log_pdf_k, log_base_k) of Euclidean sub-pdf k in the sampling direction (jf_subpdf_apply)"""
    return engine.subpdf_sample_forward(_pdf(handle), k, params_t, z_k)


@subpdf_sample.register_fake
def _(params_t, z_k, handle, k):
    B = z_k.shape[0]
    return torch.empty_like(z_k), z_k.new_empty(B), z_k.new_empty(B)


@torch.library.custom_op("jammy_b200::subpdf_sample_backward", mutates_args=(), device_types="cuda")
def subpdf_sample_backward(params_t: torch.Tensor, x_k: torch.Tensor, z_k: torch.Tensor, g_x: torch.Tensor,
                           g_logp: torch.Tensor, handle: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    return engine.subpdf_sample_backward(_pdf(handle), k, params_t, x_k, z_k, g_x, g_logp)


@subpdf_sample_backward.register_fake
def _(params_t, x_k, z_k, g_x, g_logp, handle, k):
    return torch.empty_like(params_t), torch.empty_like(x_k)


def _sample_setup(ctx, inputs, output):
    params_t, z_k, handle, k = inputs
    ctx.save_for_backward(params_t, output[0], z_k)
    ctx.handle, ctx.k = handle, k
    ctx.set_materialize_grads(False)


def _sample_bwd(ctx, g_x, g_logp, g_logbase):
    params_t, x_k, z_k = ctx.saved_tensors
    if g_x is None and g_logp is None and g_logbase is None:
        return None, None, None, None
    g_x = torch.zeros_like(x_k) if g_x is None else g_x
    g_logp = x_k.new_zeros(x_k.shape[0]) if g_logp is None else g_logp
    g_logbase = x_k.new_zeros(x_k.shape[0]) if g_logbase is None else g_logbase
    g_params, g_z = torch.ops.jammy_b200.subpdf_sample_backward(params_t, x_k, z_k, g_x.contiguous(), g_logp.contiguous(),
                                                               ctx.handle, ctx.k)
    if ctx.needs_input_grad[1]:
        # log_pdf = log N(z) + sum of the layers' log-derivatives: the base density adds -z per unit of log_pdf / log_base
        g_z = g_z - z_k * (g_logp + g_logbase).unsqueeze(1)
    return g_params, (g_z if ctx.needs_input_grad[1] else None), None, None


subpdf_sample.register_autograd(_sample_bwd, setup_context=_sample_setup)


@torch.library.custom_op("jammy_b200::mlp_params", mutates_args=(), device_types="cuda")
def mlp_params(inp: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    return engine.mlp_params_forward(inp, w1, b1, w2, b2)


@mlp_params.register_fake
def _(inp, w1, b1, w2, b2):
    return inp.new_empty(w2.shape[0], inp.shape[0])


@torch.library.custom_op("jammy_b200::mlp_params_backward", mutates_args=(), device_types="cuda")
def mlp_params_backward(inp: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, g: torch.Tensor,
                        want_inp_grad: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    g_inp, g_w1, g_b1, g_w2, g_b2 = engine.mlp_params_backward(inp, w1, b1, w2, g, want_inp_grad)
    if g_inp is None:
        g_inp = inp.new_zeros(0)
    return g_inp, g_w1, g_b1, g_w2, g_b2


@mlp_params_backward.register_fake
def _(inp, w1, b1, w2, g, want_inp_grad):
    return (torch.empty_like(inp) if want_inp_grad else inp.new_empty(0), torch.empty_like(w1), torch.empty_like(b1),
            torch.empty_like(w2), w2.new_empty(w2.shape[0]))


def _mlp_setup(ctx, inputs, output):
    inp, w1, b1, w2, b2 = inputs
    ctx.save_for_backward(inp, w1, b1, w2)


def _mlp_bwd(ctx, g):
    inp, w1, b1, w2 = ctx.saved_tensors
    want = ctx.needs_input_grad[0]
    g_inp, g_w1, g_b1, g_w2, g_b2 = torch.ops.jammy_b200.mlp_params_backward(inp, w1, b1, w2, g.contiguous(), want)
    return (g_inp if want else None), g_w1, g_b1, g_w2, g_b2


mlp_params.register_autograd(_mlp_bwd, setup_context=_mlp_setup)
