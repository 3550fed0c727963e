"""`fully_amortized_pdf`: a conditional pdf in which EVERYTHING -- the flow parameters of every sub-pdf and the weights
of the autoregressive parameter generators between them -- is predicted per row by one outer network from the
conditional input.  Same constructor, `forward`, `sample`, `init_params` and `count_parameters` as the reference
(jammy_flows/main/fully_amortized.py:22-278).

How it runs here (no CPU path; everything below launches sm_100a kernels of libjammy_b200.so):
  conditional_input [B, C] --outer generator (AmortizableMLP with permanent parameters: `jf_mlp_forward_acc`, or a plain
  Linear/tanh chain)--> amortization parameters [B, T] --> `pdf(..., amortize_everything=True)`: per sub-pdf the inner
  AmortizableMLP with PER-ROW weights (`jf_rowwise_linear`, HBM-bound: every weight is read once) --> per-row flow
  parameters --> the fused layer-chain kernels (`jf_subpdf_apply`).  Rows are processed in chunks of pdf.chunk_rows, so
  the [chunk, T] parameter block is the only large intermediate.
"""
import numpy
import torch
from torch import nn

from . import engine
from .amortizable_mlp import AmortizableMLP, list_from_str
from .pdf import pdf as _pdf


class fully_amortized_pdf(nn.Module):

    def __init__(self, pdf_defs, flow_defs, options_overwrite=dict(), conditional_input_dim=None,
                 inner_mlp_dims_sub_pdfs="128", inner_mlp_ranks=0, inner_mlp_highway_mode=1,
                 amortization_mlp_dims="128", amortization_mlp_use_custom_mode=True, amortization_mlp_ranks=5,
                 amortization_mlp_highway_mode=0, predict_log_normalization=False, skip_mlp_initialization=False):
        super().__init__()
        self.conditional_input_dim = conditional_input_dim
        assert (type(conditional_input_dim) == int), "Fully amortized PDF requires a single encoding with a single dimension!"
        assert (predict_log_normalization == False), "TODO: Still need to implement log normalization prediction here."
        if skip_mlp_initialization:
            raise NotImplementedError("jammy_flows_b200.fully_amortized_pdf: skip_mlp_initialization is not built")
        self.use_amortizable_mlp = amortization_mlp_use_custom_mode
        # reference main/fully_amortized.py:80-91
        self.pdf_to_amortize = _pdf(pdf_defs, flow_defs, options_overwrite=options_overwrite, conditional_input_dim=None,
                                    amortization_mlp_dims=inner_mlp_dims_sub_pdfs, predict_log_normalization=False,
                                    amortization_mlp_use_custom_mode=True, amortization_mlp_ranks=inner_mlp_ranks,
                                    amortization_mlp_highway_mode=inner_mlp_highway_mode, amortize_everything=True)
        for name in ("pdf_defs_list", "flow_defs_list", "total_target_dim", "target_dim_indices_intrinsic",
                     "target_dim_indices_embedded", "target_dim_indices", "base_dim_indices"):
            setattr(self, name, getattr(self.pdf_to_amortize, name))
        mlp_hidden_dims = list_from_str(amortization_mlp_dims)
        n_out = self.pdf_to_amortize.total_number_amortizable_params
        if self.use_amortizable_mlp:
            self.amortization_mlp = AmortizableMLP(conditional_input_dim, mlp_hidden_dims, n_out,
                                                   low_rank_approximations=amortization_mlp_ranks,
                                                   use_permanent_parameters=True,
                                                   highway_mode=amortization_mlp_highway_mode, svd_mode="smart")
            self.total_param_num = self.amortization_mlp.num_amortization_params
        else:
            par_counter = 0
            mlp_in_dims = [conditional_input_dim] + mlp_hidden_dims
            mlp_out_dims = mlp_hidden_dims + [n_out]
            nn_list = []
            for i in range(len(mlp_in_dims)):
                nn_list.append(torch.nn.Linear(mlp_in_dims[i], mlp_out_dims[i]))
                if i < (len(mlp_in_dims) - 1):
                    nn_list.append(nn.Tanh())
                par_counter += mlp_in_dims[i] * mlp_out_dims[i] + mlp_out_dims[i]
            self.amortization_mlp = torch.nn.Sequential(*nn_list)
            self.total_param_num = par_counter
        self.double()
        self.init_params()

    # the inner pdf holds no parameters; its chunking / RNG switches are forwarded
    @property
    def chunk_rows(self):
        return self.pdf_to_amortize.chunk_rows

    @chunk_rows.setter
    def chunk_rows(self, v):
        self.pdf_to_amortize.chunk_rows = v

    @property
    def rng_mode(self):
        """base normals of sample(): "numpy" (the reference's host RNG, default), "device", "philox" -- see pdf.rng_mode"""
        return self.pdf_to_amortize.rng_mode

    @rng_mode.setter
    def rng_mode(self, v):
        self.pdf_to_amortize.rng_mode = v

    def _chunk(self):
        """rows per pass: the [rows, T] block of amortization parameters stays below ~16 GiB (of 180 GB HBM) unless
        chunk_rows is set; the thread-per-row layer kernels want >= 3e5 rows in flight to fill 148 SMs"""
        if self.chunk_rows:
            return int(self.chunk_rows)
        t = max(1, self.pdf_to_amortize.total_number_amortizable_params)
        return int(max(1024, min(engine.DEFAULT_CHUNK_ROWS, (1 << 34) // (8 * t))))

    def kernel_status(self, reset=True):
        return self.pdf_to_amortize.kernel_status(reset=reset)

    def amortization_parameters(self, conditional_input):
        """conditional_input [B, C] -> [B, total_number_amortizable_params] (the outer generator)."""
        assert (conditional_input is not None), "This is by design a conditional PDF .. we require conditional input!"
        assert (conditional_input.dim() == 2 and conditional_input.shape[1] == self.conditional_input_dim)
        with torch.no_grad():
            if self.use_amortizable_mlp:
                return self.amortization_mlp(conditional_input)
            return engine.sequential_mlp_forward(self.amortization_mlp, conditional_input)

    def forward(self, x, conditional_input=None, force_embedding_coordinates=False, force_intrinsic_coordinates=False):
        """-> (log_pdf [B], log_pdf_base [B], base [B, D]).  Reference main/fully_amortized.py:144-177."""
        assert (conditional_input is not None), "This is by design a conditional PDF .. we require conditional input!"
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        parts = [[], [], []]
        chunk = self._chunk()
        with torch.no_grad():
            for r0 in range(0, max(x.shape[0], 1), chunk):
                c = conditional_input[r0:r0 + chunk]
                res = self.pdf_to_amortize(x[r0:r0 + chunk], amortization_parameters=self.amortization_parameters(c),
                                           force_embedding_coordinates=force_embedding_coordinates,
                                           force_intrinsic_coordinates=force_intrinsic_coordinates)
                for p_, r_ in zip(parts, res):
                    p_.append(r_)
        out = tuple(p_[0] if len(p_) == 1 else torch.cat(p_, dim=0) for p_ in parts)
        if needs_grad:
            from .pdf import _NoBackward
            out = _NoBackward.apply(next(self.parameters()), *out)
        return out

    def sample(self, conditional_input=None, samplesize=1, seed=None, allow_gradients=False,
               force_embedding_coordinates=False, force_intrinsic_coordinates=False):
        """-> (x, base sample, log_pdf, log_pdf_base).  Reference main/fully_amortized.py:180-220 (`samplesize` is
        ignored there as well: one sample per conditional row)."""
        assert (conditional_input is not None), "This is by design a conditional PDF .. we require conditional input!"
        inner, B, chunk = self.pdf_to_amortize, conditional_input.shape[0], self._chunk()
        if B <= chunk:
            return inner.sample(amortization_parameters=self.amortization_parameters(conditional_input), seed=seed,
                                allow_gradients=allow_gradients, force_embedding_coordinates=force_embedding_coordinates,
                                force_intrinsic_coordinates=force_intrinsic_coordinates)
        if allow_gradients:
            raise NotImplementedError("differentiable sampling needs the backward kernels (K8), not built yet")
        # all base normals first (one draw, so the stream does not depend on the chunking), then chunk by chunk
        z = inner._draw_base_normals(B, seed, conditional_input.dtype, conditional_input.device)
        parts = [[], [], []]
        with torch.no_grad():
            for r0 in range(0, B, chunk):
                am = self.amortization_parameters(conditional_input[r0:r0 + chunk])
                xs, _, lp, lb = inner._obtain_sample(predefined_target_input=z[r0:r0 + chunk], amortization_parameters=am,
                                                     force_embedding_coordinates=force_embedding_coordinates,
                                                     force_intrinsic_coordinates=force_intrinsic_coordinates)
                for p_, r_ in zip(parts, (xs, lp, lb)):
                    p_.append(r_)
        xs, lp, lb = (torch.cat(p_, dim=0) for p_ in parts)
        return xs, z, lp, lb

    def init_params(self, data=None, damping_factor=1000.0, mvn_min_max_sv_ratio=1e-4):
        """Reference main/fully_amortized.py:224-253: the desired value of every amortized number becomes the last bias
        of the outer generator, everything in front of it is damped."""
        global_amortization_init = self.pdf_to_amortize.init_params(data=data, damping_factor=damping_factor,
                                                                    mvn_min_max_sv_ratio=mvn_min_max_sv_ratio)
        if self.use_amortizable_mlp:
            self.amortization_mlp.initialize_uvbs(fix_final_bias=global_amortization_init,
                                                  prev_damping_factor=damping_factor)
        else:
            for internal_layer in self.amortization_mlp:
                if hasattr(internal_layer, "weight"):
                    nn.init.kaiming_uniform_(internal_layer.weight.data, a=numpy.sqrt(5))
                    fan_in, _ = nn.init._calculate_fan_in_and_fan_out(internal_layer.weight.data)
                    bound = 1 / numpy.sqrt(fan_in)
                    nn.init.uniform_(internal_layer.bias.data, -bound, bound)
                    internal_layer.weight.data /= damping_factor
                    internal_layer.bias.data /= damping_factor
            self.amortization_mlp[-1].bias.data = global_amortization_init.data.to(self.amortization_mlp[-1].bias.data)

    def count_parameters(self, verbose=False):
        if verbose:
            print("Amoritized PDF param count: \n target PDF pars predicted (not real): %d \n Total PDF (MLP) pars: %d"
                  % (self.pdf_to_amortize.total_number_amortizable_params, self.total_param_num))
        return self.total_param_num
