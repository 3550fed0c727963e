"""`jammy_flows_b200.pdf` -- drop-in for `jammy_flows.pdf` on the hot path (log_pdf / sampling / total entropy).

Mirrors the reference class `jammy_flows/main/default.py:42-151`: same constructor signature, the same string DSL
("e4+s2+e4", "gggg+f+gggg"), the same option-override semantics, index tables, MLP wiring, parameter names/shapes
(`layer_list.{sub}.{layer}.*`, `mlp_predictors.{k}.{0,2}.{weight,bias}`) and return tuples.  What differs is the
execution: instead of a Python loop over layer modules issuing hundreds of eager torch ops per layer, the layer graph
is compiled once into a static flow program (C structs, include/jammy_b200.h) and executed by fused sm_100a kernels
through the C-ABI.  There is no CPU / eager fallback: tensors must live on a CUDA device.
"""
import collections
import copy

import numpy
import torch
from torch import nn

from . import engine
from .flow_options import check_flow_option, obtain_default_options, obtain_overall_flow_info


def list_from_str(spec):
    """Reference: extra_functions.py:90-94."""
    if spec == "":
        return []
    return list(tuple(map(int, spec.split("-"))))


class _NoBackward(torch.autograd.Function):
    """Marks outputs of the inference kernels: calling backward() fails loudly instead of silently giving no grads."""

    @staticmethod
    def forward(ctx, anchor, *outs):
        return tuple(o.view_as(o) for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            "jammy_flows_b200: backward kernels exist for pdfs made of Euclidean 'g' sub-pdfs with default options "
            "(isigmoid / inormal_partly_precise stages; permanent or MLP-generated parameters; SURVEY.md K8 / cfg5); "
            "this pdf contains other layers or options, whose outputs are inference-only in this round.")


class pdf(nn.Module):

    def __init__(self, pdf_defs, flow_defs, options_overwrite=dict(), conditional_input_dim=None,
                 amortization_mlp_dims="128", predict_log_normalization=False, join_poisson_and_pdf_description=False,
                 hidden_mlp_dims_poisson="128", rank_of_mlp_mappings_poisson=0, amortization_mlp_use_custom_mode=False,
                 amortization_mlp_ranks=0, amortization_mlp_highway_mode=0, amortize_everything=False,
                 use_as_passthrough_instead_of_pdf=False, skip_mlp_initialization=False, verbose=False):
        """Same parameters as the reference constructor (main/default.py:44-100)."""
        super().__init__()
        not_built = []
        if predict_log_normalization and join_poisson_and_pdf_description and amortization_mlp_use_custom_mode:
            # the reference's own init fails for this combination (main/default.py:1897 indexes the AmortizableMLP)
            not_built.append("joint log-lambda prediction with an AmortizableMLP generator")
        if predict_log_normalization and conditional_input_dim is not None and not join_poisson_and_pdf_description:
            # the reference builds a separate log-lambda MLP here but its log_mean_poisson() raises "outdated"
            # (main/default.py:876-877)
            not_built.append("predict_log_normalization with a separate log-lambda MLP (join_poisson_and_pdf_description=False)")
        if use_as_passthrough_instead_of_pdf:
            not_built.append("use_as_passthrough_instead_of_pdf")
        if skip_mlp_initialization:
            not_built.append("skip_mlp_initialization")
        if len(not_built) > 0:
            raise NotImplementedError("jammy_flows_b200.pdf: outside the hot path built so far: " + "; ".join(not_built))

        self.amortization_mlp_use_custom_mode = amortization_mlp_use_custom_mode
        self.predict_log_normalization = predict_log_normalization
        self.join_poisson_and_pdf_description = join_poisson_and_pdf_description
        self.amortization_mlp_highway_mode = amortization_mlp_highway_mode
        self.amortize_everything = amortize_everything
        self.use_as_passthrough_instead_of_pdf = use_as_passthrough_instead_of_pdf
        self.skip_mlp_initialization = skip_mlp_initialization
        self.total_number_amortizable_params = None
        if self.amortize_everything:
            # reference main/default.py:109-120
            assert (self.predict_log_normalization == False), "Log Poisson prediction works only without full amortization in the default PDF. It can be used in the *fully_amortized_pdf*!"
            assert (self.amortization_mlp_use_custom_mode), "Amortizing all MLPs requires custom MLPs."
            self.total_number_amortizable_params = 0

        self.read_model_definition(pdf_defs, flow_defs, options_overwrite, conditional_input_dim, amortization_mlp_dims,
                                   amortization_mlp_ranks, verbose=verbose)
        self.init_flow_structure()
        self.init_encoding_structure()
        self.init_params()

        self._desc_cache = {}
        self._status_cache = {}
        # RNG used by sample(): "numpy" reproduces the reference's host RNG bit for bit (main/default.py:1661-1668);
        # "device" draws the normals with torch's Philox generator on the GPU (no host round trip).
        # base-space normals of sample(): "numpy" = the reference's host RNG (bit-identical draws for a seed), "device" =
        # torch's CUDA generator, "philox" = the library's counter-based generator (row i = f(seed, rng_first_row + i))
        self.rng_mode = "numpy"
        self.rng_first_row = 0
        self.chunk_rows = None
        from . import ops
        self._op_handle = ops.register(self)       # handle under which the torch.library ops (ops.py) find this module

    def __setstate__(self, state):
        # deepcopy / unpickle: the copy is a different module and needs a handle of its own
        super().__setstate__(state)
        from . import ops
        self._op_handle = ops.register(self)

    # ------------------------------------------------------------------------------------------------------------------
    # model definition (reference main/default.py:153-325)
    # ------------------------------------------------------------------------------------------------------------------
    def read_model_definition(self, pdf_defs, flow_defs, options_overwrite, conditional_input_dim, amortization_mlp_dims,
                              amortization_mlp_ranks, verbose=False):
        self.pdf_defs_list = pdf_defs.split("+")
        self.flow_defs_list = flow_defs.split("+")
        self.flow_opts = dict()
        for ind, cur_flow_defs in enumerate(self.flow_defs_list):
            self.flow_opts[ind] = []
            for cur_flow_index, flow_abbrv in enumerate(cur_flow_defs):
                opts = obtain_default_options(flow_abbrv)
                for opt in opts.keys():
                    check_flow_option(flow_abbrv, opt, opts[opt])
                # three specificities, most specific wins: (sub, layer) tuple > sub index > layer code
                found_specific = False
                for k in options_overwrite.keys():
                    if type(k) == tuple:
                        assert (type(k[0]) == int and type(k[1]) == int), \
                            "Require 2 ints for tuple-based flow definition! The first indexes the sub-manifold, the second the flow within the manifold."
                        assert ((k[0] >= 0) and (k[0] < len(self.flow_defs_list))), \
                            "Index of detailed options is outside allowed range of defined autoregressive structure."
                        if k[0] != ind or k[1] != cur_flow_index:
                            continue
                        assert (len(options_overwrite[k]) == 1), "We have detailed flow definition per item, require length of 1 here."
                        found_specific = True
                        for detail_abbrv in options_overwrite[k].keys():
                            assert (detail_abbrv == flow_abbrv)
                            for detail_opt, val in options_overwrite[k][detail_abbrv].items():
                                check_flow_option(flow_abbrv, detail_opt, val)
                                opts[detail_opt] = val
                if not found_specific:
                    for k in options_overwrite.keys():
                        if type(k) == int:
                            assert ((k >= 0) and (k < len(self.flow_defs_list))), \
                                "Index of detailed options is outside allowed range of defined autoregressive structure."
                            if k != ind:
                                continue
                            for detail_abbrv in options_overwrite[k].keys():
                                if detail_abbrv == flow_abbrv:
                                    found_specific = True
                                    for detail_opt, val in options_overwrite[k][detail_abbrv].items():
                                        check_flow_option(flow_abbrv, detail_opt, val)
                                        opts[detail_opt] = val
                if not found_specific:
                    for k in options_overwrite.keys():
                        if k == flow_abbrv:
                            for detail_opt, val in options_overwrite[k].items():
                                check_flow_option(flow_abbrv, detail_opt, val)
                                opts[detail_opt] = val
                self.flow_opts[ind].append(opts)
        if len(self.pdf_defs_list) != len(self.flow_defs_list):
            raise Exception("PDF defs list has to be same length as flow defs list, but ... ", self.pdf_defs_list,
                            self.flow_defs_list)
        self.conditional_input_dim = conditional_input_dim
        self.encoding_type = "single"
        self.amortization_mlp_dims = amortization_mlp_dims
        if type(self.amortization_mlp_dims) == str:
            self.amortization_mlp_dims = [self.amortization_mlp_dims] * len(self.pdf_defs_list)
        elif type(self.amortization_mlp_dims) != list:
            raise Exception("Hidden MLP dimensions must be defined either str or list, received ",
                            type(self.amortization_mlp_dims))
        self.amortization_mlp_ranks = amortization_mlp_ranks
        if type(self.amortization_mlp_ranks) in (int, str):
            self.amortization_mlp_ranks = [self.amortization_mlp_ranks] * len(self.pdf_defs_list)
        elif type(self.amortization_mlp_ranks) != list:
            raise Exception("Rank of MLP sub pdfs has to defined as an int or list type!")
        if len(self.amortization_mlp_dims) != len(self.pdf_defs_list):
            raise Exception("hidden mlp dimension definitions for sub pdfs is wrong length (%d) .. requires length (%d)"
                            % (len(self.amortization_mlp_dims), len(self.pdf_defs_list)))
        self.layer_list = nn.ModuleList()
        self.force_permanent_parameters_in_first_subpdf = 0
        if self.conditional_input_dim is None:
            if self.amortize_everything == False:          # reference main/default.py:323-325
                self.force_permanent_parameters_in_first_subpdf = 1

    # ------------------------------------------------------------------------------------------------------------------
    # layer graph (reference main/default.py:378-479)
    # ------------------------------------------------------------------------------------------------------------------
    def init_flow_structure(self):
        self.num_parameter_list = []
        flow_info = obtain_overall_flow_info()
        for subflow_index, subflow_description in enumerate(self.pdf_defs_list):
            self.num_parameter_list.append([])
            self.layer_list.append(nn.ModuleList())
            this_num_layers = len(self.flow_defs_list[subflow_index])
            for layer_ind, layer_type in enumerate(self.flow_defs_list[subflow_index]):
                if flow_info[layer_type]["type"] != subflow_description[0]:
                    raise Exception("layer type ", layer_type, " is not compatible with flow type ", subflow_description)
                this_kwargs = copy.deepcopy(self.flow_opts[subflow_index][layer_ind])
                this_kwargs["use_permanent_parameters"] = \
                    1 if (self.force_permanent_parameters_in_first_subpdf and subflow_index == 0) else 0
                if "s" in subflow_description:
                    this_kwargs["euclidean_to_sphere_as_first"] = 1 if layer_ind == 0 else 0
                elif "i" in subflow_description:
                    # reference main/default.py:414-432: boundaries from the sub-pdf definition "i1_lo_hi"
                    interval_boundaries = subflow_description.split("_")[1:]
                    if len(interval_boundaries) == 0:
                        this_kwargs["low_boundary"], this_kwargs["high_boundary"] = 0.0, 1.0
                    else:
                        this_kwargs["low_boundary"] = float(interval_boundaries[0])
                        this_kwargs["high_boundary"] = float(interval_boundaries[1])
                    this_kwargs["euclidean_to_interval_as_first"] = 1 if layer_ind == 0 else 0
                elif "e" in subflow_description:
                    # last layer models the offset; a first (non-last) "g" layer swaps isigmoid for the inverse normal
                    # CDF -- a single-layer sub-pdf therefore keeps isigmoid (reference :440-448)
                    if layer_ind == (this_num_layers - 1) and this_kwargs["skip_model_offset"] == 0:
                        this_kwargs["model_offset"] = 1
                    elif layer_ind == 0:
                        if layer_type == "g":
                            if this_kwargs["replace_first_sigmoid_with_icdf"] > 0 and \
                                    this_kwargs["inverse_function_type"] == "isigmoid":
                                this_kwargs["inverse_function_type"] = "inormal_partly_precise"
                else:
                    raise NotImplementedError("manifold type '%s' has no sm_100a kernel yet" % subflow_description)
                if "skip_model_offset" in this_kwargs:
                    del this_kwargs["skip_model_offset"]
                if layer_type == "g":
                    del this_kwargs["replace_first_sigmoid_with_icdf"]
                dim = int(subflow_description.split("_")[0][1:])
                self.layer_list[subflow_index].append(flow_info[layer_type]["module"](dim, **this_kwargs))
                self.num_parameter_list[subflow_index].append(self.layer_list[subflow_index][-1].get_total_param_num())
        # Poisson log-mean prediction (reference main/default.py:466-477): a free parameter for an unconditional pdf; for a
        # conditional one the first generator predicts it as one extra (last) output
        self.log_normalization = None
        if self.predict_log_normalization:
            assert (len(self.pdf_defs_list) == 1), "You chose to predict log-lambda, which is only allowed with a single sub-pdf (no autoregressive structure). For autoregressive PDFs with log-lambda prediction, use fully amortized PDFs."
            if self.force_permanent_parameters_in_first_subpdf:
                self.log_normalization = nn.Parameter(torch.randn(1).unsqueeze(0))
            else:
                self.log_normalization = torch.zeros(1).unsqueeze(0)
        self.update_embedding_structure()

    # reference main/default.py:481-567
    def update_embedding_structure(self):
        self.target_dims_intrinsic, self.target_dims_embedded, self.target_dims = [], [], []
        self.target_dim_indices_intrinsic, self.target_dim_indices_embedded = [], []
        self.target_dim_indices, self.base_dim_indices = [], []
        tot_i = tot_e = tot = tot_b = 0
        for ll in self.layer_list:
            intr = [l.get_layer_intrinsic_target_dimension() for l in ll]
            emb = [l.get_layer_embedded_target_dimension() for l in ll]
            base = [l.get_layer_base_dimension() for l in ll]
            use_emb = any(l.always_parametrize_in_embedding_space for l in ll)
            for i in range(len(intr) - 1):
                assert (intr[i] == intr[i + 1])
            self.target_dims_intrinsic.append(intr[-1])
            self.target_dims_embedded.append(emb[-1])
            self.target_dims.append(emb[-1] if use_emb else intr[-1])
            self.base_dim_indices.append((tot_b, tot_b + base[0]))
            tot_b += base[0]
            self.target_dim_indices_intrinsic.append((tot_i, tot_i + intr[-1]))
            tot_i += intr[-1]
            self.target_dim_indices_embedded.append((tot_e, tot_e + emb[-1]))
            tot_e += emb[-1]
            self.target_dim_indices.append((tot, tot + self.target_dims[-1]))
            tot += self.target_dims[-1]
        self.total_target_dim_intrinsic = tot_i
        self.total_target_dim_embedded = tot_e
        self.total_target_dim = tot
        self.total_base_dim = tot_b
        self._desc_cache = {}

    # ------------------------------------------------------------------------------------------------------------------
    # parameter generators (reference main/default.py:571-673): one nn.Sequential(Linear, Tanh, ..., Linear) per
    # sub-pdf that has parameters and an input; in = cond_dim + embeddings of all previous sub-pdfs.
    # ------------------------------------------------------------------------------------------------------------------
    def init_encoding_structure(self):
        self.mlp_predictors = nn.ModuleList()
        self.log_normalization_mlp = None
        prev_extra_input_num = 0
        if self.join_poisson_and_pdf_description:
            # reference main/default.py:580-584
            if len(self.pdf_defs_list) > 1:
                raise Exception("A common poisson log-lambda and flow parameter prediction is currently only supported for a PDF that has a single flow (no autoregressive structure) for simplicity! .. number of autoregressive parts here: ", len(self.pdf_defs_list))
            if self.conditional_input_dim is None:
                raise Exception("Flow does not depend on conditional input .. please set 'join_poisson_and_pdf_description' to False, currently True")
        for pdf_index, _ in enumerate(self.pdf_defs_list):
            emb_num = self.layer_list[pdf_index][-1]._embedding_conditional_return_num()
            if pdf_index == 0 and self.conditional_input_dim is None:
                self.mlp_predictors.append(None)
                prev_extra_input_num += emb_num
                if self.amortize_everything:
                    # no generator in front of the first sub-pdf: its flow parameters are amortized directly
                    # (reference main/default.py:595-598)
                    self.total_number_amortizable_params += sum(self.num_parameter_list[0])
                continue
            num_predicted_pars = sum(self.num_parameter_list[pdf_index])
            if self.predict_log_normalization and pdf_index == 0 and self.join_poisson_and_pdf_description:
                num_predicted_pars += 1             # log-lambda is the LAST output of the first generator (:624-626)
            if num_predicted_pars == 0:
                self.mlp_predictors.append(None)
                prev_extra_input_num += emb_num
                continue
            this_summary_dim = prev_extra_input_num
            if self.conditional_input_dim is not None:
                # an int: one conditional input for all sub-pdfs; a list: one per sub-pdf (reference :637-641)
                this_summary_dim += (self.conditional_input_dim if type(self.conditional_input_dim) == int
                                     else self.conditional_input_dim[pdf_index])
            hidden = list_from_str(self.amortization_mlp_dims[pdf_index])
            if self.amortization_mlp_use_custom_mode:
                # reference main/default.py:643-651: AmortizableMLP with permanent parameters
                from .amortizable_mlp import AmortizableMLP
                self.mlp_predictors.append(AmortizableMLP(
                    this_summary_dim, hidden, num_predicted_pars,
                    low_rank_approximations=self.amortization_mlp_ranks[pdf_index],
                    use_permanent_parameters=self.amortize_everything == False,
                    highway_mode=self.amortization_mlp_highway_mode, svd_mode="smart"))
                if self.amortize_everything:
                    self.total_number_amortizable_params += self.mlp_predictors[-1].num_amortization_params
                prev_extra_input_num += emb_num
                continue
            mlp_in_dims = [this_summary_dim] + hidden
            mlp_out_dims = hidden + [num_predicted_pars]
            nn_list = []
            for i in range(len(mlp_in_dims)):
                nn_list.append(torch.nn.Linear(mlp_in_dims[i], mlp_out_dims[i]))
                if i < (len(mlp_in_dims) - 1):
                    nn_list.append(nn.Tanh())
            self.mlp_predictors.append(torch.nn.Sequential(*nn_list))
            prev_extra_input_num += emb_num

    # ------------------------------------------------------------------------------------------------------------------
    # initialisation (reference main/default.py:1817-1952 and extra_functions.py:179-409 with data=None):
    # desired layer params per sub-pdf; MLP weights kaiming-uniform / damping_factor, last bias := desired params.
    # The RNG call order is kept so that equal seeds give the reference's parameters.
    # ------------------------------------------------------------------------------------------------------------------
    def init_params(self, data=None, damping_factor=1000.0, mvn_min_max_sv_ratio=1e-4):
        from .init_fns import find_init_pars_of_chained_blocks
        # amortize_everything: nothing is stored here; the desired values of the whole per-row vector are returned
        # (reference main/default.py:1828-1832, :1900-1904, :1944-1946) and end up in the last bias of the outer generator
        global_amortization_init, global_amortization_index = None, 0
        if self.amortize_everything:
            global_amortization_init = torch.zeros(self.total_number_amortizable_params)
        with torch.no_grad():
            if data is not None:
                assert (data.shape[1] == self.total_target_dim), "Initialization with data must match the target dimension of the PDF!"
            params_list = []
            this_dim_index = 0
            for subflow_index, subflow_description in enumerate(self.pdf_defs_list):
                this_layer_list = self.layer_list[subflow_index]
                this_dim = self.target_dims[subflow_index]
                if "e" in subflow_description:
                    # Euclidean chains can be initialised from data (reference extra_functions.py:179-409)
                    sub_data = data[:, this_dim_index:this_dim_index + this_dim] if data is not None else None
                    params_list.append(find_init_pars_of_chained_blocks(this_layer_list, sub_data,
                                                                        mvn_min_max_sv_ratio=mvn_min_max_sv_ratio))
                    this_dim_index += this_dim
                    continue
                this_dim_index += this_dim
                if True:
                    params_list.append(torch.cat([l.get_desired_init_parameters() for l in this_layer_list]))
            for ind, mlp_predictor in enumerate(self.mlp_predictors):
                these_params = params_list[ind]
                if len(these_params) == 0:
                    continue
                if mlp_predictor is not None and self.predict_log_normalization and \
                        self.join_poisson_and_pdf_description and ind == 0:
                    # desired initial log-lambda 0.1 behind the flow parameters (reference main/default.py:1893-1897)
                    these_params = torch.cat([these_params, torch.Tensor([0.1]).type(these_params.dtype)])
                if mlp_predictor is not None and hasattr(mlp_predictor, "initialize_uvbs") and self.amortize_everything:
                    n_uvb = mlp_predictor.num_amortization_params
                    global_amortization_init[global_amortization_index:global_amortization_index + n_uvb] = \
                        mlp_predictor.obtain_default_init_tensor(fix_final_bias=these_params,
                                                                 prev_damping_factor=damping_factor)
                    global_amortization_index += n_uvb
                elif mlp_predictor is not None and hasattr(mlp_predictor, "initialize_uvbs"):
                    # custom low-rank MLPs initialise themselves (reference main/default.py:1896-1904)
                    mlp_predictor.initialize_uvbs(fix_final_bias=these_params, prev_damping_factor=damping_factor)
                elif mlp_predictor is not None:
                    for internal_layer in mlp_predictor:
                        if hasattr(internal_layer, "weight"):
                            nn.init.kaiming_uniform_(internal_layer.weight.data, a=numpy.sqrt(5))
                            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(internal_layer.weight.data)
                            bound = 1 / numpy.sqrt(fan_in)
                            nn.init.uniform_(internal_layer.bias.data, -bound, bound)
                            internal_layer.weight.data /= damping_factor
                            internal_layer.bias.data /= damping_factor
                    mlp_predictor[-1].bias.data = these_params.data.type(mlp_predictor[-1].bias.data.dtype)
                else:
                    tot_param_index = 0
                    for layer in self.layer_list[ind]:
                        n = layer.get_total_param_num()
                        if self.amortize_everything == False:
                            layer.init_params(these_params[tot_param_index:tot_param_index + n])
                        tot_param_index += n
                    if self.amortize_everything:
                        global_amortization_init[global_amortization_index:global_amortization_index + tot_param_index] = these_params
                        global_amortization_index += tot_param_index
        self._desc_cache = {}
        return global_amortization_init

    # reference main/default.py:724-830
    def count_parameters(self, verbose=False):
        n = 0
        for p in self.parameters():
            if p.requires_grad:
                n += int(numpy.prod(p.size()))
        if verbose:
            print("total Conditional PDF pars: %d" % n)
        return n

    # ------------------------------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------------------------------
    def _desc(self, dtype):
        d = self._desc_cache.get(dtype)
        if d is None:
            d = engine.compile_pdf(self, dtype)
            self._desc_cache[dtype] = d
        return d

    def _status(self, device):
        key = (device.type, device.index)
        s = self._status_cache.get(key)
        if s is None:
            s = torch.zeros(4, dtype=torch.int64, device=device)
            self._status_cache[key] = s
        return s

    def kernel_status(self, reset=True):
        """Lazily read the device status counters (non-finite / unconverged / out-of-range / root-finder evaluations).
        This is the only place that synchronises; it replaces the reference's per-iteration host checks
        (layers/bisection_n_newton.py:95-133, main/default.py:1516)."""
        out = dict(nonfinite=0, unconverged=0, out_of_range=0, evaluations=0)
        for s in self._status_cache.values():
            v = s.cpu().tolist()
            out["nonfinite"] += v[0]
            out["unconverged"] += v[1]
            out["out_of_range"] += v[2]
            out["evaluations"] += v[3]
            if reset:
                s.zero_()
        return out

    def obtain_current_dtype_n_device(self):
        """Reference main/default.py:3275-3288."""
        try:
            first = next(self.parameters())
        except StopIteration:
            return None, None
        return first.dtype, first.device

    def get_total_embedding_dim(self):
        return sum(ll[-1]._embedding_conditional_return_num() for ll in self.layer_list)

    def _needs_transform(self):
        return any(p[0] != "e" for p in self.pdf_defs_list)

    def _anchor(self):
        p = next(self.parameters(), None)
        return p if p is not None else torch.zeros(1)

    # ------------------------------------------------------------------------------------------------------------------
    # log_pdf (reference main/default.py:1059-1117)
    # ------------------------------------------------------------------------------------------------------------------
    def forward(self, x, conditional_input=None, amortization_parameters=None, force_embedding_coordinates=False,
                force_intrinsic_coordinates=False, only_last=False):
        assert (self.use_as_passthrough_instead_of_pdf == False)
        self._check_amortization_parameters(amortization_parameters, x.shape[0])
        if type(conditional_input) == list:
            # one conditional input per sub-pdf (reference main/default.py:1092-1103)
            assert (type(self.conditional_input_dim) == list and len(self.conditional_input_dim) == len(conditional_input))
            for ci_ind, ci in enumerate(conditional_input):
                assert (self.conditional_input_dim[ci_ind] == ci.shape[1]), "Inputs of conditional input vector do not match with pre-defined input_dims!"
                assert (x.shape[0] == ci.shape[0]), "Evaluating input x and condititional input shape must be similar!"
        elif conditional_input is not None:
            assert (x.shape[0] == conditional_input.shape[0]), "Evaluating input x and condititional input shape must be similar!"
            assert (x.is_cuda == conditional_input.is_cuda), "input tensor *x* and *conditional_input* are on different devices"
            assert (self.conditional_input_dim == conditional_input.shape[1])
        else:
            assert self.conditional_input_dim is None, "conditional pdf requires conditional_input"
        chart_log_det = None
        if force_embedding_coordinates:
            # reference main/default.py:906-909: embedding -> default coordinates first, its log-det joins the flow's
            assert (x.shape[1] == self.total_target_dim_embedded), (x.shape[1], self.total_target_dim_embedded)
            if self._needs_transform():
                x, chart_log_det = engine.pdf_transform_target(self, x, None, to_embedding=False)
        elif force_intrinsic_coordinates:
            assert (x.shape[1] == self.total_target_dim_intrinsic)
        assert (x.shape[1] == self.total_target_dim), (x.shape[1], self.total_target_dim)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                                  (amortization_parameters is not None and amortization_parameters.requires_grad))
        if needs_grad and amortization_parameters is None and not only_last and engine.supports_backward(self):
            # training path: fused forward AND backward layer kernels, torch autograd only for the parameter generator
            return engine.pdf_logpdf_trainable(self, x, conditional_input)
        amort = amortization_parameters.detach() if amortization_parameters is not None else None
        if (x.is_cuda and amort is None and not only_last and not engine.uses_custom_mlp(self)
                and type(conditional_input) != list):      # (CPU tensors: engine raises, there is no CPU fallback)
            # the whole-pdf entry as a torch.library op (ops.py): opaque to torch.compile, no graph break
            from . import ops
            log_pdf, log_pdf_base, base_pos = torch.ops.jammy_b200.pdf_logpdf(x, conditional_input, self._op_handle,
                                                                              int(self.chunk_rows or 0))
        else:
            log_pdf, log_pdf_base, base_pos = engine.pdf_logpdf(self, x, conditional_input, chunk_rows=self.chunk_rows,
                                                                amort=amort, only_last=only_last)
        if chart_log_det is not None:
            log_pdf = log_pdf + chart_log_det
        if needs_grad:
            anchor = amortization_parameters if (amortization_parameters is not None and
                                                 amortization_parameters.requires_grad) else self._anchor()
            log_pdf, log_pdf_base, base_pos = _NoBackward.apply(anchor, log_pdf, log_pdf_base, base_pos)
        return log_pdf, log_pdf_base, base_pos

    def _check_amortization_parameters(self, amortization_parameters, batch):
        """reference main/default.py:925-927, :1404-1409, :1591-1595"""
        if amortization_parameters is None:
            assert (self.amortize_everything == False), "a pdf built with amortize_everything needs amortization_parameters"
            return
        assert (self.amortize_everything), "amortization_parameters require a pdf built with amortize_everything=True"
        assert (amortization_parameters.dim() == 2 and
                amortization_parameters.shape[1] == self.total_number_amortizable_params), \
            (amortization_parameters.shape, self.total_number_amortizable_params)
        assert (amortization_parameters.shape[0] == batch), "batch size of amortization_parameters must agree with the batch size of the input"

    def log_mean_poisson(self, conditional_input=None, amortization_parameters=None):
        """log-lambda of the Poisson prediction, (B,1) (or the (1,1) parameter of an unconditional pdf).
        Reference main/default.py:832-877."""
        if self.log_normalization is None:
            raise Exception("This PDF does not predict the log-mean of a Poisson distriution. Initialize with 'predict_log_normalization'=True for this possibility.")
        if amortization_parameters is not None:
            raise Exception("Currently there is no support for the prediction of log-lambda and simultanesouly passing amortization_parameters .. there is some thinking involved in what to do in this situation so it is not supported at the moment.")
        if conditional_input is None:
            return self.log_normalization
        assert (self.join_poisson_and_pdf_description)
        mlp = self.mlp_predictors[0]
        needs_grad = torch.is_grad_enabled() and (conditional_input.requires_grad or
                                                  any(q.requires_grad for q in mlp.parameters()))
        if needs_grad:
            # the Poisson term of the loss back-propagates through log-lambda (reference main/default.py:873 returns it
            # with autograd history)
            if hasattr(mlp, "u_v_b_pars"):
                raise NotImplementedError("log_mean_poisson with gradients through an AmortizableMLP generator is not "
                                          "built; evaluate under torch.no_grad() or use an nn.Sequential generator")
            return engine.sequential_mlp_forward_trainable(mlp, conditional_input)[:, -1:]
        with torch.no_grad():
            if hasattr(mlp, "u_v_b_pars"):
                return mlp(conditional_input)[:, -1:]
            return engine.sequential_mlp_forward(mlp, conditional_input)[:, -1:]

    def all_layer_inverse(self, x, log_det, data_summary, amortization_parameters=None, force_embedding_coordinates=False,
                          force_intrinsic_coordinates=False, only_last=False):
        """target -> base through every sub-pdf (reference main/default.py:879-1057)."""
        self._check_amortization_parameters(amortization_parameters, x.shape[0])
        if force_embedding_coordinates and self._needs_transform():
            x, log_det = engine.pdf_transform_target(self, x, log_det, to_embedding=False)
        logp, logp_base, base = engine.pdf_logpdf(self, x, data_summary, chunk_rows=self.chunk_rows,
                                                  amort=amortization_parameters, only_last=only_last)
        return base, log_det + (logp - logp_base)

    # ------------------------------------------------------------------------------------------------------------------
    # sampling (reference main/default.py:1300-1371, :1533-1707)
    # ------------------------------------------------------------------------------------------------------------------
    def sample(self, conditional_input=None, samplesize=1, seed=None, allow_gradients=False,
               amortization_parameters=None, force_embedding_coordinates=False, force_intrinsic_coordinates=False,
               failsafe_crosscheck_tolerance=None, dtype=None, device=None, only_last=False):
        assert (self.use_as_passthrough_instead_of_pdf == False)
        if allow_gradients:
            # reference main/default.py:1342-1355: the same call without torch.no_grad()
            if (amortization_parameters is not None or only_last or failsafe_crosscheck_tolerance
                    or not engine.supports_sample_backward(self)):
                raise NotImplementedError("sample(allow_gradients=True): this pdf has no backward path (default \"g\" layers "
                                          "and the non-Euclidean layers f / v / r / o / m have one)")
            return self._obtain_sample(conditional_input=conditional_input, seed=seed, samplesize=samplesize,
                                       force_embedding_coordinates=force_embedding_coordinates,
                                       force_intrinsic_coordinates=force_intrinsic_coordinates, device=device, dtype=dtype,
                                       _trainable=True)
        with torch.no_grad():
            return self._obtain_sample(conditional_input=conditional_input, seed=seed, samplesize=samplesize,
                                       amortization_parameters=amortization_parameters,
                                       force_embedding_coordinates=force_embedding_coordinates,
                                       force_intrinsic_coordinates=force_intrinsic_coordinates,
                                       failsafe_crosscheck_tolerance=failsafe_crosscheck_tolerance, device=device,
                                       dtype=dtype, only_last=only_last)

    def _obtain_sample(self, conditional_input=None, predefined_target_input=None, samplesize=1, seed=None,
                       amortization_parameters=None, force_embedding_coordinates=False,
                       force_intrinsic_coordinates=False, failsafe_crosscheck_tolerance=None, dtype=None, device=None,
                       only_last=False, _trainable=False):
        used_sample_size = samplesize
        if self.amortize_everything:
            # reference main/default.py:1591-1606: batch, dtype and device come from the amortization parameters
            assert (amortization_parameters is not None)
            self._check_amortization_parameters(amortization_parameters, amortization_parameters.shape[0])
            amortization_parameters = amortization_parameters.detach()
            used_sample_size = amortization_parameters.shape[0]
            data_type, used_device = amortization_parameters.dtype, amortization_parameters.device
            if conditional_input is not None:
                c0 = conditional_input[0] if type(conditional_input) == list else conditional_input
                assert (c0.shape[0] == used_sample_size and c0.device == used_device)
                assert (c0.dtype == data_type), "Dtypes between conditional_input and amortization_paramters have to agree!"
        elif amortization_parameters is not None:
            self._check_amortization_parameters(amortization_parameters, 0)
        elif type(conditional_input) == list:
            assert (type(self.conditional_input_dim) == list and len(self.conditional_input_dim) == len(conditional_input))
            for ci_ind, ci in enumerate(conditional_input):
                assert (self.conditional_input_dim[ci_ind] == ci.shape[1]), "Inputs of conditional input vector do not match with pre-defined input_dims!"
            used_sample_size = conditional_input[0].shape[0]
            data_type, used_device = conditional_input[0].dtype, conditional_input[0].device
        elif conditional_input is not None:
            used_sample_size = conditional_input.shape[0]
            data_type, used_device = conditional_input.dtype, conditional_input.device
        else:
            data_type, used_device = self.obtain_current_dtype_n_device()
            if device is not None:
                used_device = torch.device(device)
            if dtype is not None:
                data_type = dtype
        assert ((data_type is not None) and (used_device is not None))
        std_normal_samples = 0.0
        if predefined_target_input is not None:
            z = predefined_target_input
            assert (used_device == z.device)
            if type(conditional_input) == list:
                assert (z.shape[0] == conditional_input[0].shape[0] and z.dtype == conditional_input[0].dtype)
            elif conditional_input is not None:
                assert (z.shape[0] == conditional_input.shape[0] and z.dtype == conditional_input.dtype)
        else:
            z = self._draw_base_normals(used_sample_size, seed, data_type, used_device)
            std_normal_samples = z
        if _trainable and torch.is_grad_enabled():
            x, log_pdf, log_gauss = engine.pdf_sample_trainable(self, z, conditional_input)
        elif (z.is_cuda and amortization_parameters is None and not only_last and not engine.uses_custom_mlp(self)
                and type(conditional_input) != list):
            from . import ops       # the whole-pdf entry as a torch.library op (ops.py)
            x, log_pdf, log_gauss = torch.ops.jammy_b200.pdf_sample(z, conditional_input, self._op_handle,
                                                                    int(self.chunk_rows or 0))
        else:
            x, log_pdf, log_gauss = engine.pdf_sample(self, z, conditional_input, chunk_rows=self.chunk_rows,
                                                      amort=amortization_parameters, only_last=only_last)
        if force_embedding_coordinates and self._needs_transform():
            # reference main/default.py:1522-1524: default -> embedding coordinates, log p = log N(z) - (logdet + chart)
            x, chart_log_det = engine.pdf_transform_target(self, x, None, to_embedding=True)
            log_pdf = log_pdf - chart_log_det
        if failsafe_crosscheck_tolerance:
            assert (predefined_target_input is None), "Failsafe does not work with predefined input!"
            if amortization_parameters is not None or only_last:
                raise NotImplementedError("failsafe_crosscheck_tolerance together with amortization_parameters / only_last")
            x, std_normal_samples, log_pdf, log_gauss = self._recheck_sampling(
                x, std_normal_samples, log_pdf, log_gauss, failsafe_crosscheck_tolerance, conditional_input,
                force_embedding_coordinates, force_intrinsic_coordinates, data_type, used_device)
        return x, std_normal_samples, log_pdf, log_gauss

    def _draw_base_normals(self, n, seed, data_type, device):
        """[n, total_base_dim] standard normals according to `rng_mode`."""
        if self.rng_mode == "numpy":
            # reference: host numpy RNG then H2D copy (main/default.py:1661-1668)
            if seed is not None:
                numpy.random.seed(seed)
            std_normal = numpy.random.normal(size=(n, self.total_base_dim))
            return torch.from_numpy(std_normal).type(data_type).to(device)
        elif self.rng_mode == "philox":
            # counter-based device generator: row i = f(seed, first_row + i); `first_row` lets the ranks of a sharded
            # job draw slices of one global stream (jammy_flows_b200.sharding.shard_base_normals)
            used_seed = seed if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
            return engine.normal_rows(n, self.total_base_dim, used_seed, first_row=self.rng_first_row,
                                      dtype=data_type, device=device)
        else:
            gen = None
            if seed is not None:
                gen = torch.Generator(device=device)
                gen.manual_seed(seed)
            return torch.randn(n, self.total_base_dim, dtype=data_type, device=device, generator=gen)

    def _recheck_sampling(self, x, z, log_pdf, log_gauss, tol, conditional_input, force_emb, force_intr, dtype, device):
        """Round-trip check of a sample and re-draw of the rows that fail it (reference extra_functions.py:413-533,
        `recheck_sampling`, used for the iterative sphere flows): x -> base' -> x' must reproduce base, x and log p
        within `tol`; deviating rows are sampled again (recursively, with the same check)."""
        with torch.no_grad():
            lp_new, _, base_new = self.forward(x, conditional_input=conditional_input,
                                               force_embedding_coordinates=force_emb, force_intrinsic_coordinates=force_intr)
            x_new = self._obtain_sample(conditional_input=conditional_input, predefined_target_input=base_new,
                                        force_embedding_coordinates=force_emb, force_intrinsic_coordinates=force_intr,
                                        dtype=dtype, device=device)[0]
            dev = (torch.abs(lp_new - log_pdf) > tol) | (torch.abs(x_new - x) > tol).any(dim=1) \
                | (torch.abs(base_new - z) > tol).any(dim=1)
            n_dev = int(dev.sum())
            if n_dev == 0:
                return x, z, log_pdf, log_gauss
            new_c = conditional_input[dev] if conditional_input is not None else None
            rx, rz, rlp, rlg = self._obtain_sample(conditional_input=new_c, samplesize=n_dev,
                                                   force_embedding_coordinates=force_emb,
                                                   force_intrinsic_coordinates=force_intr,
                                                   failsafe_crosscheck_tolerance=tol, dtype=dtype, device=device)
            x, z, log_pdf, log_gauss = x.clone(), z.clone(), log_pdf.clone(), log_gauss.clone()
            x[dev], z[dev], log_pdf[dev], log_gauss[dev] = rx, rz, rlp, rlg
        return x, z, log_pdf, log_gauss

    def all_layer_forward(self, x, log_det, data_summary, amortization_parameters=None, force_embedding_coordinates=False,
                          force_intrinsic_coordinates=False, only_last=False):
        """base -> target through every sub-pdf (reference main/default.py:1373-1531)."""
        self._check_amortization_parameters(amortization_parameters, x.shape[0])
        xs, logp, logp_base = engine.pdf_sample(self, x, data_summary, chunk_rows=self.chunk_rows,
                                                amort=amortization_parameters, only_last=only_last)
        log_det = log_det + (logp_base - logp)
        if force_embedding_coordinates and self._needs_transform():
            xs, log_det = engine.pdf_transform_target(self, xs, log_det, to_embedding=True)
        return xs, log_det

    # ------------------------------------------------------------------------------------------------------------------
    # entropy, total path only (reference main/default.py:2263-2369)
    # ------------------------------------------------------------------------------------------------------------------
    def entropy(self, sub_manifolds=[-1], conditional_input=None, force_embedding_coordinates=True,
                force_intrinsic_coordinates=False, samplesize=100, failsafe_crosscheck_tolerance=None, dtype=None,
                device=None, _base_samples=None):
        """Monte-Carlo entropies (reference main/default.py:2263-2454): "total" = -mean log p over `samplesize` samples
        per conditional row; sub-manifold k: -mean_i log( mean_j p_k(x_k^i | x_<k^j) ), an S x S cross-evaluation of
        sub-pdf k (`engine.subpdf_logpdf`) reduced on the device (`jf_row_logmeanexp`)."""
        for subdim in sub_manifolds:
            if subdim != -1:
                assert (subdim >= 0 and subdim < len(self.layer_list))
        if force_embedding_coordinates == False:
            print("#### CAUTION: Calculating entropy without forcing embedding coordinates. This might lead to undesired "
                  "and wrong entropies when using manifold PDFs!#############")
        data_type, used_device = self.obtain_current_dtype_n_device()
        if device is not None:
            used_device = torch.device(device)
        if dtype is not None:
            data_type = dtype
        S = samplesize
        if type(conditional_input) == list:
            assert (type(self.conditional_input_dim) == list and len(self.conditional_input_dim) == len(conditional_input))
            data_type, used_device = conditional_input[0].dtype, conditional_input[0].device
            cond = [ci.repeat_interleave(S, dim=0) for ci in conditional_input]
            batch = conditional_input[0].shape[0]
        elif conditional_input is not None:
            assert (self.conditional_input_dim is not None)
            data_type, used_device = conditional_input.dtype, conditional_input.device
            cond = conditional_input.repeat_interleave(S, dim=0)
            batch = conditional_input.shape[0]
        else:
            assert (self.conditional_input_dim is None), "We require conditional input, since this is a conditional PDF."
            cond = None
            batch = 1
        n = S * batch
        use_emb = bool(force_embedding_coordinates) and self._needs_transform()
        out = dict()
        with torch.no_grad():
            if _base_samples is not None:      # test hook, like `predefined_target_input` of _obtain_sample
                z = _base_samples
                assert z.shape == (n, self.total_base_dim)
            else:
                z = torch.randn(n, self.total_base_dim, dtype=data_type, device=used_device)   # reference :2914 (device RNG)
            x, log_pdf, log_gauss = engine.pdf_sample(self, z, cond, chunk_rows=self.chunk_rows)
            if failsafe_crosscheck_tolerance:
                # reference main/default.py:2957-2976: the entropy samples go through the same round-trip check
                x, z, log_pdf, log_gauss = self._recheck_sampling(x, z, log_pdf, log_gauss, failsafe_crosscheck_tolerance,
                                                                  cond, False, False, data_type, used_device)
            emb = x
            if self._needs_transform():
                emb, chart_log_det = engine.pdf_transform_target(self, x, None, to_embedding=True)
                if use_emb:
                    log_pdf = log_pdf - chart_log_det
            for sub_mf in sub_manifolds:
                if sub_mf == -1:
                    out["total"] = -log_pdf.reshape(-1, S).mean(dim=1)
                    continue
                e0, e1 = self.target_dim_indices_embedded[sub_mf]
                t0, t1 = self.target_dim_indices[sub_mf]
                x_k = emb[:, e0:e1] if use_emb else x[:, t0:t1]
                cond_k = cond[sub_mf] if type(cond) == list else cond
                if sub_mf == 0:
                    lp = engine.subpdf_logpdf(self, 0, x_k, [cond_k] if cond_k is not None else [], use_emb)
                    out[0] = -lp.reshape(-1, S).mean(dim=1)
                    continue
                prev = emb[:, :e0]
                d_k = x_k.shape[1]
                per_b = max(1, (1 << 19) // (S * S))          # conditional rows per launch: <= 2^19 cross rows
                vals = []
                for b0 in range(0, batch, per_b):
                    b1 = min(batch, b0 + per_b)
                    nb = b1 - b0
                    rows = slice(b0 * S, b1 * S)
                    # row (b, i, j): earlier targets and conditional input of sample j, sub-pdf k coordinates of sample i
                    first = prev[rows].reshape(nb, S, e0).repeat(1, S, 1).reshape(-1, e0)
                    final = x_k[rows].reshape(nb, S, d_k).repeat_interleave(S, dim=1).reshape(-1, d_k)
                    segs = [first]
                    if cond_k is not None:
                        segs = [cond_k[rows].repeat_interleave(S, dim=0), first]
                    lp = engine.subpdf_logpdf(self, sub_mf, final, segs, use_emb)
                    vals.append(engine.row_logmeanexp(lp.reshape(-1, S)).reshape(nb, S).mean(dim=1))
                out[sub_mf] = -torch.cat(vals)
        return out

    # ------------------------------------------------------------------------------------------------------------------
    # coordinate transforms (reference main/default.py:1737-1813): identity for Euclidean-only pdfs
    # ------------------------------------------------------------------------------------------------------------------
    def transform_target_space(self, target, log_det=0, transform_from="default", transform_to="embedding"):
        """"default" coordinates are the intrinsic ones here (always_parametrize_in_embedding_space is not supported),
        so only intrinsic <-> embedding moves anything; Euclidean and interval sub-pdfs are identities."""
        for name in (transform_from, transform_to):
            if name not in ("default", "intrinsic", "embedding"):
                raise Exception("Unknown transformation space! .. ", name, "Allowed: default/intrinsic/embedding")
        squeeze = target.dim() == 1
        new_target = target.unsqueeze(0) if squeeze else target
        src_emb, dst_emb = transform_from == "embedding", transform_to == "embedding"
        assert (new_target.shape[1] == (self.total_target_dim_embedded if src_emb else self.total_target_dim_intrinsic))
        if src_emb != dst_emb and self._needs_transform():
            new_target, log_det = engine.pdf_transform_target(self, new_target, log_det, to_embedding=dst_emb)
        return (new_target.squeeze(0) if squeeze else new_target), log_det

    # ------------------------------------------------------------------------------------------------------------------
    # plain-dict program for the test oracle (no CUDA involved)
    # ------------------------------------------------------------------------------------------------------------------
    def export_program(self, dtype="float64"):
        subs = []
        for k, layers in enumerate(self.layer_list):
            ranges, off = [], 0
            for l in layers:
                ranges.append((off, off + l.total_param_num))
                off += l.total_param_num
            mlp = self.mlp_predictors[k]
            mlp_spec = None
            names = []
            if mlp is not None and hasattr(mlp, "u_v_b_pars"):
                mlp_spec = mlp.structure()
            elif mlp is not None:
                mlp_spec = dict(linear_indices=[i for i, m in enumerate(mlp) if isinstance(m, nn.Linear)])
            else:
                for li, l in enumerate(layers):
                    names += ["layer_list.%d.%d.%s" % (k, li, n) for n in l.permanent_param_names()]
            subs.append(dict(manifold=self.pdf_defs_list[k][0], layers=[l.descriptor() for l in layers],
                             layer_param_ranges=ranges, mlp=mlp_spec, permanent_param_names=names,
                             target_cols=self.target_dim_indices[k], base_cols=self.base_dim_indices[k]))
        return dict(dtype=dtype, subpdfs=subs, conditional_input_dim=self.conditional_input_dim)
