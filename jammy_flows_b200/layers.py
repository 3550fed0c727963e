"""Host-side flow-layer objects: parameter holders + static descriptors for the fused sm_100a kernels.

They mirror the reference's layer plugin API (`jammy_flows/layers/layer_base.py:4-100`): `total_param_num`,
`get_total_param_num`, `get_desired_init_parameters`, `init_params`, `flow_mapping`, `inv_flow_mapping`,
`_embedding_conditional_return(_num)`, `get_layer_{intrinsic_target,embedded_target,base}_dimension`, and keep the
reference's parameter names/shapes (the de-facto checkpoint contract) so a reference
`state_dict` loads unchanged.  No layer math lives here: `flow_mapping` / `inv_flow_mapping` run the CUDA kernels
through the C-ABI (a one-layer flow program); there is no CPU or eager-torch fallback.
"""
import math

import numpy
import torch
from torch import nn

# inverse-CDF stage codes shared with include/jammy_b200.h (JF_INV_*)
INV_TYPES = {"isigmoid": 0, "inormal_partly_precise": 1, "inormal_full_pade": 2, "inormal_partly_crude": 3}


class layer_base(nn.Module):
    """Reference: layers/layer_base.py:4-100."""

    code = "?"

    def __init__(self, dimension=1, always_parametrize_in_embedding_space=0):
        super().__init__()
        self.total_param_num = 0
        self.dimension = dimension
        self.always_parametrize_in_embedding_space = always_parametrize_in_embedding_space

    def get_total_param_num(self):
        return self.total_param_num

    def get_desired_init_parameters(self):
        return torch.randn(self.total_param_num)

    def get_layer_embedded_target_dimension(self):
        return self._embedding_conditional_return_num()

    def get_layer_intrinsic_target_dimension(self):
        return self.dimension

    def get_layer_base_dimension(self):
        return self._get_layer_base_dimension()

    # ---- things every concrete layer provides ---------------------------------------------------------------------
    def permanent_param_names(self):
        """Names (relative to this module) of the permanent tensors in `extra_inputs` order."""
        raise NotImplementedError

    def descriptor(self):
        """Plain dict consumed by program.py (-> C struct) and by the oracle."""
        raise NotImplementedError

    def packed_permanent_params(self):
        """Flat [P] vector of the permanent parameters in the reference's `extra_inputs` order."""
        names = self.permanent_param_names()
        if len(names) == 0:
            return None
        return torch.cat([getattr(self, n).reshape(-1) for n in names])

    # ---- plugin API (one-layer flow program through the C-ABI) ------------------------------------------------------
    def _run_single(self, direction, inputs, extra_inputs, **kw):
        from . import engine
        x, log_det = inputs
        return engine.run_single_layer(self, direction, x, log_det, extra_inputs, **kw)

    def flow_mapping(self, inputs, extra_inputs=None, **kw):
        """base -> target (sampling direction).  Reference: layers/layer_base.py:58-63."""
        return self._run_single("sample", inputs, extra_inputs, **kw)

    def inv_flow_mapping(self, inputs, extra_inputs=None, **kw):
        """target -> base (log_pdf direction).  Reference: layers/layer_base.py:65-70."""
        return self._run_single("logpdf", inputs, extra_inputs, **kw)


# =====================================================================================================================
# Euclidean: Gaussianization flow "g"
# =====================================================================================================================
class gf_block(layer_base):
    """Gaussianization-flow block, symbol "g".

    Reference: layers/euclidean/gaussianization_flow.py:50-386 (constructor / parameter layout),
    layers/euclidean/euclidean_base.py:9-31 (offset handling).  Parameter slice order inside `extra_inputs`:
    [offset d (last layer only)] [vs: iter*d] [means K*d] [log_widths K*d] [log_weights K*d].
    """

    code = "g"
    manifold = "e"

    def __init__(self, dimension, nonlinear_stretch_type="classic", num_kde=5, num_householder_iter=-1,
                 use_permanent_parameters=False, fit_normalization=0, inverse_function_type="inormal_partly_precise",
                 model_offset=0, softplus_for_width=0, width_smooth_saturation=1, lower_bound_for_widths=0.01,
                 upper_bound_for_widths=100, lower_bound_for_norms=1, upper_bound_for_norms=10, center_mean=0,
                 clamp_widths=0, regulate_normalization=0, add_skewness=0, rotation_mode="householder"):
        super().__init__(dimension=dimension)
        unsupported = []
        if nonlinear_stretch_type != "classic":
            unsupported.append("nonlinear_stretch_type=%s" % nonlinear_stretch_type)
        if rotation_mode != "householder":
            unsupported.append("rotation_mode=%s" % rotation_mode)
        if add_skewness:
            unsupported.append("add_skewness=1")
        if center_mean:
            unsupported.append("center_mean=1")
        if softplus_for_width or clamp_widths or not width_smooth_saturation:
            unsupported.append("non-default width regulator")
        if upper_bound_for_widths <= 0 or upper_bound_for_norms <= 0:
            unsupported.append("unbounded widths/norms")
        if len(unsupported) > 0:
            raise NotImplementedError("'g' layer options without an sm_100a kernel yet (SURVEY.md section 8f rank 1): "
                                      + ", ".join(unsupported) + " -- there is no CPU fallback")
        assert inverse_function_type in INV_TYPES
        assert lower_bound_for_widths > 0.0

        self.use_permanent_parameters = use_permanent_parameters
        self.model_offset = model_offset
        self.inverse_function_type = inverse_function_type
        self.num_kde = num_kde
        self.fit_normalization = fit_normalization
        self.regulate_normalization = regulate_normalization
        self.width_min = lower_bound_for_widths
        self.width_max = upper_bound_for_widths
        self.lower_bound_for_norms = lower_bound_for_norms
        self.upper_bound_for_norms = upper_bound_for_norms
        self.rotation_mode = rotation_mode
        self.householder_iter = dimension if num_householder_iter == -1 else num_householder_iter
        self.use_householder = self.householder_iter > 0
        self.num_householder_params = self.householder_iter * dimension if self.use_householder else 0
        self.num_params_datapoints = num_kde * dimension
        # initialisation from the Gaussianization-flow paper (reference gaussianization_flow.py:233-234)
        self.init_log_width = numpy.log((4. * numpy.sqrt(math.pi) / ((math.pi ** 4) * num_kde)) ** 0.2)

        # RNG call order below mirrors the reference constructor so that equal seeds give equal parameters
        self.offsets = None
        if self.model_offset:
            self.offsets = torch.zeros(dimension).type(torch.double).unsqueeze(0)
            if use_permanent_parameters:
                self.offsets = nn.Parameter(torch.randn(dimension).type(torch.double).unsqueeze(0))
            self.total_param_num += dimension
        if self.use_householder and use_permanent_parameters:
            self.vs = nn.Parameter(torch.randn(self.householder_iter, dimension).unsqueeze(0))
        self.total_param_num += self.num_householder_params
        if use_permanent_parameters:
            self.kde_means = nn.Parameter(torch.randn(num_kde, dimension).unsqueeze(0))
        self.total_param_num += self.num_params_datapoints
        if use_permanent_parameters:
            self.kde_log_widths = nn.Parameter(torch.ones(num_kde, dimension).unsqueeze(0) * self.init_log_width)
        self.total_param_num += self.num_params_datapoints
        if fit_normalization:
            if use_permanent_parameters:
                self.kde_log_weights = nn.Parameter(torch.randn(num_kde, dimension).unsqueeze(0))
            self.total_param_num += self.num_params_datapoints

    # ---- reference euclidean_base.py:77-104, gaussianization_flow.py:1116-1210 -----------------------------------
    def get_desired_init_parameters(self):
        par_list = []
        if self.model_offset:
            par_list.append(torch.ones(self.dimension) * 0.001)
        if self.num_householder_params > 0:
            par_list.append(torch.randn(self.householder_iter * self.dimension))
        par_list.append(torch.randn(self.num_params_datapoints))
        par_list.append(torch.ones(self.num_params_datapoints) * self.init_log_width)
        if self.fit_normalization:
            par_list.append(torch.ones(self.num_params_datapoints))
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num), (len(params), self.total_param_num)
        assert (self.use_permanent_parameters == 1)
        c = 0
        d, k = self.dimension, self.num_kde
        if self.model_offset:
            self.offsets.data = params[:d]
            c = d
        if self.use_householder:
            self.vs.data = torch.reshape(params[c:c + self.num_householder_params], [1, self.householder_iter, d])
            c += self.num_householder_params
        self.kde_means.data = torch.reshape(params[c:c + k * d], [1, k, d])
        c += k * d
        self.kde_log_widths.data = torch.reshape(params[c:c + k * d], [1, k, d])
        c += k * d
        if self.fit_normalization:
            self.kde_log_weights.data = torch.reshape(params[c:c + k * d], [1, k, d])

    def permanent_param_names(self):
        names = []
        if self.model_offset:
            names.append("offsets")
        if self.use_householder:
            names.append("vs")
        names += ["kde_means", "kde_log_widths"]
        if self.fit_normalization:
            names.append("kde_log_weights")
        return names

    def descriptor(self):
        return dict(code="g", dim=self.dimension, num_kde=self.num_kde, hh_iter=self.householder_iter,
                    inverse_function_type=self.inverse_function_type, inv_type=INV_TYPES[self.inverse_function_type],
                    fit_normalization=int(self.fit_normalization),
                    regulate_normalization=int(self.regulate_normalization), model_offset=int(self.model_offset),
                    w_min=float(self.width_min), w_max=float(self.width_max),
                    n_min=float(self.lower_bound_for_norms), n_max=float(self.upper_bound_for_norms),
                    n_params=self.total_param_num)

    def _embedding_conditional_return(self, x):
        return x

    def _embedding_conditional_return_num(self):
        return self.dimension

    def _get_layer_base_dimension(self):
        return self.dimension

    def transform_target_space(self, x, log_det=0.0, transform_from="default", transform_to="embedding"):
        return x, log_det


# =====================================================================================================================
# S2: Fisher-von-Mises layer "f" (reference defaults: Householder rotation + vMF z-scaling)
# =====================================================================================================================
class fisher_von_mises_2d(layer_base):
    """Symbol "f" (and the "n" alias).

    Reference: layers/spheres/fvm_2d.py:30-265 (constructor), layers/spheres/sphere_base.py:42-110 (rotation
    parameters, which come FIRST in the layer's slice: sphere_base.py:630/673).  Parameter slice:
    [householder iter*3] [log kappa 1] [vertical spline params] [circular spline params].
    """

    code = "f"
    manifold = "s"

    def __init__(self, dimension, euclidean_to_sphere_as_first=False, use_permanent_parameters=False,
                 add_vertical_rq_spline_flow=0, add_circular_rq_spline_flow=0, vertical_flow_defs="r",
                 circular_flow_defs="o", add_correlated_rq_spline_flow=0, correlated_max_rank=3, inverse_z_scaling=1,
                 spline_num_basis_functions=5, boundary_cos_theta_identity_region=0.0, vertical_smooth=0,
                 vertical_restrict_max_min_width_height_ratio=-1.0, vertical_fix_boundary_derivative=1,
                 vertical_fix_first_width_n_height_to_zero=0, vertical_also_fix_second_width_to_zero=0,
                 vertical_independent_width_height_parametrization=0, circular_add_rotation=1, min_kappa=1e-10,
                 kappa_prediction="direct_log_real_bounded", add_extra_rotation_inbetween=0, kappa_clamping=0,
                 add_rotation=1, rotation_mode="householder", num_householder_iter=-1):
        super().__init__(dimension=dimension)
        if dimension != 2:
            raise Exception("2-D Flow")
        unsupported = []
        if add_vertical_rq_spline_flow or add_circular_rq_spline_flow or add_correlated_rq_spline_flow:
            unsupported.append("vertical/circular/correlated rq-spline sub-flows")
        if kappa_prediction != "direct_log_real_bounded" or kappa_clamping:
            unsupported.append("kappa_prediction=%s kappa_clamping=%d" % (kappa_prediction, kappa_clamping))
        if add_rotation and rotation_mode != "householder":
            unsupported.append("rotation_mode=%s" % rotation_mode)
        if add_extra_rotation_inbetween:
            unsupported.append("add_extra_rotation_inbetween=1")
        if boundary_cos_theta_identity_region != 0.0:
            unsupported.append("boundary_cos_theta_identity_region")
        if len(unsupported) > 0:
            raise NotImplementedError("'f' layer options without an sm_100a kernel yet (SURVEY.md section 8a row a13): "
                                      + ", ".join(unsupported) + " -- there is no CPU fallback")
        self.euclidean_to_sphere_as_first = euclidean_to_sphere_as_first
        self.use_permanent_parameters = use_permanent_parameters
        self.add_rotation = add_rotation
        self.rotation_mode = rotation_mode
        self.z_scaling_factor = -1.0 if inverse_z_scaling else 1.0
        self.min_kappa = min_kappa
        self.num_householder_params = 0
        self.num_householder_iter = 0
        if add_rotation:
            self.num_householder_iter = dimension + 1 if num_householder_iter == -1 else num_householder_iter
            self.num_householder_params = self.num_householder_iter * (dimension + 1)
        # RNG order as in the reference: sphere_base (householder) first, then kappa
        if use_permanent_parameters and self.num_householder_params > 0:
            self.householder_params = nn.Parameter(torch.randn((1, self.num_householder_params)))
        self.total_param_num += self.num_householder_params
        if use_permanent_parameters:
            self.loglike_kappa = nn.Parameter(torch.randn(1).unsqueeze(0))
        self.total_param_num += 1

    # reference sphere_base.py:712-730 + fvm_2d.py:747-773
    def get_desired_init_parameters(self):
        par_list = []
        if self.num_householder_params > 0:
            par_list.append(torch.randn((self.num_householder_params)))
        par_list.append(torch.randn((1)) - 3.0)
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        n = self.num_householder_params
        if self.add_rotation:
            self.householder_params.data = params[:n].reshape(1, n)
        self.loglike_kappa.data = params[n:n + 1].reshape(1, 1)

    def permanent_param_names(self):
        return (["householder_params"] if self.num_householder_params > 0 else []) + ["loglike_kappa"]

    def descriptor(self):
        return dict(code="f", dim=2, add_rotation=int(self.add_rotation), hh_iter=self.num_householder_iter,
                    z_sign=float(self.z_scaling_factor), min_kappa=float(self.min_kappa),
                    first=int(self.euclidean_to_sphere_as_first), n_params=self.total_param_num)

    def _embedding_conditional_return(self, x):
        from . import engine
        if x.shape[1] == self.dimension:
            return engine.s2_embedding(x)
        return x

    def _embedding_conditional_return_num(self):
        return self.dimension + 1

    def _get_layer_base_dimension(self):
        if self.always_parametrize_in_embedding_space and not self.euclidean_to_sphere_as_first:
            return self.dimension + 1
        return self.dimension
