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
ROT_MODES = {"householder": 0, "none": 1, "angles": 2, "cayley": 3, "triangular_combination": 4}
WIDTH_MODES = {"smooth": 0, "exp": 1, "softplus": 2}


class gf_block(layer_base):
    """Gaussianization-flow block, symbol "g".

    Reference: layers/euclidean/gaussianization_flow.py:50-386 (constructor / parameter layout),
    layers/euclidean/euclidean_base.py:9-31 (offset handling).  Parameter slice order inside `extra_inputs`:
      [offset d (last layer only)] [rotation] then
      classic:     [means (K-center_mean)*d] [log_widths K*d] [log_weights K*d (fit_normalization)] [log_skew K*d (add_skewness)]
      rq_splines:  [log_widths d*K] [log_heights d*K] [log_derivatives d*(K+1)] [boundary_points d*4]
    rotation: householder iter*d | angles d(d-1)/2 | cayley 1 (d=2) | triangular_combination d(d-1)+d-1 | none 0.
    """

    code = "g"
    manifold = "e"

    def __init__(self, dimension, nonlinear_stretch_type="classic", num_kde=5, num_householder_iter=-1,
                 use_permanent_parameters=False, fit_normalization=0, inverse_function_type="inormal_partly_precise",
                 model_offset=0, softplus_for_width=0, width_smooth_saturation=1, lower_bound_for_widths=0.01,
                 upper_bound_for_widths=100, lower_bound_for_norms=1, upper_bound_for_norms=10, center_mean=0,
                 clamp_widths=0, regulate_normalization=0, add_skewness=0, rotation_mode="householder"):
        super().__init__(dimension=dimension)
        assert inverse_function_type in INV_TYPES
        assert lower_bound_for_widths > 0.0
        if nonlinear_stretch_type not in ("classic", "rq_splines"):
            raise Exception("Unknown non linear stretch type: %s" % nonlinear_stretch_type)
        if rotation_mode not in ROT_MODES:
            # the reference silently applies no rotation for an unknown mode; only "none" is documented for that
            raise Exception("Unknown rotation_mode: %s" % rotation_mode)
        d = dimension
        self.use_permanent_parameters = use_permanent_parameters
        self.model_offset = model_offset
        self.nonlinear_stretch_type = nonlinear_stretch_type
        self.inverse_function_type = inverse_function_type
        self.num_kde = num_kde
        self.fit_normalization = fit_normalization
        self.regulate_normalization = regulate_normalization
        self.add_skewness = add_skewness
        self.center_mean = int(center_mean)
        self.width_min = lower_bound_for_widths
        self.width_max = upper_bound_for_widths if upper_bound_for_widths > 0 else None
        self.width_smooth_saturation = width_smooth_saturation
        if self.width_smooth_saturation:
            assert (self.width_max is not None), "We require a maximum saturation level for smooth saturation!"
        self.softplus_for_width = softplus_for_width
        self.clamp_widths = clamp_widths
        self.lower_bound_for_norms = lower_bound_for_norms
        self.upper_bound_for_norms = upper_bound_for_norms
        self.rotation_mode = rotation_mode
        if fit_normalization and regulate_normalization and nonlinear_stretch_type == "classic":
            assert upper_bound_for_norms > 0, "regulate_normalization needs a positive upper_bound_for_norms"
        # width regulator variant and the clamp applied to the raw value (reference gaussianization_flow.py:264-317)
        classic = nonlinear_stretch_type == "classic"
        self.width_mode = "softplus" if softplus_for_width else ("smooth" if width_smooth_saturation else "exp")
        self.width_clamp = None
        if classic and clamp_widths:
            lo = float(numpy.log(0.01 * self.width_min))
            if self.width_mode == "smooth":
                hi = float(numpy.log(self.width_max) * 3.0)
            else:
                hi = float(numpy.log(self.width_max)) if self.width_max is not None else float("inf")
            self.width_clamp = (lo, hi)
        # initialisation from the Gaussianization-flow paper (reference gaussianization_flow.py:233-234)
        self.init_log_width = numpy.log((4. * numpy.sqrt(math.pi) / ((math.pi ** 4) * num_kde)) ** 0.2)
        self.num_params_datapoints = num_kde * d

        # RNG call order below mirrors the reference constructor so that equal seeds give equal parameters
        self.offsets = None
        if self.model_offset:
            self.offsets = torch.zeros(d).type(torch.double).unsqueeze(0)
            if use_permanent_parameters:
                self.offsets = nn.Parameter(torch.randn(d).type(torch.double).unsqueeze(0))
            self.total_param_num += d
        # ---- rotation (reference :141-205) ----
        self.householder_iter = 0
        self.use_householder = False
        self.num_householder_params = 0
        self.num_rotation_params = 0
        if rotation_mode == "triangular_combination":
            self.num_triangle_params = int(d - 1 + d * (d - 1))
            self.num_rotation_params = self.num_triangle_params
            if use_permanent_parameters and d > 1:
                self.triangle_trafo_pars = nn.Parameter(torch.randn(self.num_triangle_params).unsqueeze(0))
        elif rotation_mode == "householder":
            self.householder_iter = d if num_householder_iter == -1 else num_householder_iter
            self.use_householder = self.householder_iter > 0
            if self.use_householder:
                if use_permanent_parameters:
                    self.vs = nn.Parameter(torch.randn(self.householder_iter, d).unsqueeze(0))
                self.num_householder_params = self.householder_iter * d
            self.num_rotation_params = self.num_householder_params
        elif rotation_mode == "angles":
            self.num_angle_pars = 0
            if d > 1:
                self.num_angle_pars = int(d * (d - 1) / 2)
                if use_permanent_parameters:
                    self.angle_pars = nn.Parameter(torch.randn((1, self.num_angle_pars)))
            self.num_rotation_params = self.num_angle_pars
        elif rotation_mode == "cayley":
            self.num_cayley_pars = 0
            if d > 1:
                self.num_cayley_pars = 1
                assert (d == 2), "Cayley requires 2 dims at the moment"
                if use_permanent_parameters:
                    # the reference constructor creates `cayley_pars` and then fails in init_params
                    # (gaussianization_flow.py:1193 indexes a 1-d tensor with two indices): cayley only works amortised
                    raise IndexError("rotation_mode='cayley' with permanent parameters fails in the reference's "
                                     "init_params (gaussianization_flow.py:1193); use it in a conditional sub-pdf")
            self.num_rotation_params = self.num_cayley_pars
        self.total_param_num += self.num_rotation_params
        # ---- non-linear stretch (reference :218-386) ----
        if classic:
            self.total_param_num_means = (num_kde - self.center_mean) * d
            if use_permanent_parameters:
                self.kde_means = nn.Parameter(torch.randn(num_kde - self.center_mean, d).unsqueeze(0))
            self.total_param_num += self.total_param_num_means
            if use_permanent_parameters:
                self.kde_log_widths = nn.Parameter(torch.ones(num_kde, d).unsqueeze(0) * self.init_log_width)
            self.total_param_num += self.num_params_datapoints
            if fit_normalization:
                if use_permanent_parameters:
                    self.kde_log_weights = nn.Parameter(torch.randn(num_kde, d).unsqueeze(0))
                self.total_param_num += self.num_params_datapoints
            if add_skewness:
                if use_permanent_parameters:
                    self.kde_log_skew_exponents = nn.Parameter(torch.randn(num_kde, d).unsqueeze(0))
                self.total_param_num += self.num_params_datapoints
        else:
            if use_permanent_parameters:
                self.log_widths = nn.Parameter(torch.randn(d, num_kde).unsqueeze(0))
                self.log_heights = nn.Parameter(torch.randn(d, num_kde).unsqueeze(0))
                self.log_derivatives = nn.Parameter(torch.randn(d, num_kde + 1).unsqueeze(0))
                self.boundary_points = nn.Parameter(torch.randn(d, 4).unsqueeze(0))
            self.total_param_num += (num_kde * d) * 2 + (num_kde + 1) * d + 4 * d
        if num_kde > 32:
            raise NotImplementedError("'g' with num_kde > 32 (JF_MAX_KDE)")

    @property
    def is_default_kernel_config(self):
        """True when the layer runs on the specialised (fast) chain kernel; the other options use the general one."""
        return (self.nonlinear_stretch_type == "classic" and self.rotation_mode == "householder"
                and not self.add_skewness and not self.center_mean and self.width_mode == "smooth"
                and self.width_clamp is None)

    # ---- reference euclidean_base.py:77-104, gaussianization_flow.py:1116-1210 -----------------------------------
    def get_desired_init_parameters(self):
        par_list = []
        d = self.dimension
        if self.model_offset:
            par_list.append(torch.ones(d) * 0.001)
        if self.rotation_mode == "householder":
            if self.num_householder_params > 0:
                par_list.append(torch.randn(self.householder_iter * d))
        elif self.rotation_mode != "none":
            par_list.append(torch.zeros(self.num_rotation_params))
        if self.nonlinear_stretch_type == "classic":
            par_list.append(torch.randn(self.total_param_num_means))
            par_list.append(torch.ones(self.num_params_datapoints) * self.init_log_width)
            if self.fit_normalization:
                par_list.append(torch.ones(self.num_params_datapoints))
            if self.add_skewness:
                par_list.append(torch.zeros(self.num_params_datapoints))
        else:
            par_list.append(torch.ones(self.num_kde * d))
            par_list.append(torch.ones(self.num_kde * d))
            par_list.append(torch.ones((self.num_kde + 1) * d) * 0.54135)   # softplus^-1(1)
            par_list.append(torch.Tensor(d * [-1.0, 1.0, -1.0, 1.0]))
        return torch.cat(par_list)

    def _rotation_param_name(self):
        if self.num_rotation_params == 0:
            return None
        return {"householder": "vs", "angles": "angle_pars", "cayley": "cayley_pars",
                "triangular_combination": "triangle_trafo_pars"}[self.rotation_mode]

    def init_params(self, params):
        assert (len(params) == self.total_param_num), (len(params), self.total_param_num)
        assert (self.use_permanent_parameters == 1)
        c = 0
        d, k = self.dimension, self.num_kde
        if self.model_offset:
            self.offsets.data = params[:d]
            c = d
        if self.rotation_mode == "householder":
            if self.use_householder:
                self.vs.data = torch.reshape(params[c:c + self.num_householder_params], [1, self.householder_iter, d])
                c += self.num_householder_params
        elif self.num_rotation_params > 0:
            getattr(self, self._rotation_param_name()).data = torch.reshape(
                params[c:c + self.num_rotation_params], [1, self.num_rotation_params])
            c += self.num_rotation_params
        if self.nonlinear_stretch_type == "classic":
            n = self.total_param_num_means
            self.kde_means.data = torch.reshape(params[c:c + n], [1, k - self.center_mean, d])
            c += n
            self.kde_log_widths.data = torch.reshape(params[c:c + k * d], [1, k, d])
            c += k * d
            if self.fit_normalization:
                self.kde_log_weights.data = torch.reshape(params[c:c + k * d], [1, k, d])
                c += k * d
            if self.add_skewness:
                self.kde_log_skew_exponents.data = torch.reshape(params[c:c + k * d], [1, k, d])
                c += k * d
        else:
            self.log_widths.data = torch.reshape(params[c:c + k * d], [1, d, k])
            c += k * d
            self.log_heights.data = torch.reshape(params[c:c + k * d], [1, d, k])
            c += k * d
            self.log_derivatives.data = torch.reshape(params[c:c + (k + 1) * d], [1, d, k + 1])
            c += (k + 1) * d
            self.boundary_points.data = torch.reshape(params[c:c + 4 * d], [1, d, 4])

    def permanent_param_names(self):
        names = []
        if self.model_offset:
            names.append("offsets")
        rot = self._rotation_param_name()
        if rot is not None:
            names.append(rot)
        if self.nonlinear_stretch_type == "classic":
            names += ["kde_means", "kde_log_widths"]
            if self.fit_normalization:
                names.append("kde_log_weights")
            if self.add_skewness:
                names.append("kde_log_skew_exponents")
        else:
            names += ["log_widths", "log_heights", "log_derivatives", "boundary_points"]
        return names

    def descriptor(self):
        return dict(code="g", dim=self.dimension, num_kde=self.num_kde, hh_iter=self.householder_iter,
                    inverse_function_type=self.inverse_function_type, inv_type=INV_TYPES[self.inverse_function_type],
                    fit_normalization=int(self.fit_normalization),
                    regulate_normalization=int(self.regulate_normalization), model_offset=int(self.model_offset),
                    w_min=float(self.width_min), w_max=float(self.width_max if self.width_max is not None else -1.0),
                    n_min=float(self.lower_bound_for_norms), n_max=float(self.upper_bound_for_norms),
                    rotation_mode=self.rotation_mode, n_rot=int(self.num_rotation_params),
                    width_mode=self.width_mode, width_clamp=self.width_clamp,
                    add_skewness=int(bool(self.add_skewness)), center_mean=self.center_mean,
                    stretch=self.nonlinear_stretch_type, default_kernel=bool(self.is_default_kernel_config),
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
# Euclidean: affine layer "t"
# =====================================================================================================================
V_POTENTIALS = {"exponential": 0, "linear": 1, "quadratic": 2, "splines": 3}     # JF_POT_* (exponential_map_s2.py:285-344)
COV_TYPES = {"identity": 0, "diagonal_symmetric": 1, "diagonal": 2, "full": 3}


class mvn_block(layer_base):
    """Affine flow (multivariate normal), symbol "t".

    Reference: layers/euclidean/multivariate_normal.py:48-190 (constructor / parameters), layers/matrix_fns.py:4-145
    (lower-triangular matrix), layers/euclidean/euclidean_base.py:9-31 (offset).  Parameter slice:
    [offset d (last layer only)] [log-diagonal: 1 | d] [lower-triangular entries d(d-1)/2, cov_type="full" only]."""

    code = "t"
    manifold = "e"

    def __init__(self, dimension, cov_type="full", use_permanent_parameters=False, model_offset=0,
                 width_smooth_saturation=1, lower_bound_for_widths=0.01, upper_bound_for_widths=100,
                 softplus_for_width=0, clamp_widths=0):
        super().__init__(dimension=dimension)
        if softplus_for_width or clamp_widths or not width_smooth_saturation or upper_bound_for_widths <= 0:
            raise NotImplementedError("'t' layer with a non-default width regulator has no sm_100a kernel "
                                      "(SURVEY.md section 8f rank 1) -- there is no CPU fallback")
        assert cov_type in COV_TYPES, cov_type
        assert lower_bound_for_widths > 0.0
        self.use_permanent_parameters = use_permanent_parameters
        self.model_offset = model_offset
        self.cov_type = cov_type
        self.width_min, self.width_max = lower_bound_for_widths, upper_bound_for_widths
        # RNG call order as in the reference: offsets (euclidean_base), then the covariance parameters
        self.offsets = None
        if self.model_offset:
            self.offsets = torch.zeros(dimension).type(torch.double).unsqueeze(0)
            if use_permanent_parameters:
                self.offsets = nn.Parameter(torch.randn(dimension).type(torch.double).unsqueeze(0))
            self.total_param_num += dimension
        n_low = int(dimension * (dimension - 1) / 2)
        if cov_type == "diagonal_symmetric":
            if use_permanent_parameters:
                self.single_diagonal_log = nn.Parameter(torch.randn(1, 1).type(torch.double))
            self.total_param_num += 1
        elif cov_type == "diagonal":
            if use_permanent_parameters:
                self.full_diagonal_log = nn.Parameter(torch.randn(1, dimension).type(torch.double))
            self.total_param_num += dimension
        elif cov_type == "full":
            if use_permanent_parameters:
                self.full_diagonal_log = nn.Parameter(torch.randn(1, dimension).type(torch.double))
                self.lower_triangular_entries = nn.Parameter(torch.randn(1, n_low).type(torch.double))
            self.total_param_num += dimension + n_low

    def _n_cov(self):
        return self.total_param_num - (self.dimension if self.model_offset else 0)

    def get_desired_init_parameters(self):
        par_list = []
        if self.model_offset:
            par_list.append(torch.ones(self.dimension) * 0.001)
        par_list.append(torch.zeros(self._n_cov()))
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num), (len(params), self.total_param_num)
        assert (self.use_permanent_parameters == 1)
        d = self.dimension
        if self.model_offset:
            self.offsets.data = params[:d]
            params = params[d:]
        if self.cov_type == "diagonal_symmetric":
            self.single_diagonal_log.data = torch.reshape(params[:1], [1, 1])
        elif self.cov_type == "diagonal":
            self.full_diagonal_log.data = torch.reshape(params[:d], [1, d])
        elif self.cov_type == "full":
            self.full_diagonal_log.data = torch.reshape(params[:d], [1, d])
            self.lower_triangular_entries.data = torch.reshape(params[d:], [1, int(d * (d - 1) / 2)])

    def permanent_param_names(self):
        names = ["offsets"] if self.model_offset else []
        if self.cov_type == "diagonal_symmetric":
            names.append("single_diagonal_log")
        elif self.cov_type == "diagonal":
            names.append("full_diagonal_log")
        elif self.cov_type == "full":
            names += ["full_diagonal_log", "lower_triangular_entries"]
        return names

    def descriptor(self):
        return dict(code="t", dim=self.dimension, cov_type=self.cov_type, cov=COV_TYPES[self.cov_type],
                    model_offset=int(self.model_offset), w_min=float(self.width_min), w_max=float(self.width_max),
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
KAPPA_MODES = {"direct_log_real_bounded": 0, "softplus_real_bounded": 1, "log_bounded": 2, "mu": 3, "mu_squared": 4,
               "quatvec": 5, "quatvec_squared": 6}


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
        if add_correlated_rq_spline_flow:
            unsupported.append("correlated rq-spline sub-flow (needs the AmortizableMLP path, SURVEY.md section 8f rank 4)")
        if kappa_prediction not in KAPPA_MODES:
            raise Exception("Unknown kappa_prediction: %s" % kappa_prediction)
        if add_rotation and rotation_mode not in ("householder", "angles", "xyz", "quaternion"):
            raise Exception("Unknown rotation mode for spheres: ", rotation_mode)
        # reference fvm_2d.py:133-138
        if kappa_prediction in ("mu", "mu_squared"):
            assert (add_rotation)
            assert (rotation_mode == "xyz")
        if kappa_prediction in ("quatvec", "quatvec_squared"):
            assert (add_rotation)
            assert (rotation_mode == "quaternion"), ("ROTATION MODE?!", rotation_mode)
        if not (0.0 <= boundary_cos_theta_identity_region < 1.0):
            raise Exception("boundary_cos_theta_identity_region must be in [0, 1)")
        if add_circular_rq_spline_flow and circular_add_rotation:
            # reference fvm_2d.py:207 asserts the same
            raise AssertionError("Currently not allowing additional S-1 rotations (circular_add_rotation must be 0)")
        if len(unsupported) > 0:
            raise NotImplementedError("'f' layer options without an sm_100a kernel yet (SURVEY.md section 8a row a13): "
                                      + ", ".join(unsupported) + " -- there is no CPU fallback")
        if spline_num_basis_functions == -1:
            assert (vertical_smooth == 1), "num_basis_functions=-1 means alternating 2/3 as basis functions and requires smooth splines."
        self.euclidean_to_sphere_as_first = euclidean_to_sphere_as_first
        self.use_permanent_parameters = use_permanent_parameters
        self.add_rotation = add_rotation
        self.rotation_mode = rotation_mode
        self.z_scaling_factor = -1.0 if inverse_z_scaling else 1.0
        self.min_kappa = min_kappa
        self.num_householder_params = 0
        self.num_householder_iter = 0
        self.kappa_prediction, self.kappa_clamping = kappa_prediction, kappa_clamping
        self.add_extra_rotation_inbetween = add_extra_rotation_inbetween
        self.boundary_cos_theta_identity_region = boundary_cos_theta_identity_region
        if add_rotation:
            # reference sphere_base.py:79-105 ("householder params stands for any rotation params here")
            if rotation_mode == "angles":
                self.num_householder_params = int(((dimension + 1) * dimension) / 2)
            elif rotation_mode == "xyz":
                self.num_householder_params = 3
            elif rotation_mode == "quaternion":
                self.num_householder_params = 4
            else:
                self.num_householder_iter = dimension + 1 if num_householder_iter == -1 else num_householder_iter
                self.num_householder_params = self.num_householder_iter * (dimension + 1)
        # RNG order as in the reference: sphere_base (rotation) first, then kappa
        if use_permanent_parameters and self.num_householder_params > 0:
            self.householder_params = nn.Parameter(torch.randn((1, self.num_householder_params)))
        self.total_param_num += self.num_householder_params
        # reference fvm_2d.py:140-146: the norm-of-the-rotation-vector predictions have no kappa parameter of their own
        self.num_loglike_kappa_params = 1 if KAPPA_MODES[kappa_prediction] <= 2 else 0
        if self.num_loglike_kappa_params:
            if use_permanent_parameters:
                self.loglike_kappa = nn.Parameter(torch.randn(1).unsqueeze(0))
            self.total_param_num += 1

        # nested pass-through sub-flows (reference fvm_2d.py:158-225: `pdf("i1_-1.00_1.00", vertical_flow_defs, ...)`
        # and `pdf("s1", circular_flow_defs, ...)` with amortize_everything / use_as_passthrough_instead_of_pdf).
        # Only their static spline configuration is needed here; they hold no parameters and draw no random numbers.
        self.add_vertical_rq_spline_flow = add_vertical_rq_spline_flow
        self.add_circular_rq_spline_flow = add_circular_rq_spline_flow
        self.vertical_layers, self.circular_layers = [], []
        self.total_num_vertical_params = 0
        if add_vertical_rq_spline_flow:
            bound = float("%.2f" % (1.0 - boundary_cos_theta_identity_region))
            for cur_r, code in enumerate(vertical_flow_defs):
                assert code == "r", "vertical_flow_defs may only contain 'r' layers"
                nb = spline_num_basis_functions
                if spline_num_basis_functions == -1:
                    nb = 3 if cur_r % 2 == 1 else 2
                self.vertical_layers.append(rational_quadratic_spline(
                    1, num_basis_functions=nb, low_boundary=-bound, high_boundary=bound,
                    fix_boundary_derivatives=-1.0 if vertical_fix_boundary_derivative == 0 else 1.0,
                    smooth_second_derivative=vertical_smooth,
                    restrict_max_min_width_height_ratio=vertical_restrict_max_min_width_height_ratio,
                    fix_first_width_n_height_to_zero=vertical_fix_first_width_n_height_to_zero,
                    also_fix_second_width_to_zero=vertical_also_fix_second_width_to_zero,
                    independent_width_height_parametrization=vertical_independent_width_height_parametrization))
            self.total_num_vertical_params = sum(l.total_param_num for l in self.vertical_layers)
            self.total_param_num += self.total_num_vertical_params
            if use_permanent_parameters:
                self.vertical_flow_params = nn.Parameter(torch.randn(1, self.total_num_vertical_params))
        self.total_num_circular_params = 0
        if add_circular_rq_spline_flow:
            for code in circular_flow_defs:
                assert code == "o", "circular_flow_defs may only contain 'o' layers"
                self.circular_layers.append(spline_1d(
                    1, euclidean_to_sphere_as_first=False, add_rotation=0, num_basis_functions=2,
                    smooth_second_derivative=1,
                    fix_first_width_n_height_to_zero=vertical_fix_first_width_n_height_to_zero,
                    also_fix_second_width_to_zero=vertical_also_fix_second_width_to_zero,
                    independent_width_height_parametrization=vertical_independent_width_height_parametrization))
            self.total_num_circular_params = sum(l.total_param_num for l in self.circular_layers)
            self.total_param_num += self.total_num_circular_params
            if use_permanent_parameters:
                self.circular_flow_params = nn.Parameter(torch.randn(1, self.total_num_circular_params))
        if len(self.vertical_layers) + len(self.circular_layers) > 4:
            raise NotImplementedError("more than 4 nested spline sub-flows in one 'f' layer (JF_MAX_NESTED)")

    # reference sphere_base.py:712-730 + fvm_2d.py:747-773
    def get_desired_init_parameters(self):
        par_list = []
        if self.num_householder_params > 0:
            par_list.append(torch.randn((self.num_householder_params)))
        if self.num_loglike_kappa_params:
            par_list.append(torch.randn((1)) - 3.0)
        par_list += [l.get_desired_init_parameters() for l in self.vertical_layers]
        par_list += [l.get_desired_init_parameters() for l in self.circular_layers]
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        n = self.num_householder_params
        if self.add_rotation:
            self.householder_params.data = params[:n].reshape(1, n)
        k = self.num_loglike_kappa_params
        if k:
            self.loglike_kappa.data = params[n:n + 1].reshape(1, 1)
        nv, nc = self.total_num_vertical_params, self.total_num_circular_params
        if self.add_vertical_rq_spline_flow:
            self.vertical_flow_params.data = params[n + k:n + k + nv].reshape(1, nv)
        if self.add_circular_rq_spline_flow:
            self.circular_flow_params.data = params[n + k + nv:n + k + nv + nc].reshape(1, nc)

    def permanent_param_names(self):
        names = (["householder_params"] if self.num_householder_params > 0 else []) + \
                (["loglike_kappa"] if self.num_loglike_kappa_params else [])
        if self.add_vertical_rq_spline_flow:
            names.append("vertical_flow_params")
        if self.add_circular_rq_spline_flow:
            names.append("circular_flow_params")
        return names

    def descriptor(self):
        off, vertical, circular = 0, [], []
        for l in self.vertical_layers:
            vertical.append(dict(l.spline_spec(), param_offset=off))
            off += l.total_param_num
        for l in self.circular_layers:
            circular.append(dict(l.spline_spec(), param_offset=off))
            off += l.total_param_num
        return dict(code="f", dim=2, add_rotation=int(self.add_rotation), hh_iter=self.num_householder_iter,
                    rotation_mode=self.rotation_mode, n_rot=int(self.num_householder_params),
                    kappa_mode=KAPPA_MODES[self.kappa_prediction], kappa_clamping=int(self.kappa_clamping),
                    extra_rotation=int(self.add_extra_rotation_inbetween),
                    identity_region=float(self.boundary_cos_theta_identity_region),
                    z_sign=float(self.z_scaling_factor), min_kappa=float(self.min_kappa),
                    first=int(self.euclidean_to_sphere_as_first), vertical=vertical, circular=circular,
                    n_params=self.total_param_num)

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


# =====================================================================================================================
# rational-quadratic splines: option bookkeeping shared by "r" and "o"
# =====================================================================================================================
class _spline_options:
    """Parameter counts and the static spline descriptor.  Reference: layers/intervals/rational_quadratic_spline.py:98-178
    and layers/spheres/splines_1d.py:38-109 (the two constructors differ only in the derivative bookkeeping)."""

    def _setup_spline(self, periodic, num_basis_functions, min_width, min_height, min_derivative,
                      fix_boundary_derivatives, smooth_second_derivative, restrict_max_min_width_height_ratio,
                      fix_first_width_n_height_to_zero, also_fix_second_width_to_zero,
                      independent_width_height_parametrization, use_permanent_parameters):
        self.num_basis_functions = num_basis_functions
        self.num_width_params = num_basis_functions
        self.num_height_params = num_basis_functions
        self.fix_first_width_n_height_to_zero = fix_first_width_n_height_to_zero
        self.also_fix_second_width_to_zero = also_fix_second_width_to_zero
        if fix_first_width_n_height_to_zero:
            self.num_width_params = num_basis_functions - 1
            self.num_height_params = num_basis_functions - 1
            if also_fix_second_width_to_zero > 0:
                self.num_width_params -= 1
        self.fix_boundary_derivatives = fix_boundary_derivatives
        self.smooth_second_derivative = smooth_second_derivative
        self.boundary_log_derivs_fixed_value = 0.0
        if fix_boundary_derivatives > 0.0:
            self.boundary_log_derivs_fixed_value = float(numpy.log(numpy.exp(fix_boundary_derivatives - min_derivative) - 1.0))
        self.deriv_num_bd_subtraction = 0
        if periodic:
            # widths/heights Parameters are created BEFORE the derivative bookkeeping in splines_1d.py:56-63
            if use_permanent_parameters:
                self.rel_log_widths = nn.Parameter(torch.randn(self.num_width_params).type(torch.double).unsqueeze(0))
                self.rel_log_heights = nn.Parameter(torch.randn(self.num_height_params).type(torch.double).unsqueeze(0))
            if smooth_second_derivative == 1:
                assert (num_basis_functions == 2), "Only support 2 basis functions for smooth derivative!"
                self.deriv_num_bd_subtraction = 3
            elif fix_boundary_derivatives > 0.0:
                self.deriv_num_bd_subtraction = 2
                assert (fix_boundary_derivatives > min_derivative), "Fixed boundary derivative should be larger than min derivative!"
            else:
                self.deriv_num_bd_subtraction = 1
        else:
            if smooth_second_derivative == 1:
                assert ((num_basis_functions == 2) or (num_basis_functions == 3)), "Only support 2/3 basis functions for smooth derivative!"
                if num_basis_functions == 2:
                    self.deriv_num_bd_subtraction = 3 if fix_boundary_derivatives > 0.0 else 1
                else:
                    self.deriv_num_bd_subtraction = 4 if fix_boundary_derivatives > 0.0 else 2
            elif fix_boundary_derivatives > 0.0:
                self.deriv_num_bd_subtraction = 2
                assert (fix_boundary_derivatives > min_derivative)
        self.num_derivative_params = num_basis_functions + 1 - self.deriv_num_bd_subtraction
        if smooth_second_derivative and num_basis_functions == 3:
            self.num_width_params -= 1
            self.num_height_params -= 1
        if use_permanent_parameters:
            if not periodic:
                self.rel_log_widths = nn.Parameter(torch.randn(self.num_width_params).type(torch.double).unsqueeze(0))
                self.rel_log_heights = nn.Parameter(torch.randn(self.num_height_params).type(torch.double).unsqueeze(0))
            if self.num_derivative_params > 0:
                self.rel_log_derivatives = nn.Parameter(torch.randn(self.num_derivative_params).type(torch.double).unsqueeze(0))
        self.total_param_num += self.num_width_params + self.num_height_params + self.num_derivative_params
        self.min_width, self.min_height, self.min_derivative = min_width, min_height, min_derivative
        self.restrict_max_min_width_height_ratio = restrict_max_min_width_height_ratio
        self.independent_width_height_parametrization = independent_width_height_parametrization
        self._periodic = periodic

    def _spline_spec(self, lo, hi, natural_direction):
        if self.smooth_second_derivative:
            kind = "circular" if self._periodic else "smooth"
        else:
            kind = "plain"
        if self.fix_boundary_derivatives > 0.0 and kind != "circular":
            bd_mode = 1
        elif self._periodic and kind == "plain":
            bd_mode = 2
        else:
            bd_mode = 0
        return dict(kind=kind, n_bins=self.num_basis_functions, n_w=self.num_width_params, n_h=self.num_height_params,
                    n_d=self.num_derivative_params, fix_first=int(self.fix_first_width_n_height_to_zero),
                    fix_second=int(self.also_fix_second_width_to_zero),
                    indep=int(self.independent_width_height_parametrization), bd_mode=bd_mode,
                    bd_fixed=float(self.boundary_log_derivs_fixed_value), lo=float(lo), hi=float(hi),
                    min_w=float(self.min_width), min_h=float(self.min_height), min_d=float(self.min_derivative),
                    max_ratio=float(self.restrict_max_min_width_height_ratio), natural_direction=int(natural_direction),
                    param_offset=0)

    def _spline_init(self):
        n = self.num_width_params + self.num_height_params + self.num_derivative_params
        return torch.zeros(n) if self.smooth_second_derivative else torch.ones(n) * 0.54

    def _spline_init_params(self, params):
        c = 0
        self.rel_log_widths.data[0, :] = params[c:c + self.num_width_params]
        c += self.num_width_params
        self.rel_log_heights.data[0, :] = params[c:c + self.num_height_params]
        c += self.num_height_params
        if self.num_derivative_params > 0:
            self.rel_log_derivatives.data[0, :] = params[c:c + self.num_derivative_params]

    def _spline_names(self):
        return ["rel_log_widths", "rel_log_heights"] + (["rel_log_derivatives"] if self.num_derivative_params > 0 else [])


# =====================================================================================================================
# Interval: rational-quadratic spline "r"
# =====================================================================================================================
class rational_quadratic_spline(layer_base, _spline_options):
    """Symbol "r".  Reference: layers/intervals/rational_quadratic_spline.py:62-178 (constructor),
    layers/intervals/interval_base.py:8-31.  Parameter slice: [widths][heights][derivatives]."""

    code = "r"
    manifold = "i"

    def __init__(self, dimension, num_basis_functions=10, euclidean_to_interval_as_first=0, use_permanent_parameters=0,
                 low_boundary=0, high_boundary=1.0, min_width=1e-4, min_height=1e-4, min_derivative=1e-4,
                 fix_boundary_derivatives=-1.0, smooth_second_derivative=0, restrict_max_min_width_height_ratio=-1.0,
                 fix_first_width_n_height_to_zero=0, also_fix_second_width_to_zero=0,
                 independent_width_height_parametrization=0):
        super().__init__(dimension=dimension)
        assert (self.dimension == 1)
        if num_basis_functions > 32:
            raise NotImplementedError("'r' with more than 32 basis functions (JF_MAX_BINS)")
        self.use_permanent_parameters = use_permanent_parameters
        self.low_boundary, self.high_boundary = low_boundary, high_boundary
        self.interval_width = high_boundary - low_boundary
        self.euclidean_to_interval_as_first = euclidean_to_interval_as_first
        assert (self.high_boundary > self.low_boundary)
        self._setup_spline(False, num_basis_functions, min_width, min_height, min_derivative, fix_boundary_derivatives,
                           smooth_second_derivative, restrict_max_min_width_height_ratio,
                           fix_first_width_n_height_to_zero, also_fix_second_width_to_zero,
                           independent_width_height_parametrization, use_permanent_parameters)

    def get_desired_init_parameters(self):
        return self._spline_init()

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        self._spline_init_params(params)

    def permanent_param_names(self):
        return self._spline_names()

    def spline_spec(self):
        return self._spline_spec(self.low_boundary, self.high_boundary, 1)

    def descriptor(self):
        return dict(code="r", dim=1, first=int(self.euclidean_to_interval_as_first), lo=float(self.low_boundary),
                    hi=float(self.high_boundary), spline=self.spline_spec(), n_params=self.total_param_num)

    def _embedding_conditional_return(self, x):
        return x

    def _embedding_conditional_return_num(self):
        return self.dimension

    def _get_layer_base_dimension(self):
        return self.dimension

    def transform_target_space(self, x, log_det=0.0, transform_from="default", transform_to="embedding"):
        return x, log_det


# =====================================================================================================================
# S1 layers: "o" (circular spline) and "m" (Moebius)
# =====================================================================================================================
class _s1_base(layer_base):
    """Reference: layers/spheres/sphere_base.py:42-110 with dimension 1 (rotation = Householder reflections in R^2,
    parameters first in the layer slice)."""

    manifold = "s"

    def _setup_s1(self, euclidean_to_sphere_as_first, add_rotation, use_permanent_parameters):
        self.euclidean_to_sphere_as_first = euclidean_to_sphere_as_first
        self.use_permanent_parameters = use_permanent_parameters
        self.add_rotation = add_rotation
        self.num_householder_iter = 0
        self.num_householder_params = 0
        if add_rotation:
            self.num_householder_iter = 2
            self.num_householder_params = 4
        if use_permanent_parameters and self.num_householder_params > 0:
            self.householder_params = nn.Parameter(torch.randn((1, self.num_householder_params)))
        self.total_param_num += self.num_householder_params

    def _embedding_conditional_return(self, x):
        if x.shape[1] == self.dimension:
            return torch.cat([torch.cos(x), torch.sin(x)], dim=1)
        return x

    def _embedding_conditional_return_num(self):
        return self.dimension + 1

    def _get_layer_base_dimension(self):
        if self.always_parametrize_in_embedding_space and not self.euclidean_to_sphere_as_first:
            return self.dimension + 1
        return self.dimension


class spline_1d(_s1_base, _spline_options):
    """Symbol "o".  Reference: layers/spheres/splines_1d.py:9-109."""

    code = "o"

    def __init__(self, dimension=1, euclidean_to_sphere_as_first=True, add_rotation=1, natural_direction=1,
                 use_permanent_parameters=False, num_basis_functions=2, min_width=1e-4, min_height=1e-4,
                 min_derivative=1e-4, fix_boundary_derivatives=-1.0, smooth_second_derivative=0,
                 fix_first_width_n_height_to_zero=0, also_fix_second_width_to_zero=0,
                 independent_width_height_parametrization=0):
        super().__init__(dimension=1)
        if dimension != 1:
            raise Exception("The moebius flow is defined for dimension 1, but dimension %d is handed over" % (dimension))
        if num_basis_functions > 32:
            raise NotImplementedError("'o' with more than 32 basis functions (JF_MAX_BINS)")
        self._setup_s1(euclidean_to_sphere_as_first, add_rotation, use_permanent_parameters)
        self.natural_direction = natural_direction
        self._setup_spline(True, num_basis_functions, min_width, min_height, min_derivative, fix_boundary_derivatives,
                           smooth_second_derivative, -1.0, fix_first_width_n_height_to_zero,
                           also_fix_second_width_to_zero, independent_width_height_parametrization,
                           use_permanent_parameters)

    def get_desired_init_parameters(self):
        par_list = []
        if self.num_householder_params > 0:
            par_list.append(torch.randn((self.num_householder_params)))
        par_list.append(self._spline_init())
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        n = self.num_householder_params
        if self.add_rotation:
            self.householder_params.data = params[:n].reshape(1, n)
        self._spline_init_params(params[n:])

    def permanent_param_names(self):
        return (["householder_params"] if self.num_householder_params > 0 else []) + self._spline_names()

    def spline_spec(self):
        return self._spline_spec(0.0, 2 * math.pi, self.natural_direction)

    def descriptor(self):
        return dict(code="o", dim=1, add_rotation=int(self.add_rotation), hh_iter=self.num_householder_iter,
                    first=int(self.euclidean_to_sphere_as_first), natural_direction=int(self.natural_direction),
                    spline=self.spline_spec(), n_params=self.total_param_num)


class moebius(_s1_base):
    """Symbol "m".  Reference: layers/spheres/moebius_1d.py:8-55.  Parameter slice: [householder 4][K x (x, y, log-radius,
    log-weight)]."""

    code = "m"

    def __init__(self, dimension=1, euclidean_to_sphere_as_first=True, add_rotation=0, natural_direction=0,
                 use_permanent_parameters=False, use_moebius_xyz_parametrization=True, num_basis_functions=5):
        super().__init__(dimension=1)
        if dimension != 1:
            raise Exception("The moebius flow is defined for dimension 1, but dimension %d is handed over" % (dimension))
        if not use_moebius_xyz_parametrization:
            raise NotImplementedError("'m' with use_moebius_xyz_parametrization=0 has no sm_100a kernel")
        if num_basis_functions > 16:
            raise NotImplementedError("'m' with more than 16 basis functions")
        self._setup_s1(euclidean_to_sphere_as_first, add_rotation, use_permanent_parameters)
        self.num_basis_functions = num_basis_functions
        self.num_omega_pars = 4
        self.total_param_num += self.num_basis_functions * self.num_omega_pars
        if use_permanent_parameters:
            self.moebius_pars = nn.Parameter(torch.randn(self.num_basis_functions, self.num_omega_pars).type(torch.double).unsqueeze(0))
        self.natural_direction = natural_direction

    def get_desired_init_parameters(self):
        par_list = []
        if self.num_householder_params > 0:
            par_list.append(torch.randn((self.num_householder_params)))
        par_list.append(torch.randn((self.num_basis_functions * self.num_omega_pars)))
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        n = self.num_householder_params
        if self.add_rotation:
            self.householder_params.data = params[:n].reshape(1, n)
        self.moebius_pars.data = params[n:].reshape(1, self.num_basis_functions, self.num_omega_pars)

    def permanent_param_names(self):
        return (["householder_params"] if self.num_householder_params > 0 else []) + ["moebius_pars"]

    def descriptor(self):
        return dict(code="m", dim=1, add_rotation=int(self.add_rotation), hh_iter=self.num_householder_iter,
                    first=int(self.euclidean_to_sphere_as_first), natural_direction=int(self.natural_direction),
                    K=self.num_basis_functions, n_params=self.total_param_num)


# =====================================================================================================================
# S2: exponential-map flow "v"
# =====================================================================================================================
class exponential_map_s2(layer_base):
    """Symbol "v" with the exponential potential.  Reference: layers/spheres/exponential_map_s2.py:64-151.
    Parameter slice: [householder iter*3][potential_pars 5 x K: mean xyz, log-weight, log-beta]."""

    code = "v"
    manifold = "s"

    def __init__(self, dimension, euclidean_to_sphere_as_first=False, use_permanent_parameters=False,
                 exp_map_type="linear", natural_direction=0, num_components=10, add_rotation=0,
                 max_num_newton_iter=1000, mean_parametrization="old"):
        super().__init__(dimension=dimension)
        if dimension != 2:
            raise Exception("The moebius flow should be used for dimension 2!")
        unsupported = []
        if exp_map_type not in V_POTENTIALS:
            # "nn" raises in the reference itself ("only used for testing"); anything else is unknown to the reference too
            unsupported.append("exp_map_type=%s" % exp_map_type)
        if mean_parametrization != "old":
            # (the reference's own evaluation of "householder" raises a TypeError: exponential_map_s2.py:270 calls
            #  compute_householder_matrix without its hh_iter argument)
            unsupported.append("mean_parametrization=%s" % mean_parametrization)
        if num_components > 16:
            unsupported.append("num_components > 16")
        if len(unsupported) > 0:
            raise NotImplementedError("'v' layer options without an sm_100a kernel yet (SURVEY.md section 8a row a15): "
                                      + ", ".join(unsupported) + " -- there is no CPU fallback")
        self.euclidean_to_sphere_as_first = euclidean_to_sphere_as_first
        self.use_permanent_parameters = use_permanent_parameters
        self.add_rotation = add_rotation
        self.num_householder_iter = 0
        self.num_householder_params = 0
        if add_rotation:
            self.num_householder_iter = dimension + 1
            self.num_householder_params = self.num_householder_iter * (dimension + 1)
        if use_permanent_parameters and self.num_householder_params > 0:
            self.householder_params = nn.Parameter(torch.randn((1, self.num_householder_params)))
        self.total_param_num += self.num_householder_params
        self.num_components = num_components
        self.exp_map_type = exp_map_type
        self.natural_direction = natural_direction
        self.max_num_newton_iter = max_num_newton_iter
        # mean direction (3) + log weight, + log beta for the exponential potential (exponential_map_s2.py:124-127)
        # "splines": + 10 widths, 10 heights, 11 derivatives of the spline whose integral is the potential (:129-131)
        self.num_spline_basis_functions = 10
        self.num_potential_pars = 3 + (2 if exp_map_type == "exponential" else
                                       (1 + 3 * self.num_spline_basis_functions + 1 if exp_map_type == "splines" else 1))
        if use_permanent_parameters:
            self.potential_pars = nn.Parameter(torch.randn(self.num_potential_pars, self.num_components).unsqueeze(0))
        self.total_param_num += self.num_potential_pars * self.num_components

    def get_desired_init_parameters(self):
        par_list = []
        if self.num_householder_params > 0:
            par_list.append(torch.randn((self.num_householder_params)))
        par_list.append(torch.randn((self.num_potential_pars * self.num_components)))
        return torch.cat(par_list)

    def init_params(self, params):
        assert (len(params) == self.total_param_num)
        n = self.num_householder_params
        if self.add_rotation:
            self.householder_params.data = params[:n].reshape(1, n)
        self.potential_pars.data = params[n:].reshape(1, self.num_potential_pars, self.num_components)

    def permanent_param_names(self):
        return (["householder_params"] if self.num_householder_params > 0 else []) + ["potential_pars"]

    def descriptor(self):
        return dict(code="v", dim=2, add_rotation=int(self.add_rotation), hh_iter=self.num_householder_iter,
                    first=int(self.euclidean_to_sphere_as_first), natural_direction=int(self.natural_direction),
                    K=self.num_components, max_iter=int(self.max_num_newton_iter), n_params=self.total_param_num,
                    exp_map_type=self.exp_map_type, potential=V_POTENTIALS[self.exp_map_type])

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
