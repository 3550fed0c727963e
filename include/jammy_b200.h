/*
 * jammy_b200.h -- C-ABI of the B200-native (sm_100a) jammy_flows hot path.
 *
 * One shared library (libjammy_b200.so), plain pointers and sizes, no torch/C++ types in any signature.
 * Every entry point replaces one interface of the reference (thoglu/jammy_flows v1.1.0, paths relative to
 * /root/reference/jammy_flows/):
 *
 *   jf_subpdf_apply      <- layer plugin API  layers/layer_base.py:58-70 (`flow_mapping` / `inv_flow_mapping`
 *                           with `[x, log_det]` and `extra_inputs`), chained over all layers of one sub-pdf exactly as
 *                           the loops in main/default.py:998-1031 (log_pdf) and :1482-1506 (sampling) do, including
 *                           the sub-pdf's base chart and the N(0,1) base log-density (main/default.py:1110-1117).
 *   jf_mlp_forward       <- the parameter generator nn.Sequential(Linear,Tanh,...,Linear), main/default.py:654-670,
 *                           called at main/default.py:956 / :1438 with cat([cond] + embedded previous targets).
 *   jf_pdf_logpdf        <- pdf.forward / all_layer_inverse, main/default.py:1059-1117 and :879-1057.
 *   jf_pdf_sample        <- pdf._obtain_sample / all_layer_forward, main/default.py:1533-1707 and :1373-1531
 *                           (base normals are supplied by the caller, like `predefined_target_input`).
 *   jf_pdf_logpdf_host / jf_pdf_sample_host
 *                        <- the same two calls with HOST buffers: pipelined H2D -> kernels -> D2H (the end-to-end path).
 *
 * Ownership: the caller allocates every buffer (inputs, outputs, workspace) and passes raw device pointers
 * (host pointers only for the *_host entries); the library never allocates or frees device memory.
 * Threading: re-entrant; all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*).  No entry
 * point synchronises the device except the *_host entries: they run on two internal NON-BLOCKING streams (created once
 * per host thread and device, then kept) and synchronise those before returning; whatever the caller prepared on other
 * streams (parameters, workspace, status words) must be complete when a *_host entry is called.
 * ABI history: v2 added JfSplineDesc and the spline / S1 / "v" / "t" layer kinds; v3 added the non-default "g" options
 * (rotation_mode, width_mode, width_clamp, skew, center_mean, stretch, clamp_lo/hi in JfLayerDesc).  v4 added
 * jf_subpdf_apply_generated (fused generator + layer chain), jf_mlp_backward (tensor-core generator gradient) and the
 * "f" rotation modes / kappa link functions (JF_ROT_XYZ, JF_ROT_QUATERNION, JF_KAPPA_* in rotation_mode / width_mode).
 * Errors: return value 0 = ok, <0 = invalid/unsupported descriptor (JF_ERR_*), >0 = CUDA runtime error code.
 * Numerical conditions (non-finite values, unconverged root finds, out-of-range inputs) are counted in the device
 * int64 array `status[JF_STATUS_WORDS]` and are read lazily by the caller (no implicit sync), mirroring the reference's
 * fail-safes in layers/bisection_n_newton.py:84-133.
 */
#ifndef JAMMY_B200_H
#define JAMMY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JF_ABI_VERSION 5

#define JF_MAX_LAYERS 16
#define JF_MAX_SUBPDFS 8
#define JF_MAX_MLP_LINEAR 6
#define JF_MAX_MLP_SEGMENTS 10
#define JF_MAX_DIM 16      /* max intrinsic dimension of one Euclidean sub-pdf */
#define JF_MAX_KDE 32      /* max num_kde of a "g" layer */
#define JF_STATUS_WORDS 4
#define JF_MAX_NESTED 4    /* spline sub-flows (vertical + circular) of one "f" layer */
#define JF_MAX_BINS 32     /* max num_basis_functions of a rational-quadratic spline */

/* dtype of all floating-point buffers of a call */
#define JF_F32 0
#define JF_F64 1

/* direction */
#define JF_DIR_LOGPDF 0 /* target -> base, reference `inv_flow_mapping` */
#define JF_DIR_SAMPLE 1 /* base -> target, reference `flow_mapping` */

/* layer kinds (reference layer codes, flow_options.py:25-240) */
#define JF_LAYER_GF 1  /* "g" gf_block */
#define JF_LAYER_FVM 2 /* "f" fisher_von_mises_2d ("n" alias), optionally with vertical/circular spline sub-flows */
#define JF_LAYER_RQS 3      /* "r" rational_quadratic_spline on an interval */
#define JF_LAYER_S1SPLINE 4 /* "o" spline_1d on the circle */
#define JF_LAYER_MOEBIUS 5  /* "m" moebius on the circle */
#define JF_LAYER_EXPMAP 6   /* "v" exponential_map_s2 (exponential potential) */
#define JF_LAYER_MVN 7      /* "t" mvn_block: affine layer with a lower-triangular matrix; cov_type in inv_type */

/* covariance type of "t" (layers/euclidean/multivariate_normal.py:56), stored in JfLayerDesc.inv_type */
#define JF_COV_IDENTITY 0
#define JF_COV_DIAGONAL_SYMMETRIC 1
#define JF_COV_DIAGONAL 2
#define JF_COV_FULL 3

/* spline variants (layers/spline_fns.py:45-186 / :361-559 / :561-760) */
#define JF_SPLINE_PLAIN 0
#define JF_SPLINE_SMOOTH 1
#define JF_SPLINE_CIRCULAR 2
/* where the knot derivatives of a spline come from */
#define JF_BD_PARAMS 0   /* all from parameters */
#define JF_BD_FIXED 1    /* boundary derivatives fixed (fix_boundary_derivatives > 0) */
#define JF_BD_PERIODIC 2 /* "o", not smooth: first derivative copied to the end (splines_1d.py:176) */

/* inverse-CDF stage of "g" (gaussianization_flow.py:480-671) */
#define JF_INV_ISIGMOID 0
#define JF_INV_PARTLY_PRECISE 1
#define JF_INV_FULL_PADE 2
#define JF_INV_PARTLY_CRUDE 3

/* normalisation handling of "g" */
#define JF_NORM_NONE 0      /* fit_normalization=0: uniform weights */
#define JF_NORM_RAW 1       /* fit_normalization=1, regulate_normalization=0: log-softmax of the raw values */
#define JF_NORM_REGULATED 2 /* regulated into [n_min, n_min+n_max] first (gaussianization_flow.py:342) */

/* rotation of "g" (gaussianization_flow.py:141-205, :711-800) */
#define JF_ROT_HOUSEHOLDER 0 /* hh_iter reflections (default) */
#define JF_ROT_NONE 1
#define JF_ROT_ANGLES 2      /* chain of d(d-1)/2 Givens rotations */
#define JF_ROT_CAYLEY 3      /* d = 2, one parameter */
#define JF_ROT_TRIANGULAR 4  /* triangular_combination: unit lower x zero-sum diagonal x unit upper, d(d-1)+d-1 parameters */
/* rotation of "f" in the embedding space of S2 (layers/spheres/sphere_base.py:79-91, :112-216): JF_ROT_HOUSEHOLDER,
 * JF_ROT_ANGLES (3 Givens angles), and */
#define JF_ROT_XYZ 5         /* 3 parameters: the z axis is rotated onto the normalised vector */
#define JF_ROT_QUATERNION 6  /* 4 parameters: rotation of the (unnormalised) quaternion */
/* concentration of "f" (layers/spheres/fvm_2d.py:108-138, :289-330), stored in JfLayerDesc.width_mode; width_clamp =
 * kappa_clamping (raw parameter clamped at -5 from below) */
#define JF_KAPPA_DIRECT_LOG 0       /* exp(raw) + min_kappa (default) */
#define JF_KAPPA_SOFTPLUS 1         /* softplus(raw) + min_kappa */
#define JF_KAPPA_LOG_BOUNDED 2      /* exp(softplus(raw) + log min_kappa) */
#define JF_KAPPA_MU 3               /* |rotation parameters| (rotation mode xyz); no kappa parameter */
#define JF_KAPPA_MU_SQUARED 4
#define JF_KAPPA_QUATVEC 5          /* |vector part of the quaternion| (rotation mode quaternion); no kappa parameter */
#define JF_KAPPA_QUATVEC_SQUARED 6

/* width regulator of "g" (gaussianization_flow.py:264-317) */
#define JF_WIDTH_SMOOTH 0   /* w = w_min + 1/(1/w_max + exp(-raw))   (width_smooth_saturation=1, default) */
#define JF_WIDTH_EXP 1      /* w = w_min + exp(raw) */
#define JF_WIDTH_SOFTPLUS 2 /* w = w_min + softplus(raw) */

/* non-linear stretch of "g" */
#define JF_STRETCH_CLASSIC 0 /* logistic mixture CDF + inverse-CDF stage */
#define JF_STRETCH_RQS 1     /* rational-quadratic spline with linear tails (spline_fns.py:188-358) */

/* potential of "v" (exponential_map_s2.py:285-344), stored in JfLayerDesc.inv_type */
#define JF_POT_EXPONENTIAL 0 /* grad = sum_k w_k mu_k exp(beta_k (x.mu_k - 1)); parameters [5, K] */
#define JF_POT_LINEAR 1      /* grad = sum_k w_k mu_k;                         parameters [4, K] */
#define JF_POT_QUADRATIC 2   /* grad = sum_k w_k mu_k (x.mu_k);                parameters [4, K] */
#define JF_POT_SPLINES 3   /* v: integral-of-a-spline potential, 10-bin rational-quadratic spline per component (exponential_map_s2.py:346-388) */

/* status words */
#define JF_STATUS_NONFINITE 0
#define JF_STATUS_UNCONVERGED 1
#define JF_STATUS_OUT_OF_RANGE 2
#define JF_STATUS_ITERATIONS 3 /* total root-finder function evaluations (diagnostics) */

/* error codes */
#define JF_OK 0
#define JF_ERR_BAD_DESC -1
#define JF_ERR_UNSUPPORTED -2
#define JF_ERR_BAD_ARG -3
#define JF_ERR_WORKSPACE -4

/* One rational-quadratic spline transformation as an "r" / "o" layer configures it
 * (layers/intervals/rational_quadratic_spline.py:62-178, layers/spheres/splines_1d.py:9-109). */
typedef struct JfSplineDesc {
    int32_t kind;              /* JF_SPLINE_* */
    int32_t n_bins;            /* num_basis_functions */
    int32_t n_w, n_h, n_d;     /* raw width / height / derivative parameters in the slice, in this order */
    int32_t fix_first;         /* fix_first_width_n_height_to_zero */
    int32_t fix_second;        /* also_fix_second_width_to_zero */
    int32_t indep;             /* independent_width_height_parametrization: heights += widths */
    int32_t bd_mode;           /* JF_BD_* */
    int32_t natural_direction; /* "o": which direction is the closed-form one (splines_1d.py:162-164) */
    int32_t param_offset;      /* nested in "f": offset from the first sub-flow parameter; else 0 */
    int32_t reserved;
    double lo, hi;             /* support = image */
    double min_w, min_h, min_d;
    double bd_fixed;           /* softplus^-1 of the fixed boundary derivative (rational_quadratic_spline.py:117-120) */
    double max_ratio;          /* restrict_max_min_width_height_ratio, <= 0: off */
    double reserved1;
} JfSplineDesc;

typedef struct JfLayerDesc {
    int32_t kind;         /* JF_LAYER_* */
    int32_t dim;          /* intrinsic dimension of the layer */
    int32_t n_params;     /* length of the layer's slice in the sub-pdf parameter vector */
    int32_t param_offset; /* start of that slice (layers are stored in flow order, main/default.py:1488) */
    int32_t K;            /* g: num_kde */
    int32_t hh_iter;      /* number of Householder reflections (g: in R^d; f: in R^3); 0 = no rotation */
    int32_t inv_type;     /* g: JF_INV_* ; t: JF_COV_* ; v: JF_POT_* */
    int32_t norm_mode;    /* g: JF_NORM_* */
    int32_t has_offset;   /* g: model_offset (euclidean_base.py:34-75); offset params come first in the slice */
    int32_t first;        /* s1/s2/interval: layer also applies the base chart of the sub-pdf (sphere_base.py:637-648,
                             interval_base.py:61-79) */
    int32_t natural_direction; /* o, m, v */
    int32_t max_iter;     /* v: max_num_newton_iter */
    int32_t n_vertical;   /* f: number of nested "r" sub-flows, spline[0 .. n_vertical) */
    int32_t n_circular;   /* f: number of nested "o" sub-flows, spline[n_vertical .. n_vertical+n_circular) */
    int32_t rotation_mode; /* g, f: JF_ROT_* (0 = Householder, the default) */
    int32_t width_mode;    /* g: JF_WIDTH_* ; f: JF_KAPPA_* */
    int32_t width_clamp;   /* g: clamp_widths: raw width parameter clamped into [clamp_lo, clamp_hi] first; f: kappa_clamping */
    int32_t skew;          /* g: add_skewness (one more K*d block of log skew exponents at the end of the slice);
                              f: add_extra_rotation_inbetween (fvm_2d.py:381-402) */
    int32_t center_mean;   /* g: only K-1 means per dimension are parameters (gaussianization_flow.py:841-848) */
    int32_t stretch;       /* g: JF_STRETCH_* */
    double w_min, w_max; /* g: width bounds (gaussianization_flow.py:300-317) */
    double n_min, n_max; /* g: norm bounds (gaussianization_flow.py:342) */
    double z_sign;       /* f: z_scaling_factor (+1/-1, fvm_2d.py:96-99) */
    double min_kappa;    /* f: kappa = exp(raw) + min_kappa (fvm_2d.py:123) */
    double lo, hi;       /* r: interval boundaries */
    double clamp_lo, clamp_hi; /* g: see width_clamp (clamp_hi may be +inf); f: clamp_lo = boundary_cos_theta_identity_region */
    JfSplineDesc spline[JF_MAX_NESTED]; /* r, o: spline[0]; f: nested sub-flows */
} JfLayerDesc;

typedef struct JfSubPdfDesc {
    int32_t manifold;   /* 'e', 's' or 'i' (ASCII) */
    int32_t dim;        /* intrinsic dimension */
    int32_t n_layers;
    int32_t n_params;   /* sum of the layers' n_params */
    JfLayerDesc layers[JF_MAX_LAYERS];
} JfSubPdfDesc;

typedef struct JfMlpDesc {
    int32_t n_linear;                     /* number of Linear layers; tanh between them */
    int32_t dims[JF_MAX_MLP_LINEAR + 1];  /* in, hidden..., out */
    int32_t n_segments;                   /* input = concatenation of column blocks */
    int32_t seg_cols[JF_MAX_MLP_SEGMENTS];
} JfMlpDesc;

/*
 * Apply all layers of ONE sub-pdf to B rows.
 *   in / out        [B, in_cols] / [B, out_cols] row-major with leading dimensions ld_in / ld_out (elements).
 *                   LOGPDF: in = target coordinates (e: d; s2: theta,phi; s1: angle; interval: x), out = base coordinates.
 *                   SAMPLE: in = base normals, out = target coordinates.
 *   params          raw parameters in the reference's `extra_inputs` order; element (j,row) is
 *                   params[j*p_stride_param + row*p_stride_row].  p_stride_row == 0 means one shared vector
 *                   (the reference's permanent nn.Parameters, broadcast over the batch).
 *   logdet_in/out   [B]; in may be NULL (= 0).  out = in +/- the log-det terms (functional, in is not modified
 *                   unless the two pointers alias -- reference tests/test_general.py:533-550).
 *   logbase_in/out  [B] optional (NULL to skip): out = in + sum_j log N(z_j; 0,1) over this sub-pdf's base coords.
 *   emb_out         optional [B, ld_emb]: embedding of the TARGET coordinates handed to later MLPs
 *                   (`_embedding_conditional_return`, main/default.py:1050-1053): e/interval -> identity (d),
 *                   s2 -> (x,y,z), s1 -> (cos, sin).
 */
int jf_subpdf_apply(const JfSubPdfDesc* desc, int dtype, int direction,
                    const void* in, int64_t ld_in,
                    const void* params, int64_t p_stride_param, int64_t p_stride_row,
                    const void* logdet_in, void* logdet_out,
                    const void* logbase_in, void* logbase_out,
                    void* out, int64_t ld_out,
                    void* emb_out, int64_t ld_emb,
                    int64_t B, int64_t* status, void* stream);

/*
 * Training: gradient of the log_pdf of ONE Euclidean "g" sub-pdf with respect to its PER-ROW raw parameters (what the
 * reference obtains from autograd through gaussianization_flow.py:995-1057 / :699-861 / :389-454); the caller chains it
 * into the parameter generator's backward (main/default.py:956).
 *   x            [B, d] target coordinates (ld_x), as passed to jf_subpdf_apply in the LOGPDF direction
 *   params       per-row raw parameters, element (j,row) at params[j*p_stride_param + row*p_stride_row], p_stride_row != 0
 *   grad_logp    [B] upstream gradient of log p = log N(base) + logdet per row, or NULL (= 1)
 *   grad_params  out, indexed like params
 * Rows whose base point lies in the Pade tails of an inverse-normal stage (|z| > 5.33, ~1e-7 of all rows) get a zero
 * gradient for that stage and are counted in status[JF_STATUS_OUT_OF_RANGE].
 */
int jf_subpdf_backward(const JfSubPdfDesc* desc, int dtype,
                       const void* x, int64_t ld_x,
                       const void* params, int64_t p_stride_param, int64_t p_stride_row,
                       const void* grad_logp, void* grad_params,
                       int64_t B, int64_t* status, void* stream);

/*
 * Training, one pass: the LOGPDF direction of jf_subpdf_apply (base point, logdet, log N(base)) AND the gradients of
 * jf_subpdf_backward from ONE kernel (csrc/gf_fb.cuh, worker = (row, dimension)); the backward needs the recomputed
 * forward anyway, so a training step that calls this instead of apply + backward saves the forward kernel.  With
 * grad_logp = NULL the gradient buffers hold the per-row JACOBIAN d log_pdf[row] / d params[., row], which a caller can
 * scale by the upstream gradient later (jf_mlp_backward's row_scale) -- that is how the autograd path of
 * jammy_flows_b200.pdf.forward uses it.
 *   grad_x       optional out [B, d] (ld_gx): d log_pdf / d x (the reference differentiates through the evaluation points too)
 *   base_out     optional out [B, d] (ld_out); logdet_out, logbase_out: optional out [B]
 * Other arguments and the status counters as in jf_subpdf_backward.
 */
int jf_subpdf_forward_backward(const JfSubPdfDesc* desc, int dtype,
                               const void* x, int64_t ld_x,
                               const void* params, int64_t p_stride_param, int64_t p_stride_row,
                               const void* grad_logp, void* grad_params, void* grad_x, int64_t ld_gx,
                               void* base_out, int64_t ld_out, void* logdet_out, void* logbase_out,
                               int64_t B, int64_t* status, void* stream);

/*
 * Training through SAMPLES (reference: pdf.sample(allow_gradients=True), main/default.py:1342, which differentiates through
 * the bisection / Newton iterations of layers/bisection_n_newton.py): backward of the sampling direction of a Euclidean
 * "g" sub-pdf at the sample x = T(z; params) that jf_subpdf_apply(JF_DIR_SAMPLE) returned.  The layer inputs are
 * recovered by running the closed-form log_pdf direction from x (no root finder), then every element is differentiated
 * in implicit-function form (csrc/gf_fb.cuh, MODE 1).
 *   x            [B, d] the samples (ld_x)              params  as in jf_subpdf_backward
 *   grad_x       [B, d] (ld_gx) cotangent of x, or NULL (= 0);   grad_logp  [B] cotangent of log_pdf(x), or NULL (= 0)
 *   grad_params  out, indexed like params
 *   grad_z       optional out [B, d] (ld_gz): cotangent of the base point z
 */
int jf_subpdf_sample_backward(const JfSubPdfDesc* desc, int dtype,
                              const void* x, int64_t ld_x,
                              const void* params, int64_t p_stride_param, int64_t p_stride_row,
                              const void* grad_x, int64_t ld_gx, const void* grad_logp,
                              void* grad_params, void* grad_z, int64_t ld_gz,
                              int64_t B, int64_t* status, void* stream);

/*
 * Training, non-Euclidean sub-pdfs (S2 "f" / "v", S1 "o" / "m", interval "r"): per-row JACOBIAN of
 * log_pdf = log N(base) + logdet with respect to the raw parameters and the coordinates -- what the reference obtains
 * from autograd through layers/spheres/fvm_2d.py, exponential_map_s2.py, splines_1d.py, moebius_1d.py,
 * layers/intervals/rational_quadratic_spline.py and layers/spline_fns.py.  Forward-mode sweep of the value-path device
 * code over dual numbers (csrc/jac_sweep.cuh): one pass per parameter / coordinate, every option of the value path covered
 * by construction.  The values themselves come from jf_subpdf_apply.
 *   params       element (j,row) at params[j*p_stride_param + row*p_stride_row]; p_stride_row == 0: shared parameters
 *   jac_params   out, element (j,row) at jac_params[j*jac_stride_param + row]   (at most 160 parameters per sub-pdf)
 *   jac_x        out, optional [B, dim] (ld_jx): d log_pdf / d x
 *   jac_base     out, optional [n_params + dim][B][dim] (needs jac_x): derivative of the BASE point with respect to every
 *                parameter, then every coordinate -- with it the caller differentiates the SAMPLING direction implicitly
 *                (x = T(z): dx/dtheta = -(dbase/dx)^-1 dbase/dtheta), the reference's sample(allow_gradients=True)
 * "v" layers are differentiated in their closed-form direction only (natural_direction = 0), JF_ERR_UNSUPPORTED otherwise.
 */
int jf_subpdf_jacobian(const JfSubPdfDesc* desc, int dtype,
                       const void* x, int64_t ld_x,
                       const void* params, int64_t p_stride_param, int64_t p_stride_row,
                       void* jac_params, int64_t jac_stride_param, void* jac_x, int64_t ld_jx, void* jac_base,
                       int64_t B, int64_t* status, void* stream);

/*
 * params = W_L * tanh(... tanh(W_1 * concat(segments) + b_1) ...) + b_L for B rows.
 *   seg_ptrs[i]/seg_ld[i]  column block i: [B, seg_cols[i]] with leading dimension seg_ld[i].
 *   weights[l]             weight of Linear l in torch layout: [dims[l+1], dims[l]] row-major (= Linear.weight).
 *   biases[l]              [dims[l+1]].
 *   out                    element (j,row) at out[j*out_stride_param + row*out_stride_row].
 */
int jf_mlp_forward(const JfMlpDesc* desc, int dtype,
                   const void* const* seg_ptrs, const int64_t* seg_ld,
                   const void* const* weights, const void* const* biases,
                   void* out, int64_t out_stride_param, int64_t out_stride_row,
                   int64_t B, void* stream);

/* Same chain of Linear/tanh layers, optionally ACCUMULATING into `out` (out += result).  This is the building block of
 * the reference's AmortizableMLP connectivity modes (amortizable_mlp.py:508-682: a sum of sub-MLP outputs and a linear
 * highway, later sub-MLPs reading the running sum): the host composes the modes out of these calls. */
int jf_mlp_forward_acc(const JfMlpDesc* desc, int dtype,
                       const void* const* seg_ptrs, const int64_t* seg_ld,
                       const void* const* weights, const void* const* biases,
                       void* out, int64_t out_stride_param, int64_t out_stride_row,
                       int64_t B, int accumulate, void* stream);

/* Same, with caller-provided device workspace of jf_mlp_workspace_bytes(desc, dtype) bytes (0: none needed).  With a
 * workspace the fp64 one-hidden-layer (128) case runs on the tcgen05 tensor cores: the last-layer weights are split
 * once into int8 slices (kept in the workspace; `prepared` != 0 reuses the slices of the previous call, i.e. the same
 * weights) and the 128 x N contraction is evaluated exactly in int8 x int8 -> int32 on TMEM accumulators. */
int64_t jf_mlp_workspace_bytes(const JfMlpDesc* desc, int dtype);
int jf_mlp_forward_ws(const JfMlpDesc* desc, int dtype,
                      const void* const* seg_ptrs, const int64_t* seg_ld,
                      const void* const* weights, const void* const* biases,
                      void* out, int64_t out_stride_param, int64_t out_stride_row,
                      int64_t B, void* workspace, int64_t workspace_bytes, int prepared, void* stream);

/* Training: gradient of the parameter generator  params = W2 tanh(W1 inp + b1) + b2  (what the reference gets from
 * autograd through nn.Sequential(Linear, Tanh, Linear), main/default.py:654-670, called at :956) given the gradient with
 * respect to its PARAM-MAJOR output, i.e. what jf_subpdf_backward writes.  Eligible: fp32, one hidden layer of 128, at
 * most 96 inputs (jf_mlp_backward_workspace_bytes returns -1 otherwise, jf_mlp_backward JF_ERR_UNSUPPORTED).  The two
 * large products run as tcgen05 kind::tf32 MMAs (10-bit operand mantissas, fp32 accumulation: relative error of the
 * weight gradients ~1e-3 / sqrt(rows)); csrc/mlp_bwd.cuh.
 *   inp          [B, dims[0]] (ld_inp)         weights / biases as in jf_mlp_forward (biases[1] is not read)
 *   grad_out     element (j,row) at grad_out[j*go_stride_param + row]; go_stride_row must be 1, 16-byte aligned rows
 *   row_scale    [B] or NULL: the upstream gradient is grad_out[j, row] * row_scale[row] (grad_out = the per-row Jacobian of
 *                jf_subpdf_forward_backward, row_scale = d loss / d log_pdf); applied on the small operands, grad_out is
 *                read as it is
 *   grad_w1/b1/w2/b2   outputs in torch layout, ZEROED BY THE CALLER (partial sums are added with red.global)
 *   grad_inp     [B, dims[0]] (ld_ginp) or NULL
 *   workspace    jf_mlp_backward_workspace_bytes(desc, dtype, B) bytes, 256-byte aligned */
int64_t jf_mlp_backward_workspace_bytes(const JfMlpDesc* desc, int dtype, int64_t B);
int jf_mlp_backward(const JfMlpDesc* desc, int dtype, const void* inp, int64_t ld_inp,
                    const void* const* weights, const void* const* biases,
                    const void* grad_out, int64_t go_stride_param, int64_t go_stride_row, const void* row_scale,
                    void* grad_w1, void* grad_b1, void* grad_w2, void* grad_b2,
                    void* grad_inp, int64_t ld_ginp, int64_t B,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Parameter generator + layer chain of ONE conditional Euclidean sub-pdf in a single kernel (csrc/gf_fused.cuh): replaces
 * the hand-off main/default.py:956 (MLP call) -> :998-1029 (layer loop over `extra_inputs` slices), and for sampling
 * :1438 -> :1482-1506, without the [rows, n_params] block ever existing in HBM.  Eligible: fp64, manifold 'e' with
 * dim <= 4, default-option "g" layers with K = 10, generator Linear(<=16) -> tanh(128) -> Linear;
 * jf_subpdf_generated_workspace_bytes returns -1 otherwise (and jf_subpdf_apply_generated JF_ERR_UNSUPPORTED): callers
 * then chain jf_mlp_forward_ws and jf_subpdf_apply.  `prepared` != 0 reuses the weight slices a previous call with the
 * same weights and direction left in the workspace.  seg_ptrs / weights / biases as in jf_mlp_forward, the remaining
 * arguments as in jf_subpdf_apply. */
int64_t jf_subpdf_generated_workspace_bytes(const JfSubPdfDesc* desc, const JfMlpDesc* mlp, int dtype);
int jf_subpdf_apply_generated(const JfSubPdfDesc* desc, const JfMlpDesc* mlp, int dtype, int direction,
                              const void* const* seg_ptrs, const int64_t* seg_ld,
                              const void* const* weights, const void* const* biases,
                              const void* in, int64_t ld_in, const void* logdet_in, void* logdet_out,
                              const void* logbase_in, void* logbase_out, void* out, int64_t ld_out, int64_t B,
                              void* workspace, int64_t workspace_bytes, int prepared, int64_t* status, void* stream);

/* ---- whole-pdf entries -------------------------------------------------------------------------------------------- */
typedef struct JfPdfDesc {
    int32_t abi_version; /* JF_ABI_VERSION */
    int32_t dtype;       /* JF_F32 / JF_F64 */
    int32_t n_sub;
    int32_t cond_dim;    /* 0 = unconditional */
    int32_t total_target_dim;
    int32_t total_base_dim;
    int32_t reserved0;
    int32_t reserved1;
    JfSubPdfDesc sub[JF_MAX_SUBPDFS];
    int32_t target_col[JF_MAX_SUBPDFS]; /* first column of the sub-pdf in x (main/default.py:481-567) */
    int32_t base_col[JF_MAX_SUBPDFS];   /* first column in the base-space tensor */
    int32_t emb_dim[JF_MAX_SUBPDFS];    /* _embedding_conditional_return_num() */
    int32_t has_mlp[JF_MAX_SUBPDFS];    /* 1: parameters come from mlp[k]; 0: shared vector */
    JfMlpDesc mlp[JF_MAX_SUBPDFS];      /* n_segments/seg_cols are filled in by the library */
} JfPdfDesc;

typedef struct JfPdfParams {
    const void* shared[JF_MAX_SUBPDFS];                       /* raw shared vector [n_params] or NULL */
    const void* weights[JF_MAX_SUBPDFS][JF_MAX_MLP_LINEAR]; /* Linear.weight [out,in] (device) */
    const void* biases[JF_MAX_SUBPDFS][JF_MAX_MLP_LINEAR];
} JfPdfParams;

/* bytes of device workspace needed to process `chunk_rows` rows at a time */
int64_t jf_pdf_workspace_bytes(const JfPdfDesc* desc, int64_t chunk_rows);

/* log_pdf of B rows, all buffers on the device.  x [B, ldx], cond [B, ldc] or NULL.
 * Outputs (any may be NULL): logp [B], logp_base [B], base [B, ld_base]. */
int jf_pdf_logpdf(const JfPdfDesc* desc, const JfPdfParams* params,
                  const void* x, int64_t ldx, const void* cond, int64_t ldc,
                  void* logp, void* logp_base, void* base, int64_t ld_base,
                  int64_t B, void* workspace, int64_t workspace_bytes, int64_t chunk_rows,
                  int64_t* status, void* stream);

/* sampling: z [B, ldz] base normals -> x [B, ldx]; logp = log N(z) - logdet, logp_base = log N(z). */
int jf_pdf_sample(const JfPdfDesc* desc, const JfPdfParams* params,
                  const void* z, int64_t ldz, const void* cond, int64_t ldc,
                  void* x, int64_t ldx, void* logp, void* logp_base,
                  int64_t B, void* workspace, int64_t workspace_bytes, int64_t chunk_rows,
                  int64_t* status, void* stream);

/* Same two calls with HOST buffers (pinned for full speed).  The library pipelines H2D copy, kernels and D2H copy
 * over two internal streams in chunks of `chunk_rows`; `workspace` is device memory of
 * jf_pdf_host_workspace_bytes(desc, chunk_rows) bytes.  Blocks until the results are in host memory. */
int64_t jf_pdf_host_workspace_bytes(const JfPdfDesc* desc, int64_t chunk_rows);
int jf_pdf_logpdf_host(const JfPdfDesc* desc, const JfPdfParams* params,
                       const void* x_host, int64_t ldx, const void* cond_host, int64_t ldc,
                       void* logp_host, void* logp_base_host, void* base_host, int64_t ld_base,
                       int64_t B, void* workspace, int64_t workspace_bytes, int64_t chunk_rows,
                       int64_t* status);
int jf_pdf_sample_host(const JfPdfDesc* desc, const JfPdfParams* params,
                       const void* z_host, int64_t ldz, const void* cond_host, int64_t ldc,
                       void* x_host, int64_t ldx, void* logp_host, void* logp_base_host,
                       int64_t B, void* workspace, int64_t workspace_bytes, int64_t chunk_rows,
                       int64_t* status);

/* Target-space charts of all sub-pdfs in one pass: what the reference's `force_embedding_coordinates` does before the
 * log_pdf chain and after the sampling chain (pdf.transform_target_space, main/default.py:1737-1813 with
 * layers/spheres/sphere_base.py:242-332; Euclidean and interval sub-pdfs are copied).
 *   to_embedding != 0: in [B, total intrinsic dim] (s2: theta,phi; s1: angle) -> out [B, total embedded dim]
 *                      (s2: x,y,z; s1: cos,sin); logdet_out = logdet_in + sum over s2 sub-pdfs of log sin(theta)
 *   to_embedding == 0: the inverse charts; logdet_out = logdet_in - sum log sin(theta)
 * logdet_in may be NULL (= 0), logdet_out may be NULL. */
int jf_pdf_transform_target(const JfPdfDesc* desc, int to_embedding,
                            const void* in, int64_t ld_in, void* out, int64_t ld_out,
                            const void* logdet_in, void* logdet_out, int64_t B, void* stream);

/* out[r] = log(mean_c exp(in[r*cols + c])) for r < rows: the S x S cross-evaluation reduce of the marginal entropies
 * (pdf.entropy with sub_manifolds, main/default.py:2444-2448).  Device buffers. */
int jf_row_logmeanexp(int dtype, const void* in, int64_t rows, int64_t cols, void* out, void* stream);

/* Base-space standard normals on the device, out [B, ld_out] (first `dim` columns): Philox4x32-10 keyed by `seed`, row i
 * is a function of (seed, first_row + i) only.  Replaces the reference's host numpy RNG + H2D copy in pdf.sample
 * (main/default.py:1661-1668); ranks of a sharded job pass their first global row and draw slices of one stream. */
int jf_normal_rows(int dtype, uint64_t seed, uint64_t first_row, int64_t B, int32_t dim, void* out, int64_t ld_out,
                   void* stream);

/* One Linear layer whose weights differ from row to row -- the "being amortised" mode of the reference's AmortizableMLP
 * (amortizable_mlp.py:470-585: `_adaptive_matmul`, `_apply_amortized_mlp` with use_permanent_parameters=False), run for
 * every sub-pdf by pdf(..., amortize_everything=True) / fully_amortized_pdf (main/fully_amortized.py:22-278):
 *   out[r, o] (+)= act( sum_i W_r[o, i] * in[r, i] + b_r[o] )
 *   W_r[o, i] = params[r*ld_params + off_w + o*n_in + i],   b_r[o] = params[r*ld_params + off_b + o]  (off_b < 0: no bias)
 *   act: 0 = identity, 1 = tanh;  accumulate != 0: out += ...;  out element (o, r) at out[o*out_stride_param + r*out_stride_row].
 * A rank-k factorised layer U (V^T x) is two calls (n_out = k, then n_in = k).  HBM-bound: every weight is read once. */
int jf_rowwise_linear(int dtype, const void* params, int64_t ld_params, int64_t off_w, int64_t off_b,
                      const void* in, int64_t ld_in, int32_t n_in, int32_t n_out, int act, int accumulate,
                      void* out, int64_t out_stride_param, int64_t out_stride_row, int64_t B, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------------------------------------- */
int jf_abi_version(void);
/* FP64 DFMA / FP32 FFMA peak probe used as the roofline denominator of the compute-bound layer kernels:
 * runs `iters` dependent FMA chains on every SM and returns the elapsed milliseconds (CUDA events) in *ms and the
 * FMA count in *fma_count. */
int jf_probe_fma_peak(int dtype, int iters, float* ms, double* fma_count, void* scratch /* >= 256 bytes, device */, void* stream);
/* number of kernel launches issued by this library since process start (for bench.py's gpu_launches) */
int64_t jf_launch_count(void);
/* sizeof() of the ABI structs as compiled into the library (0: JfLayerDesc, 1: JfSubPdfDesc, 2: JfMlpDesc,
 * 3: JfPdfDesc, 4: JfPdfParams, 5: JfSplineDesc); bindings assert these against their own mirrors. */
int64_t jf_struct_size(int which);

#ifdef __cplusplus
}
#endif
#endif /* JAMMY_B200_H */
