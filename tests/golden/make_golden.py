"""Generate golden input/output vectors by running the UNMODIFIED reference (thoglu/jammy_flows at /root/reference).

Run in the build container only:   python tests/golden/make_golden.py [case_name ...]

The reference has no stored numeric fixtures of its own (its tests are self-consistency tests,
reference: tests/test_general.py:393-556), so absolute parity is pinned by executing the reference here on seeded
inputs and committing inputs, parameters (state_dict) and outputs as small .npz files.  Every `-m "not gpu"` oracle test
and every `-m gpu` CUDA parity test reads these files; nothing reads /root/reference at test time.

Per case the file holds
  meta                json: pdf_defs, flow_defs, options_overwrite, conditional_input_dim, dtype, param-set description
  param/<name>        reference state_dict entries (names are the de-facto checkpoint contract)
  x, cond             evaluation inputs            -> logp, logp_base, base   (reference pdf.forward, main/default.py:1059)
  z                   base-space normals           -> samp_x, samp_logp, samp_logp_base
                      (reference pdf._obtain_sample(predefined_target_input=z), main/default.py:1533)
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from refshim import import_reference  # noqa: E402


def s2_uniform(n, gen):
    u = torch.rand(n, generator=gen, dtype=torch.float64)
    theta = torch.acos(1.0 - 2.0 * u)
    phi = torch.rand(n, generator=gen, dtype=torch.float64) * 2 * np.pi
    return torch.stack([theta, phi], dim=1)


def make_inputs(pdf_defs, n, gen, tails=False):
    cols = []
    for sub in pdf_defs.split("+"):
        kind, dim = sub[0], int(sub.split("_")[0][1:])
        if kind == "e":
            v = 1.5 * torch.randn(n, dim, generator=gen, dtype=torch.float64)
            if tails:
                # a block of far-tail points +-(5..30): exercises the Pade tails / underflow handling
                m = min(n // 10, 200)
                mag = 5.0 + 25.0 * torch.rand(m, dim, generator=gen, dtype=torch.float64)
                sgn = torch.where(torch.rand(m, dim, generator=gen) < 0.5, -1.0, 1.0).double()
                v[:m] = mag * sgn
            cols.append(v)
        elif kind == "s" and dim == 2:
            cols.append(s2_uniform(n, gen))
        elif kind == "s" and dim == 1:
            cols.append(torch.rand(n, 1, generator=gen, dtype=torch.float64) * 2 * np.pi)
        elif kind == "i":
            parts = sub.split("_")
            lo, hi = (0.0, 1.0) if len(parts) == 1 else (float(parts[1]), float(parts[2]))
            cols.append(lo + (hi - lo) * torch.rand(n, dim, generator=gen, dtype=torch.float64))
        else:
            raise ValueError(sub)
    return torch.cat(cols, dim=1)


CFG3_F = {"add_vertical_rq_spline_flow": 1, "spline_num_basis_functions": -1, "vertical_smooth": 1,
          "vertical_flow_defs": "rr", "circular_flow_defs": "oo", "vertical_fix_boundary_derivative": 1,
          "add_circular_rq_spline_flow": 1, "circular_add_rotation": 0, "vertical_fix_first_width_n_height_to_zero": 1,
          "vertical_also_fix_second_width_to_zero": 1, "vertical_independent_width_height_parametrization": 1}

CASES = {
    # name: dict(pdf_defs, flow_defs, opts, cond_dim, dtype, n, perturb, tails)
    # BASELINE.json configs[0]: e2 "gg" unconditional fp64
    "cfg1_e2_gg": dict(pdf_defs="e2", flow_defs="gg", n=2000, tails=True),
    "cfg1_e2_gg_perturbed": dict(pdf_defs="e2", flow_defs="gg", n=2000, tails=True, perturb=0.3),
    # BASELINE.json configs[1]: README e4+s2+e4; "n" does not exist in this snapshot -> "f" defaults
    "cfg2_e4s2e4_f": dict(pdf_defs="e4+s2+e4", flow_defs="gggg+f+gggg", n=1000),
    "cfg2_e4s2e4_f_perturbed": dict(pdf_defs="e4+s2+e4", flow_defs="gggg+f+gggg", n=1000, perturb=0.3),
    # g-layer unit cases: every inverse-CDF variant, d=1 (Q=-1 quirk), d=3, conditional (per-row params)
    "g_e1_isigmoid": dict(pdf_defs="e1", flow_defs="g", n=500, tails=True, perturb=0.3),
    "g_e3_ggg_cond": dict(pdf_defs="e3", flow_defs="ggg", n=500, cond_dim=3, perturb=0.2),
    "g_e2_partly_precise_all": dict(pdf_defs="e2", flow_defs="gg", n=500, tails=True, perturb=0.3,
                                    opts={"g": {"inverse_function_type": "inormal_partly_precise"}}),
    "g_e2_full_pade": dict(pdf_defs="e2", flow_defs="gg", n=500, tails=True, perturb=0.3,
                           opts={"g": {"inverse_function_type": "inormal_full_pade"}}),
    "g_e2_partly_crude": dict(pdf_defs="e2", flow_defs="gg", n=500, tails=True, perturb=0.3,
                              opts={"g": {"inverse_function_type": "inormal_partly_crude"}}),
    "g_e2_nonorm": dict(pdf_defs="e2", flow_defs="gg", n=500, perturb=0.3,
                        opts={"g": {"fit_normalization": 0}}),
    "g_e5_k7_cond_f32": dict(pdf_defs="e5", flow_defs="gg", n=500, cond_dim=2, perturb=0.03, dtype="float32",
                             opts={"g": {"num_kde": 7}}),
    # "f": rotation modes of sphere_base.py:79-91 and the kappa predictions of fvm_2d.py:108-138 -- the reference's own
    # sweep (tests/test_general.py:123-172) plus the two remaining link functions with kappa_clamping
    "f_rot_angles": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.5, opts={"f": {"rotation_mode": "angles"}}),
    "f_rot_xyz_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3, opts={"f": {"rotation_mode": "xyz"}}),
    "f_rot_xyz_mu": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.5,
                         opts={"f": {"rotation_mode": "xyz", "kappa_prediction": "mu"}}),
    "f_rot_xyz_mu_squared_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3,
                                      opts={"f": {"rotation_mode": "xyz", "kappa_prediction": "mu_squared"}}),
    "f_rot_quat": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.5, opts={"f": {"rotation_mode": "quaternion"}}),
    "f_rot_quat_quatvec_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3,
                                    opts={"f": {"rotation_mode": "quaternion", "kappa_prediction": "quatvec"}}),
    "f_rot_quat_quatvec_squared": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.5,
                                       opts={"f": {"rotation_mode": "quaternion", "kappa_prediction": "quatvec_squared"}}),
    "f_kappa_softplus_clamp": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.5,
                                   opts={"f": {"kappa_prediction": "softplus_real_bounded", "kappa_clamping": 1}}),
    "f_kappa_log_bounded_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3,
                                     opts={"f": {"kappa_prediction": "log_bounded", "min_kappa": 1e-3}}),
    "f_extra_rotation": dict(pdf_defs="s2", flow_defs="f", n=500, perturb=0.3,
                             opts={"f": {"add_extra_rotation_inbetween": 1, "add_vertical_rq_spline_flow": 1,
                                         "add_circular_rq_spline_flow": 1, "circular_add_rotation": 0}}),
    "f_identity_region_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3,
                                   opts={"f": {"boundary_cos_theta_identity_region": 0.1, "add_vertical_rq_spline_flow": 1,
                                               "add_circular_rq_spline_flow": 1, "circular_add_rotation": 0}}),
    "s2_f_uncond": dict(pdf_defs="s2", flow_defs="f", n=1000, perturb=0.5),
    "s2_f_cond": dict(pdf_defs="s2", flow_defs="f", n=1000, cond_dim=2, perturb=0.3),
    # BASELINE.json configs[2]: s2 "f" with smooth vMF-scaled spline sub-flows + i1 "r" (docs/suggested_settings.rst:52-73)
    "cfg3_s2i1_fr": dict(pdf_defs="s2+i1", flow_defs="f+r", n=1000, opts={"f": CFG3_F}),
    "cfg3_s2i1_fr_perturbed": dict(pdf_defs="s2+i1", flow_defs="f+r", n=1000, perturb=0.3, opts={"f": CFG3_F}),
    # spline unit cases: interval "r" (plain / fixed boundary derivs / smooth 2 and 3 bins / restricted ratio), custom
    # interval boundaries, conditional (per-row parameters)
    "r_i1_rr_cond": dict(pdf_defs="i1_-0.5_0.8", flow_defs="rr", n=500, cond_dim=2, perturb=0.5),
    "r_i1_fixed_bd": dict(pdf_defs="i1", flow_defs="rr", n=500, perturb=0.5,
                          opts={"r": {"fix_boundary_derivatives": 1.0, "num_basis_functions": 8}}),
    "r_i1_smooth23": dict(pdf_defs="i1", flow_defs="rr", n=500, perturb=0.5,
                          opts={(0, 0): {"r": {"smooth_second_derivative": 1, "num_basis_functions": 2}},
                                (0, 1): {"r": {"smooth_second_derivative": 1, "num_basis_functions": 3,
                                               "fix_boundary_derivatives": 1.0}}}),
    "r_i1_ratio": dict(pdf_defs="i1", flow_defs="r", n=500, perturb=0.5,
                       opts={"r": {"restrict_max_min_width_height_ratio": 50.0, "fix_first_width_n_height_to_zero": 1,
                                   "independent_width_height_parametrization": 1}}),
    # S1: circular splines (default smooth + rotation; plain periodic; plain fixed; reversed direction) and Moebius
    "o_s1_default": dict(pdf_defs="s1", flow_defs="oo", n=500, perturb=0.5),
    "o_s1_plain_cond": dict(pdf_defs="s1", flow_defs="oo", n=500, cond_dim=2, perturb=0.3,
                            opts={(0, 0): {"o": {"smooth_second_derivative": 0, "num_basis_functions": 5}},
                                  (0, 1): {"o": {"smooth_second_derivative": 0, "num_basis_functions": 4,
                                                 "fix_boundary_derivatives": 1.0, "natural_direction": 0,
                                                 "add_rotation": 0}}}),
    "m_s1_default": dict(pdf_defs="s1", flow_defs="mm", n=500, perturb=0.3),
    "m_s1_natural_cond": dict(pdf_defs="s1", flow_defs="m", n=500, cond_dim=2, perturb=0.3,
                              opts={"m": {"natural_direction": 1, "add_rotation": 1}}),
    # f with plain (non-smooth) 5-bin sub-flows, conditional
    "s2_f_plain_subflows_cond": dict(pdf_defs="s2", flow_defs="f", n=500, cond_dim=2, perturb=0.3,
                                     opts={"f": {"add_vertical_rq_spline_flow": 1, "add_circular_rq_spline_flow": 1,
                                                 "vertical_flow_defs": "r", "circular_flow_defs": "o"}}),
    # "v": exponential-map flow (fp64 only in the reference), unconditional and as the conditional tail of cfg4
    "s2_v_uncond": dict(pdf_defs="s2", flow_defs="v", n=500, perturb=0.0),
    "s2_vv_rot": dict(pdf_defs="s2", flow_defs="vv", n=300, perturb=0.0, opts={"v": {"add_rotation": 1, "num_components": 4}}),
    "s2_v_natural": dict(pdf_defs="s2", flow_defs="v", n=300, perturb=0.0, opts={"v": {"natural_direction": 1}}),
    "cfg4_e6s2_gv_small": dict(pdf_defs="e6+s2", flow_defs="gggggg+v", n=300, cond_dim=64, perturb=0.02),
    "s2_v_splines": dict(pdf_defs="s2", flow_defs="v", n=300, perturb=0.3, opts={"v": {"exp_map_type": "splines"}}),
    "s2_v_splines_natural_cond": dict(pdf_defs="e2+s2", flow_defs="gg+v", n=200, cond_dim=2, perturb=0.1,
                              opts={"v": {"exp_map_type": "splines", "natural_direction": 1, "num_components": 4}}),
    "s2_v_linear": dict(pdf_defs="s2", flow_defs="v", n=300, perturb=0.0, opts={"v": {"exp_map_type": "linear"}}),
    "s2_v_quadratic_natural": dict(pdf_defs="s2", flow_defs="vv", n=300, perturb=0.0,
                                   opts={"v": {"exp_map_type": "quadratic", "natural_direction": 1}}),
    "s2_v_quadratic_cond": dict(pdf_defs="e2+s2", flow_defs="gg+v", n=300, cond_dim=2, perturb=0.0,
                                opts={"v": {"exp_map_type": "quadratic", "num_components": 6}}),
    # "t": affine layer (the docs' recommended Euclidean recipe is "g...gt" with cov_type="full", suggested_settings.rst:14-41)
    "t_e3_ggt_full": dict(pdf_defs="e3", flow_defs="ggt", n=500, tails=True, perturb=0.3, opts={"t": {"cov_type": "full"}}),
    "t_e4_gt_cond_diag": dict(pdf_defs="e4", flow_defs="gt", n=500, cond_dim=2, perturb=0.2),
    "t_e2e3_sym_identity": dict(pdf_defs="e2+e3", flow_defs="t+gt", n=500, perturb=0.3,
                                opts={0: {"t": {"cov_type": "diagonal_symmetric"}}, 1: {"t": {"cov_type": "identity"}}}),
    # non-default "g" options (the reference's own option sweep: tests/test_general.py:307-320)
    "g_e3_angles_cond": dict(pdf_defs="e3", flow_defs="gg", n=400, cond_dim=2, perturb=0.2, tails=True,
                             opts={"g": {"rotation_mode": "angles"}}),
    "g_e4_angles": dict(pdf_defs="e4", flow_defs="gg", n=400, perturb=0.3, opts={"g": {"rotation_mode": "angles"}}),
    "g_e2_cayley": dict(pdf_defs="e2", flow_defs="gg", n=400, perturb=0.3, tails=True,
                        opts={"g": {"rotation_mode": "cayley"}}),
    "g_e2_cayley_cond": dict(pdf_defs="e2", flow_defs="g", n=300, cond_dim=2, perturb=0.2,
                             opts={"g": {"rotation_mode": "cayley"}}),
    "g_e3_tri_cond": dict(pdf_defs="e3", flow_defs="gg", n=400, cond_dim=2, perturb=0.2,
                          opts={"g": {"rotation_mode": "triangular_combination"}}),
    "g_e4_tri": dict(pdf_defs="e4", flow_defs="gg", n=400, perturb=0.3, tails=True,
                     opts={"g": {"rotation_mode": "triangular_combination"}}),
    "g_e2_softplus_width": dict(pdf_defs="e2", flow_defs="gg", n=400, perturb=0.3, tails=True,
                                opts={"g": {"softplus_for_width": 1}}),
    "g_e2_clamp_widths": dict(pdf_defs="e2", flow_defs="gg", n=400, perturb=0.3, opts={"g": {"clamp_widths": 1}}),
    "g_e2_unbounded_clamp_cond": dict(pdf_defs="e2", flow_defs="gg", n=400, cond_dim=2, perturb=0.2,
                                      opts={"g": {"upper_bound_for_widths": -1, "clamp_widths": 1,
                                                  "width_smooth_saturation": 0}}),
    "g_e2_softplus_clamp": dict(pdf_defs="e2", flow_defs="g", n=300, perturb=0.3,
                                opts={"g": {"softplus_for_width": 1, "clamp_widths": 1}}),
    "g_e2_skew": dict(pdf_defs="e2", flow_defs="gg", n=400, perturb=0.3, tails=True, opts={"g": {"add_skewness": 1}}),
    "g_e3_skew_cond": dict(pdf_defs="e3", flow_defs="gg", n=400, cond_dim=2, perturb=0.2,
                           opts={"g": {"add_skewness": 1}}),
    "g_e2_rqs": dict(pdf_defs="e2", flow_defs="gg", n=400, perturb=0.3, tails=True,
                     opts={"g": {"nonlinear_stretch_type": "rq_splines"}}),
    "g_e3_rqs_cond": dict(pdf_defs="e3", flow_defs="gg", n=400, cond_dim=2, perturb=0.2,
                          opts={"g": {"nonlinear_stretch_type": "rq_splines"}}),
    "g_e2_center_mean": dict(pdf_defs="e2", flow_defs="gg", n=300, perturb=0.3, opts={"g": {"center_mean": 1}}),
    "g_e2_center_mean_cond": dict(pdf_defs="e2", flow_defs="gg", n=300, cond_dim=2, perturb=0.2,
                                  opts={"g": {"center_mean": 1}}),
    # force_embedding_coordinates (what pdf.entropy uses by default): charts before / after the chain
    "emb_e2s2e2_cond": dict(pdf_defs="e2+s2+e2", flow_defs="gg+f+gg", n=300, cond_dim=2, perturb=0.3, emb=True),
    "emb_s1s2i1": dict(pdf_defs="s1+s2+i1", flow_defs="m+v+r", n=300, perturb=0.0, emb=True),
    # data-driven initialisation pdf.init_params(data=x) (extra_functions.py:179-409): state_dict AFTER the init
    "init_e3_ggt_data": dict(pdf_defs="e3", flow_defs="ggt", n=400, data_init=True, opts={"t": {"cov_type": "full"}}),
    "init_e2e2_cond_data": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=400, cond_dim=2, data_init=True),
    # AmortizableMLP parameter generators (amortization_mlp_use_custom_mode, amortizable_mlp.py): low-rank factors and
    # the five connectivity modes; dims/ranks as in the reference's own sweep (tests/test_general.py:302-304)
    "amlp_e2e2_cond_mode0": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=300, cond_dim=3, perturb=0.05,
                                 pdf_kw=dict(amortization_mlp_use_custom_mode=True, amortization_mlp_dims="64-30",
                                             amortization_mlp_ranks="2-10-1000")),
    "amlp_e2s2_mode1": dict(pdf_defs="e2+s2", flow_defs="gg+f", n=300, perturb=0.05,
                            pdf_kw=dict(amortization_mlp_use_custom_mode=True, amortization_mlp_dims="64-30",
                                        amortization_mlp_ranks="2-10-1000-3", amortization_mlp_highway_mode=1)),
    "amlp_e2e2_cond_mode2": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=300, cond_dim=3, perturb=0.05,
                                 pdf_kw=dict(amortization_mlp_use_custom_mode=True, amortization_mlp_dims="64-30",
                                             amortization_mlp_ranks="2-10-5-0-4", amortization_mlp_highway_mode=2)),
    "amlp_e2e2_cond_mode3": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=300, cond_dim=3, perturb=0.05,
                                 pdf_kw=dict(amortization_mlp_use_custom_mode=True, amortization_mlp_dims="64-30",
                                             amortization_mlp_ranks="2-10-5-0-4", amortization_mlp_highway_mode=3)),
    "amlp_e2e2_cond_mode4": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=300, cond_dim=3, perturb=0.05,
                                 pdf_kw=dict(amortization_mlp_use_custom_mode=True, amortization_mlp_dims="64-30",
                                             amortization_mlp_ranks=0, amortization_mlp_highway_mode=4)),
    # only_last=True (main/default.py:1018-1024, :1490-1502): only the LAST layer of every sub-pdf is applied, sphere
    # layers with their base chart forced on (fix_euclidean_to_sphere_first)
    "last_e3s2e2": dict(pdf_defs="e3+s2+e2", flow_defs="ggg+vf+gg", n=300, perturb=0.3, only_last=True),
    "last_e2s1s2_cond": dict(pdf_defs="e2+s1+s2", flow_defs="gg+mo+fv", n=300, cond_dim=3, perturb=0.2, only_last=True),
    # Poisson log-mean prediction (main/default.py:466-477, :624-626, :832-877): joint with the flow parameters (last
    # output of the generator) for a conditional pdf, a free parameter for an unconditional one
    "poisson_e2_joint_cond": dict(pdf_defs="e2", flow_defs="gg", n=300, cond_dim=3, perturb=0.2, poisson=True,
                                  pdf_kw=dict(predict_log_normalization=True, join_poisson_and_pdf_description=True)),
    "poisson_s2_joint_cond": dict(pdf_defs="s2", flow_defs="f", n=300, cond_dim=2, perturb=0.2, poisson=True,
                                  pdf_kw=dict(predict_log_normalization=True, join_poisson_and_pdf_description=True)),
    "poisson_e2_uncond": dict(pdf_defs="e2", flow_defs="gg", n=300, perturb=0.2, poisson=True,
                              pdf_kw=dict(predict_log_normalization=True)),
    # one conditional input per sub-pdf (conditional_input_dim as a list, main/default.py:286-296, :944-949)
    "condlist_e2s2e1": dict(pdf_defs="e2+s2+e1", flow_defs="gg+f+g", n=300, cond_dim=[3, 2, 4], perturb=0.2),
    # training (BASELINE.json configs[4] structure at fixture size): gradients of mean(log_pdf) w.r.t. every MLP tensor
    "train_e3_ggg_cond": dict(pdf_defs="e3", flow_defs="ggg", n=200, cond_dim=3, perturb=0.2, grads=True),
    "train_e10_gg_cond": dict(pdf_defs="e10", flow_defs="gg", n=100, cond_dim=4, perturb=0.05, grads=True),
    "train_e2_gg_uncond": dict(pdf_defs="e2", flow_defs="gg", n=300, perturb=0.3, grads=True),
    "train_e3e2_uncond": dict(pdf_defs="e3+e2", flow_defs="gg+gg", n=200, perturb=0.2, grads=True),
    "train_e2e2_cond": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=200, cond_dim=2, perturb=0.2, grads=True),
    # gradients through SAMPLES (sample(allow_gradients=True), main/default.py:1342): loss = sum(x * w) + 0.3 sum(log_pdf)
    "strain_e3_ggg_cond": dict(pdf_defs="e3", flow_defs="ggg", n=100, cond_dim=3, perturb=0.2, sgrads=True),
    "strain_e2e2_uncond": dict(pdf_defs="e2+e2", flow_defs="gg+gg", n=100, perturb=0.2, sgrads=True),
    "strain_e2s2e2_f": dict(pdf_defs="e2+s2+e2", flow_defs="gg+f+gg", n=100, perturb=0.1, sgrads=True),
    "strain_s1i1_or_cond": dict(pdf_defs="s1+i1_-0.5_0.8", flow_defs="o+r", n=100, cond_dim=2, perturb=0.1, sgrads=True),
    "strain_e3_ggt_cond": dict(pdf_defs="e3", flow_defs="ggt", n=100, cond_dim=2, perturb=0.2, sgrads=True,
                               opts={"t": {"cov_type": "full"}}),
    # the docs' recommended Euclidean recipe "g...gt" (suggested_settings.rst:14-41) in the training path
    "train_e3_ggt_full_cond": dict(pdf_defs="e3", flow_defs="ggt", n=150, cond_dim=2, perturb=0.2, grads=True,
                                   opts={"t": {"cov_type": "full"}}),
    "train_e2e3_t_variants": dict(pdf_defs="e2+e3", flow_defs="tg+gt", n=150, perturb=0.2, grads=True,
                                  opts={"t": {"cov_type": "diagonal"}}),
    "train_e2_gt_symmetric": dict(pdf_defs="e2", flow_defs="gt", n=150, cond_dim=2, perturb=0.2, grads=True,
                                  opts={"t": {"cov_type": "diagonal_symmetric"}}),
    "train_e3_gg_variants_cond": dict(pdf_defs="e3+e2", flow_defs="gg+gg", n=150, cond_dim=2, perturb=0.2, grads=True,
                                      opts={"g": {"rotation_mode": "none", "fit_normalization": 0, "num_kde": 6,
                                                  "inverse_function_type": "isigmoid"}}),
    "train_e2_gg_rawnorm": dict(pdf_defs="e2", flow_defs="gg", n=150, perturb=0.2, grads=True,
                                opts={"g": {"regulate_normalization": 0}}),
    # non-Euclidean sub-pdfs in the training path (README-style mixed flow, every manifold layer kind)
    "train_e2s2e2_f": dict(pdf_defs="e2+s2+e2", flow_defs="gg+f+gg", n=120, perturb=0.1, grads=True),
    "train_s2_f_splines_cond": dict(pdf_defs="s2", flow_defs="f", n=120, cond_dim=2, perturb=0.1, grads=True,
                                    opts={"f": CFG3_F}),
    "train_s2_v_cond": dict(pdf_defs="s2", flow_defs="v", n=120, cond_dim=2, perturb=0.1, grads=True),
    "train_s1i1_or_cond": dict(pdf_defs="s1+i1_-0.5_0.8", flow_defs="o+r", n=120, cond_dim=2, perturb=0.1, grads=True),
    "train_s1_m_uncond": dict(pdf_defs="s1", flow_defs="m", n=120, perturb=0.1, grads=True),
}


def build_case(jf, name, spec):
    seed = 1
    torch.manual_seed(seed)
    np.random.seed(seed)
    dtype = getattr(torch, spec.get("dtype", "float64"))
    opts = spec.get("opts", {})
    cond_dim = spec.get("cond_dim", None)
    pdf = jf.pdf(spec["pdf_defs"], spec["flow_defs"], options_overwrite=opts, conditional_input_dim=cond_dim,
                 **spec.get("pdf_kw", {}))
    pdf = pdf.to(dtype)
    gen = torch.Generator().manual_seed(1234)
    if spec.get("perturb", 0.0) > 0:
        # leave the near-identity init regime (default MLP weights are /1000: main/default.py:1924)
        with torch.no_grad():
            for _, p in pdf.named_parameters():
                p.add_(spec["perturb"] * torch.randn(p.shape, generator=gen, dtype=torch.float64).to(p.dtype))
    n = spec["n"]
    x = make_inputs(spec["pdf_defs"], n, gen, tails=spec.get("tails", False)).to(dtype)
    cond = None
    if type(cond_dim) == list:
        cond = [torch.randn(n, cd, generator=gen, dtype=torch.float64).to(dtype) for cd in cond_dim]
    elif cond_dim is not None:
        cond = torch.randn(n, cond_dim, generator=gen, dtype=torch.float64).to(dtype)
    z = torch.randn(n, pdf.total_base_dim, generator=gen, dtype=torch.float64).to(dtype)
    if spec.get("data_init", False):
        torch.manual_seed(5)
        np.random.seed(5)
        # correlated, shifted data so that the PCA / percentile / covariance fits have something to find
        mix = torch.randn(x.shape[1], x.shape[1], generator=gen, dtype=torch.float64)
        x = (x @ mix + 0.7).to(dtype)
        pdf.init_params(data=x)
    with torch.no_grad():
        logp, logp_base, base = pdf(x, conditional_input=cond)
        samp_x, _, samp_logp, samp_logp_base = pdf._obtain_sample(conditional_input=cond, predefined_target_input=z)
        # the reference's own round trip (sample -> forward) as the yardstick for sampling parity
        rt_logp, _, rt_base = pdf(samp_x, conditional_input=cond)
    out = {
        "meta": json.dumps(dict(name=name, pdf_defs=spec["pdf_defs"], flow_defs=spec["flow_defs"],
                                options_overwrite={str(k): v for k, v in opts.items()},
                                conditional_input_dim=cond_dim, dtype=spec.get("dtype", "float64"),
                                pdf_kw=spec.get("pdf_kw", {}),
                                perturb=spec.get("perturb", 0.0), seed=seed,
                                reference="thoglu/jammy_flows v1.1.0 @ /root/reference, torch %s CPU" % torch.__version__)),
        "x": x.numpy(), "z": z.numpy(),
        "logp": logp.numpy(), "logp_base": logp_base.numpy(), "base": base.numpy(),
        "samp_x": samp_x.numpy(), "samp_logp": samp_logp.numpy(), "samp_logp_base": samp_logp_base.numpy(),
        "ref_roundtrip_base_err": np.nanmax(np.abs((rt_base - z).numpy())),
        "ref_roundtrip_logp_err": np.nanmax(np.abs((rt_logp - samp_logp).numpy())),
    }
    if spec.get("poisson", False):
        with torch.no_grad():
            out["log_lambda"] = pdf.log_mean_poisson(conditional_input=cond).numpy()
    if spec.get("only_last", False):
        with torch.no_grad():
            l_logp, l_logp_base, l_base = pdf(x, conditional_input=cond, only_last=True)
            l_sx, _, l_slogp, _ = pdf._obtain_sample(conditional_input=cond, predefined_target_input=z, only_last=True)
        out.update({"last_logp": l_logp.numpy(), "last_logp_base": l_logp_base.numpy(), "last_base": l_base.numpy(),
                    "last_samp_x": l_sx.numpy(), "last_samp_logp": l_slogp.numpy()})
    if spec.get("emb", False):
        with torch.no_grad():
            x_emb, _ = pdf.transform_target_space(x, 0.0, transform_from="default", transform_to="embedding")
            logp_e, _, base_e = pdf(x_emb, conditional_input=cond, force_embedding_coordinates=True)
            sx_e, _, slogp_e, _ = pdf._obtain_sample(conditional_input=cond, predefined_target_input=z,
                                                     force_embedding_coordinates=True)
        # entropies incl. marginal ones (main/default.py:2263-2454); the base normals are the global-RNG draw of
        # all_layer_forward_individual_subdims_incl_sampling (:2914), reproduced here from the same seed
        S, nb = 12, (3 if cond is not None else 1)
        subs = [-1] + list(range(len(spec["pdf_defs"].split("+"))))
        c_small = cond[:nb] if cond is not None else None
        for flag, tag in ((True, "emb"), (False, "intr")):
            torch.manual_seed(77)
            with torch.no_grad():
                ent = pdf.entropy(sub_manifolds=subs, conditional_input=c_small, samplesize=S,
                                  force_embedding_coordinates=flag)
            for k_, v_ in ent.items():
                out["ent_%s_%s" % (tag, k_)] = v_.numpy()
        torch.manual_seed(77)
        out["ent_z"] = torch.randn(size=(S * nb, pdf.total_base_dim), dtype=dtype).numpy()
        out["ent_S"] = np.int64(S)
        out.update({"x_emb": x_emb.numpy(), "logp_emb": logp_e.numpy(), "base_emb": base_e.numpy(),
                    "samp_x_emb": sx_e.numpy(), "samp_logp_emb": slogp_e.numpy()})
    if type(cond) == list:
        for i_, c_ in enumerate(cond):
            out["cond%d" % i_] = c_.numpy()
    elif cond is not None:
        out["cond"] = cond.numpy()
    for k, v in pdf.state_dict().items():
        out["param/" + k] = v.numpy()
    if spec.get("sgrads", False):
        # reference autograd through the sampling direction (bisection + Newton iterations unrolled by autograd)
        pdf.zero_grad()
        w = torch.linspace(-1.0, 1.5, x.shape[1], dtype=dtype).unsqueeze(0) * torch.linspace(0.5, 1.5, n, dtype=dtype).unsqueeze(1)
        sx, _, slp, _ = pdf._obtain_sample(conditional_input=cond, predefined_target_input=z)
        ((sx * w).sum() + 0.3 * slp.sum()).backward()
        out["sgrad_w"] = w.numpy()
        for k, p_ in pdf.named_parameters():
            if p_.grad is not None:
                out["sgrad/" + k] = p_.grad.detach().numpy()
    if spec.get("grads", False):
        # reference autograd: d mean(log_pdf) / d (every parameter)
        pdf.zero_grad()
        lp, _, _ = pdf(x, conditional_input=cond)
        lp.mean().backward()
        for k, p_ in pdf.named_parameters():
            out["grad/" + k] = p_.grad.detach().numpy()
    return out


# fully_amortized_pdf (main/fully_amortized.py:22-278): one outer generator predicts, per row, the flow parameters AND the
# weights of every inner AmortizableMLP (pdf(..., amortize_everything=True); amortizable_mlp.py:586-611 with
# extra_inputs).  Inner connectivity modes 0-4, dense and factorised per-row layers, inner widths below and above 32.
FA_CASES = {
    "fa_e2s2e2_lowrank_mode1": dict(pdf_defs="e2+s2+e2", flow_defs="gg+f+gg", n=120, perturb=0.05,
                                    fa_kw=dict(conditional_input_dim=3, inner_mlp_dims_sub_pdfs="16", inner_mlp_ranks=3,
                                               inner_mlp_highway_mode=1, amortization_mlp_dims="32",
                                               amortization_mlp_ranks=5, amortization_mlp_highway_mode=0)),
    "fa_e2e1_dense_mode4_seq": dict(pdf_defs="e2+e1", flow_defs="gg+g", n=120, perturb=0.05,
                                    fa_kw=dict(conditional_input_dim=2, inner_mlp_dims_sub_pdfs="8-6", inner_mlp_ranks=0,
                                               inner_mlp_highway_mode=4, amortization_mlp_dims="16",
                                               amortization_mlp_use_custom_mode=False)),
    "fa_e3i1_wide_mode0": dict(pdf_defs="e3+i1", flow_defs="gg+r", n=60, perturb=0.05,
                               fa_kw=dict(conditional_input_dim=4, inner_mlp_dims_sub_pdfs="40", inner_mlp_ranks="3-0",
                                          inner_mlp_highway_mode=0, amortization_mlp_dims="24",
                                          amortization_mlp_ranks=4, amortization_mlp_highway_mode=1)),
    "fa_e1e2_mode3": dict(pdf_defs="e1+e2", flow_defs="g+gg", n=60, perturb=0.05,
                          fa_kw=dict(conditional_input_dim=2, inner_mlp_dims_sub_pdfs="34-5", inner_mlp_ranks=0,
                                     inner_mlp_highway_mode=3, amortization_mlp_dims="16", amortization_mlp_ranks=3,
                                     amortization_mlp_highway_mode=2)),
    "fa_e1s1_lowrank_mode2": dict(pdf_defs="e1+s1", flow_defs="gg+m", n=120, perturb=0.05,
                                  fa_kw=dict(conditional_input_dim=2, inner_mlp_dims_sub_pdfs="12-7",
                                             inner_mlp_ranks="2-3-2-2-1", inner_mlp_highway_mode=2,
                                             amortization_mlp_dims="16", amortization_mlp_ranks=3)),
}


def build_fa_case(jf, name, spec):
    seed = 1
    torch.manual_seed(seed)
    np.random.seed(seed)
    fa = jf.fully_amortized_pdf(spec["pdf_defs"], spec["flow_defs"], **spec["fa_kw"])
    # with a plain nn.Sequential outer generator the reference's init leaves the last bias in float32
    # (main/fully_amortized.py:253 assigns the float32 init vector): cast back as a user has to
    fa = fa.double()
    gen = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for _, p in fa.named_parameters():
            p.add_(spec["perturb"] * torch.randn(p.shape, generator=gen, dtype=torch.float64).to(p.dtype))
    n = spec["n"]
    x = make_inputs(spec["pdf_defs"], n, gen)
    cond = torch.randn(n, spec["fa_kw"]["conditional_input_dim"], generator=gen, dtype=torch.float64)
    inner = fa.pdf_to_amortize
    z = torch.randn(n, inner.total_base_dim, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        amort = fa.amortization_mlp(cond)
        logp, logp_base, base = fa(x, conditional_input=cond)
        samp_x, _, samp_logp, samp_logp_base = inner._obtain_sample(amortization_parameters=amort,
                                                                    predefined_target_input=z)
        rt_logp, _, rt_base = fa(samp_x, conditional_input=cond)
    out = {
        "meta": json.dumps(dict(name=name, pdf_defs=spec["pdf_defs"], flow_defs=spec["flow_defs"], fa_kw=spec["fa_kw"],
                                dtype="float64", perturb=spec["perturb"], seed=seed,
                                total_number_amortizable_params=int(inner.total_number_amortizable_params),
                                total_param_num=int(fa.total_param_num),
                                reference="thoglu/jammy_flows v1.1.0 @ /root/reference, torch %s CPU" % torch.__version__)),
        "x": x.numpy(), "z": z.numpy(), "cond": cond.numpy(), "amort_head": amort[:8].numpy(),
        "logp": logp.numpy(), "logp_base": logp_base.numpy(), "base": base.numpy(),
        "samp_x": samp_x.numpy(), "samp_logp": samp_logp.numpy(), "samp_logp_base": samp_logp_base.numpy(),
        "ref_roundtrip_base_err": np.nanmax(np.abs((rt_base - z).numpy())),
        "ref_roundtrip_logp_err": np.nanmax(np.abs((rt_logp - samp_logp).numpy())),
    }
    for k, v in fa.state_dict().items():
        out["param/" + k] = v.numpy()
    return out


def main():
    jf = import_reference()
    names = sys.argv[1:] or (list(CASES.keys()) + list(FA_CASES.keys()))
    for name in names:
        out = build_fa_case(jf, name, FA_CASES[name]) if name in FA_CASES else build_case(jf, name, CASES[name])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-32s logp[0:2]=%s rt_base_err=%.2e rt_logp_err=%.2e  %.1f KB" % (
            name, out["logp"][:2], out["ref_roundtrip_base_err"], out["ref_roundtrip_logp_err"],
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
