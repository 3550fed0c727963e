"""CPU: host-side mirror of the reference interface (constructor, DSL, option overrides, parameter contract)."""
import numpy as np
import pytest
import torch

import jammy_flows_b200 as jfb
from helpers import build_pdf, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_contract_and_seeded_init(name):
    """Names/shapes equal the reference's state_dict; with the same seed the constructor reproduces the reference's
    initial parameters bit for bit (RNG call order is mirrored) -- checked on the unperturbed goldens."""
    meta, params, _ = load_golden(name)
    p = build_pdf(meta, seed=meta["seed"])
    sd = p.state_dict()
    assert set(sd.keys()) == set(params.keys())
    for k in sd:
        assert tuple(sd[k].shape) == params[k].shape, k
    if meta["perturb"] == 0 and not name.startswith("init_"):     # init_*: parameters after init_params(data=...)
        for k in sd:
            assert np.array_equal(sd[k].numpy(), params[k]), k


def test_readme_headline_constructs_with_n_alias():
    """README.md:15-17 of the reference: pdf("e4+s2+e4", "gggg+n+gggg"); 'n' is an alias of 'f' (SURVEY.md F2)."""
    p = jfb.pdf("e4+s2+e4", "gggg+n+gggg")
    assert p.total_target_dim == 10 and p.total_base_dim == 10
    assert [sum(n) for n in p.num_parameter_list] == [548, 10, 548]
    assert p.mlp_predictors[0] is None
    assert p.mlp_predictors[1][0].in_features == 4 and p.mlp_predictors[1][-1].out_features == 10
    assert p.mlp_predictors[2][0].in_features == 7 and p.mlp_predictors[2][-1].out_features == 548
    assert p.count_parameters() == 74194


def test_first_layer_icdf_rule_and_offset():
    p = jfb.pdf("e2", "gg")
    l0, l1 = p.layer_list[0]
    assert l0.inverse_function_type == "inormal_partly_precise" and l0.model_offset == 0
    assert l1.inverse_function_type == "isigmoid" and l1.model_offset == 1
    single = jfb.pdf("e3", "g")
    assert single.layer_list[0][0].inverse_function_type == "isigmoid"     # offset branch wins (main/default.py:442-448)
    assert single.layer_list[0][0].total_param_num == 3 + 9 + 90
    assert p.count_parameters() == 130


def test_option_override_specificity():
    opts = {"g": {"num_kde": 7}, 0: {"g": {"num_kde": 5}}, (0, 1): {"g": {"num_kde": 3}}}
    p = jfb.pdf("e2+e2", "gg+gg", options_overwrite=opts)
    assert [l.num_kde for l in p.layer_list[0]] == [5, 3]
    assert [l.num_kde for l in p.layer_list[1]] == [7, 7]
    with pytest.raises(AssertionError):
        jfb.pdf("e2", "g", options_overwrite={"g": {"num_kde": -1}})
    with pytest.raises(AssertionError):
        jfb.pdf("e2", "g", options_overwrite={"g": {"no_such_option": 1}})


def test_conditional_wiring_dims():
    p = jfb.pdf("e6+s2", "gggggg+f", conditional_input_dim=64)
    assert p.mlp_predictors[0][0].in_features == 64 and p.mlp_predictors[0][-1].out_features == 6 * 216 + 6
    assert p.mlp_predictors[1][0].in_features == 70 and p.mlp_predictors[1][-1].out_features == 10
    assert len(list(p.layer_list.parameters())) == 0      # conditional: no permanent layer parameters


def test_unsupported_fails_loudly():
    for kw in (dict(skip_mlp_initialization=True),
               # a separate log-lambda MLP: the reference's own log_mean_poisson() calls it outdated and raises
               dict(predict_log_normalization=True, conditional_input_dim=2),
               # the reference's init fails for a joint prediction through an AmortizableMLP
               dict(predict_log_normalization=True, join_poisson_and_pdf_description=True, conditional_input_dim=2,
                    amortization_mlp_use_custom_mode=True)):
        with pytest.raises(NotImplementedError):
            jfb.pdf("e2", "gg", **kw)
    with pytest.raises(AssertionError):      # log-lambda prediction needs a single sub-pdf (main/default.py:471-472)
        jfb.pdf("e2+e1", "gg+g", predict_log_normalization=True)
    with pytest.raises(Exception):           # joint prediction needs a conditional input (main/default.py:583-584)
        jfb.pdf("e2", "gg", predict_log_normalization=True, join_poisson_and_pdf_description=True)
    with pytest.raises(AssertionError):      # reference main/default.py:118-119: amortizing everything needs custom MLPs
        jfb.pdf("e2", "gg", amortize_everything=True)
    with pytest.raises(NotImplementedError):
        jfb.pdf("e2", "gc")          # out of scope layer code
    with pytest.raises(NotImplementedError):
        jfb.pdf("e2", "gx")          # not built yet: must not fall back to anything
    with pytest.raises(NotImplementedError):
        jfb.pdf("s2", "v", options_overwrite={"v": {"mean_parametrization": "householder"}})   # (fails in the reference too)
    with pytest.raises(Exception):
        jfb.pdf("e2", "f")           # layer/manifold mismatch (main/default.py:397-398)


def test_cpu_tensors_are_rejected_no_fallback():
    p = jfb.pdf("e2", "gg").double()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p(torch.zeros(4, 2, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.sample(samplesize=3)


def test_amortize_everything_structure():
    """pdf(..., amortize_everything=True) owns no parameters; the amortizable count is the first sub-pdf's flow
    parameters plus every inner generator's flat vector (reference main/default.py:595-651)."""
    p = jfb.pdf("e2+s2", "gg+f", amortization_mlp_use_custom_mode=True, amortization_mlp_dims="16",
                amortization_mlp_ranks=3, amortization_mlp_highway_mode=1, amortize_everything=True)
    assert len(list(p.parameters())) == 0
    assert p.mlp_predictors[0] is None and p.mlp_predictors[1].use_permanent_parameters is False
    n_first = sum(p.num_parameter_list[0])
    assert p.total_number_amortizable_params == n_first + p.mlp_predictors[1].num_amortization_params
    init = p.init_params()
    assert init.shape == (p.total_number_amortizable_params,)
    with pytest.raises(Exception):           # CPU tensors are rejected: there is no fallback
        import torch
        p(torch.zeros(4, 4, dtype=torch.float64), amortization_parameters=init.double().unsqueeze(0).repeat(4, 1))


def test_backward_path_selection():
    """which pdfs take the differentiable path of pdf.forward / pdf.sample(allow_gradients=True) (host logic only): default
    "g" chains incl. rotation none and "t" layers, every non-Euclidean layer in its closed-form direction; everything else
    raises NotImplementedError at backward time instead of falling back"""
    from jammy_flows_b200 import engine
    yes = [("e3", "ggg", {}), ("e3", "ggt", {}), ("e2+s2+e2", "gg+f+gg", {}), ("s2", "v", {}), ("s1+i1", "o+r", {}), ("s1", "m", {}),
           ("e2", "gg", {"g": {"rotation_mode": "none", "fit_normalization": 0}}),
           ("s2", "f", {"f": {"add_vertical_rq_spline_flow": 1, "add_circular_rq_spline_flow": 1}})]
    no = [("e2", "gg", {"g": {"rotation_mode": "angles"}}), ("e2", "gg", {"g": {"add_skewness": 1}}),
          ("e2", "gg", {"g": {"nonlinear_stretch_type": "rq_splines"}}), ("e2", "gg", {"g": {"inverse_function_type": "inormal_full_pade"}}),
          ("s2", "v", {"v": {"natural_direction": 1}}), ("s1", "m", {"m": {"natural_direction": 1}})]
    for pd, fd, opts in yes:
        p = jfb.pdf(pd, fd, options_overwrite=opts)
        assert engine.supports_backward(p) and engine.supports_sample_backward(p), (pd, fd, opts)
    for pd, fd, opts in no:
        p = jfb.pdf(pd, fd, options_overwrite=opts)
        assert not engine.supports_backward(p), (pd, fd, opts)
    p = jfb.pdf("e2", "gg", conditional_input_dim=2, amortization_mlp_use_custom_mode=True)
    assert not engine.supports_backward(p)          # AmortizableMLP generators: staged inference path only
