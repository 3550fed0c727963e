"""CPU: the oracle (oracle/jf_oracle.py) against the golden vectors produced by the unmodified reference.
This is what pins the oracle (the reference has no stored fixtures of its own)."""
import numpy as np
import pytest

from helpers import build_fa, build_pdf, fa_golden_names, fa_outer_spec, golden_names, load_golden, rel_err
from oracle.jf_oracle import OraclePdf, fully_amortized_parameters

# the oracle uses the reference's own arithmetic (same torch ops), so it reproduces the goldens to rounding
TOL = {"float64": 1e-12, "float32": 2e-6}


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference(name):
    meta, params, data = load_golden(name)
    pdf = build_pdf(meta)
    o = OraclePdf(pdf.export_program(meta["dtype"]), params)
    cond = data.get("cond")
    tol = TOL[meta["dtype"]]
    if "natural" in name:
        # natural_direction=1: the log_pdf direction itself is an iterative inverse (stop criterion 1e-12 on the step),
        # so two implementations agree only to the reference's own round-trip error
        tol = max(tol, 10 * float(data["ref_roundtrip_base_err"]))
    logp, logp_base, base = o.log_pdf(data["x"], cond)
    ok = np.isfinite(data["logp"])
    assert rel_err(logp.numpy(), data["logp"])[ok].max() < tol
    assert rel_err(logp_base.numpy(), data["logp_base"])[ok].max() < tol
    assert rel_err(base.numpy(), data["base"])[ok].max() < tol
    xs, slogp, slogp_base = o.sample(data["z"], cond)
    ok = np.isfinite(data["samp_logp"]) & np.isfinite(data["samp_x"]).all(axis=1)
    # sampling goes through 25 bisections + <=20 Newton steps; allow the reference's own round-trip error
    stol = max(tol, 10 * float(np.nan_to_num(data["ref_roundtrip_base_err"])))
    assert rel_err(xs.numpy(), data["samp_x"])[ok].max() < max(stol, 1e-9 if meta["dtype"] == "float64" else 1e-4)
    assert rel_err(slogp_base.numpy(), data["samp_logp_base"])[ok].max() < tol


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("emb_")])
def test_oracle_embedding_coordinates(name):
    """force_embedding_coordinates: charts before the log_pdf chain / after the sampling chain."""
    meta, params, data = load_golden(name)
    pdf = build_pdf(meta)
    o = OraclePdf(pdf.export_program(meta["dtype"]), params)
    cond = data.get("cond")
    xe, _ = o.transform_target(data["x"], to_embedding=True)
    assert np.abs(xe.numpy() - data["x_emb"]).max() < 1e-14
    back, _ = o.transform_target(data["x_emb"], to_embedding=False)
    assert np.abs(back.numpy() - data["x"]).max() < 1e-7          # acos near the poles
    logp, _, base = o.log_pdf_embedding(data["x_emb"], cond)
    assert rel_err(logp.numpy(), data["logp_emb"]).max() < 1e-12
    assert rel_err(base.numpy(), data["base_emb"]).max() < 1e-12
    xs, slogp, _ = o.sample_embedding(data["z"], cond)
    stol = max(1e-9, 10 * float(data["ref_roundtrip_base_err"]))
    assert rel_err(xs.numpy(), data["samp_x_emb"]).max() < stol
    assert rel_err(slogp.numpy(), data["samp_logp_emb"]).max() < stol
    # entropies (total + every marginal), both coordinate conventions, on the reference's own base normals
    nb = 3 if cond is not None else 1
    S = int(data["ent_S"])
    subs = [-1] + list(range(len(meta["pdf_defs"].split("+"))))
    for flag, tag in ((True, "emb"), (False, "intr")):
        ent = o.entropy(data["ent_z"], None if cond is None else cond[:nb], S, subs, embedding=flag)
        for k_, v_ in ent.items():
            ref = data["ent_%s_%s" % (tag, k_)]
            assert rel_err(v_.numpy(), ref).max() < max(1e-10, 10 * float(data["ref_roundtrip_base_err"])), (tag, k_)


@pytest.mark.parametrize("name", fa_golden_names())
def test_oracle_fully_amortized(name):
    """fully_amortized_pdf (main/fully_amortized.py): outer generator -> per-row flow parameters and per-row inner MLP
    weights (pdf(amortize_everything=True)); also pins the parameter counts and the bit-identical seeded init."""
    meta, params, data = load_golden(name)
    fa = build_fa(meta, seed=meta["seed"])
    assert fa.pdf_to_amortize.total_number_amortizable_params == meta["total_number_amortizable_params"]
    assert fa.count_parameters() == meta["total_param_num"]
    assert sorted(fa.state_dict().keys()) == sorted(params.keys())
    inner = fa.pdf_to_amortize
    assert len(list(inner.parameters())) == 0          # everything is amortized: the inner pdf owns nothing
    am = fully_amortized_parameters(fa_outer_spec(fa), params, data["cond"])
    assert am.shape[1] == meta["total_number_amortizable_params"]
    assert rel_err(am[:8].numpy(), data["amort_head"]).max() < 1e-12
    o = OraclePdf(inner.export_program("float64"), {})
    logp, logp_base, base = o.log_pdf(data["x"], None, amort=am)
    assert rel_err(logp.numpy(), data["logp"]).max() < 1e-11
    assert rel_err(logp_base.numpy(), data["logp_base"]).max() < 1e-11
    assert rel_err(base.numpy(), data["base"]).max() < 1e-11
    xs, slogp, slogp_base = o.sample(data["z"], None, amort=am)
    stol = max(1e-9, 10 * float(data["ref_roundtrip_base_err"]))
    assert rel_err(xs.numpy(), data["samp_x"]).max() < stol
    assert rel_err(slogp.numpy(), data["samp_logp"]).max() < stol
    assert rel_err(slogp_base.numpy(), data["samp_logp_base"]).max() < 1e-12


def test_fully_amortized_seeded_init_matches_reference_layout():
    """the init puts the desired flow / inner-generator values into the LAST bias of the outer generator
    (main/fully_amortized.py:224-253, main/default.py:1828-1946): the tail of u_v_b_pars is the inner init vector"""
    import torch
    meta, params, _ = load_golden("fa_e2s2e2_lowrank_mode1")
    fa = build_fa(meta, seed=3)
    t = fa.pdf_to_amortize.total_number_amortizable_params
    torch.manual_seed(5)
    init = fa.pdf_to_amortize.init_params()
    assert init.shape == (t,)
    fa.amortization_mlp.initialize_uvbs(fix_final_bias=init)
    assert torch.equal(fa.amortization_mlp.u_v_b_pars.data[0, -t:], init.double())


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("last_")])
def test_oracle_only_last(name):
    """forward(..., only_last=True) / _obtain_sample(..., only_last=True): main/default.py:1015-1024, :1490-1502"""
    meta, params, data = load_golden(name)
    pdf = build_pdf(meta)
    o = OraclePdf(pdf.export_program(meta["dtype"]), params)
    cond = data.get("cond")
    logp, logp_base, base = o.log_pdf(data["x"], cond, only_last=True)
    assert rel_err(logp.numpy(), data["last_logp"]).max() < 1e-12
    assert rel_err(logp_base.numpy(), data["last_logp_base"]).max() < 1e-12
    assert rel_err(base.numpy(), data["last_base"]).max() < 1e-12
    xs, slogp, _ = o.sample(data["z"], cond, only_last=True)
    # the "v" layer's sampling direction is an iterative inverse: agreement to the reference's own round-trip error
    stol = max(1e-9, 10 * float(data["ref_roundtrip_base_err"]))
    assert rel_err(xs.numpy(), data["last_samp_x"]).max() < stol
    assert rel_err(slogp.numpy(), data["last_samp_logp"]).max() < stol


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("poisson_")])
def test_poisson_log_normalization_structure(name):
    """predict_log_normalization (main/default.py:466-477, :624-626, :1893-1897): the generator predicts one extra
    (last) output, initialised to 0.1; an unconditional pdf owns a (1,1) parameter.  The flow itself is unchanged, which
    test_oracle_reproduces_reference checks on the same files (the oracle reads the first n_par generator outputs)."""
    import torch
    meta, params, data = load_golden(name)
    pdf = build_pdf(meta, seed=meta["seed"])
    assert sorted(pdf.state_dict().keys()) == sorted(params.keys())
    for k, v in pdf.state_dict().items():
        assert tuple(v.shape) == params[k].shape, k
    n_par = sum(pdf.num_parameter_list[0])
    if meta["conditional_input_dim"] is None:
        assert pdf.log_normalization.shape == (1, 1) and data["log_lambda"].shape == (1, 1)
        pdf.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        assert np.array_equal(pdf.log_mean_poisson().detach().numpy(), data["log_lambda"])
    else:
        last = pdf.mlp_predictors[0][-1]
        assert last.out_features == n_par + 1
        assert abs(float(last.bias.data[-1]) - 0.1) < 1e-7
        assert data["log_lambda"].shape == (data["x"].shape[0], 1)
