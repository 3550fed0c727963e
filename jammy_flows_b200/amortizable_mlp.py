"""`AmortizableMLP` with permanent parameters: the custom parameter generator of the reference
(`amortization_mlp_use_custom_mode=True`), jammy_flows/amortizable_mlp.py:9-682.

What it is: a flat parameter vector `u_v_b_pars` [1, N] that encodes one or several chains of Linear/tanh layers, every
weight matrix either dense or a rank-r product U V^T, combined in one of five connectivity ("highway") modes:
  0  one chain input -> hidden... -> output
  1  mode 0 + a linear map input -> output (which carries the final bias)
  2  a sum of one-hidden-layer chains, all reading the input, + the linear map
  3  like 2, but chain i > 0 reads the running output sum
  4  like 3, but chain i > 0 reads [input, running output sum]
The layout of the vector (per layer: U, V, bias; chains in order; the linear map last), the rank rules of svd_mode
"smart"/"naive" and the initialisation (including its RNG call order) follow the reference, so `state_dict`s are
interchangeable and equal seeds give equal parameters.

How it runs here: every chain is a plain Linear/tanh chain for the sm_100a MLP kernel (`jf_mlp_forward_acc`); low-rank
factors are multiplied out once per call (out x r times r x in, tiny) so the kernel sees dense weights, and the
connectivity modes are composed on the host from accumulating kernel launches (jammy_flows_b200/engine.py).

"Being amortised" (`use_permanent_parameters=False`, reference amortizable_mlp.py:233-246, :586-611): the module owns no
parameters; `forward(i, extra_inputs=[B, num_amortization_params])` applies a DIFFERENT network to every row, its flat
vector being that row of `extra_inputs` (what `pdf(..., amortize_everything=True)` / `fully_amortized_pdf` do for every
sub-pdf).  Every layer is one `jf_rowwise_linear` launch (two for a factorised layer: V^T x, then U (.)), composed by
`engine.amortized_mlp_forward` in the same connectivity modes.
"""
import math

import numpy
import torch
from torch import nn


def list_from_str(spec):
    if spec == "":
        return []
    return list(tuple(map(int, spec.split("-"))))


class _Chain:
    """One Linear/tanh chain inside the flat parameter vector."""

    def __init__(self, inputs, outputs, ranks, final_bias, svd_mode):
        self.inputs, self.outputs = list(inputs), list(outputs)
        self.final_bias = final_bias
        self.layers = []          # dicts: n_in, n_out, rank, full, n_u, n_v, n_b
        n = len(self.inputs)
        for i in range(n):
            n_in, n_out, lr = self.inputs[i], self.outputs[i], ranks[i]
            max_rank = min(n_in, n_out)
            if lr > 0:
                used = min(max_rank, lr)
            else:
                used = 0 if svd_mode == "naive" else max_rank
            if svd_mode == "naive":
                low = used > 0
            elif svd_mode == "smart":
                # a factorisation that needs more numbers than the dense matrix is stored dense
                low = (used * (n_in + n_out) < n_in * n_out) and (lr > 0)
            else:
                raise Exception("unknown svd mode", svd_mode)
            last = i == n - 1
            self.layers.append(dict(n_in=n_in, n_out=n_out, rank=used, full=0 if low else 1,
                                    n_u=used * n_out if low else n_in * n_out, n_v=used * n_in if low else 0,
                                    n_b=n_out if (not last or final_bias) else 0))
        self.num_params = sum(l["n_u"] + l["n_v"] + l["n_b"] for l in self.layers)

    def segments(self, flat):
        """flat [num_params] -> list of Linear/tanh chains [[(W, b), ...], ...] to be run one after the other (no
        activation between two chains).  A factorised layer whose factors are at least 4x smaller than the dense
        matrix is applied as the reference does, U (V^T x): the chain is cut after V^T (a `rank`-wide intermediate).
        With an output as wide as the parameter vector of a fully amortized pdf that is the difference between
        rank*(in+out) and in*out multiply-adds per row; smaller layers are multiplied out (one launch)."""
        segs, cur, pos = [], [], 0
        for l in self.layers:
            u = flat[pos:pos + l["n_u"]]
            pos += l["n_u"]
            v = flat[pos:pos + l["n_v"]]
            pos += l["n_v"]
            b = flat[pos:pos + l["n_b"]]
            pos += l["n_b"]
            if l["n_b"] == 0:
                b = torch.zeros(l["n_out"], dtype=flat.dtype, device=flat.device)
            if l["full"]:
                cur.append((u.reshape(l["n_out"], l["n_in"]).contiguous(), b.contiguous()))
            elif 4 * l["rank"] * (l["n_in"] + l["n_out"]) <= l["n_in"] * l["n_out"]:
                cur.append((v.reshape(l["rank"], l["n_in"]).contiguous(),
                            torch.zeros(l["rank"], dtype=flat.dtype, device=flat.device)))
                segs.append(cur)
                cur = [(u.reshape(l["n_out"], l["rank"]).contiguous(), b.contiguous())]
            else:
                w = torch.matmul(u.reshape(l["n_out"], l["rank"]), v.reshape(l["rank"], l["n_in"]))
                cur.append((w.contiguous(), b.contiguous()))
        segs.append(cur)
        return segs

    def dense(self, flat):
        """flat [num_params] -> [(W [out,in], b [out])] with low-rank factors multiplied out."""
        out, pos = [], 0
        for l in self.layers:
            u = flat[pos:pos + l["n_u"]]
            pos += l["n_u"]
            v = flat[pos:pos + l["n_v"]]
            pos += l["n_v"]
            b = flat[pos:pos + l["n_b"]]
            pos += l["n_b"]
            if l["full"]:
                w = u.reshape(l["n_out"], l["n_in"])
            else:
                w = torch.matmul(u.reshape(l["n_out"], l["rank"]), v.reshape(l["rank"], l["n_in"]))
            if l["n_b"] == 0:
                b = torch.zeros(l["n_out"], dtype=flat.dtype, device=flat.device)
            out.append((w.contiguous(), b.contiguous()))
        return out


class AmortizableMLP(nn.Module):

    def __init__(self, input_dim, hidden_dims, output_dim, highway_mode=0, low_rank_approximations=0,
                 nonlinearity="tanh", use_permanent_parameters=True, svd_mode="smart", precise_mlp_structure=dict()):
        super().__init__()
        if nonlinearity != "tanh":
            raise NotImplementedError("AmortizableMLP: only the tanh nonlinearity has an sm_100a kernel")
        if len(precise_mlp_structure.keys()) > 0:
            raise NotImplementedError("AmortizableMLP: precise_mlp_structure is not supported")
        assert 0 <= highway_mode <= 4
        self.input_dim, self.output_dim = input_dim, output_dim
        self.highway_mode = highway_mode
        self.nonlinearity = nonlinearity
        self.svd_mode = svd_mode
        self.use_permanent_parameters = use_permanent_parameters
        if type(hidden_dims) == str:
            self.hidden_dims = list_from_str(hidden_dims)
        elif type(hidden_dims) == int:
            self.hidden_dims = [hidden_dims]
        elif type(hidden_dims) == list:
            self.hidden_dims = hidden_dims
        else:
            raise Exception("Unsupported type ", type(hidden_dims), " for hidden_dims .. can be int/str/list of ints")
        h = self.hidden_dims
        n_mat = len(h) + 1 if highway_mode == 0 else (len(h) + 2 if highway_mode == 1 else 2 * len(h) + 1)
        ranks = low_rank_approximations
        if type(ranks) == int:
            ranks = n_mat * [ranks]
        elif type(ranks) == str:
            ranks = list_from_str(ranks)
        assert len(ranks) == n_mat, (len(ranks), n_mat)
        self.total_low_rank_approximations = ranks
        self.chains, self.highway = [], None
        if highway_mode < 2:
            if highway_mode == 0:
                self.chains.append(_Chain([input_dim] + h, h + [output_dim], ranks, True, svd_mode))
            else:
                if len(h) > 0:
                    self.chains.append(_Chain([input_dim] + h, h + [output_dim], ranks[:-1], False, svd_mode))
                self.highway = _Chain([input_dim], [output_dim], ranks[-1:], True, svd_mode)
        else:
            start = {2: input_dim, 3: output_dim, 4: input_dim + output_dim}[highway_mode]
            for ind in range(len(h)):
                first = input_dim if ind == 0 else start
                self.chains.append(_Chain([first, h[ind]], [h[ind], output_dim], ranks[2 * ind:2 * ind + 2], False, svd_mode))
            self.highway = _Chain([input_dim], [output_dim], ranks[-1:], True, svd_mode)
        self.num_amortization_params = sum(c.num_params for c in self.chains) + (self.highway.num_params if self.highway else 0)
        if use_permanent_parameters:
            self.u_v_b_pars = nn.Parameter(torch.randn(self.num_amortization_params).type(torch.double).unsqueeze(0))
            self.initialize_uvbs()
        else:
            self.u_v_b_pars = None      # the flat vector arrives per row through forward(extra_inputs=...)

    # ---- initialisation (reference amortizable_mlp.py:375-466) -----------------------------------------------------
    def obtain_default_init_tensor(self, fix_final_bias=None, prev_damping_factor=1000.0):
        init = torch.randn(self.num_amortization_params, dtype=torch.float64).unsqueeze(0)
        index = 0
        for ch in self.chains:
            for l in ch.layers:
                if l["full"] == 1:       # dense layers get torch's Linear init; factorised ones keep the normal draw
                    gain = nn.init.calculate_gain("leaky_relu", numpy.sqrt(5))
                    bound = math.sqrt(3.0) * gain / math.sqrt(l["n_in"])
                    with torch.no_grad():
                        init[:, index:index + l["n_u"]].uniform_(-bound, bound)
                    bound = 1 / numpy.sqrt(l["n_in"])
                    if l["n_b"] > 0:
                        nn.init.uniform_(init[:, index + l["n_u"]:index + l["n_u"] + l["n_b"]], -bound, bound)
                index += l["n_u"] + l["n_v"] + l["n_b"]
        if self.highway is not None:
            l = self.highway.layers[0]
            gain = nn.init.calculate_gain("leaky_relu", numpy.sqrt(5))
            bound = math.sqrt(3.0) * gain / math.sqrt(l["n_in"])
            with torch.no_grad():
                init[:, index:index + l["n_u"]].uniform_(-bound, bound)
            bound = 1 / numpy.sqrt(l["n_in"])
            nn.init.uniform_(init[:, index + l["n_u"]:index + l["n_u"] + l["n_b"]], -bound, bound)
        if fix_final_bias is not None:
            init = init / prev_damping_factor
            rel = self.highway if self.highway is not None else self.chains[-1]
            init[0, -rel.layers[-1]["n_b"]:] = fix_final_bias
        return init.squeeze(0)

    def initialize_uvbs(self, fix_total=None, fix_final_bias=None, prev_damping_factor=1000.0):
        if fix_total is not None:
            self.u_v_b_pars.data[0, ...] = fix_total
        else:
            self.u_v_b_pars.data[0, ...] = self.obtain_default_init_tensor(fix_final_bias=fix_final_bias,
                                                                           prev_damping_factor=prev_damping_factor)

    # ---- structure for the engine / the test oracle ----------------------------------------------------------------
    def structure(self):
        conv = lambda c: dict(layers=[dict(l) for l in c.layers], num_params=c.num_params)
        return dict(custom=True, amortised=not self.use_permanent_parameters, highway_mode=self.highway_mode, input_dim=self.input_dim, output_dim=self.output_dim,
                    chains=[conv(c) for c in self.chains], highway=conv(self.highway) if self.highway else None)

    def dense_weights(self, dtype, device):
        """-> ([chain: [(W, b), ...]], highway [(W, b)] or None), dense tensors on `device`."""
        flat = self.u_v_b_pars.detach().to(device=device, dtype=dtype).reshape(-1)
        pos, chains = 0, []
        for c in self.chains:
            chains.append(c.dense(flat[pos:pos + c.num_params]))
            pos += c.num_params
        hw = self.highway.dense(flat[pos:pos + self.highway.num_params]) if self.highway is not None else None
        return chains, hw

    def chain_segments(self, dtype, device):
        """like dense_weights, every chain as a list of consecutive Linear/tanh chains (`_Chain.segments`)"""
        flat = self.u_v_b_pars.detach().to(device=device, dtype=dtype).reshape(-1)
        pos, chains = 0, []
        for c in self.chains:
            chains.append(c.segments(flat[pos:pos + c.num_params]))
            pos += c.num_params
        hw = self.highway.segments(flat[pos:pos + self.highway.num_params]) if self.highway is not None else None
        return chains, hw

    def forward(self, i, extra_inputs=None):
        """[B, input_dim] (CUDA) -> [B, output_dim].  Reference: amortizable_mlp.py:586-682."""
        from . import engine
        if extra_inputs is not None:
            if self.use_permanent_parameters:
                raise Exception("MLP uses permanent parameters but extra inputs are given in forward. This is not allowed!")
            assert (extra_inputs.shape[1] == self.num_amortization_params), \
                ("Extra inputs dimension (%d) does not match number of amortization params of MLP (%d) "
                 % (extra_inputs.shape[1], self.num_amortization_params))
            return engine.amortized_mlp_forward(self, [i], extra_inputs, i.shape[0])
        assert (self.use_permanent_parameters)
        return engine.custom_mlp_forward(self, [i], i.shape[0])
