// Per-row Jacobian of the log_pdf of a NON-EUCLIDEAN sub-pdf (S2: "f" with every option and nested spline sub-flows, "v";
// S1: "o", "m"; interval: "r") with respect to its raw parameters and its coordinates -- what the reference gets from
// autograd through layers/spheres/*.py, layers/intervals/*.py, layers/spline_fns.py.
//
// Forward-mode sweep: the layer code of csrc/s2.cuh / chain1.cuh / spline.cuh is instantiated with dual numbers
// (csrc/dual.cuh), thread (row, seed) runs the log_pdf direction with ONE input seeded and writes one Jacobian entry.
// These layers cost ~1 % of a step in the value path and have 10-100 parameters, so (parameters + coordinates) passes
// stay cheap next to the "g" chains (which have their own closed-form reverse pass, csrc/gf_fb.cuh), and every branch,
// clamp and parametrisation option of the value path is differentiated by construction.
#pragma once
#include "dual.cuh"
#include "subpdf_kernels.cuh"

namespace jf {

constexpr int kJacMaxParams = 160;      // raw parameters of one manifold sub-pdf (a thread holds them as dual numbers)

template <typename F>
struct JacIO {
    int n_params, d;                    // seeds 0 .. n_params-1: parameters; n_params .. n_params+d-1: coordinates
    int64_t B;
    const F* in; int64_t ld_in;
    const F* params; int64_t sj, sr;    // element (i,row) at params[i*sj + row*sr]; sr == 0: shared
    F* jac; int64_t jac_sj;             // out: element (i,row) at jac[i*jac_sj + row]
    F* jx; int64_t ld_jx;               // out, optional: [B, d]
    F* jbase;                           // out, optional: [n_params + d][B][d] derivative of the BASE point per seed
};

template <typename F>
JF_DEVINL void jac_load_params(const JacIO<F>& io, int64_t row, int seed, Dual* p) {
    const F* prow = io.params + row * io.sr;
    for (int i = 0; i < io.n_params; ++i) p[i] = Dual((double)prow[(int64_t)i * io.sj], i == seed ? 1.0 : 0.0);
}
template <typename F>
JF_DEVINL void jac_store(const JacIO<F>& io, int64_t row, int seed, double v) {
    if (seed < io.n_params) io.jac[(int64_t)seed * io.jac_sj + row] = (F)v;
    else io.jx[row * io.ld_jx + (seed - io.n_params)] = (F)v;
}
template <typename F>
JF_DEVINL void jac_store_base(const JacIO<F>& io, int64_t row, int seed, int i, double v) {
    if (io.jbase != nullptr) io.jbase[((int64_t)seed * io.B + row) * io.d + i] = (F)v;
}

template <typename F>
__global__ void __launch_bounds__(128) s2_jac_kernel(const __grid_constant__ JacIO<F> io, const __grid_constant__ S2Args<Dual> g) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int seed = blockIdx.y;
    if (row >= io.B || (seed >= io.n_params && io.jx == nullptr)) return;
    Dual p[kJacMaxParams];
    jac_load_params(io, row, seed, p);
    Dual c0((double)io.in[row * io.ld_in + 0], seed == io.n_params ? 1.0 : 0.0);
    Dual c1((double)io.in[row * io.ld_in + 1], seed == io.n_params + 1 ? 1.0 : 0.0);
    Dual logdet(0.0);
    int oor = 0, evals = 0, unconv = 0;
    for (int l = g.a.n_layers - 1; l >= 0; --l) {
        if (g.layers[l].kind == JF_LAYER_EXPMAP) v_layer<Dual>(true, c0, c1, logdet, g.layers[l], p, 1, evals, unconv);
        else fvm_logpdf<Dual>(c0, c1, logdet, g.layers[l], g.splines, p, 1, oor);
    }
    const Dual total = logdet - Dual(0.5) * (c0 * c0 + c1 * c1);
    jac_store(io, row, seed, total.d);
    jac_store_base(io, row, seed, 0, c0.d);
    jac_store_base(io, row, seed, 1, c1.d);
}

template <typename F>
__global__ void __launch_bounds__(128) chain1_jac_kernel(const __grid_constant__ JacIO<F> io, const __grid_constant__ Chain1Args<Dual> g) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int seed = blockIdx.y;
    if (row >= io.B || (seed >= io.n_params && io.jx == nullptr)) return;
    Dual p[kJacMaxParams];
    jac_load_params(io, row, seed, p);
    Dual x((double)io.in[row * io.ld_in], seed == io.n_params ? 1.0 : 0.0);
    Dual logdet(0.0);
    int oor = 0, evals = 0, unconv = 0;
    for (int l = g.a.n_layers - 1; l >= 0; --l)
        x = layer1_logpdf<Dual>(g.layers[l], g.manifold, x, logdet, p, 1, oor, evals, unconv);
    const Dual total = logdet - Dual(0.5) * x * x;
    jac_store(io, row, seed, total.d);
    jac_store_base(io, row, seed, 0, x.d);
}

}  // namespace jf
