// Parameter-generator MLP: params = W_L tanh(... tanh(W_1 [cond | embeddings] + b_1) ...) + b_L
// (reference main/default.py:654-670, called at :956 / :1438).
//
// v1 kernel: one CTA owns TM rows end to end (input gather -> every Linear+tanh in shared memory -> last Linear
// streamed to the caller's parameter buffer), so hidden activations never touch HBM.  The last layer writes
// "param-major" ([P, rows]) so the consuming layer kernel reads coalesced.
#pragma once
#include "common.cuh"

namespace jf {

template <typename T>
struct MlpArgs {
    int n_linear;
    int dims[JF_MAX_MLP_LINEAR + 1];
    int n_segments;
    int seg_cols[JF_MAX_MLP_SEGMENTS];
    const T* seg_ptr[JF_MAX_MLP_SEGMENTS];
    int64_t seg_ld[JF_MAX_MLP_SEGMENTS];
    const T* wt[JF_MAX_MLP_LINEAR];   // torch Linear.weight layout: [out, in] row-major
    const T* bias[JF_MAX_MLP_LINEAR];
    T* out; int64_t so_p, so_r;       // out[j*so_p + row*so_r]
    int64_t B;
    int lda;                          // odd leading dimension of the activation tiles
    int acc;                          // generic kernel only: out += result instead of out = result
};

constexpr int kMlpTN = 64;   // output columns per pass
constexpr int kMlpKC = 32;   // k-chunk of the weight tile
constexpr int kMlpLDW = kMlpTN + 1;   // padded row stride of the weight tile (conflict-free transposing store)

template <typename T, int TM>
__global__ void __launch_bounds__(256) mlp_kernel(const __grid_constant__ MlpArgs<T> m) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* act0 = reinterpret_cast<T*>(smem_raw);
    T* act1 = act0 + (size_t)TM * m.lda;
    T* sW = act1 + (size_t)TM * m.lda;   // [kMlpKC][kMlpLDW]
    constexpr int RPT = TM / 16;          // rows per thread
    const int tid = threadIdx.x;
    const int tx = tid & 15;              // row lane: rows tx + 16*i
    const int ty = tid >> 4;              // column group: cols ty*4 .. ty*4+3
    const int64_t row0 = (int64_t)blockIdx.x * TM;

    // gather the concatenated input rows
    {
        const int in_dim = m.dims[0];
        for (int e = tid; e < TM * in_dim; e += blockDim.x) {
            const int r = e / in_dim;
            int c = e - r * in_dim;
            const int64_t row = row0 + r;
            T v = 0;
            if (row < m.B) {
                int s = 0;
                while (c >= m.seg_cols[s]) { c -= m.seg_cols[s]; ++s; }
                v = m.seg_ptr[s][row * m.seg_ld[s] + c];
            }
            act0[(size_t)r * m.lda + (e - r * in_dim)] = v;
        }
    }
    __syncthreads();

    T* src = act0;
    T* dst = act1;
    for (int l = 0; l < m.n_linear; ++l) {
        const int Kin = m.dims[l], N = m.dims[l + 1];
        const bool last = (l == m.n_linear - 1);
        const T* __restrict__ W = m.wt[l];
        for (int n0 = 0; n0 < N; n0 += kMlpTN) {
            T acc[RPT][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = n0 + ty * 4 + c;
                const T b = (n < N) ? m.bias[l][n] : T(0);
#pragma unroll
                for (int i = 0; i < RPT; ++i) acc[i][c] = b;
            }
            for (int k0 = 0; k0 < Kin; k0 += kMlpKC) {
                __syncthreads();   // previous tile fully consumed
                for (int e = tid; e < kMlpKC * kMlpTN; e += blockDim.x) {
                    const int nn = e / kMlpKC, kk = e - nn * kMlpKC;     // k fastest: coalesced reads of W[n][k]
                    const int k = k0 + kk, n = n0 + nn;
                    sW[kk * kMlpLDW + nn] = (k < Kin && n < N) ? W[(size_t)n * Kin + k] : T(0);
                }
                __syncthreads();
                const int kend = min(kMlpKC, Kin - k0);
#pragma unroll 4
                for (int kk = 0; kk < kend; ++kk) {
                    T bv[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) bv[c] = sW[kk * kMlpLDW + ty * 4 + c];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const T av = src[(size_t)(tx + 16 * i) * m.lda + k0 + kk];
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[i][c] = fma(av, bv[c], acc[i][c]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = tx + 16 * i;
                const int64_t row = row0 + r;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int n = n0 + ty * 4 + c;
                    if (n >= N) continue;
                    if (last) {
                        if (row < m.B) {
                            T* o = m.out + (int64_t)n * m.so_p + row * m.so_r;
                            *o = m.acc ? (*o + acc[i][c]) : acc[i][c];
                        }
                    } else {
                        dst[(size_t)r * m.lda + n] = tanh(acc[i][c]);
                    }
                }
            }
        }
        __syncthreads();
        T* t = src; src = dst; dst = t;
    }
}


// One Linear layer with a narrow input (K <= 16) and a wide output: the U (.) half of a factorised layer of the outer
// generator of a fully amortized pdf (rank -> thousands of amortization parameters per row).  Write-bound: 8 N bytes per
// row against K N multiply-adds.  One thread per output column keeps its weight row in registers, the CTA's input rows
// sit in shared memory (broadcast reads), stores are coalesced along the row-major output.
constexpr int kExpandK = 16;
constexpr int kExpandRows = 64;

template <typename T> struct ExpandPair;
template <> struct ExpandPair<double> { typedef double2 type; };
template <> struct ExpandPair<float> { typedef float2 type; };

// KT: the input width padded to 4 / 8 / 16 (zero weights and zero inputs in the padding: no predicated instructions in
// the inner loop).  Every thread owns TWO output columns (c, c + 256) so that one shared-memory read of an input pair
// feeds four multiply-adds: ncu of the first version (one column per thread, one 8-byte LDS per multiply-add) showed
// the shared-memory pipe at 82 % and short-scoreboard stalls dominating at 2.9 TB/s of stores.
template <typename T, int KT>
__global__ void __launch_bounds__(256) mlp_expand_kernel(const __grid_constant__ MlpArgs<T> m) {
    // a CTA keeps its 512 weight rows in registers and walks over row tiles (blockIdx.y, stride gridDim.y): the weight
    // fetch and its latency are paid once per CTA, the next tile's inputs are staged while the current one is computed
    typedef typename ExpandPair<T>::type P2;
    __shared__ __align__(16) T s_in[2][kExpandRows][KT];
    const int K = m.dims[0], N = m.dims[1];
    const int c0 = blockIdx.x * 512 + threadIdx.x, c1 = c0 + 256;
    const bool live0 = c0 < N, live1 = c1 < N;
    const int64_t n_tiles = (m.B + kExpandRows - 1) / kExpandRows;
    auto stage = [&](int64_t tile, int buf) {
        const int64_t r0 = tile * kExpandRows;
        const int nr = (int)min((int64_t)kExpandRows, m.B - r0);
        for (int idx = threadIdx.x; idx < nr * KT; idx += 256) {
            const int r = idx / KT, kk = idx - r * KT;
            T v = T(0);
            if (kk < K) {
                int k = kk, s = 0;
                while (k >= m.seg_cols[s]) { k -= m.seg_cols[s]; ++s; }
                v = m.seg_ptr[s][(r0 + r) * m.seg_ld[s] + k];
            }
            s_in[buf][r][kk] = v;
        }
    };
    T w0[KT], w1[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        w0[k] = (live0 && k < K) ? m.wt[0][(int64_t)c0 * K + k] : T(0);
        w1[k] = (live1 && k < K) ? m.wt[0][(int64_t)c1 * K + k] : T(0);
    }
    const T b0 = live0 ? m.bias[0][c0] : T(0), b1 = live1 ? m.bias[0][c1] : T(0);
    int64_t tile = blockIdx.y;
    int buf = 0;
    if (tile < n_tiles) stage(tile, 0);
    __syncthreads();
    for (; tile < n_tiles; tile += gridDim.y) {
        if (tile + gridDim.y < n_tiles) stage(tile + gridDim.y, buf ^ 1);
        const int64_t r0 = tile * kExpandRows;
        const int nr = (int)min((int64_t)kExpandRows, m.B - r0);
        T* o0 = m.out + (int64_t)c0 * m.so_p + r0 * m.so_r;
        T* o1 = m.out + (int64_t)c1 * m.so_p + r0 * m.so_r;
#pragma unroll 2
        for (int r = 0; r < nr; ++r) {
            const P2* in2 = reinterpret_cast<const P2*>(&s_in[buf][r][0]);
            T a0 = b0, a1 = b1, e0 = T(0), e1 = T(0);
#pragma unroll
            for (int k = 0; k < KT; k += 2) {
                const P2 x = in2[k >> 1];
                a0 = fma(w0[k], x.x, a0);
                e0 = fma(w0[k + 1], x.y, e0);
                a1 = fma(w1[k], x.x, a1);
                e1 = fma(w1[k + 1], x.y, e1);
            }
            T v0 = a0 + e0, v1 = a1 + e1;
            if (live0) {
                if (m.acc) v0 += o0[r * m.so_r];
                o0[r * m.so_r] = v0;
            }
            if (live1) {
                if (m.acc) v1 += o1[r * m.so_r];
                o1[r * m.so_r] = v1;
            }
        }
        __syncthreads();
        buf ^= 1;
    }
}

}  // namespace jf

namespace jf {

// ---------------------------------------------------------------------------------------------------------------------
// Narrow generator: Linear(in <= 16) -> tanh(H <= 128) -> Linear(out <= 16), one THREAD per row.
// (reference main/default.py:654-670; the README flow's S2 sub-pdf: 4 -> 128 -> 10.)  With ten outputs the tensor-core
// kernel pays a 64-column MMA tile and a full TMEM drain for 10 numbers and is bound by its prologue (the 128 tanh per
// row); here the whole row lives in registers, the weights are broadcast from shared memory with 16-byte loads, four
// hidden units are in flight per thread, and two thirds of the executed instructions are FP64.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSmMaxIn = 16, kSmMaxOut = 16;

JF_DEVINL double sm_tanh(double x) {
    const double e = exp_neg(-2.0 * fabs(x));
    const double t = (1.0 - e) * rcp_1to2(1.0 + e);
    return x < 0.0 ? -t : t;
}
JF_DEVINL float sm_tanh(float x) { return tanhf(x); }

template <typename T, int IN, int OUT>     // IN, OUT: padded to even compile-time sizes (4/8/16)
__global__ void __launch_bounds__(128) mlp_small_kernel(const __grid_constant__ MlpArgs<T> m) {
    extern __shared__ __align__(16) unsigned char smem_small[];
    const int H = m.dims[1], Kin = m.dims[0], N = m.dims[2];
    T* sW1 = reinterpret_cast<T*>(smem_small);          // [H][IN]   (rows of W1, zero padded)
    T* sW2 = sW1 + (size_t)H * IN;                      // [H][OUT]  (W2 transposed, zero padded)
    T* sB1 = sW2 + (size_t)H * OUT;                     // [H]
    for (int e = threadIdx.x; e < H * IN; e += blockDim.x) {
        const int u = e / IN, i = e - u * IN;
        sW1[e] = i < Kin ? m.wt[0][(size_t)u * Kin + i] : T(0);
    }
    for (int e = threadIdx.x; e < H * OUT; e += blockDim.x) {
        const int u = e / OUT, j = e - u * OUT;
        sW2[e] = j < N ? m.wt[1][(size_t)j * H + u] : T(0);
    }
    for (int e = threadIdx.x; e < H; e += blockDim.x) sB1[e] = m.bias[0][e];
    __syncthreads();
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < m.B; row += (int64_t)gridDim.x * blockDim.x) {
        T x[IN], o[OUT];
#pragma unroll
        for (int c = 0; c < IN; ++c) {                   // (static c: the row stays in registers)
            T v = T(0);
            int off = 0;
            for (int sg = 0; sg < m.n_segments; ++sg) {
                const int cols = m.seg_cols[sg];
                if (c >= off && c < off + cols) v = m.seg_ptr[sg][row * m.seg_ld[sg] + (c - off)];
                off += cols;
            }
            x[c] = v;
        }
#pragma unroll
        for (int j = 0; j < OUT; ++j) o[j] = j < N ? m.bias[1][j] : T(0);
#pragma unroll 1
        for (int u0 = 0; u0 < H; u0 += 4) {
            T h[4];
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int u = u0 + uu;
                T z = u < H ? sB1[u] : T(0);
                const T* w = sW1 + (size_t)(u < H ? u : 0) * IN;
#pragma unroll
                for (int i = 0; i < IN; ++i) z = fma(x[i], w[i], z);
                h[uu] = u < H ? sm_tanh(z) : T(0);
            }
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const T* w = sW2 + (size_t)(u0 + uu < H ? u0 + uu : 0) * OUT;
#pragma unroll
                for (int j = 0; j < OUT; ++j) o[j] = fma(h[uu], w[j], o[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < OUT; ++j)
            if (j < N) {
                T* dst = m.out + (int64_t)j * m.so_p + row * m.so_r;
                *dst = m.acc ? *dst + o[j] : o[j];
            }
    }
}

}  // namespace jf
