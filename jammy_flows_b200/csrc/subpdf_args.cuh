// Arguments common to every sub-pdf kernel (one thread per row).
#pragma once
#include "common.cuh"

namespace jf {

template <typename T>
struct SubPdfArgs {
    // geometry
    int n_layers, d;
    int64_t B;
    // io
    const T* in;  int64_t ld_in;
    T* out;       int64_t ld_out;
    const T* params; int64_t sj, sr;   // element (j,row) at params[j*sj + row*sr]; sr == 0: shared
    const T* logdet_in;  T* logdet_out;
    const T* logbase_in; T* logbase_out;
    T* emb_out; int64_t ld_emb;
    int64_t* status;
    int tab_total;   // size of the processed table (elements) in shared mode
};

}  // namespace jf
