// backward of the sampling direction of the "g" chain: the MODE 1 instantiations of csrc/gf_fb.cuh (own translation unit)
#define JF_FB_MODE 1
#include "gf_fb_inst.cu"
