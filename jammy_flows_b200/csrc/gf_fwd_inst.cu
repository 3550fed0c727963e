// log_pdf forward with (row, dimension) workers: the MODE 2 instantiations of csrc/gf_fb.cuh (own translation unit)
#define JF_FB_MODE 2
#include "gf_fb_inst.cu"
