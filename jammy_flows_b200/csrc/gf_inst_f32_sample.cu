// explicit instantiation unit: "g"-chain kernels, float, JF_DIR_SAMPLE
#include "gf_launch.cuh"
namespace jf {
JF_GF_LAUNCH_DIR_BODY(float, JF_DIR_SAMPLE)
}
