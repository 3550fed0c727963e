// explicit instantiation unit: "g"-chain kernels, double, JF_DIR_SAMPLE
#include "gf_launch.cuh"
namespace jf {
JF_GF_LAUNCH_DIR_BODY(double, JF_DIR_SAMPLE)
}
