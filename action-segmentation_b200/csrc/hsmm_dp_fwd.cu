// Instantiates the register-resident DP kernels (hsmm_dp_reg.cuh) for one (mode, precision) pair.
#include "hsmm_dp_reg.cuh"
namespace hsmm {
HSMM_DP_DEFINE_LAUNCHER(dp_launch_fwd, 1, false)
}
