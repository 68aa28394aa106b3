// Instantiates the linear-window backward kernels (hsmm_dp_lin.cuh).
#include "hsmm_dp_lin.cuh"
namespace hsmm {
int dp_lin_launch_bwd(DpParams p, cudaStream_t st) {
    const RegChoice ch = choose(p.C, p.L, 2, p.trans_succ != nullptr, p.xp != 0);
    p.W = ch.W;
    p.VPB = 4;
    return launch_lin<2>(p, ch, st);
}
}  // namespace hsmm
