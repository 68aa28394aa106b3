// Instantiates the linear-window forward kernels (hsmm_dp_lin.cuh).
#include "hsmm_dp_lin.cuh"
namespace hsmm {
bool dp_lin_eligible(int C, int L, int mode, bool sparse, bool xp) { return lin_eligible(choose(C, L, mode, sparse, xp), xp); }
int dp_lin_launch_fwd(DpParams p, cudaStream_t st) {
    const RegChoice ch = choose(p.C, p.L, 1, p.trans_pred != nullptr, p.xp != 0);
    p.W = ch.W;
    p.VPB = 4;
    return launch_lin<1>(p, ch, st);
}
}  // namespace hsmm
