// Instantiates the deferred-arg-max Viterbi kernels (hsmm_dp_vit2.cuh).
#include "hsmm_dp_vit2.cuh"
namespace hsmm {
bool dp_vit2_eligible(int C, int L, bool sparse) { return vit2_eligible(choose(C, L, 0, sparse, false), L); }
int dp_vit2_launch(DpParams p, cudaStream_t st) {
    const RegChoice ch = choose(p.C, p.L, 0, p.trans_pred != nullptr, false);
    p.W = ch.W;
    p.VPB = 4;
    return launch_vit2(p, ch, st);
}
}  // namespace hsmm
