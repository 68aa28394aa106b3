// Instantiates the two-videos-per-warp kernels (hsmm_dp_pair.cuh).
#include "hsmm_dp_pair.cuh"
namespace hsmm {
bool dp_pair_eligible(int C, int L, bool sparse, bool xp) { return pair_shape_ok(C, L, sparse, xp); }
int dp_pair_launch(DpParams p, int mode, cudaStream_t st) {
    if (mode == 0) return launch_pair<0>(p, st);
    if (mode == 1) return launch_pair<1>(p, st);
    return launch_pair<2>(p, st);
}
}  // namespace hsmm
