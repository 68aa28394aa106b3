// Host-side dispatch of the register-resident DP: shape support, variant names, launcher selection.
#include "hsmm_dp_reg.cuh"

namespace hsmm {

int dp_launch_vit(DpParams p, cudaStream_t st);
int dp_launch_fwd(DpParams p, cudaStream_t st);
int dp_launch_fwd_xp(DpParams p, cudaStream_t st);
int dp_launch_bwd(DpParams p, cudaStream_t st);
int dp_launch_bwd_xp(DpParams p, cudaStream_t st);

bool dp_reg_supported(int C, int L, int mode, bool sparse, bool xp) { return choose(C, L, mode, sparse, xp).v >= 0; }

const char* dp_reg_name(int C, int L, int mode, bool sparse, bool xp) {
    static thread_local char buf[128];
    RegChoice ch = choose(C, L, mode, sparse, xp);
    if (ch.v < 0) return "none";
    static const char* tmn[] = {"trans-reg", "trans-smem", "trans-sparse"};
    snprintf(buf, sizeof(buf), "reg<KR=%d,S=%d>/%s/%s/W=%d/VPB=%d/smem=%zu%s", kVariants[ch.v].KR, kVariants[ch.v].S,
             tmn[ch.tm], kVariants[ch.v].lreg ? "len-reg" : "len-smem", ch.W, ch.VPB, ch.smem, xp ? "/f64-state" : "");
    return buf;
}

int dp_reg_launch(DpParams p, int mode, cudaStream_t st) {
    if (mode == 0) return dp_launch_vit(p, st);
    if (mode == 1) return p.xp ? dp_launch_fwd_xp(p, st) : dp_launch_fwd(p, st);
    return p.xp ? dp_launch_bwd_xp(p, st) : dp_launch_bwd(p, st);
}

}  // namespace hsmm
