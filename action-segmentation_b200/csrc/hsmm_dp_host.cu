// Host-side dispatch of the register-resident DP: shape support, variant names, launcher selection.
#include <stdlib.h>

#include "hsmm_dp_reg.cuh"

namespace hsmm {

int dp_launch_vit(DpParams p, cudaStream_t st);
int dp_launch_fwd(DpParams p, cudaStream_t st);
int dp_launch_fwd_xp(DpParams p, cudaStream_t st);
int dp_launch_bwd(DpParams p, cudaStream_t st);
int dp_launch_bwd_xp(DpParams p, cudaStream_t st);
bool dp_lin_eligible(int C, int L, int mode, bool sparse, bool xp);
int dp_lin_launch_fwd(DpParams p, cudaStream_t st);
int dp_lin_launch_bwd(DpParams p, cudaStream_t st);
bool dp_vit2_eligible(int C, int L, bool sparse);
int dp_vit2_launch(DpParams p, cudaStream_t st);
bool dp_pair_eligible(int C, int L, bool sparse, bool xp);
int dp_pair_launch(DpParams p, int mode, cudaStream_t st);

// HSMM_DISABLE_LIN=1 keeps every video on the log-domain kernels (A/B comparisons, debugging)
static std::atomic<int> g_lin{-1};
static bool lin_enabled() {
    int v = g_lin.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("HSMM_DISABLE_LIN");
        v = (e && e[0] == '1') ? 0 : 1;
        g_lin.store(v, std::memory_order_relaxed);
    }
    return v == 1;
}
bool dp_lin_enabled() { return lin_enabled(); }
int dp_lin_set_enabled(int on) {
    const int prev = lin_enabled() ? 1 : 0;
    g_lin.store(on ? 1 : 0, std::memory_order_relaxed);
    return prev;
}
// Two videos per warp (hsmm_dp_pair.cuh) pays off when the call has enough videos to keep every SM's issue slots busy
// (throughput regime: 1.6-1.7x fewer instructions per frame); with fewer videos a launch lasts as long as its longest
// video, and the pair kernels' longer per-frame dependency chain (a 20-deep window per lane instead of 10) makes that
// worse.  Threshold on the videos of the call (of the whole group for hsmm_dp_grouped); HSMM_PAIR_MIN_VIDEOS or
// hsmm_set_pair_min_videos() override it (0 = always, negative = never).
static std::atomic<int> g_pair_min{-2};
static int pair_min_videos() {
    int v = g_pair_min.load(std::memory_order_relaxed);
    if (v == -2) {
        const char* e = getenv("HSMM_PAIR_MIN_VIDEOS");
        v = e ? atoi(e) : 4096;
        if (getenv("HSMM_DISABLE_PAIR") && getenv("HSMM_DISABLE_PAIR")[0] == '1') v = -1;
        g_pair_min.store(v, std::memory_order_relaxed);
    }
    return v;
}
int dp_pair_set_min_videos(int n) {
    const int prev = pair_min_videos();
    g_pair_min.store(n < 0 ? -1 : n, std::memory_order_relaxed);
    return prev;
}
// Grouped launches (hsmm_dp_grouped): from this many videos on, the C <= 16 tasks of a float-state group run two videos per
// warp INSIDE the same launch as the other tasks (hsmm_dp_group.cu: dp_mixed_*).  HSMM_MIXED_MIN_VIDEOS /
// hsmm_set_mixed_min_videos() override it (0 = always, negative = never); never when the pair kernels are disabled.
static std::atomic<int> g_mixed_min{-2};
int dp_mixed_min_videos() {
    int v = g_mixed_min.load(std::memory_order_relaxed);
    if (v == -2) {
        const char* e = getenv("HSMM_MIXED_MIN_VIDEOS");
        v = e ? atoi(e) : 1024;
        g_mixed_min.store(v < 0 ? -1 : v, std::memory_order_relaxed);
    }
    return pair_min_videos() < 0 ? -1 : v;
}
int dp_mixed_set_min_videos(int n) {
    const int prev = dp_mixed_min_videos();
    g_mixed_min.store(n < 0 ? -1 : n, std::memory_order_relaxed);
    return prev;
}
bool dp_pair_enabled_for(int videos) {
    const int m = pair_min_videos();
    return m >= 0 && videos >= m;
}
bool dp_pair_used(int C, int L, int mode, bool sparse, bool xp, int videos) {
    return lin_enabled() && dp_pair_enabled_for(videos) && dp_pair_eligible(C, L, sparse, mode == 0 ? false : xp);
}
bool dp_lin_used(int C, int L, int mode, bool sparse, bool xp) {
    if (!lin_enabled()) return false;
    if (dp_pair_eligible(C, L, sparse, mode == 0 ? false : xp)) return true;  // pair-eligible shapes are lin-eligible too
    return mode == 0 ? dp_vit2_eligible(C, L, sparse) : dp_lin_eligible(C, L, mode, sparse, xp);
}

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
    p.only_flagged = 0;
    if (mode == 0) {
        if (dp_lin_used(p.C, p.L, 0, p.trans_pred != nullptr, false)) {
            // deferred-arg-max kernel; behind it the generic kernel decodes the videos whose sparse hint was degenerate
            const int rc = dp_pair_used(p.C, p.L, 0, p.trans_pred != nullptr, false, p.B) ? dp_pair_launch(p, 0, st) : dp_vit2_launch(p, st);
            if (rc || p.trans_pred == nullptr) return rc;
            p.only_flagged = 1;
        }
        return dp_launch_vit(p, st);
    }
    const bool sparse = (mode == 2) ? (p.trans_succ != nullptr) : (p.trans_pred != nullptr);
    if (dp_lin_used(p.C, p.L, mode, sparse, p.xp != 0)) {
        // linear-window kernel first; the log-domain kernel behind it recomputes the videos it flagged
        const int rc = dp_pair_used(p.C, p.L, mode, sparse, p.xp != 0, p.B) ? dp_pair_launch(p, mode, st)
                       : (mode == 1) ? dp_lin_launch_fwd(p, st) : dp_lin_launch_bwd(p, st);
        if (rc) return rc;
        p.only_flagged = 1;
    }
    if (mode == 1) return p.xp ? dp_launch_fwd_xp(p, st) : dp_launch_fwd(p, st);
    return p.xp ? dp_launch_bwd_xp(p, st) : dp_launch_bwd(p, st);
}

}  // namespace hsmm
