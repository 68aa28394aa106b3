// Grouped launches of the CrossTask-envelope DP kernels: ONE kernel per kernel family over the batches of several tasks.
//
// A step of the unsupervised trainer touches 18 task-homogeneous batches (data/corpus.py:633-636), each its own class set
// and parameters.  Launched one by one, every DP kernel runs ~32 CTAs for as long as ITS longest video lasts -- the GPU is
// never full and every task pays that latency three times.  Grouped, the same kernels see all the videos of the step at
// once (parameter blocks indexed by blockIdx, hsmm_common.cuh: DpGroup): a launch lasts as long as the longest video of
// the STEP and the SMs stay full while shorter videos retire.
//
// Envelope (what the chain-constrained CrossTask models need; anything else goes through the per-task entry points):
// sparse transition lists, L = K-1 <= 20, C <= 32, one precision (float state or HSMM_FLAG_F64_STATE) for the group.
#include "hsmm_dp_pair.cuh"
#include "hsmm_dp_vit2.cuh"

namespace hsmm {

bool dp_lin_enabled();
bool dp_pair_enabled_for(int videos);
int dp_mixed_min_videos();

// Forward and backward pass of a video back to back in ONE launch: a video's backward pass starts when ITS forward pass
// ends instead of when the longest video of the launch has finished its forward pass, so the SMs stay full while the
// long videos are still going forward (the saved planes it reads back were written by the same warp: L2-coherent loads).
// A video the forward pass flags (fflag) is skipped by the backward body and left, with its backward pass, to the
// log-domain kernels launched behind.
template <bool XP, int MINB>
__global__ void __launch_bounds__(128, MINB) dp_lin_fb_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_lin_forward_kernel_body<XP, 20, 1, 2>(g.t[t], local);
    __syncwarp();
    dp_lin_backward_kernel_body<XP, 20, 1, 2>(g.t[t], local);
}
__global__ void __launch_bounds__(128, 4) dp_pair_fb_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    dp_pair_forward_kernel_body<PAIR_KR>(g.t[t], local);
    __syncwarp();
    dp_pair_backward_kernel_body<PAIR_KR>(g.t[t], local);
}

// Both families in ONE launch (float state): tasks with C <= 16 run two videos per warp (VPB == 8 marks them), the others
// one video per warp.  Two launches would run one after the other on the call's stream, each as long as its longest
// video; one launch has a third fewer warps for the same videos (configs[1]: 1664 instead of 2304), so every scheduler
// juggles ~3 instead of ~4 of these latency-bound instruction chains.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) dp_mixed_fb_kernel_grouped(const __grid_constant__ DpGroup g) {
    int local;
    const int t = group_find(g, blockIdx.x, local);
    if (g.t[t].VPB == 8) {
        dp_pair_forward_kernel_body<PAIR_KR>(g.t[t], local);
        __syncwarp();
        dp_pair_backward_kernel_body<PAIR_KR>(g.t[t], local);
    } else {
        dp_lin_forward_kernel_body<false, 20, 1, 2>(g.t[t], local);
        __syncwarp();
        dp_lin_backward_kernel_body<false, 20, 1, 2>(g.t[t], local);
    }
}
// (The same mix for Viterbi was measured, r02q: alone it is slower -- 1.96 instead of 1.63 ms for the 2304 videos of
// configs[1], the two-videos-per-warp chain is longer per frame and a lone Viterbi launch is latency-bound -- and beside
// the forward+backward launch it gains 1 %: Viterbi groups stay on one kernel family.)

// per-task videos per block: 8 for the tasks in `pair8` (may be null), else 4
static int fill_group_mixed(DpGroup& g, const DpParams* ps, int n, const bool* pair8) {
    g.n = n;
    int first = 0;
    for (int i = 0; i < n; ++i) {
        const int vpb = (pair8 && pair8[i]) ? 8 : 4;
        g.t[i] = ps[i];
        g.t[i].W = 1;
        g.t[i].VPB = vpb;
        g.t[i].only_flagged = 0;
        g.first[i] = first;
        first += (ps[i].B + vpb - 1) / vpb;
    }
    g.first[n] = first;
    return first;
}

static int fill_group(DpGroup& g, const DpParams* ps, const int* idx, int n, int videos_per_block, int only_flagged) {
    g.n = n;
    int first = 0;
    for (int i = 0; i < n; ++i) {
        g.t[i] = ps[idx[i]];
        g.t[i].W = 1;
        g.t[i].VPB = 4;
        g.t[i].only_flagged = only_flagged;
        g.first[i] = first;
        first += (ps[idx[i]].B + videos_per_block - 1) / videos_per_block;
    }
    g.first[n] = first;
    return first;
}

template <bool XP>
static int launch_generic(DpGroup& g, int blocks, int mode, size_t smem, cudaStream_t st) {
    constexpr int MAXT = kMaxThreadsSmall1;
    if (mode == 0)
        dp_forward_kernel_grouped<true, false, 20, 1, 2, true, MAXT><<<blocks, 128, smem, st>>>(g);
    else if (mode == 1)
        dp_forward_kernel_grouped<false, XP, 20, 1, 2, true, MAXT><<<blocks, 128, smem, st>>>(g);
    else
        dp_backward_kernel_grouped<XP, 20, 1, 2, true, MAXT><<<blocks, 128, smem, st>>>(g);
    return check_launch("grouped log-domain DP kernel");
}

template <bool XP>
static int launch_lin_group(DpGroup& g, int blocks, int mode, cudaStream_t st) {
    const size_t smem = 4 * (2 * 32 + 2) * sizeof(float);
    if (mode == 3) {
        // 4 CTAs per SM for the float-state pair (128 registers); forcing 5 spills the windows and doubles the time (r02p A/B)
        if constexpr (XP)
            dp_lin_fb_kernel_grouped<XP, 3><<<blocks, 128, smem, st>>>(g);
        else if (blocks <= 3 * dp_num_sms())
            dp_lin_fb_kernel_grouped<false, 3><<<blocks, 128, smem, st>>>(g);
        else
            dp_lin_fb_kernel_grouped<false, 4><<<blocks, 128, smem, st>>>(g);
    } else if (mode == 0)
        dp_vit2_kernel_grouped<20, 1, 2><<<blocks, 128, smem, st>>>(g);
    else if (mode == 1)
        dp_lin_forward_kernel_grouped<XP, 20, 1, 2><<<blocks, 128, smem, st>>>(g);
    else
        dp_lin_backward_kernel_grouped<XP, 20, 1, 2><<<blocks, 128, smem, st>>>(g);
    return check_launch("grouped linear-window DP kernel");
}

static int launch_pair_group(DpGroup& g, int blocks, int mode, cudaStream_t st) {
    if (mode == 3)
        dp_pair_fb_kernel_grouped<<<blocks, 128, 0, st>>>(g);
    else if (mode == 0)
        dp_pair_vit_kernel_grouped<PAIR_KR><<<blocks, 128, 0, st>>>(g);
    else if (mode == 1)
        dp_pair_forward_kernel_grouped<PAIR_KR><<<blocks, 128, 0, st>>>(g);
    else
        dp_pair_backward_kernel_grouped<PAIR_KR><<<blocks, 128, 0, st>>>(g);
    return check_launch("grouped two-videos-per-warp DP kernel");
}

int dp_group_launch(const DpParams* ps, int n, int mode, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n > GROUP_MAX) {
        set_error("grouped DP launch: at most %d tasks per call (got %d)", GROUP_MAX, n);
        return -2;
    }
    const bool xp = mode != 0 && ps[0].xp != 0;
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
        const DpParams& p = ps[i];
        const int32_t* list = mode == 2 ? p.trans_succ : p.trans_pred;
        if (mode == 3 && !p.trans_succ) list = nullptr;
        if (!list || p.L > 20 || p.C > 32 || (mode != 0 && (p.xp != 0) != xp)) {
            set_error("grouped DP launch: task %d is outside the envelope (sparse transition lists, K-1 <= 20, C <= 32, one precision): "
                      "C=%d K=%d list=%p", i, p.C, p.L + 1, (const void*)list);
            return -2;
        }
        const size_t sm = smem_bytes(kVariants[1], p.C, 1, 4, 2, mode == 3 ? 2 : mode, xp);
        if (sm > smem) smem = sm;
    }
    static thread_local DpGroup g;  // 9.6 KB: kept off the stack; the launch copies it
    int all[GROUP_MAX], small[GROUP_MAX], large[GROUP_MAX];
    int ns = 0, nl = 0;
    const bool lin = dp_lin_enabled();
    int videos = 0;
    for (int i = 0; i < n; ++i) videos += ps[i].B;
    const bool pair = dp_pair_enabled_for(videos);  // throughput regime only: the two kernels run one after the other
    for (int i = 0; i < n; ++i) {
        all[i] = i;
        if (lin && pair && pair_shape_ok(ps[i].C, ps[i].L, true, xp))
            small[ns++] = i;
        else
            large[nl++] = i;
    }
    int rc = 0;
    // mixed launch: float state, forward+backward, enough videos for the launch to be issue-bound, and at least one task
    // that can run two videos per warp (otherwise the one-video kernel below is the same thing)
    bool mixed = false;
    const int mixed_min = dp_mixed_min_videos();
    if (lin && !xp && !pair && mode == 3 && mixed_min >= 0 && videos >= mixed_min) {
        bool p8[GROUP_MAX];
        int n8 = 0;
        for (int i = 0; i < n; ++i) n8 += (p8[i] = pair_shape_ok(ps[i].C, ps[i].L, true, false));
        if (n8 > 0) {
            const int blocks = fill_group_mixed(g, ps, n, p8);
            const size_t sm_lin = 4 * (2 * 32 + 2) * sizeof(float);
            // up to 3 x SMs CTAs: compiled for three CTAs per SM (152 instead of 128 registers: the backward loops lose a
            // tenth of their instructions; r02u: the call 2.52 -> 2.15 ms, the configs[1] step 4.52 -> 4.20 ms)
            if (blocks <= 3 * dp_num_sms())
                dp_mixed_fb_kernel_grouped<3><<<blocks, 128, sm_lin, st>>>(g);
            else
                dp_mixed_fb_kernel_grouped<4><<<blocks, 128, sm_lin, st>>>(g);
            rc = check_launch("grouped mixed-family DP kernel");
            if (rc) return rc;
            mixed = true;
        }
    }
    if (lin && !mixed) {
        if (ns) {
            const int blocks = fill_group(g, ps, small, ns, 8, 0);
            rc = launch_pair_group(g, blocks, mode, st);
            if (rc) return rc;
        }
        if (nl) {
            const int blocks = fill_group(g, ps, large, nl, 4, 0);
            rc = xp ? launch_lin_group<true>(g, blocks, mode, st) : launch_lin_group<false>(g, blocks, mode, st);
            if (rc) return rc;
        }
    }
    // the log-domain kernels: the whole job when the linear-window kernels are switched off, otherwise only the videos
    // those kernels flagged (their warps exit at once for every other video)
    const int blocks = fill_group(g, ps, all, n, 4, lin ? 1 : 0);
    if (mode == 3) {
        rc = xp ? launch_generic<true>(g, blocks, 1, smem, st) : launch_generic<false>(g, blocks, 1, smem, st);
        if (rc) return rc;
        return xp ? launch_generic<true>(g, blocks, 2, smem, st) : launch_generic<false>(g, blocks, 2, smem, st);
    }
    return xp ? launch_generic<true>(g, blocks, mode, smem, st) : launch_generic<false>(g, blocks, mode, smem, st);
}

}  // namespace hsmm
