#!/usr/bin/env python
"""Benchmark of the HSMM hot path on B200: video frames/s for emission scoring + log-semiring
forward/backward (logZ and all expected counts) + max-plus Viterbi on synthetic data of the shapes
BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W          # driver default: configs[1]; one rank per GPU under torchrun
    python bench.py --config {0,1,2,3,4} [--max-span K]    # the other BASELINE configs (table in DESIGN.md section 6)
    python bench.py --impl reference ...                   # the reference's own code on the host cores (oracle/_ref)

  configs[0]  S6-shape: one task, C = 11, D = 200, K = 100, T ~ U[1000, 3000], dense transitions
  configs[1]  U7-shape HSMM EM: 18 CrossTask-like tasks (133 steps, C = 2s+1 in 7..23), D = 200, K = 20, chain-constrained
  configs[2]  configs[1] + the -1e4 narration penalty tensor
  configs[3]  Breakfast-shape: C = 48, D = 64, T ~ U[500, 10000], K = 200 (--max-span 500 for the large span)
  configs[4]  decode-only sweep: 10 000 videos, T ~ U[500, 2000], D = 200, C in {16, 64, 133} x K in {50, 100, 200}

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, library kernels only.  `e2e`: the public module
API with HOST buffers -- the live frames of every batch are copied from pinned memory (hsmm_upload_ragged) and the
loss, spans and labels are read back inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# the step keeps ~36 independent kernel chains in flight (one or two streams per task): give the driver enough hardware
# work queues, or streams share a queue and serialise (the default of 8 caps the step at 8 concurrent kernels)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# number of steps of the 18 primary CrossTask tasks (sum = 133 step classes)
CROSSTASK_STEPS = [6, 5, 8, 11, 6, 6, 6, 11, 8, 11, 3, 7, 5, 8, 11, 5, 9, 7]
METRIC = "video frames/sec for HSMM fwd-bwd+Viterbi"
UNIT = "frames/s"
SWEEP_C, SWEEP_K = (16, 64, 133), (50, 100, 200)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4], help="index into BASELINE.json configs")
    ap.add_argument("--videos-per-task", type=int, default=None, help="videos of each task per step and per GPU")
    ap.add_argument("--max-span", type=int, default=None, help="--sm_max_span_length (default: the config's)")
    ap.add_argument("--feature-dim", type=int, default=None)
    ap.add_argument("--tmin", type=int, default=None)
    ap.add_argument("--tmax", type=int, default=None)
    ap.add_argument("--narration", action="store_true", help="same as --config 2")
    ap.add_argument("--sweep-videos", type=int, default=10000, help="configs[4]: videos per (C, K) cell and GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    if args.narration and args.config == 1:
        args.config = 2
    return args


def config_of(args):
    """Shape parameters of the chosen BASELINE config (command-line overrides applied)."""
    c = args.config
    if c == 0:
        cfg = dict(tasks=[11], chain=False, K=100, D=200, tmin=1000, tmax=3000, V=512, narration=False)
    elif c in (1, 2):
        cfg = dict(tasks=[2 * s + 1 for s in CROSSTASK_STEPS], chain=True, K=20, D=200, tmin=1000, tmax=3000, V=128,
                   narration=(c == 2))
    elif c == 3:
        cfg = dict(tasks=[48], chain=False, K=200, D=64, tmin=500, tmax=10000, V=256, narration=False)
    else:
        cfg = dict(tasks=[16], chain=False, K=50, D=200, tmin=500, tmax=2000, V=1000, narration=False)
    for key, val in (("K", args.max_span), ("D", args.feature_dim), ("tmin", args.tmin), ("tmax", args.tmax),
                     ("V", args.videos_per_task)):
        if val is not None:
            cfg[key] = val
    return cfg


def workload_name(args, cfg):
    c = args.config
    if c == 0:
        return "configs[0] S6-shape: one task, C=11 (10 steps + background) + EOS, dense transitions, D=%d, K=%d, T~U[%d,%d], " \
               "%d videos/GPU" % (cfg["D"], cfg["K"], cfg["tmin"], cfg["tmax"], cfg["V"])
    if c in (1, 2):
        return "configs[1] U7-shape HSMM EM: 18 CrossTask-like tasks (133 steps, C=2s+1 in 7..23 + EOS, chain-constrained " \
               "transitions), D=%d, K=%d, T~U[%d,%d], %d videos/task/GPU%s" % (
                   cfg["D"], cfg["K"], cfg["tmin"], cfg["tmax"], cfg["V"],
                   ", narration penalty -1e4 (configs[2])" if cfg["narration"] else "")
    if c == 3:
        return "configs[3] Breakfast-shape: C=48 + EOS, dense transitions, D=%d, K=%d, T~U[%d,%d], %d videos/GPU" % (
            cfg["D"], cfg["K"], cfg["tmin"], cfg["tmax"], cfg["V"])
    return "configs[4] decode-only sweep: %d videos/GPU per cell, T~U[%d,%d], D=%d, C in %s x K in %s, dense transitions" % (
        args.sweep_videos, cfg["tmin"], cfg["tmax"], cfg["D"], list(SWEEP_C), list(SWEEP_K))


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
class Task:
    pass


def chain_masks(C):
    """Ordering constraint of data/crosstask.py:328-388 + self loops (semimarkov.py:50-54): [to, from]."""
    mask = torch.ones(C, C, dtype=torch.bool)
    for c in range(C):
        mask[c, c] = False
        if c + 1 < C:
            mask[c + 1, c] = False
    init_mask = torch.ones(C, dtype=torch.bool)
    init_mask[0] = False
    return mask, init_mask


def make_params(tk, C, K, D, chain, gen, device):
    """Model parameters of one class set and the score tensors the DP consumes (parameter-only work)."""
    from action_segmentation_b200 import hsmm
    tk.C, tk.K, tk.D, tk.chain = C, K, D, chain
    means = torch.randn(C, D, generator=gen) * 0.35
    if chain:
        means[0::2] = means[0]  # merged backgrounds (--annotate_background_with_previous)
    tk.means = means.to(device)
    tk.cov_diag = (torch.rand(D, generator=gen) + 0.5).to(device)
    tk.trans_logits = torch.randn(C, C, generator=gen) * 0.1
    tk.init_logits = torch.rand(C, generator=gen)
    tk.log_rates = torch.log(torch.rand(C, generator=gen) * 8 + 4) if K <= 20 else \
        torch.log(torch.rand(C, generator=gen) * (K / 4.0) + K / 8.0)
    if chain:
        tmask, imask = chain_masks(C)
        tk.trans_mask, tk.init_mask = tmask, imask
        trans = torch.log_softmax(tk.trans_logits.masked_fill(tmask, -1e9), dim=0)
        init = torch.log_softmax(tk.init_logits.masked_fill(imask, -1e9), dim=0)
        tk.pred, tk.succ = hsmm.sparse_transition_lists(~tmask, torch.device(device))
    else:
        tk.trans_mask = tk.init_mask = None
        trans = torch.log_softmax(tk.trans_logits, dim=0)
        init = torch.log_softmax(tk.init_logits, dim=0)
        tk.pred = tk.succ = None
    k = torch.arange(K, dtype=torch.float32).unsqueeze(-1)
    lenp = k * tk.log_rates.unsqueeze(0) - torch.exp(tk.log_rates).unsqueeze(0) - torch.lgamma(k + 1)
    tk.trans, tk.init, tk.lenp = trans.to(device), init.to(device), lenp.to(device)
    tk.class_ids = torch.arange(C + 1, device=device, dtype=torch.int32)
    tk.eparams = hsmm.emission_params(tk.means, tk.cov_diag)  # w, bias, 1/var, row constant


def make_task(C, cfg, gen, device, V=None, X=None, lengths=None):
    from action_segmentation_b200 import hsmm
    tk = Task()
    V = V or cfg["V"]
    tk.V = V
    make_params(tk, C, cfg["K"], cfg["D"], cfg["chain"], gen, device)
    D = cfg["D"]
    if cfg["chain"]:
        end = torch.full((V, C), -1e9)
        end[:, C - 1] = 0
        tk.end = end.to(device)
    else:
        tk.end = None
    tk.lengths = lengths if lengths is not None else torch.randint(cfg["tmin"], cfg["tmax"] + 1, (V,), generator=gen)
    if lengths is None:
        tk.lengths[0] = cfg["tmax"]
        if os.environ.get("HSMM_BENCH_BUCKETS"):
            tk.lengths = torch.sort(tk.lengths, descending=True)[0]  # experiment: length buckets are contiguous slices
    Tmax = int(tk.lengths.max())
    tk.Tmax = Tmax
    tk.penalty = None
    if X is not None:
        tk.X = X
    else:
        # labels: the classes in order (cyclic when there are more segments than classes), random cut points
        nseg = C if cfg["chain"] else max(C, 24)
        cuts = torch.sort((torch.rand(V, nseg - 1, generator=gen) * (tk.lengths[:, None] - 1)).long() + 1, dim=1)[0].to(device)
        pos = torch.arange(Tmax, device=device).unsqueeze(0).expand(V, Tmax).contiguous()
        labels = torch.searchsorted(cuts, pos, right=True) % C
        dgen = torch.Generator(device=device).manual_seed(int(torch.randint(0, 2 ** 31, (1,), generator=gen)))
        X = torch.randn(V, Tmax, D, device=device, generator=dgen)
        X += tk.means[labels]
        live = (pos < tk.lengths.to(device)[:, None])
        X *= live.unsqueeze(-1)
        tk.X = X.contiguous()
        if cfg["narration"]:
            # every step gets one window around its true span; outside it the step costs -1e4 per frame
            pen = torch.zeros(V, Tmax, C, device=device)
            for j in range(1, C, 2):
                inside = (labels == j)
                lo = torch.where(inside, pos, Tmax).min(dim=1)[0] - 40
                hi = torch.where(inside, pos, -1).max(dim=1)[0] + 40
                allowed = (pos >= lo[:, None]) & (pos <= hi[:, None])
                pen[:, :, j] = (~allowed).float() * -1e4
            tk.penalty = pen
    tk.lengths_i32, tk.order = hsmm.prepare_lengths(tk.lengths, torch.device(device))
    tk.frames = int(tk.lengths.sum())
    tk.gradw = torch.full((V,), 1.0 / V, device=device)
    return tk


def make_workload(args, cfg, rank, device):
    gen = torch.Generator().manual_seed(args.seed + rank)
    classes = cfg["tasks"]
    sel = os.environ.get("HSMM_BENCH_TASKS")  # experiment switch (echoed in config.env_switches): "small" = C <= 16, "large" = C > 16
    if sel:
        classes = [C for C in classes if (C <= 16) == (sel == "small")]
    return [make_task(C, cfg, gen, device) for C in classes]


def step_weights(tk, world):
    """d loss / d logZ_b of the step: 1 / (videos of the task over all ranks); built once, not inside the timed step."""
    cache = tk.__dict__.setdefault("_gradw_world", {})
    if world not in cache:
        cache[world] = tk.gradw if world == 1 else (tk.gradw / world).contiguous()
    return cache[world]


def packed_layout(tasks):
    """Offsets of every task's [d_means | d_trans | d_len | d_init | sum logZ] slice of the packed buffer."""
    off, lay = 0, []
    for tk in tasks:
        sizes = [tk.C * tk.D, tk.C * tk.C, tk.K * tk.C, tk.C, tk.C, 1]  # wx, d_trans, d_len, d_init, wsum, logz
        lay.append((off, sizes))
        off += sum(sizes)
    return lay, off


def env_switches():
    return {k: os.environ[k] for k in ("HSMM_BENCH_SKIP", "HSMM_BENCH_TASKS", "HSMM_BENCH_BUCKETS", "HSMM_BENCH_GROUPS", "HSMM_BENCH_FUSED", "HSMM_BENCH_ORDERED_GROUPS", "HSMM_EMISSION_TWO_CTAS", "HSMM_BENCH_COPY_STREAMS", "HSMM_BENCH_UPLOAD", "HSMM_UPLOAD_CTAS", "HSMM_BENCH_PRIO", "HSMM_BENCH_VIT_LAST", "HSMM_BENCH_NO_OVERLAP", "HSMM_BENCH_NO_REDUCE", "HSMM_DISABLE_LIN", "HSMM_DISABLE_PAIR", "HSMM_PAIR_MIN_VIDEOS",
                                              "HSMM_FORCE_GENERIC") if os.environ.get(k)}


# ---------------------------------------------------------------------------------------------
# one step of the hot path, inputs resident in HBM, library kernels only
# ---------------------------------------------------------------------------------------------
def device_step(tasks, streams, packed, layout, world, reduce=True, decode_only=False):
    """One pass of the hot path over every task's batch.  Per task: emission scoring -> {forward -> backward ->
    class-weighted feature sums} on the task's stream and, concurrently on a second stream, Viterbi (which
    needs the emission scores only).  `decode_only` (configs[4]): emission + Viterbi."""
    from action_segmentation_b200 import hsmm
    lib = hsmm._lib.load()
    skip = set(filter(None, os.environ.get("HSMM_BENCH_SKIP", "").split(",")))  # ablation only: marked invalid in the line
    cur = torch.cuda.current_stream()
    if not decode_only:
        packed.zero_()
    fork = torch.cuda.Event()
    fork.record(cur)
    n, ns = len(tasks), len(streams)
    outs = []
    used, forked = [], set()
    for i, tk in enumerate(tasks):
        st, st2 = streams[i % ns], streams[(n + i) % ns]
        for x in (st, st2):
            if id(x) not in forked:
                forked.add(id(x))
                used.append(x)
                x.wait_event(fork)
        with torch.cuda.stream(st):
            if "em" in skip and getattr(tk, "em_cache", None) is not None:
                em, rowterm, offset = tk.em_cache
            else:
                em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32,
                                                           params=tk.eparams)
                if "em" in skip:
                    tk.em_cache = (em, rowterm, offset)
            em_ready = torch.cuda.Event()
            em_ready.record(st)
        st2.wait_event(em_ready)
        with torch.cuda.stream(st2):
            if "vit" in skip:
                outs.append((None, None, em, offset))
            else:
                spans, labels, score = hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32,
                                                           tk.order, tk.class_ids, want_labels=True, want_score=False,
                                                           trans_pred=tk.pred)
                outs.append((spans, labels, em, offset))
        if decode_only:
            continue
        off, sizes = layout[i]
        v, o = [], off
        for m in sizes:
            v.append(packed[o:o + m])
            o += m
        wx, d_trans, d_len, d_init, wsum, lz = v
        nb = int(os.environ.get("HSMM_BENCH_BUCKETS", "0"))
        xp = tk.penalty is not None
        g = step_weights(tk, world)
        if nb > 1:
            # experiment: the task's forward/backward in nb length-homogeneous sub-batches, each its own chain
            ldc = em.shape[2]
            with torch.cuda.stream(st):
                d_em = torch.empty(tk.V, tk.Tmax, ldc, device=em.device)
            lz_parts, evs = [], []
            step_v = (tk.V + nb - 1) // nb
            for j in range(nb):
                a, b = j * step_v, min(tk.V, (j + 1) * step_v)
                if a >= b:
                    continue
                sj = streams[(2 * n + i * nb + j) % ns] if j else st
                if id(sj) not in forked:
                    forked.add(id(sj))
                    used.append(sj)
                    sj.wait_event(fork)
                sj.wait_event(em_ready)
                with torch.cuda.stream(sj):
                    sl = slice(a, b)
                    lzj, savedj = hsmm.logz_forward(em[sl], tk.C, tk.init, tk.trans, tk.lenp, None if tk.end is None else tk.end[sl],
                                                    offset[sl], tk.lengths_i32[sl], None, trans_pred=tk.pred, f64_state=xp)
                    hsmm.logz_backward(em[sl], tk.C, tk.init, tk.trans, tk.lenp, None if tk.end is None else tk.end[sl],
                                       tk.lengths_i32[sl], None, g[sl], savedj, out=(d_init, d_trans.view(tk.C, tk.C), d_len.view(tk.K, tk.C)),
                                       trans_succ=tk.succ, f64_state=xp, d_em=d_em[sl])
                    lz_parts.append(lzj.sum())
                    ev = torch.cuda.Event()
                    ev.record(sj)
                    evs.append(ev)
            with torch.cuda.stream(st):
                for ev in evs:
                    st.wait_event(ev)
                hsmm._lib.check(lib.hsmm_weighted_feature_sums(hsmm._p(tk.X), hsmm._p(d_em), d_em.shape[2],
                                                               hsmm._p(tk.lengths_i32), tk.V, tk.Tmax, tk.D, tk.C, hsmm._p(wx),
                                                               hsmm._p(wsum), hsmm._stream()), "hsmm_weighted_feature_sums")
                lz.copy_(torch.stack(lz_parts).sum().float().reshape(1))
            continue
        with torch.cuda.stream(st):
            logz, saved = hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order,
                                            trans_pred=tk.pred, f64_state=xp)
            _, _, _, d_em = hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, tk.lengths_i32, tk.order, g,
                                               saved, out=(d_init, d_trans.view(tk.C, tk.C), d_len.view(tk.K, tk.C)),
                                               trans_succ=tk.succ, f64_state=xp)
            if "ws" not in skip:
                hsmm._lib.check(lib.hsmm_weighted_feature_sums(hsmm._p(tk.X), hsmm._p(d_em), d_em.shape[2],
                                                               hsmm._p(tk.lengths_i32), tk.V, tk.Tmax, tk.D, tk.C, hsmm._p(wx),
                                                               hsmm._p(wsum), hsmm._stream()), "hsmm_weighted_feature_sums")
            lz.copy_(logz.sum().float().reshape(1))
    for st in used:  # join only the streams that were forked off this step (a graph capture rejects anything else)
        ev = torch.cuda.Event()
        ev.record(st)
        cur.wait_event(ev)
    if world > 1 and reduce and not decode_only:
        torch.distributed.all_reduce(packed)
    return outs


def make_streams(n_tasks):
    """The step's streams: one per task (emission, weighted sums), then the DP streams of the task groups.
    HSMM_BENCH_PRIO=1 (experiment) gives the streams behind the per-task ones a higher priority."""
    n = min(2 * n_tasks, 36) + 1 + 18 * int(os.environ.get("HSMM_BENCH_BUCKETS", "0"))
    prio = os.environ.get("HSMM_BENCH_PRIO", "0") == "1"
    return [torch.cuda.Stream(priority=-1 if (prio and i >= n_tasks) else 0) for i in range(n)]


def grouped_eligible(tasks):
    """hsmm_dp_grouped's envelope: sparse transition lists, K - 1 <= 20, C <= 32, one precision."""
    return all(tk.chain and tk.K - 1 <= 20 and tk.C <= 32 for tk in tasks) and len({tk.penalty is not None for tk in tasks}) == 1


def device_step_grouped(tasks, streams, packed, layout, world, reduce=True, n_groups=1, fused=True):
    """The same work as `device_step`, with the DP kernels of several tasks in ONE launch per kernel family
    (hsmm_dp_grouped): per task emission scoring on the task's stream; per group of tasks {forward -> backward} on the
    group's stream and Viterbi on a second one; per task the class-weighted feature sums once its group's backward
    pass is done."""
    from action_segmentation_b200 import hsmm
    lib = hsmm._lib.load()
    cur = torch.cuda.current_stream()
    packed.zero_()
    fork = torch.cuda.Event()
    fork.record(cur)
    n, ns = len(tasks), len(streams)
    assert ns >= n + 2 * n_groups
    em_out, em_ev = [None] * n, [None] * n
    # HSMM_BENCH_ORDERED_GROUPS=1 (experiment): the emission launches of group g+1 wait for those of group g, so that they
    # run under group g's DP instead of beside group g's emission
    ordered = os.environ.get("HSMM_BENCH_ORDERED_GROUPS", "0") == "1" and n_groups > 1
    prev_evs = []
    for gi in range(n_groups if ordered else 1):
        mine = list(range(gi, n, n_groups)) if ordered else list(range(n))
        for i in mine:
            tk, st = tasks[i], streams[i]
            st.wait_event(fork)
            for ev in prev_evs:
                st.wait_event(ev)
            with torch.cuda.stream(st):
                em_out[i] = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32, params=tk.eparams)
                ev = torch.cuda.Event()
                ev.record(st)
                em_ev[i] = ev
        prev_evs = [em_ev[i] for i in mine]
    xp = tasks[0].penalty is not None
    outs = [None] * n
    used = list(streams[:n])
    for gi in range(n_groups):
        idx = list(range(gi, n, n_groups))  # interleaved: every group gets small and large class sets
        s_dp, s_vit = streams[n + gi], streams[n + n_groups + gi]
        used += [s_dp, s_vit]
        base = []
        for i in idx:
            tk = tasks[i]
            em, rowterm, offset = em_out[i]
            base.append(dict(em=em, C=tk.C, init=tk.init, trans=tk.trans, lenp=tk.lenp, end=tk.end, offset=offset,
                             lengths_i32=tk.lengths_i32, order=tk.order, f64_state=xp))
        for s_x in (s_dp, s_vit):
            s_x.wait_event(fork)
            for i in idx:
                s_x.wait_event(em_ev[i])
        def decode():
            with torch.cuda.stream(s_vit):
                res = hsmm.grouped_dp(0, [dict(b, trans_list=tasks[i].pred, class_ids=tasks[i].class_ids) for b, i in zip(base, idx)])
                for i, (spans, labels, _) in zip(idx, res):
                    outs[i] = (spans, labels, em_out[i][0], em_out[i][2])

        vit_last = os.environ.get("HSMM_BENCH_VIT_LAST", "0") == "1"
        if not vit_last:
            decode()
        with torch.cuda.stream(s_dp):
            fb_in = []
            for b, i in zip(base, idx):
                tk = tasks[i]
                off, sizes = layout[i]
                v, o = [], off
                for m in sizes:
                    v.append(packed[o:o + m])
                    o += m
                wx, d_trans, d_len, d_init, wsum, lz = v
                g = step_weights(tk, world)
                fb_in.append(dict(b, trans_list=tk.pred, trans_list2=tk.succ, grad=g,
                                  out=(d_init, d_trans.view(tk.C, tk.C), d_len.view(tk.K, tk.C))))
            if fused:  # forward and backward of every video back to back in one launch
                fb = hsmm.grouped_dp(3, fb_in)
                fw = [(r[0], r[1]) for r in fb]
                bw = [r[2:] for r in fb]
            else:
                fw = hsmm.grouped_dp(1, fb_in)
                bw = hsmm.grouped_dp(2, [dict(b, trans_list=b["trans_list2"], saved=f[1]) for b, f in zip(fb_in, fw)])
            bwd_done = torch.cuda.Event()
            bwd_done.record(s_dp)
        if vit_last:
            decode()
        for i, (logz, saved), (_, _, _, d_em) in zip(idx, fw, bw):
            tk = tasks[i]
            off, sizes = layout[i]
            wx = packed[off:off + sizes[0]]
            o_ws = off + sum(sizes[:4])
            wsum = packed[o_ws:o_ws + sizes[4]]
            lz = packed[o_ws + sizes[4]:o_ws + sizes[4] + 1]
            st = streams[i]
            st.wait_event(bwd_done)
            with torch.cuda.stream(st):
                hsmm._lib.check(lib.hsmm_weighted_feature_sums(hsmm._p(tk.X), hsmm._p(d_em), d_em.shape[2],
                                                               hsmm._p(tk.lengths_i32), tk.V, tk.Tmax, tk.D, tk.C, hsmm._p(wx),
                                                               hsmm._p(wsum), hsmm._stream()), "hsmm_weighted_feature_sums")
                lz.copy_(logz.sum().float().reshape(1))
    for st in used:
        ev = torch.cuda.Event()
        ev.record(st)
        cur.wait_event(ev)
    if world > 1 and reduce:
        torch.distributed.all_reduce(packed)
    return outs


# ---------------------------------------------------------------------------------------------
# end to end through the public module API, HOST buffers
# ---------------------------------------------------------------------------------------------
class HostBatch:
    """Pinned host copy of one task's padded batch (what `padding_colate` hands the wrapper) and TWO sets of device
    landing buffers (zeroed once: only live rows are ever copied), so that the next step's inputs can be uploaded while
    this step computes -- the input prefetch every training loop does."""

    def __init__(self, tk):
        self.features = tk.X.cpu().pin_memory()
        self.penalty = None if tk.penalty is None else tk.penalty.cpu().pin_memory()
        self.lengths_i32 = tk.lengths.to(torch.int32)
        self.dev_features = [torch.zeros_like(tk.X) for _ in range(2)]
        self.dev_penalty = [None if tk.penalty is None else torch.zeros_like(tk.penalty) for _ in range(2)]
        live = int(tk.lengths.sum())
        self.h2d_bytes = live * tk.D * 4 + (0 if tk.penalty is None else live * tk.C * 4)


def e2e_upload(tasks, host, copy_streams, slot):
    """Host -> device copy of one step's inputs (live frames only); returns the events to wait for.  The tasks alternate
    between the copy streams: a ragged upload is one ~1.6 MB copy per video, and back to back on ONE stream such copies
    reach 46-50 GB/s of the 55 GB/s a single large copy gets on this box (~3.8 us of set-up each, r02q)."""
    from action_segmentation_b200 import hsmm
    mode = os.environ.get("HSMM_BENCH_UPLOAD", "kernel")  # kernel: hsmm_upload_ragged_mapped; copy: hsmm_upload_ragged; hybrid
    for i, (tk, hb) in enumerate(zip(tasks, host)):
        by_kernel = mode == "kernel" or (mode == "hybrid" and i % 2 == 0)
        ld = tk.lengths_i32 if by_kernel else None
        with torch.cuda.stream(copy_streams[i % len(copy_streams)]):
            hsmm.upload_ragged(hb.features, hb.dev_features[slot], hb.lengths_i32, ld)
            if hb.penalty is not None:
                hsmm.upload_ragged(hb.penalty, hb.dev_penalty[slot], hb.lengths_i32, ld if tk.C % 4 == 0 else None)
    evs = []
    for cs in copy_streams:
        ev = torch.cuda.Event()
        ev.record(cs)
        evs.append(ev)
    return evs


def e2e_step(models, tasks, host, streams, copy_streams, slot, ready, decode_only=False):
    """Public API with host buffers.  The inputs of THIS step (landing-buffer set `slot`) were uploaded while the previous
    step computed (`ready` = their event); this call first enqueues the upload of the NEXT step's inputs into the other
    set, then runs the module API on this step's and reads loss + predictions back.  Every step therefore contains one
    full host -> device copy of a step's inputs and the device -> host read of its results."""
    nxt = e2e_upload(tasks, host, copy_streams, 1 - slot)
    lls, preds = [], []
    ns = len(streams)
    for i, (m, tk, hb) in enumerate(zip(models, tasks, host)):
        st = streams[i % ns]
        for ev in ready:
            st.wait_event(ev)
        with torch.cuda.stream(st):
            ends = None if tk.end is None else [[] for _ in range(tk.V)]
            feats, pen = hb.dev_features[slot], hb.dev_penalty[slot]
            if decode_only:
                spans, labels = m.viterbi(feats, tk.lengths, None, additional_allowed_ends_per_instance=ends,
                                          constraints=pen, return_labels=True, non_blocking=True)
            else:
                m.zero_grad()
                ll, _, spans, labels = m.log_likelihood_and_viterbi(feats, tk.lengths, None,
                                                                    additional_allowed_ends_per_instance=ends,
                                                                    constraints=pen, non_blocking=True)
                (-ll).backward()
                h = torch.empty((), dtype=torch.float32, pin_memory=True)
                h.copy_(ll.detach(), non_blocking=True)
                lls.append(h)
            preds.append((spans, labels))
    for st in streams:  # this step's results are in host memory; the next step's upload may still be in flight
        st.synchronize()
    return float(sum(float(h) for h in lls)), preds, nxt


def build_models(tasks):
    import action_segmentation_b200 as pkg
    from action_segmentation_b200.args import HsmmArgs
    models = []
    for tk in tasks:
        C = tk.C
        kw = {}
        if tk.chain:
            trans = {c: ({c, c + 1} if c + 1 < C else {c}) for c in range(C)}
            kw = dict(allowed_starts={0}, allowed_transitions=trans, allowed_ends={C - 1})
        m = pkg.SemiMarkovModule(HsmmArgs(sm_max_span_length=tk.K), C, tk.D, allow_self_transitions=True, **kw).cuda()
        with torch.no_grad():
            m.gaussian_means.copy_(tk.means)
            m.gaussian_cov.copy_(torch.diag(tk.cov_diag))
            m.transition_logits.copy_(tk.trans_logits)
            m.init_logits.copy_(tk.init_logits)
            m.poisson_log_rates.copy_(tk.log_rates)
        models.append(m)
    return models


def allreduce_all_gradients(models):
    """ONE all-reduce of every model's gradients, packed (the e2e leg of N > 1)."""
    grads = [p.grad for m in models for p in m.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    torch.distributed.all_reduce(flat)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    # The gradients were allocated on the per-task streams and are written here on the current one: the next step frees
    # them (zero_grad) on ITS stream, whose allocator may hand the block out again at once -- e.g. to an index tensor --
    # while these copies are still queued.  The reduced gradients are what the optimiser step consumes next anyway: wait.
    torch.cuda.current_stream().synchronize()


# ---------------------------------------------------------------------------------------------
# per-kernel durations (serialised on one stream, CUDA events) -> dominant kernel roofline
# ---------------------------------------------------------------------------------------------
def kernel_breakdown(tasks, reps=3, decode_only=False):
    """Per-kernel device time of one step with every launch serialised on one stream (CUDA events around each
    call; best of `reps` per launch, summed over the step's launches)."""
    from action_segmentation_b200 import hsmm
    names = ["emission", "viterbi"] if decode_only else ["emission", "logz_forward", "logz_backward", "weighted_feature_sums", "viterbi"]
    best = {}
    launches = {n: 0 for n in names}

    def timed(key, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        b.synchronize()
        best[key] = min(best.get(key, 1e30), a.elapsed_time(b))
        return r

    for rep in range(reps):
        for i, tk in enumerate(tasks):
            em, rowterm, offset = timed(("emission", i), lambda: hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty,
                                                                                         tk.lengths_i32, params=tk.eparams))
            if not decode_only:
                xp = tk.penalty is not None
                logz, saved = timed(("logz_forward", i), lambda: hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end,
                                                                                   offset, tk.lengths_i32, tk.order,
                                                                                   trans_pred=tk.pred, f64_state=xp))
                d = timed(("logz_backward", i), lambda: hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end,
                                                                           tk.lengths_i32, tk.order, tk.gradw, saved,
                                                                           trans_succ=tk.succ, f64_state=xp))
                timed(("weighted_feature_sums", i), lambda: hsmm.weighted_feature_sums(tk.X, d[3], tk.C, tk.lengths_i32))
                del saved, d
            timed(("viterbi", i), lambda: hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset,
                                                              tk.lengths_i32, tk.order, tk.class_ids, want_score=False,
                                                              trans_pred=tk.pred))
            if rep == 0:
                for n in names:
                    launches[n] += 1
    return {n: sum(v for (k, _), v in best.items() if k == n) for n in names}, launches


def kernel_breakdown_grouped(tasks, reps=3, fused=True):
    """As `kernel_breakdown` for the grouped step: emission and weighted sums per task, each DP pass as ONE grouped call.
    Every family is captured into its own CUDA graph after an eager warm-up and the graph's replays are timed with CUDA
    events on the launching stream: device time of the family's kernels, serialised, without the host time of the Python
    calls between them (which a pair of events around the eager calls would include)."""
    from action_segmentation_b200 import hsmm
    xp = tasks[0].penalty is not None
    keep = {}

    def f_emission():
        keep["ems"] = [hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32, params=tk.eparams) for tk in tasks]
        keep["base"] = [dict(em=e[0], C=tk.C, init=tk.init, trans=tk.trans, lenp=tk.lenp, end=tk.end, offset=e[2],
                             lengths_i32=tk.lengths_i32, order=tk.order, f64_state=xp) for tk, e in zip(tasks, keep["ems"])]

    def f_fb():
        fb = hsmm.grouped_dp(3, [dict(b, trans_list=tk.pred, trans_list2=tk.succ, grad=tk.gradw) for b, tk in zip(keep["base"], tasks)])
        keep["bw"] = [r[2:] for r in fb]

    def f_fwd():
        keep["fw"] = hsmm.grouped_dp(1, [dict(b, trans_list=tk.pred) for b, tk in zip(keep["base"], tasks)])

    def f_bwd():
        keep["bw"] = hsmm.grouped_dp(2, [dict(b, trans_list=tk.succ, saved=f[1], grad=tk.gradw)
                                         for b, tk, f in zip(keep["base"], tasks, keep["fw"])])

    def f_wsums():
        keep["ws"] = [hsmm.weighted_feature_sums(tk.X, r[3], tk.C, tk.lengths_i32) for tk, r in zip(tasks, keep["bw"])]

    def f_vit():
        keep["vit"] = hsmm.grouped_dp(0, [dict(b, trans_list=tk.pred, class_ids=tk.class_ids) for b, tk in zip(keep["base"], tasks)])

    fams = [("emission", f_emission, len(tasks))]
    fams += [("logz_forward_backward", f_fb, 1)] if fused else [("logz_forward", f_fwd, 1), ("logz_backward", f_bwd, 1)]
    fams += [("weighted_feature_sums", f_wsums, len(tasks)), ("viterbi", f_vit, 1)]
    kms, launches = {}, {}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for name, fn, n in fams:
            fn()  # eager warm-up (allocations, lazy module loads)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn()
            best = 1e30
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(side)
                g.replay()
                b.record(side)
                b.synchronize()
                best = min(best, a.elapsed_time(b))
            kms[name], launches[name] = best, n
            keep["graph_" + name] = g  # the later families read this one's outputs: keep its pool alive
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return kms, launches


def compute_ceiling(tasks, frames_per_s_per_gpu, sm_mhz, decode_only):
    """Secondary (compute) ceiling of SURVEY.md section 8(d): C*(L + C') semiring operations per frame and pass
    (C' = 2 for the chain-constrained transition lists, else C; Viterbi 1 pass, forward 1, backward + counts 2) against
    the FP32 lane-operation peak 148 SMs x 128 lanes x SM clock."""
    tot_frames = float(sum(tk.frames for tk in tasks))
    passes = 1 if decode_only else 4
    ops = sum(tk.frames * tk.C * ((tk.K - 1) + (2 if tk.chain else tk.C)) for tk in tasks) / tot_frames * passes
    clock = (sm_mhz or 1965.0) * 1e6
    peak = 148 * 128 * clock
    return {"bound": "fp32_issue", "semiring_ops_per_frame": ops, "peak_lane_ops_per_s": peak, "sm_mhz_used": clock / 1e6,
            "achieved_lane_ops_per_s": frames_per_s_per_gpu * ops, "frac": frames_per_s_per_gpu * ops / peak,
            "note": "algorithmic operation count; every semiring operation costs several instructions (add, max/ex2, fma)"}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        rows = [r for r in self.rows[max(0, lo - 1):hi] if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()][-3:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(int(r[0]) for r in rows), "sm_max_mhz": int(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref, imported unmodified over oracle/torch_struct_shim.py)
# ---------------------------------------------------------------------------------------------
def cpu_sample(args, cfg):
    """Bounded sample of the config's workload for the CPU legs -- the SAME sample in `cpu_baseline` and in
    `--impl reference`.  The reference materialises (B, T, K, C+1, C+1) float32 potentials and autograd keeps
    about three copies, so T (and for configs[3]/[4] the shape itself) is reduced where that would not fit in
    host memory; the reduction is stated in `sample`."""
    gen = torch.Generator().manual_seed(args.seed)
    c = args.config
    reduced = ""
    if c == 0:
        C, K, D, B, T = 11, cfg["K"], cfg["D"], 1, 800
        reduced = " (T reduced from <=%d: ~20 s of CPU work per step)" % cfg["tmax"]
    elif c in (1, 2):
        C, K, D, B, T = 13, cfg["K"], cfg["D"], 2, 1000
        reduced = " (one 6-step task, T reduced from <=%d: ~20 s of CPU work per step)" % cfg["tmax"]
    elif c == 3:
        C, K, D, B, T = 48, 60, cfg["D"], 1, 120
        reduced = " (reduced from T<=%d, K=%d: the potentials of ONE full video are %.0f GB, and this sample already costs " \
                  "~20 s per step)" % (cfg["tmax"], cfg["K"], cfg["tmax"] * cfg["K"] * 49 * 49 * 4 / 1e9)
    else:
        C, K, D, B, T = 64, 100, cfg["D"], 1, 200
        reduced = " (one mid cell of the sweep, reduced from T<=%d: C=133/K=200 needs 29 GB per video)" % cfg["tmax"]
    chain = cfg["chain"]
    lengths = torch.randint(max(2 * C, T // 2), T + 1, (B,), generator=gen)
    lengths[0] = T
    means = torch.randn(C, D, generator=gen) * 0.35
    cuts = torch.sort((torch.rand(B, C - 1, generator=gen) * (lengths[:, None] - 1)).long() + 1, dim=1)[0]
    pos = torch.arange(T).unsqueeze(0).expand(B, T).contiguous()
    labels = torch.searchsorted(cuts, pos, right=True)
    X = (torch.randn(B, T, D, generator=gen) + means[labels]) * (pos < lengths[:, None]).unsqueeze(-1)
    pen = None
    if cfg["narration"]:
        pen = torch.zeros(B, T, C)
        for j in range(1, C, 2):
            inside = labels == j
            lo = torch.where(inside, pos, T).min(dim=1)[0] - 40
            hi = torch.where(inside, pos, -1).max(dim=1)[0] + 40
            pen[:, :, j] = (~((pos >= lo[:, None]) & (pos <= hi[:, None]))).float() * -1e4
    tmask, imask = chain_masks(C) if chain else (None, None)
    return dict(X=X, lengths=lengths, means=means, cov=torch.rand(D, generator=gen) + 0.5, tmask=tmask, imask=imask, chain=chain,
                trans_logits=torch.randn(C, C, generator=gen) * 0.1, init_logits=torch.rand(C, generator=gen),
                log_rates=torch.log(torch.rand(C, generator=gen) * 8 + 4), C=C, K=K, frames=int(lengths.sum()), penalty=pen,
                decode_only=(c == 4), reduced=reduced)


def reference_module(s):
    """The UNMODIFIED reference `SemiMarkovModule` (oracle/_ref/src/models/semimarkov/semimarkov_modules.py) with the
    sample's parameters, or None when the reference sources are not available."""
    from action_segmentation_b200.args import HsmmArgs
    from tests.golden import ref_import
    if not ref_import.reference_available():
        return None
    mods, _ = ref_import.load_reference()
    C = s["C"]
    kw = {}
    if s["chain"]:
        kw = dict(allowed_starts={0}, allowed_transitions={c: ({c, c + 1} if c + 1 < C else {c}) for c in range(C)},
                  allowed_ends={C - 1})
    m = mods.SemiMarkovModule(HsmmArgs(sm_max_span_length=s["K"]), C, s["X"].shape[2], allow_self_transitions=True, **kw)
    with torch.no_grad():
        m.gaussian_means.copy_(s["means"])
        m.gaussian_cov.copy_(torch.diag(s["cov"]))
        m.transition_logits.copy_(s["trans_logits"])
        m.init_logits.copy_(s["init_logits"])
        m.poisson_log_rates.copy_(s["log_rates"])
    return m


def cpu_reference_step(s, module):
    B = s["X"].shape[0]
    ends = [[] for _ in range(B)] if s["chain"] else None
    if module is not None:  # the reference's own code path: log_likelihood().backward() + viterbi()
        vc = [torch.arange(s["C"]) for _ in range(B)]  # every class is valid (the reference needs the list with allowed_ends)
        if not s["decode_only"]:
            module.zero_grad()
            ll, _ = module.log_likelihood(s["X"], s["lengths"], vc, additional_allowed_ends_per_instance=ends, constraints=s["penalty"])
            (-ll).backward()
        with torch.no_grad():
            module.viterbi(s["X"], s["lengths"], vc, additional_allowed_ends_per_instance=ends, constraints=s["penalty"])
        return
    from oracle.reference_port import ReferencePort  # builder's port of the same algorithm (only without oracle/_ref)
    rp = ReferencePort(s["means"], s["cov"], s["trans_logits"], s["init_logits"], s["log_rates"], s["K"], s["tmask"], s["imask"])
    pends = [[s["C"] - 1] for _ in range(B)] if s["chain"] else None
    if not s["decode_only"]:
        rp.train_step(s["X"], s["lengths"], constraints=s["penalty"], allowed_ends=pends)
    with torch.no_grad():
        rp.viterbi(s["X"], s["lengths"], constraints=s["penalty"], allowed_ends=pends)


def run_cpu_baseline(args, cfg, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    s = cpu_sample(args, cfg)
    module = reference_module(s)
    for _ in range(warmup):
        cpu_reference_step(s, module)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(s, module)
    dt = (time.perf_counter() - t0) / steps
    what = "log_likelihood().backward() + viterbi()" if not s["decode_only"] else "viterbi()"
    sample = "%d videos, C=%d+EOS, T<=%d, D=%d, K=%d, %s transitions%s%s: %d frames/step; %s of the reference's SemiMarkovModule " \
             "(materialised potentials, sequential DP, autograd marginals) %s" % (
                 s["X"].shape[0], s["C"], s["X"].shape[1], s["X"].shape[2], s["K"], "chain-constrained" if s["chain"] else "dense",
                 ", narration penalty" if s["penalty"] is not None else "", s["reduced"], s["frames"], what,
                 "imported unmodified from oracle/_ref over oracle/torch_struct_shim.py" if module is not None
                 else "restated in oracle/reference_port.py (oracle/_ref absent)")
    return dict(value=s["frames"] / dt, unit=UNIT, cores=torch.get_num_threads(), kind="reference" if module is not None else "port",
                sample=sample), dt


# ---------------------------------------------------------------------------------------------
def bind_near_gpu(local):
    """Pin this rank's host threads (and hence its pinned allocations, first-touch) to the CPUs NVML reports as local to
    its GPU: with 8 ranks on one NUMA node host->device copies collapse to ~20 GB/s per GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        mine = cpus & allowed
        if mine:
            os.sched_setaffinity(0, mine)
            return "%d CPUs local to GPU %d" % (len(mine), local)
        return "GPU %d's local CPUs are outside this process's cpuset (%d CPUs allowed)" % (local, len(allowed))
    except Exception as e:  # noqa: BLE001
        return "unavailable (%s)" % type(e).__name__


def measure(one_step, steps, barrier, drain=None):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        one_step()
    if drain is not None:
        drain()  # outstanding all-reduces belong to the timed steps
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1)


def reduce_timing(ms, frames, world, device):
    t = torch.tensor([ms, float(frames)], device=device, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        tsum = t.clone()
        torch.distributed.all_reduce(tsum, op=torch.distributed.ReduceOp.SUM)
        return float(tmax[0]), float(tsum[1])
    return ms, float(frames)


def peak_hbm():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def run_train_decode(args, cfg, rank, world, device, barrier, sampler):
    from action_segmentation_b200 import _lib
    tasks = make_workload(args, cfg, rank, device)
    frames = sum(tk.frames for tk in tasks)
    layout, total = packed_layout(tasks)
    packed = torch.zeros(total, device=device)
    # forward + backward in one launch for the float-state kernels; the f64-state pair (162 registers: three CTAs per SM, not
    # all 576 CTAs resident) runs as two launches over three task groups (r02u: unfused 3 / 2 / 1 groups 6.25 / 6.35 / 6.65 ms,
    # fused 1 / 2 groups 6.74 / 6.67 ms)
    fused = os.environ.get("HSMM_BENCH_FUSED", "0" if cfg["narration"] else "1") == "1"
    n_groups = int(os.environ.get("HSMM_BENCH_GROUPS", "1" if fused else "3"))
    grouped = n_groups > 0 and grouped_eligible(tasks) and len(tasks) > 1 and not os.environ.get("HSMM_BENCH_SKIP")

    def device_step(tasks, streams, packed, layout, world, reduce=True):  # noqa: F811 (the step of this run)
        if grouped:
            return device_step_grouped(tasks, streams, packed, layout, world, reduce=reduce, n_groups=n_groups, fused=fused)
        return globals()["device_step"](tasks, streams, packed, layout, world, reduce=reduce)

    streams = make_streams(len(tasks))
    D = cfg["D"]

    # ---- device-resident throughput -------------------------------------------------------------
    # The step is ~100 library kernels on 36 streams: launched from Python it is host-bound, so after the eager
    # warm-up it is captured ONCE into a CUDA graph (same kernels, same streams/dependencies, allocations from
    # the graph's private pool) and the timed region replays it; the packed all-reduce follows each replay.
    for _ in range(args.warmup):
        device_step(tasks, streams, packed, layout, world)
    barrier()
    graph, launches_per_step = None, None
    graphs, packs, comm, pending = [], [packed], None, {}
    if not args.no_graph:
        l_cap = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            graph_outs = device_step(tasks, streams, packed, layout, world, reduce=False)  # noqa: F841 (kept alive)
        launches_per_step = _lib.launch_count() - l_cap
        graphs = [graph]
        if world > 1 and os.environ.get("HSMM_BENCH_NO_OVERLAP") != "1":
            # double-buffered statistics: step i's packed all-reduce runs on a side stream while step i+1's graph
            # (which accumulates into the OTHER buffer) already computes; a buffer is reused two steps later, after
            # its all-reduce has finished
            packed_b = torch.zeros_like(packed)
            graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_b):
                graph_outs_b = device_step(tasks, streams, packed_b, layout, world, reduce=False)  # noqa: F841
            graphs.append(graph_b)
            packs.append(packed_b)
            comm = torch.cuda.Stream()
    step_no = [0]

    def one_step():
        if graph is None:
            device_step(tasks, streams, packed, layout, world)
            return
        i = step_no[0] % len(graphs)
        step_no[0] += 1
        cur = torch.cuda.current_stream()
        if i in pending:
            cur.wait_event(pending.pop(i))  # the buffer's previous all-reduce is done (it is zeroed by the graph)
        graphs[i].replay()
        if os.environ.get("HSMM_BENCH_NO_REDUCE") == "1":  # experiment: how much of the N > 1 step is the collective?
            return
        if world > 1 and comm is None:
            torch.distributed.all_reduce(packs[i])
        elif world > 1:
            done = torch.cuda.Event()
            done.record(cur)
            comm.wait_event(done)
            with torch.cuda.stream(comm):
                torch.distributed.all_reduce(packs[i])
                ev = torch.cuda.Event()
                ev.record(comm)
            pending[i] = ev

    def drain():
        cur = torch.cuda.current_stream()
        for ev in pending.values():
            cur.wait_event(ev)
        pending.clear()

    for _ in range(args.warmup):
        one_step()
    drain()
    barrier()
    mark0 = sampler.mark()
    l0 = _lib.launch_count()
    ms = measure(one_step, args.steps, barrier, drain)
    launches = (_lib.launch_count() - l0) if graph is None else launches_per_step * args.steps
    ms_own = ms
    ms, frames_all = reduce_timing(ms, frames, world, device)
    ms_per_step = ms / args.steps
    value = frames_all / (ms_per_step * 1e-3)
    if world > 1 and os.environ.get("HSMM_BENCH_VERBOSE"):
        print("rank %d: %.3f ms/step on %d frames" % (rank, ms_own / args.steps, frames), file=sys.stderr)

    # ---- sustained: the same step back to back for >= 2 s (the DP kernels are issue-bound and follow the SM clock) ----
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / ms_per_step) + 1)
        barrier()
        m0 = sampler.mark()
        ms_s = measure(one_step, n_sus, barrier, drain)
        m1 = sampler.mark()
        ms_s, _ = reduce_timing(ms_s, frames, world, device)
        sustained = {"value": frames_all / (ms_s / n_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_s / 1e3,
                     "ms_per_step": ms_s / n_sus, "clocks": sampler.summary(m0, m1 + 1) if rank == 0 else None}
    # the timed region is ~0.03 s, shorter than one nvidia-smi period: the clocks line covers the timed steps AND the
    # sustained loop that follows them
    clocks = sampler.summary(mark0, sampler.mark() + 1) if rank == 0 else None

    # ---- per-kernel durations and the roofline of the dominant kernel ---------------------------
    peak_gbs, peak_src = peak_hbm()
    kms, klaunch = kernel_breakdown_grouped(tasks, fused=fused) if grouped else kernel_breakdown(tasks)
    dom = max(kms, key=kms.get)
    meanC = sum(tk.C * tk.frames for tk in tasks) / float(frames)
    pen_bytes = 4.0 * meanC if cfg["narration"] else 0.0
    bpf = (4 * D + 8 + pen_bytes) if dom == "viterbi" else (8 * D + pen_bytes)  # decode / train step
    achieved = frames * bpf / (kms[dom] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom) if args.config in (1, 2) else None
    except Exception:
        pass
    step_bytes = frames * (8 * D + 8 + pen_bytes)
    step_frac = step_bytes / (ms_per_step * 1e-3) / 1e9 / peak_gbs
    comp = compute_ceiling(tasks, value / world, (clocks or {}).get("sm_mhz"), False)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_frame": bpf,
                "kernel_ms": {k: round(v, 3) for k, v in kms.items()}, "launches_per_step": klaunch,
                "kernel_ms_note": "each family timed alone: its API calls captured into one CUDA graph, replays timed with CUDA events "
                                  "on the launching stream (a grouped DP call = its two or three kernels over all tasks; the "
                                  "per-task calls of a family run one after the other); in the step they overlap across streams",
                # whole step against the same peak, per GPU (step_bytes counts this rank's frames)
                "step_achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9, "step_frac": step_frac,
                "compute_ceiling": comp, "binding": "fp32_issue" if comp["frac"] > step_frac else "hbm"}

    # ---- end to end through the public API -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(tasks, streams, frames_all, args, world, device, barrier, False)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = run_cpu_baseline(args, cfg, 1, 1)

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_step_per_gpu": frames, "videos_per_step_per_gpu":
                   sum(tk.V for tk in tasks), "parallelism": "dp%d over videos, 1 packed all-reduce/step%s" % (
                       world, " on a side stream, overlapping the next step (double-buffered statistics)" if comm is not None else ""),
                   "launch": "eager (Python)" if graph is None else "CUDA-graph replay of the step (captured after eager warm-up)",
                   "dp_launches": ("hsmm_dp_grouped: one launch per kernel family over the %d tasks (%d group%s%s)" % (
                       len(tasks), n_groups, "" if n_groups == 1 else "s", "; forward + backward fused into one launch" if fused else ""))
                   if grouped else "one launch per task and kernel family",
                   "l2_policy": "inputs larger than L2 (%.1f GB of features per step per GPU)" % (frames * D * 4 / 1e9),
                   "dp_variants": sorted(set("%s | %s | %s" % tuple(_lib.dp_variant(tk.C, tk.K, m, tk.chain, tk.penalty is not None)
                                                                    for m in (0, 1, 2)) for tk in tasks))},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "sustained": sustained, "gpu_launches": int(launches), "clocks": clocks,
    }


def run_e2e(tasks, streams, frames_all, args, world, device, barrier, decode_only):
    models = build_models(tasks)
    host = [HostBatch(tk) for tk in tasks]
    h2d = sum(h.h2d_bytes for h in host)
    d2h = sum(tk.V * (tk.Tmax + 1) * 8 + tk.V * tk.Tmax * 8 + 4 for tk in tasks)
    copy_streams = [torch.cuda.Stream() for _ in range(int(os.environ.get("HSMM_BENCH_COPY_STREAMS", "3")))]
    ready = e2e_upload(tasks, host, copy_streams, 0)  # prime the pipeline: inputs of the first (warm-up) step
    slot = 0
    for _ in range(max(1, min(args.warmup, 2))):
        _, _, ready = e2e_step(models, tasks, host, streams, copy_streams, slot, ready, decode_only)
        slot = 1 - slot
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        _, _, ready = e2e_step(models, tasks, host, streams, copy_streams, slot, ready, decode_only)
        slot = 1 - slot
        if world > 1 and not decode_only:
            allreduce_all_gradients(models)
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    tt = torch.tensor([dt], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    padded = sum(h.features.numel() * 4 + (0 if h.penalty is None else h.penalty.numel() * 4) for h in host)
    return {"value": frames_all / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
            "ms_per_step": float(tt[0]) * 1e3, "steps": n_e2e, "h2d_gbs_achieved": h2d / float(tt[0]) / 1e9,
            "note": "every timed step: one upload of a step's live frames from pinned memory (hsmm_upload_ragged: %.2f of the padded "
                    "%.2f GB per GPU; it is the NEXT step's input, prefetched into a second set of landing buffers while this step "
                    "computes), the module API on this step's inputs (one emission pass per batch: log_likelihood_and_viterbi), loss + "
                    "spans + labels read back; one packed gradient all-reduce per step when N > 1" % (h2d / 1e9, padded / 1e9)}


def run_sweep(args, cfg, rank, world, device, barrier, sampler):
    """configs[4]: decode-only throughput over C x K, the same videos for every cell."""
    from action_segmentation_b200 import _lib
    gen = torch.Generator().manual_seed(args.seed + rank)
    nv = args.sweep_videos
    chunk = min(nv, cfg["V"])
    peak_gbs, peak_src = peak_hbm()
    D = cfg["D"]
    # the videos: built once (features around 16 class means), shared by every cell
    base = [make_task(16, cfg, gen, device, V=min(chunk, nv - i)) for i in range(0, nv, chunk)]
    frames = sum(tk.frames for tk in base)
    streams = [torch.cuda.Stream() for _ in range(min(2 * len(base), 20) + 1)]
    cells, tot_ms, tot_frames, launches = [], 0.0, 0.0, 0
    mark0 = sampler.mark()
    steps = max(1, min(args.steps, 2))
    for C in SWEEP_C:
        for K in SWEEP_K:
            ccfg = dict(cfg, K=K)
            tasks = [make_task(C, ccfg, gen, device, V=b.V, X=b.X, lengths=b.lengths) for b in base]
            for _ in range(max(1, min(args.warmup, 3))):
                device_step(tasks, streams, None, None, world, decode_only=True)
            barrier()
            l0 = _lib.launch_count()
            ms = measure(lambda: device_step(tasks, streams, None, None, world, decode_only=True), steps, barrier)
            launches += _lib.launch_count() - l0
            ms, frames_all = reduce_timing(ms, frames, world, device)
            kms, _ = kernel_breakdown(tasks, reps=2, decode_only=True)  # min of two: the first pays the allocations
            v = frames_all / (ms / steps * 1e-3)
            comp = compute_ceiling(tasks, v / world, None, True)
            hbm_frac = (v / world) * (4 * D + 8) / 1e9 / peak_gbs
            cells.append({"C": C, "K": K, "frames_per_s": v, "ms_per_step": ms / steps, "hbm_frac": hbm_frac,
                          "fp32_issue_frac": comp["frac"], "binding": "fp32_issue" if comp["frac"] > hbm_frac else "hbm",
                          "viterbi_variant": _lib.dp_variant(C, K, 0), "kernel_ms": {k: round(x, 3) for k, x in kms.items()}})
            tot_ms += ms / steps
            tot_frames += frames_all
            del tasks
            torch.cuda.empty_cache()
    clocks = sampler.summary(mark0, sampler.mark() + 1) if rank == 0 else None
    value = tot_frames / (tot_ms * 1e-3)
    worst = max(cells, key=lambda c: c["ms_per_step"])
    roofline = {"bound": "hbm", "kernel": "viterbi", "achieved": (value / world) * (4 * D + 8) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                "frac": (value / world) * (4 * D + 8) / 1e9 / peak_gbs, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": 4 * D + 8,
                "note": "whole sweep (emission + Viterbi of all nine cells); the slowest cell is C=%d K=%d, bound by %s" % (
                    worst["C"], worst["K"], worst["binding"])}
    e2e = None
    if not args.no_e2e:
        # end to end on the first cell (C=16, K=50): live frames from pinned host memory, labels + spans back
        sub = base[:4]
        etasks = [make_task(16, dict(cfg, K=50), gen, device, V=b.V, X=b.X, lengths=b.lengths) for b in sub]
        fr = float(sum(b.frames for b in sub)) * world
        e2e = run_e2e(etasks, streams, fr, args, world, device, barrier, True)
        e2e["note"] += "; cell C=16 K=50 on %d videos/GPU" % sum(b.V for b in sub)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = run_cpu_baseline(args, cfg, 1, 1)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": tot_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg), "frames_per_cell_per_gpu": frames, "videos_per_cell_per_gpu": nv,
                   "parallelism": "dp%d over videos, no collective (decode)" % world,
                   "launch": "eager (Python), %d batches of %d videos" % (len(base), chunk),
                   "l2_policy": "inputs larger than L2 (%.1f GB of features per cell per GPU)" % (frames * D * 4 / 1e9),
                   "step": "one step = the nine (C, K) cells once; value = frames of all cells / summed cell times"},
        "sweep": cells, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }


def emit(result):
    """The ONE JSON line, on the process's original stdout."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(result) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    cfg = config_of(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: whatever the libraries print meanwhile (NCCL's version banner, warnings of the
    # reference's modules) goes to stderr -- file descriptor 1 points there until the line is written
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        if rank != 0:
            return
        cb, dt = run_cpu_baseline(args, cfg, max(1, args.steps), max(0, min(args.warmup, 1)))
        emit(({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args, cfg), "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the HSMM path)"
    affinity = bind_near_gpu(local) if world > 1 else "not applied (single rank)"
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    import action_segmentation_b200  # noqa: F401

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.config == 4:
        result = run_sweep(args, cfg, rank, world, device, barrier, sampler)
    else:
        result = run_train_decode(args, cfg, rank, world, device, barrier, sampler)
    if rank == 0:
        sampler.stop()
        result["config"]["host_affinity"] = affinity
        sw = env_switches()
        if sw:
            result["config"]["env_switches"] = sw
            if "HSMM_BENCH_SKIP" in sw or "HSMM_BENCH_TASKS" in sw or "HSMM_BENCH_NO_REDUCE" in sw:
                result["invalid"] = "ablation run: HSMM_BENCH_SKIP / HSMM_BENCH_TASKS drop work from the timed region"
        emit(result)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
