#!/usr/bin/env python
"""Benchmark of the HSMM hot path on B200: video frames/s for emission scoring + log-semiring
forward/backward (logZ and all expected counts) + max-plus Viterbi, on synthetic CrossTask-shaped data
(BASELINE.json configs[1]: unsupervised HSMM, --mix_tasks --task_specific_steps
--sm_constrain_transitions: 18 tasks, 133 step classes, per-task chains of 2s+1 classes, 200-dim
features, videos sharded over the GPUs, one all-reduce of the packed sufficient statistics per step).

    python bench.py --gpus N --steps K --warmup W          # one rank per GPU under torchrun for N > 1
    python bench.py --impl reference ...                   # CPU port of the reference's own algorithm

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, library kernels only.  `e2e`: the
public module API (`log_likelihood().backward()`, `viterbi()`), features copied from pinned host
memory and results read back inside the timed region.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# number of steps of the 18 primary CrossTask tasks (sum = 133 step classes)
CROSSTASK_STEPS = [6, 5, 8, 11, 6, 6, 6, 11, 8, 11, 3, 7, 5, 8, 11, 5, 9, 7]
METRIC = "video frames/sec for HSMM fwd-bwd+Viterbi"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos-per-task", type=int, default=128, help="videos of each task per step and per GPU")
    ap.add_argument("--max-span", type=int, default=20, help="--sm_max_span_length (reference default 20)")
    ap.add_argument("--feature-dim", type=int, default=200)
    ap.add_argument("--tmin", type=int, default=1000)
    ap.add_argument("--tmax", type=int, default=3000)
    ap.add_argument("--narration", action="store_true", help="configs[2]: add the -1e4 narration penalty tensor")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-videos", type=int, default=5)
    ap.add_argument("--cpu-sample-frames", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic CrossTask-shaped workload
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


def make_task(t, steps, V, D, K, tmin, tmax, narration, gen, device):
    C = 2 * steps + 1
    tk = Task()
    tk.C, tk.K, tk.D, tk.V = C, K, D, V
    means = torch.randn(C, D, generator=gen) * 0.35
    means[0::2] = means[0]  # merged backgrounds (--annotate_background_with_previous)
    tk.means = means.to(device)
    tk.cov_diag = (torch.rand(D, generator=gen) + 0.5).to(device)
    tmask, imask = chain_masks(C)
    tk.trans_mask, tk.init_mask = tmask, imask
    tk.trans_logits = torch.randn(C, C, generator=gen) * 0.1
    tk.init_logits = torch.rand(C, generator=gen)
    tk.log_rates = torch.log(torch.rand(C, generator=gen) * 8 + 4)
    trans = torch.log_softmax(tk.trans_logits.masked_fill(tmask, -1e9), dim=0)
    init = torch.log_softmax(tk.init_logits.masked_fill(imask, -1e9), dim=0)
    k = torch.arange(K, dtype=torch.float32).unsqueeze(-1)
    lenp = k * tk.log_rates.unsqueeze(0) - torch.exp(tk.log_rates).unsqueeze(0) - torch.lgamma(k + 1)
    tk.trans, tk.init, tk.lenp = trans.to(device), init.to(device), lenp.to(device)
    end = torch.full((V, C), -1e9)
    end[:, C - 1] = 0
    tk.end = end.to(device)
    tk.lengths = torch.randint(tmin, tmax + 1, (V,), generator=gen)
    tk.lengths[0] = tmax
    Tmax = int(tk.lengths.max())
    tk.Tmax = Tmax
    # labels: the chain in order, random cut points
    cuts = torch.sort((torch.rand(V, C - 1, generator=gen) * (tk.lengths[:, None] - 1)).long() + 1, dim=1)[0].to(device)
    pos = torch.arange(Tmax, device=device).unsqueeze(0).expand(V, Tmax).contiguous()
    labels = torch.searchsorted(cuts, pos, right=True)
    dgen = torch.Generator(device=device).manual_seed(int(torch.randint(0, 2 ** 31, (1,), generator=gen)))
    X = torch.randn(V, Tmax, D, device=device, generator=dgen)
    X += tk.means[labels]
    live = (pos < tk.lengths.to(device)[:, None])
    X *= live.unsqueeze(-1)
    tk.X = X.contiguous()
    tk.penalty = None
    if narration:
        # every step gets one window around its true span; outside it the step costs -1e4 per frame
        pen = torch.zeros(V, Tmax, C, device=device)
        for j in range(1, C, 2):
            inside = (labels == j)
            lo = torch.where(inside, pos, Tmax).min(dim=1)[0] - 40
            hi = torch.where(inside, pos, -1).max(dim=1)[0] + 40
            allowed = (pos >= lo[:, None]) & (pos <= hi[:, None])
            pen[:, :, j] = (~allowed).float() * -1e4
        tk.penalty = pen
    from action_segmentation_b200 import hsmm
    tk.lengths_i32, tk.order = hsmm.prepare_lengths(tk.lengths, torch.device(device))
    tk.frames = int(tk.lengths.sum())
    tk.class_ids = torch.arange(C + 1, device=device, dtype=torch.int32)
    tk.gradw = torch.full((V,), 1.0 / V, device=device)
    tk.eparams = hsmm.emission_params(tk.means, tk.cov_diag)  # w, bias, 1/var, row constant (parameter-only work)
    # ordering constraints leave <= 2 unmasked transitions per class: hint for the sparse-transition kernels
    tk.pred, tk.succ = hsmm.sparse_transition_lists(~tmask, torch.device(device))
    return tk


def make_workload(args, rank, device):
    gen = torch.Generator().manual_seed(args.seed + rank)
    tasks = [make_task(t, s, args.videos_per_task, args.feature_dim, args.max_span, args.tmin, args.tmax, args.narration,
                       gen, device) for t, s in enumerate(CROSSTASK_STEPS)]
    return tasks


def packed_layout(tasks):
    """Offsets of every task's [d_means | d_trans | d_len | d_init | sum logZ] slice of the packed buffer."""
    off, lay = 0, []
    for tk in tasks:
        sizes = [tk.C * tk.D, tk.C * tk.C, tk.K * tk.C, tk.C, tk.C, 1]  # wx, d_trans, d_len, d_init, wsum, logz
        lay.append((off, sizes))
        off += sum(sizes)
    return lay, off


# ---------------------------------------------------------------------------------------------
# one step of the hot path, inputs resident in HBM, library kernels only
# ---------------------------------------------------------------------------------------------
def device_step(tasks, streams, packed, layout, world, reduce=True, sched=None):
    """One pass of the hot path over every task's batch.  Per task: emission scoring -> {forward -> backward ->
    class-weighted feature sums} on the task's stream and, concurrently on a second stream, Viterbi (which
    needs the emission scores only).  `sched` = "interleaved": every task's emission on its own stream;
    "emfirst": all emissions back to back on one stream before any DP kernel (the emission kernel is a
    persistent one-CTA-per-SM kernel and otherwise waits for whole SMs to drain)."""
    from action_segmentation_b200 import hsmm
    lib = hsmm._lib.load()
    sched = sched or os.environ.get("HSMM_BENCH_SCHED", "interleaved")
    skip = set(filter(None, os.environ.get("HSMM_BENCH_SKIP", "").split(",")))  # ablation only: invalid as a bench number
    cur = torch.cuda.current_stream()
    packed.zero_()
    fork = torch.cuda.Event()
    fork.record(cur)
    n = len(tasks)
    ems, em_evs = [], []
    if sched == "emfirst":
        s_stream = streams[2 * n]
        s_stream.wait_event(fork)
        with torch.cuda.stream(s_stream):
            for tk in tasks:
                ems.append(hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32, params=tk.eparams))
            ev = torch.cuda.Event()
            ev.record(s_stream)
        em_evs = [ev] * n
    outs = []
    for i, tk in enumerate(tasks):
        st, st2 = streams[i], streams[n + i]
        st.wait_event(fork)
        off, sizes = layout[i]
        v = []
        o = off
        for m in sizes:
            v.append(packed[o:o + m])
            o += m
        wx, d_trans, d_len, d_init, wsum, lz = v
        if sched == "emfirst":
            em, rowterm, offset = ems[i]
            st.wait_event(em_evs[i])
            em_ready = em_evs[i]
        else:
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
        with torch.cuda.stream(st):
            xp = tk.penalty is not None
            logz, saved = hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order,
                                            trans_pred=tk.pred, f64_state=xp)
            g = tk.gradw if world == 1 else tk.gradw / world
            _, _, _, d_em = hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, tk.lengths_i32, tk.order, g,
                                               saved, out=(d_init, d_trans.view(tk.C, tk.C), d_len.view(tk.K, tk.C)),
                                               trans_succ=tk.succ, f64_state=xp)
            if "ws" not in skip:
                hsmm._lib.check(lib.hsmm_weighted_feature_sums(hsmm._p(tk.X), hsmm._p(d_em), d_em.shape[2],
                                                               hsmm._p(tk.lengths_i32), tk.V, tk.Tmax, tk.D, tk.C, hsmm._p(wx),
                                                               hsmm._p(wsum), hsmm._stream()), "hsmm_weighted_feature_sums")
            lz.copy_(logz.sum().float().reshape(1))
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        cur.wait_event(ev)
    if world > 1 and reduce:
        torch.distributed.all_reduce(packed)
    return outs


def e2e_step(models, tasks, host, streams):
    """Public API with host buffers: H2D of the step's features, loss + predictions read back."""
    lls = []
    preds = []
    dev_in = []
    # the step's inputs: every task's features (and penalties) leave pinned host memory back to back, so that the
    # copy engine never idles while the host prepares the next call
    for i, tk in enumerate(tasks):
        with torch.cuda.stream(streams[i % len(streams)]):
            feats = host[i]["features"].cuda(non_blocking=True)
            pen = None if host[i]["penalty"] is None else host[i]["penalty"].cuda(non_blocking=True)
            dev_in.append((feats, pen))
    for i, (m, tk) in enumerate(zip(models, tasks)):
        st = streams[i % len(streams)]
        with torch.cuda.stream(st):
            feats, pen = dev_in[i]
            m.zero_grad()
            ll, _ = m.log_likelihood(feats, tk.lengths, None, additional_allowed_ends_per_instance=[[] for _ in range(tk.V)],
                                     constraints=pen)
            (-ll).backward()
            spans, labels = m.viterbi(feats, tk.lengths, None, additional_allowed_ends_per_instance=[[] for _ in range(tk.V)],
                                      constraints=pen, return_labels=True, non_blocking=True)
            h = torch.empty((), dtype=torch.float32, pin_memory=True)
            h.copy_(ll.detach(), non_blocking=True)
            lls.append(h)
            preds.append((spans, labels))
    torch.cuda.synchronize()  # every task's loss, spans and labels are now in host memory
    return float(sum(float(h) for h in lls)), preds


def build_models(tasks, args):
    import action_segmentation_b200 as pkg
    from action_segmentation_b200.args import HsmmArgs as RefArgs
    models = []
    for tk in tasks:
        C = tk.C
        trans = {c: ({c, c + 1} if c + 1 < C else {c}) for c in range(C)}
        m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=args.max_span), C, tk.D, allow_self_transitions=True,
                                 allowed_starts={0}, allowed_transitions=trans, allowed_ends={C - 1}).cuda()
        with torch.no_grad():
            m.gaussian_means.copy_(tk.means)
            m.gaussian_cov.copy_(torch.diag(tk.cov_diag))
            m.transition_logits.copy_(tk.trans_logits)
            m.init_logits.copy_(tk.init_logits)
            m.poisson_log_rates.copy_(tk.log_rates)
        models.append(m)
    return models


# ---------------------------------------------------------------------------------------------
# per-kernel durations (serialised on one stream, CUDA events) -> dominant kernel roofline
# ---------------------------------------------------------------------------------------------
def kernel_breakdown(tasks, reps=3):
    """Per-kernel device time of one step with every launch serialised on one stream (CUDA events around each
    call; best of `reps` per launch, summed over the step's launches)."""
    from action_segmentation_b200 import hsmm
    names = ["emission", "logz_forward", "logz_backward", "weighted_feature_sums", "viterbi"]
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
            xp = tk.penalty is not None
            logz, saved = timed(("logz_forward", i), lambda: hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset,
                                                                               tk.lengths_i32, tk.order, trans_pred=tk.pred,
                                                                               f64_state=xp))
            g = torch.full((tk.V,), 1.0 / tk.V, device=em.device)
            d = timed(("logz_backward", i), lambda: hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end,
                                                                       tk.lengths_i32, tk.order, g, saved, trans_succ=tk.succ,
                                                                       f64_state=xp))
            timed(("weighted_feature_sums", i), lambda: hsmm.weighted_feature_sums(tk.X, d[3], tk.C, tk.lengths_i32))
            timed(("viterbi", i), lambda: hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset,
                                                              tk.lengths_i32, tk.order, tk.class_ids, want_score=False,
                                                              trans_pred=tk.pred))
            if rep == 0:
                for n in names:
                    launches[n] += 1
    return {n: sum(v for (k, _), v in best.items() if k == n) for n in names}, launches


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
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(int(r[0]) for r in rows), "sm_max_mhz": int(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (oracle/reference_port.py) on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_sample(args, passes=1):
    """Bounded sample of the workload for the CPU legs: about 6 s per video and pass on 16 cores, so the number of videos
    shrinks with the number of passes (warm-up + timed steps) to keep the whole run within a few minutes."""
    gen = torch.Generator().manual_seed(args.seed)
    steps = 6
    C, D, K = 2 * steps + 1, args.feature_dim, args.max_span
    B = max(1, min(args.cpu_sample_videos, 25 // max(1, passes)))
    T = args.cpu_sample_frames
    lengths = torch.randint(max(2 * C, T // 2), T + 1, (B,), generator=gen)
    lengths[0] = T
    means = torch.randn(C, D, generator=gen) * 0.35
    cuts = torch.sort((torch.rand(B, C - 1, generator=gen) * (lengths[:, None] - 1)).long() + 1, dim=1)[0]
    pos = torch.arange(T).unsqueeze(0).expand(B, T).contiguous()
    labels = torch.searchsorted(cuts, pos, right=True)
    X = (torch.randn(B, T, D, generator=gen) + means[labels]) * (pos < lengths[:, None]).unsqueeze(-1)
    tmask, imask = chain_masks(C)
    return dict(X=X, lengths=lengths, means=means, cov=torch.rand(D, generator=gen) + 0.5, tmask=tmask, imask=imask,
                trans_logits=torch.randn(C, C, generator=gen) * 0.1, init_logits=torch.rand(C, generator=gen),
                log_rates=torch.log(torch.rand(C, generator=gen) * 8 + 4), C=C, K=K, frames=int(lengths.sum()))


def cpu_reference_step(s):
    from oracle.reference_port import ReferencePort
    rp = ReferencePort(s["means"], s["cov"], s["trans_logits"], s["init_logits"], s["log_rates"], s["K"], s["tmask"], s["imask"])
    ends = [[s["C"] - 1] for _ in range(s["X"].shape[0])]
    rp.train_step(s["X"], s["lengths"], allowed_ends=ends)
    with torch.no_grad():
        rp.viterbi(s["X"], s["lengths"], allowed_ends=ends)


def run_cpu_baseline(args, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    s = cpu_sample(args, steps + warmup)
    for _ in range(warmup):
        cpu_reference_step(s)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(s)
    dt = (time.perf_counter() - t0) / steps
    sample = "%d videos of one 6-step task (C=%d+EOS), T<=%d, D=%d, K=%d: %d frames/step; materialised potentials + " \
             "sequential DP + autograd backward + argmax decode (oracle/reference_port.py)" % (
                 s["X"].shape[0], s["C"], s["X"].shape[1], s["X"].shape[2], s["K"], s["frames"])
    return dict(value=s["frames"] / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port", sample=sample), dt


def workload_name(args):
    return "configs[1] U7-shape HSMM EM: 18 CrossTask-like tasks (133 steps, C=2s+1 in 7..23 + EOS, chain-constrained " \
           "transitions), D=%d, K=%d, T~U[%d,%d], %d videos/task/GPU%s" % (
               args.feature_dim, args.max_span, args.tmin, args.tmax, args.videos_per_task,
               ", narration penalty -1e4 (configs[2])" if args.narration else "")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb, dt = run_cpu_baseline(args, max(1, args.steps), max(0, min(args.warmup, 1)))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args), "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the HSMM path)"
    torch.cuda.set_device(local)
    device = "cuda:%d" % local
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    import action_segmentation_b200 as pkg
    from action_segmentation_b200 import _lib

    tasks = make_workload(args, rank, device)
    frames = sum(tk.frames for tk in tasks)
    layout, total = packed_layout(tasks)
    packed = torch.zeros(total, device=device)
    streams = [torch.cuda.Stream() for _ in range(2 * len(tasks) + 1)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    # The step is ~90 library kernels on 36 streams: launched from Python it is host-bound, so after the eager
    # warm-up it is captured ONCE into a CUDA graph (same kernels, same streams/dependencies, allocations from
    # the graph's private pool) and the timed region replays it; the packed all-reduce follows each replay.
    for _ in range(args.warmup):
        device_step(tasks, streams, packed, layout, world)
    barrier()
    graph = None
    launches_per_step = None
    if not args.no_graph:
        l_cap = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            graph_outs = device_step(tasks, streams, packed, layout, world, reduce=False)  # noqa: F841 (kept alive)
        launches_per_step = _lib.launch_count() - l_cap

    def one_step():
        if graph is None:
            device_step(tasks, streams, packed, layout, world)
        else:
            graph.replay()
            if world > 1:
                torch.distributed.all_reduce(packed)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = (_lib.launch_count() - l0) if graph is None else launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, float(frames)], device=device, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        tsum = t.clone()
        torch.distributed.all_reduce(tsum, op=torch.distributed.ReduceOp.SUM)
        ms, frames_all = float(tmax[0]), float(tsum[1])
    else:
        frames_all = float(frames)
    ms_per_step = ms / args.steps
    value = frames_all / (ms_per_step * 1e-3)

    # ---- per-kernel durations and the roofline of the dominant kernel ---------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kms, klaunch = kernel_breakdown(tasks)
    dom = max(kms, key=kms.get)
    D = args.feature_dim
    meanC = sum(tk.C * tk.frames for tk in tasks) / float(frames)
    pen_bytes = 4.0 * meanC if args.narration else 0.0
    bytes_per_frame = {"viterbi": 4 * D + 8 + pen_bytes}
    train_bpf = 8 * D + pen_bytes
    bpf = bytes_per_frame.get(dom, train_bpf)
    achieved = frames * bpf / (kms[dom] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
    except Exception:
        pass
    step_bytes = frames * (8 * D + 8 + pen_bytes)
    # secondary (compute) ceiling of the dominant kernel, SURVEY 8(d): the DP kernels are bound by FP32 issue, not HBM.
    # Issue-slot utilisation comes from the committed ncu summaries (a profiler number, never a bench value).
    compute = None
    try:
        import csv
        names = {"logz_backward": "dp_lin_backward", "logz_forward": "dp_lin_forward", "viterbi": "dp_vit2",
                 "emission": "emission_tc", "weighted_feature_sums": "weighted_sums"}
        def issue_pct(path):
            rows = list(csv.reader(open(os.path.join(ROOT, "profiles", path))))
            col = rows[0].index("smsp__issue_active.avg.pct_of_peak_sustained_active")
            v = [float(r[col]) for r in rows[2:] if names[dom] in r[0]]
            return sum(v) / len(v) if v else None
        compute = {"bound": "fp32_issue", "issue_active_pct_in_step_launch": issue_pct("r01m_ncu_bench_top_summary.csv"),
                   "issue_active_pct_saturated_launch": issue_pct("r01d_ncu_dp_saturated_summary.csv"),
                   "source": "profiles/r01m_ncu_bench_top_summary.csv, profiles/r01d_ncu_dp_saturated_summary.csv (ncu)"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_frame": bpf,
                "kernel_ms": {k: round(v, 3) for k, v in kms.items()}, "launches_per_step": klaunch,
                # whole step against the same peak, per GPU (step_bytes counts this rank's frames)
                "step_achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak_gbs, "compute_ceiling": compute}

    # ---- end to end through the public API -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        models = build_models(tasks, args)
        host = [{"features": tk.X.cpu().pin_memory(), "penalty": None if tk.penalty is None else tk.penalty.cpu().pin_memory()}
                for tk in tasks]
        h2d = sum(h["features"].numel() * 4 + (0 if h["penalty"] is None else h["penalty"].numel() * 4) for h in host)
        d2h = sum(tk.V * (tk.Tmax + 1) * 8 + tk.V * tk.Tmax * 8 + 4 for tk in tasks)
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step(models, tasks, host, streams)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step(models, tasks, host, streams)
            if world > 1:
                from action_segmentation_b200 import distributed as hd
                for m in models:
                    hd.allreduce_gradients(m.parameters(), torch.zeros((), device=device))
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        e2e = {"value": frames_all / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": float(tt[0]) * 1e3, "steps": n_e2e, "h2d_gbs_achieved": h2d / float(tt[0]) / 1e9}
        del host

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = run_cpu_baseline(args, 1, 1)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "frames_per_step_per_gpu": frames, "videos_per_step_per_gpu":
                       sum(tk.V for tk in tasks), "parallelism": "dp%d over videos, 1 packed all-reduce/step" % world,
                       "launch": "eager (Python)" if graph is None else "CUDA-graph replay of the step (captured after eager warm-up)",
                       "l2_policy": "inputs larger than L2 (%.1f GB of features per step per GPU)" % (frames * D * 4 / 1e9),
                       "dp_variants": sorted(set(_lib.dp_variant(tk.C, tk.K, 1, True) for tk in tasks))},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
