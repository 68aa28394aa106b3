"""Per-task stage timeline of one bench step (CUDA events on the task streams): where does the step time go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from action_segmentation_b200 import hsmm  # noqa: E402


def main():
    em_first = "--em-first" in sys.argv
    sys.argv = [a for a in sys.argv if a != "--em-first"]
    args = bench.parse()
    torch.cuda.set_device(0)
    tasks = bench.make_workload(args, 0, "cuda:0")
    layout, total = bench.packed_layout(tasks)
    packed = torch.zeros(total, device="cuda:0")
    streams = [torch.cuda.Stream() for _ in range(2 * len(tasks) + 1)]
    for _ in range(3):
        bench.device_step(tasks, streams, packed, layout, 1)
    torch.cuda.synchronize()
    # instrumented copy of device_step
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    cur = torch.cuda.current_stream()
    t0 = ev()
    t0.record(cur)
    marks = []
    n = len(tasks)
    import time
    h0 = time.perf_counter()
    host = []
    pre = []
    if em_first:
        sA = streams[2 * n]
        sA.wait_event(t0)
        with torch.cuda.stream(sA):
            for tk in tasks:
                a = ev(); a.record(sA)
                r = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32, params=tk.eparams)
                b = ev(); b.record(sA)
                pre.append((a, r, b))
            all_em = ev(); all_em.record(sA)
    for i, tk in enumerate(tasks):
        st, st2 = streams[i], streams[n + i]
        st.wait_event(t0)
        m = {}
        if em_first:
            m["em0"], (em, rowterm, offset), m["em1"] = pre[i]
            st.wait_event(all_em)
            m["em1"] = all_em
        else:
            with torch.cuda.stream(st):
                m["em0"] = ev(); m["em0"].record(st)
                em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32, params=tk.eparams)
                m["em1"] = ev(); m["em1"].record(st)
        with torch.cuda.stream(st2):
            st2.wait_event(m["em1"])
            m["vit0"] = ev(); m["vit0"].record(st2)
            hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order, tk.class_ids,
                                want_labels=True, want_score=False, trans_pred=tk.pred)
            m["vit1"] = ev(); m["vit1"].record(st2)
        with torch.cuda.stream(st):
            m["fwd0"] = ev(); m["fwd0"].record(st)
            logz, saved = hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order,
                                            trans_pred=tk.pred)
            m["fwd1"] = ev(); m["fwd1"].record(st)
            _, _, _, d_em = hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, tk.lengths_i32, tk.order, tk.gradw,
                                               saved, trans_succ=tk.succ)
            m["bwd1"] = ev(); m["bwd1"].record(st)
            hsmm.weighted_feature_sums(tk.X, d_em, tk.C, tk.lengths_i32)
            m["wfs1"] = ev(); m["wfs1"].record(st)
        marks.append(m)
        host.append((time.perf_counter() - h0) * 1e3)
    torch.cuda.synchronize()
    print("task  C  host_ms |  em: start-end | fwd start-end | bwd end | wfs end | vit: start-end")
    for i, (tk, m) in enumerate(zip(tasks, marks)):
        t = {k: t0.elapsed_time(v) for k, v in m.items()}
        print("%3d %3d %7.2f | %6.2f-%6.2f | %6.2f-%6.2f | %6.2f | %6.2f | %6.2f-%6.2f" % (
            i, tk.C, host[i], t["em0"], t["em1"], t["fwd0"], t["fwd1"], t["bwd1"], t["wfs1"], t["vit0"], t["vit1"]))


if __name__ == "__main__":
    main()
