"""C1-sized scans are bound by parallelism, not throughput: time one step (both ends) for different unit-shape
masks, with the two ends on one stream or on two, with and without graph replay."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from approx_counter_b200 import ApproxCounter, host
wl = sys.argv[1] if len(sys.argv) > 1 else "C1"
import bench
w = bench.WORKLOADS[wl]
n, sl, k, lim = w["n"], w["sl"], w["k"], w["lim"]
dev = torch.device("cuda", 0)
ends = [host.synth_ends(w["seed"], 0, n, sl, b) for b in (False, True)]
thr = host.adjust_threshold(1.0, 16, k)
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
for mask, graph, two in [(0x3FFFFF, 0, 0), (0x3FFFFF, 1, 0), (0x3FFFFF, 0, 1), (0x3FFFFF, 1, 1), (0x3F0000, 0, 0), (0x3F0000, 0, 1), (0x380000, 0, 1), (0, 0, 0), (0, 0, 1)]:
    ctxs = [ApproxCounter(0), ApproxCounter(0)]
    outs = []
    for c, s, st in zip(ctxs, ends, streams):
        c.set_stream((st if two else streams[0]).cuda_stream)
        c.set_option("scan_graph", graph)
        c.set_option("shape_mask", mask)
        c.upload_sample(s)
        km = c.count_kmers_topn(k, thr, lim)[0]
        c.set_queries(km, k)
        outs.append(torch.zeros(len(km), dtype=torch.int64, device=dev))
    def step():
        if two:
            ev = torch.cuda.Event(); ev.record(streams[0]); streams[1].wait_event(ev)
        for c, o in zip(ctxs, outs):
            c.scan(o.data_ptr())
        if two:
            ev2 = torch.cuda.Event(); ev2.record(streams[1]); streams[0].wait_event(ev2)
    for _ in range(20): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 300
    e0.record(streams[0])
    for _ in range(steps): step()
    e1.record(streams[0])
    torch.cuda.synchronize()
    print(json.dumps({"workload": wl, "shape_mask": hex(mask), "graph": graph, "two_streams": two, "ms_per_step": e0.elapsed_time(e1) / steps}), flush=True)
    for c in ctxs: c.close()
