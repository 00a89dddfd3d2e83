"""Dev probe: throughput of L concurrent graph replays (independent referrals on L streams, one model copy
+ workspaces per lane) vs the serial replay."""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np, torch
import bench
from instancerefer_b200 import synthetic

model, dev = bench._forward_setup(0)
K = 200
for L in (1, 2, 3, 4, 6, 8):
    lanes = []
    for l in range(L):
        m = copy.deepcopy(model)
        d = bench._resident_dict(synthetic.make_batch(1000 + 7 * l, batch_size=1, **bench.WORKLOAD), dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(2):
                m(dict(d))
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                o = m(dict(d))
        lanes.append((m, d, st, g, o))
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        t0.record(main)
        for _, _, st, _, _ in lanes:
            st.wait_event(t0)
        for i in range(K):
            _, _, st, g, _ = lanes[i % L]
            with torch.cuda.stream(st):
                g.replay()
        for _, _, st, _, _ in lanes:
            main.wait_stream(st)
        t1.record(main)
        torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print(f'lanes {L}: {ms / K * 1e3:.1f} us/referral, {K / ms * 1e3:.0f} referrals/s', flush=True)
    del lanes
    torch.cuda.empty_cache()
