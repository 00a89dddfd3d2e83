"""Graph-replay time of each branch of the forward in isolation (which chain is the critical path?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import bench, weights
from instancerefer_b200 import SparseTensor, synthetic
from instancerefer_b200.candidates import KEY, CandidatePack
from instancerefer_b200.instancerefer import InstanceRefer
args = bench.make_args()
m = InstanceRefer(7, args); m.load_state_dict(weights.make_state_dict(123)); m = m.cuda().eval()
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
d0 = synthetic.to_data_dict(b, SparseTensor, 'cuda')
pack = CandidatePack(d0, d0['object_cat'], 'cuda'); pack.resident = True
d0[KEY] = pack; d0['_ir_lang_len_max'] = 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

def timeit(fn, name, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print(f'{name:28s} median {ts[len(ts)//2]:7.1f} us  min {ts[0]:7.1f} us')

full = lambda: m(dict(d0))
m.concurrent = True
timeit(full, 'full (4 streams)')
m.concurrent = False
timeit(full, 'full (1 stream)')
dl = lambda: m.lang(dict(d0))
timeit(dl, 'lang')
def attr():
    d = dict(d0); m.attribute.encode_candidates(d, 'cuda', pack)
timeit(attr, 'attribute.encode_candidates')
def scene():
    d = dict(d0); m.scene.encode_scene(d, torch.device('cuda'))
timeit(scene, 'scene.encode_scene')
def rel():
    d = dict(d0); m.relation.encode_graph(d, torch.device('cuda'))
timeit(rel, 'relation.encode_graph')
def heads():
    d = dict(full_out)
    m.attribute.match(d); m.relation.match(d); m.scene.match(d)
m.concurrent = True
full_out = m(dict(d0))
timeit(heads, 'match heads (7 mlp + attn)')
from instancerefer_b200 import ops
ws = m.scene.net.workspace(d0['lidar'].F.shape[0], 'cuda')
timeit(lambda: ops.encoder_build_maps(ws, d0['lidar'].C), 'scene maps only')
prep = m.scene.net.prepared(); out = torch.empty(ws.n_max, 128, device='cuda')
timeit(lambda: ops.encoder_features(prep['params'], ws, d0['lidar'].F, out), 'scene 13 conv layers only')

# two encoders concurrently on two streams (how much do they overlap?)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d = dict(d0); m.attribute.encode_candidates(d, 'cuda', pack)
    with torch.cuda.stream(s2):
        d = dict(d0); m.scene.encode_scene(d, torch.device('cuda'))
    cur.wait_stream(s1); cur.wait_stream(s2)
timeit(both, 'attr || scene (2 streams)')
def three():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d = dict(d0); m.attribute.encode_candidates(d, 'cuda', pack)
    with torch.cuda.stream(s2):
        d = dict(d0); m.scene.encode_scene(d, torch.device('cuda'))
    m.lang(dict(d0))
    cur.wait_stream(s1); cur.wait_stream(s2)
timeit(three, 'attr || scene || lang')
ws_a = m.attribute.net.workspace(32 * 1024, 'cuda')
prep_a = m.attribute.net.prepared(); out_a = torch.empty(ws_a.n_max, 128, device='cuda')
def convs2():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ops.encoder_features(prep_a['params'], ws_a, None, out_a)
    with torch.cuda.stream(s2):
        ops.encoder_features(prep['params'], ws, d0['lidar'].F, out)
    cur.wait_stream(s1); cur.wait_stream(s2)
timeit(lambda: ops.encoder_features(prep_a['params'], ws_a, None, out_a), 'attr 13 conv layers only')
timeit(convs2, '13 convs attr || 13 convs scene')
