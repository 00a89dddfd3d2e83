"""Host-side profile of the resident forward (where does the Python time go?)."""
import cProfile, pstats, sys, os, types, time
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
for _ in range(5): m(dict(d0))
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20): m(dict(d0))
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host ms/step', (t1 - t) / 20 * 1e3, 'with sync', (t2 - t) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): m(dict(d0))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
