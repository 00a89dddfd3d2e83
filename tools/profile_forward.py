"""Forward of ONE resident bench scene, eager launches, for ncu (launch list / --set full): 3 warm-up forwards, then
`reps` forwards.  Writes the per-layer pair / row counts of that scene to gpurun_out/profile_forward_layers.json so that
tools/ncu_conv_table.py can turn ncu's per-launch durations into per-layer GB/s.
usage: [IR_GATHER=tma] [IR_ENCODER=persist] python tools/profile_forward.py [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from instancerefer_b200 import synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model, dev = bench._forward_setup(0)
if os.environ.get('IR_PAIR') == '1':
    model.pair_encoders = True
b = synthetic.make_batch(1000, batch_size=1, **bench.WORKLOAD)
d = bench._resident_dict(b, dev)
for _ in range(3):
    model(dict(d))
torch.cuda.synchronize()
for _ in range(reps):
    model(dict(d))
torch.cuda.synchronize()
maps = [0, 5, 1, 1, 6, 2, 2, 7, 3, 3, 8, 4, 4]
lvl_of = [0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4]
chans = [(7, 32, 27)] + [(32, 64, 8), (64, 64, 27), (64, 64, 27), (64, 128, 8), (128, 128, 27), (128, 128, 27)] + \
        [(128, 128, 8), (128, 128, 27), (128, 128, 27)] * 2
layers = []
for enc, net in (('instance', model.attribute.net), ('scene', model.scene.net)):
    kc, nl = net._last_ws.kcount().cpu().numpy(), net._last_ws.nlvl().cpu().numpy()
    for l, (cin, cout, K) in enumerate(chans):
        P = int(kc[maps[l]][:K].sum())
        layers.append(dict(encoder=enc, layer=l, cin=cin, cout=cout, K=K, rows_out=int(nl[lvl_of[l]]), pairs=P,
                           bytes_gather=P * (4 * cin + 4), bytes_scatter=P * (4 * cout + 4), bytes_weights=4 * K * cin * cout))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(dict(reps=reps, layers=layers), open(os.path.join(ROOT, 'gpurun_out', 'profile_forward_layers.json'), 'w'))
print('ok', len(layers), 'layers')
