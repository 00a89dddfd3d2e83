"""Dev aid: the one-call training passes of ONE sparse encoder (ir_encoder_train_forward / _backward, capacity mode) on
the bench's configs[2] batch, alone on one stream: CUDA-event time per direction, and — under
`ncu --profile-from-start off --metrics gpu__time_duration.sum` — one forward + backward between profiler start/stop,
whose launch list is then in layer order.
usage: python tools/encoder_train_probe.py scene|attr [reps]"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from instancerefer_b200 import SparseTensor, _lib, ops, synthetic, training as T
from instancerefer_b200.candidates import CandidatePack
from instancerefer_b200.instancerefer import InstanceRefer

which = sys.argv[1] if len(sys.argv) > 1 else 'scene'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda', 0)
ops.check_device(0)
model = InstanceRefer(7, bench.make_args())
model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)
model = model.to(dev).train()
b = bench.train_batches(0, n=1)[0]
d = synthetic.to_data_dict(b, SparseTensor, dev)
pack = CandidatePack(d, d['object_cat'], dev)
if which == 'scene':
    net = model.scene.net
    ws, F0, C0 = T.prepare_scene_maps(model, d, dev)
else:
    net = model.attribute.net
    ws, F0 = T.prepare_attribute_maps(model, pack), None
torch.cuda.synchronize()
print(which, 'n_max', ws.n_max, 'levels', ws.nlvl().tolist())
G = T.EncoderGraph(ws, [ws.n_max] * 5)
flat = [t for conv, bn in net._layers() for t in (conv.kernel, bn.weight, bn.bias)]
st = T.EncoderTrainGraphed._state(net, G, flat, F0.shape[1] if F0 is not None else 0)
if F0 is not None:
    st.f0[:F0.shape[0]].copy_(F0)
st.dout.normal_()


def fwd():
    _lib.call("ir_encoder_train_forward", C.byref(st.P), ops._p(st.f0), ws.ptr, ws.n_max, None, ops._p(st.arena), ops._stream())


def bwd():
    _lib.call("ir_encoder_train_backward", C.byref(st.P), ops._p(st.f0), ws.ptr, ws.n_max, None, ops._p(st.arena),
              ops._p(st.dout), C.byref(st.Gr), ops._stream())


for _ in range(2):
    fwd(); bwd()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fwd(); bwd()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(reps):
    ev[0].record(); fwd(); ev[1].record(); bwd(); ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
print(f'{which} encoder alone: forward {tf / reps * 1e3:.1f} us, backward {tb / reps * 1e3:.1f} us')
