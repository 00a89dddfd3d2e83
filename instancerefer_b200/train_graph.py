"""One training iteration of the reference's solver (lib/solver.py:195-205: forward, get_loss, backward, optimiser
step) replayed from ONE CUDA graph (B200-first: explicit stream capture instead of a tracing compiler).

Eagerly the step is host-bound: ~600 launches behind ~45 library calls, 21 autograd nodes and the optimiser, 6.4 ms of
Python for 5.5 ms of GPU work (profiles/r2_train_host_phases.log).  ``GraphedTrainStep(model, optimizer, config)(batch)``
keeps on the host only what cannot be captured — the class filter over the Python lists, packing the instances and the
boxes into pinned buffers, async H2D copies into STATIC device buffers, three per-step scalars of Adam and the dropout
step counter — and replays everything else: train-mode forward of all four modules, the three loss terms with their
gradients, the whole backward (torch's autograd engine runs inside the capture; every node is a call into the CUDA
library), the bucketed gradient all-reduce (NCCL, captured on its own stream beside the backward) and the fused Adam.

What makes the step capturable (each is a deviation from the eager path, none changes the arithmetic):
  * capacity mode (``_ir_capacity``): both sparse encoders lay their buffers out for the workspace capacity and read
    the level row counts on the device, so the forward has no read-back (eager: one D2H of ten counts);
  * Adam's lr / bias corrections come from a device buffer (``FlatAdam.stage_step`` -> ``ir_adam_step_dev``), computed
    on the host exactly as ``ir_adam_step`` computes them, so a scheduler (lib/solver.py:119-125) still works;
  * dropout folds a device-side step counter into its seed (``ir_dropout_seed_step``): a fresh mask per replay.

Graphs are cached by shape signature: scenes, token count, per-scene instance and candidate counts, lidar row bucket,
optimiser constants.  A signature is run EAGERLY the first ``min_hits`` times it is seen (that also warms every lazy
allocation) and captured on the next; batches whose signature never repeats therefore cost exactly the eager step.
With the reference's loader (32 scenes per batch, free instance counts) signatures rarely repeat — the graph pays off
for fixed-shape feeds (the bench's BASELINE configs[2], bucketed / padded loaders)."""
import numpy as np
import torch

from . import ops
from .candidates import KEY, CandidatePack
from .loss_helper import BOXES, _gt_obb, boxes_host, boxes_views, get_loss, stash_host_labels
from .sparse_tensor import SparseTensor

DEVICE_KEYS = ('lang_feat', 'lang_len', 'object_cat', 'point_min', 'point_max', 'ref_center_label')
RESULT_KEYS = ('loss', 'ref_loss', 'lang_loss', 'seg_loss', 'seg_acc')


class _Slot:
    pass


class StepResult:
    """The five scalars of an iteration (loss, ref_loss, lang_loss, seg_loss, seg_acc) on their way to the host: one
    async D2H copy into a pinned ring entry queued right behind the step.  ``get()`` waits for THAT copy only, so a
    training loop can log iteration i after it has queued iteration i+1 and the host never idles the GPU."""

    def __init__(self, host, event):
        self._host, self._event, self._vals = host, event, None

    def get(self):
        if self._vals is None:
            self._event.synchronize()
            self._vals = dict(zip(RESULT_KEYS, self._host.tolist()))
        return self._vals

    def loss(self):
        return self.get()['loss']


class GraphedTrainStep:
    def __init__(self, model, optimizer, config, min_hits=1, max_graphs=8, depth=2):
        a = model.args
        if not (a.use_gt_lang and a.attribute_module and a.relation_module and a.scene_module):
            raise NotImplementedError("graph replay of the training step needs the full model with use_gt_lang: True "
                                      "(the class filter runs on the host before the language branch)")
        self.model, self.opt, self.config = model, optimizer, config
        self.min_hits, self.max_graphs = max(1, min_hits), max_graphs    # first sight is always eager: it warms lazy state
        self.depth = max(1, depth)       # slots (static inputs + graph) per signature, used alternately: the H2D copies of
                                         # iteration i+1 run on a copy stream beside the replay of iteration i
        self.cache, self.hits = {}, {}
        self.replays = self.eager_steps = self.launches_replayed = 0
        self.profile = None              # set to a list: (host seconds of staging, event before, event after the replay)
        self.h2d_bytes = 0
        self._seed_dev = None
        self._bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        # Everything — eager first sights, staging, capture, replay — runs on ONE private stream: autograd pins each
        # parameter's AccumulateGrad node to the stream of the iteration that created it, and a node living on the
        # legacy default stream (which cannot capture) would invalidate the capture of the backward.
        self.stream = torch.cuda.Stream(device=optimizer.flat.device)
        self.copy_stream = torch.cuda.Stream(device=optimizer.flat.device)
        self._ring = torch.zeros(8, len(RESULT_KEYS), dtype=torch.float32).pin_memory()
        self._ring_ev = [None] * 8
        self._n_results = 0

    # ------------------------------------------------------------------ host side of every step
    @staticmethod
    def _lidar(batch):
        if 'lidar' in batch:
            return batch['lidar'].F, batch['lidar'].C
        return batch['lidar_feats'], batch['lidar_coords']

    def _signature(self, batch, target):
        cls = batch['instance_class']
        n_c = tuple(sum(1 for c in cl if int(c) == int(target[i])) for i, cl in enumerate(cls))
        F, _ = self._lidar(batch)
        pts0 = batch['instance_points'][0][0]
        return (tuple(batch['lang_feat'].shape), int(batch['lang_len'].max()), tuple(len(cl) for cl in cls), n_c,
                tuple(pts0.shape), ops.round_rows(F.shape[0]), F.shape[1], self.opt.graph_key(),
                tuple(m.momentum for m in self._bns),          # launch arguments: a BN-momentum schedule re-captures
                tuple(tuple(batch[k].shape) + (str(batch[k].dtype),) for k in DEVICE_KEYS))

    def _eager(self, batch):
        dev = self.opt.flat.device
        d = stash_host_labels(dict(batch))
        d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
        if 'lidar' not in d:
            d['lidar'] = SparseTensor(d.pop('lidar_feats'), d.pop('lidar_coords'))
        else:
            d['lidar'] = SparseTensor(d['lidar'].F.to(dev, non_blocking=True), d['lidar'].C.to(dev, non_blocking=True))
        self.opt.zero_grad()
        d = get_loss(self.model(d), self.config)
        d['loss'].backward()
        self.opt.step()
        self.eager_steps += 1
        # like a replayed step: results without autograd history (a caller holding the dict would otherwise keep this
        # iteration's AccumulateGrad nodes — pinned to this stream — alive into the capture of the next one)
        d = {k: (v.detach() if torch.is_tensor(v) else [t.detach() for t in v] if k == 'cluster_label' else v) for k, v in d.items()}
        d['result'] = self._result(d)
        return d

    def _new_slot(self, batch, target, tag):
        dev = self.opt.flat.device
        s = _Slot()
        s.tag = tag
        s.dev = {k: torch.empty(tuple(batch[k].shape), dtype=batch[k].dtype, device=dev) for k in DEVICE_KEYS}
        F, C = self._lidar(batch)
        rows = ops.round_rows(F.shape[0])
        s.lidar_F = torch.zeros(rows, F.shape[1], dtype=torch.float32, device=dev)
        s.lidar_C = torch.zeros(rows, 4, dtype=torch.int32, device=dev)
        s.n0_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        s.n0_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        pack0 = CandidatePack(batch, target, dev)
        s.pack_static = {k: torch.empty_like(v) for k, v in pack0.static_buffers().items()}
        B = len(batch['instance_class'])
        s.scene_ofs = pack0.scene_ofs().clone()                        # a function of the signature: constant per slot
        f64, i32, counts, obb_ofs = boxes_host(pack0.pred_obb_batch, np.zeros((B, 7)))
        s.box_host = torch.zeros(f64.size, dtype=torch.float64).pin_memory()
        s.box_dev = torch.zeros(f64.size, dtype=torch.float64, device=dev)
        s.box_ints = torch.from_numpy(i32).to(dev)                     # offsets: constant per slot as well
        s.boxes = dict(boxes_views(s.box_dev, s.box_ints, counts, obb_ofs), static=True)
        s.staged, s.graph_done = torch.cuda.Event(), torch.cuda.Event()
        s.staged.record()
        s.graph_done.record()
        s.graph = None
        return s

    def _stage(self, s, batch, target):
        """Host filter + pack + every H2D copy of the step into the slot's static buffers, on the copy stream: they
        wait for the slot's previous replay (which read these buffers) and run beside whatever the step stream does."""
        dev = self.opt.flat.device
        s.staged.synchronize()                       # earlier copies out of this slot's pinned buffers are done
        d = stash_host_labels(dict(batch))
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            cs.wait_event(s.graph_done)
            pack = CandidatePack(d, target, dev, static=s.pack_static, tag=s.tag)
            pack._scene_ofs = s.scene_ofs
            pack.resident = True
            for k in DEVICE_KEYS:
                s.dev[k].copy_(batch[k], non_blocking=True)
            F, C = self._lidar(batch)
            n0 = F.shape[0]
            s.lidar_F[:n0].copy_(F, non_blocking=True)
            s.lidar_C[:n0].copy_(C, non_blocking=True)
            s.n0_host[0] = n0
            s.n0_dev.copy_(s.n0_host, non_blocking=True)
            f64, _, _, _ = boxes_host(pack.pred_obb_batch, _gt_obb(d, self.config))
            s.box_host.numpy()[:] = f64
            s.box_dev.copy_(s.box_host, non_blocking=True)
            s.staged.record(cs)
        # per-step scalars read by the replay itself: on the step stream, behind the previous replay
        if self._seed_dev is None:
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._seed_host = torch.zeros(16, dtype=torch.int64).pin_memory()
            self._seed_n = 0
        self._seed_n += 1
        h = self._seed_host[self._seed_n % 16:self._seed_n % 16 + 1]
        h[0] = self._seed_n
        self._seed_dev.copy_(h, non_blocking=True)
        self.h2d_bytes = (pack.h2d_bytes + sum(t.numel() * t.element_size() for t in s.dev.values()) +
                          n0 * (F.shape[1] + 4) * 4 + 4 + f64.size * 8 + 8 + 16)
        return pack, d

    def _result(self, d):
        """Queue the D2H copy of the step's scalars (current stream) -> StepResult."""
        res = d.get('_ir_result')
        if res is None:
            res = torch.stack([d[k].detach().reshape(-1)[0].float() for k in RESULT_KEYS])
        i = self._n_results % 8
        self._n_results += 1
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()           # the entry's previous copy landed (8 iterations ago)
        host = self._ring[i]
        host.copy_(res, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[i] = ev
        return StepResult(host, ev)

    def _device_step(self, s, pack, lmax):
        d = dict(s.dev)
        d['lidar'] = SparseTensor(s.lidar_F, s.lidar_C)
        d['_ir_lidar_rows'] = s.n0_dev
        d['_ir_lang_len_max'] = lmax
        d['_ir_capacity'] = True
        d[KEY] = pack
        d[BOXES] = s.boxes
        d = get_loss(self.model(d), self.config)
        ops.stamp('loss')
        d['loss'].backward()
        ops.stamp('bwd:joined')
        self.opt.step()
        ops.stamp('adam')
        d['_ir_result'] = torch.stack([d[k].detach().reshape(-1)[0].float() for k in RESULT_KEYS])
        return d

    # ------------------------------------------------------------------ public API
    def __call__(self, batch):
        """batch: the HOST dict of one iteration (CPU tensors, pinned for async copies; 'lidar' SparseTensor or
        lidar_feats / lidar_coords; instance_* lists) -> data_dict with the keys get_loss writes (device tensors;
        for a replayed step they are views of the slot's static buffers, valid until its next replay)."""
        if not self.model.training:
            raise RuntimeError("GraphedTrainStep needs model.train()")
        ops.check_device()
        caller = torch.cuda.current_stream(self.opt.flat.device)
        self.stream.wait_stream(caller)
        with torch.cuda.stream(self.stream):
            out = self._step(batch)
        caller.wait_stream(self.stream)              # the caller's stream sees the step's results and new weights
        return out

    def _step(self, batch):
        tgt = batch['object_cat']
        target = (tgt.detach().to('cpu') if tgt.is_cuda else tgt).tolist()
        key = self._signature(batch, target)
        e = self.cache.get(key)
        if e is None:
            n = self.hits.get(key, 0)
            if n < self.min_hits:
                if len(self.hits) > 4096:
                    self.hits.clear()
                self.hits[key] = n + 1
                return self._eager(batch)
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            e = dict(slots=[self._new_slot(batch, target, f't{len(self.cache)}s{j}') for j in range(self.depth)], i=0)
            self.cache[key] = e
        s = e['slots'][e['i'] % self.depth]
        e['i'] += 1
        import time
        t0 = time.perf_counter()
        pack, host = self._stage(s, batch, target)
        self.opt.stage_step()
        t_stage = time.perf_counter() - t0
        if s.graph is None:
            ops.dropout_seed_step(self._seed_dev)
            self.opt.zero_grad()                     # .grad = None: the captured backward creates, never accumulates
            import gc
            gc.collect()                             # autograd graphs of earlier iterations still waiting for the collector
            torch.cuda.synchronize()                 # would pin AccumulateGrad nodes to the stream they ran on
            g = torch.cuda.CUDAGraph()
            from . import _lib
            c0 = _lib.load().ir_launch_count()
            try:
                with torch.cuda.graph(g, stream=self.stream):
                    s.out = self._device_step(s, pack, key[1])
            finally:
                ops.dropout_seed_step(None)
            s.launches = _lib.load().ir_launch_count() - c0          # this library's kernels inside one replay
            # results without the capture's autograd graph (its nodes would otherwise live as long as the slot)
            s.out = {k: (v.detach() if torch.is_tensor(v) else [t.detach() for t in v] if k == 'cluster_label' else v)
                     for k, v in s.out.items()}
            # buffers allocated outside the capture that the graph points into must outlive it
            s.keep = [dict(n.__dict__.get('_train_graphs', {})) for n in (self.model.attribute.net, self.model.scene.net)]
            s.graph = g
            self.opt._staged = True                  # the capture consumed the flag, the replay reads the same scalars
        self.stream.wait_event(s.staged)
        if self.profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.graph.replay()
            e1.record()
            self.profile.append((t_stage, e0, e1))
        else:
            s.graph.replay()
        s.graph_done.record()
        self.opt.after_replay()
        self.replays += 1
        self.launches_replayed += s.launches
        out = dict(s.out)
        out['result'] = self._result(out)
        out['num_filtered_objs'] = pack.num_filtered
        out['pred_obb_batch'] = pack.pred_obb_batch
        out['_ir_host_labels'] = host.get('_ir_host_labels', {})
        return out
