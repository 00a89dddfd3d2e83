"""CUDA-graph replay of the whole forward (B200-first: no tracing compiler, explicit stream capture).

``GraphedInstanceRefer(model)(data_dict)`` has the contract of ``InstanceRefer.forward`` (eval mode,
``use_gt_lang: True``).  Per step the host does what cannot be captured — the class filter over the
Python lists, packing every instance into one pinned buffer, async H2D copies into STATIC device
buffers — and then replays one captured graph: ~120 kernels on four streams (language branch,
instance encoder, scene encoder, relation graph) with every row / pair count kept on the device.

Graphs are cached by shape signature (B, max tokens, #instances, #candidates, #active scenes, lidar
row bucket).  Outputs are views of static buffers: they stay valid until the next call with the same
signature (clone them to keep them)."""
import torch

from . import ops
from .candidates import KEY, CandidatePack, target_classes
from .sparse_tensor import SparseTensor

OUT_KEYS = ('lang_feat', 'atten_attr', 'atten_rel', 'atten_scene', 'lang_cls_feats', 'lang_attr_feats',
            'lang_rel_feats', 'lang_scene_feats', 'lang_scores', 'obj_feats', 'attribute_scores',
            'relation_scores', 'scene_scores', 'seg_scores', 'vis_atten', 'ref_probs', 'ref_pred')


class _Entry:
    pass


class GraphedInstanceRefer:
    def __init__(self, model, max_graphs=8):
        if model.training:
            raise NotImplementedError("graph replay is an eval-mode path")
        self.model = model
        self.max_graphs = max_graphs
        self.cache = {}
        self.h2d_bytes = 0

    # --- static input staging ------------------------------------------------------------
    def _stage(self, e, data_dict, lmax_host):
        m = self.model
        dev = e.device
        pack = CandidatePack(data_dict, data_dict['_ir_target_host'], dev, static=e.pack_static)
        e.lang_feat.copy_(data_dict['lang_feat'], non_blocking=True)
        e.lang_len.copy_(data_dict['lang_len'], non_blocking=True)
        e.point_min.copy_(data_dict['point_min'], non_blocking=True)
        lid = data_dict['lidar']
        n0 = lid.F.shape[0]
        e.lidar_F[:n0].copy_(lid.F, non_blocking=True)
        e.lidar_C[:n0].copy_(lid.C, non_blocking=True)
        e.n0_host[0] = n0
        e.n0_dev.copy_(e.n0_host, non_blocking=True)
        self.h2d_bytes = (pack.h2d_bytes + e.lang_feat.numel() * 4 + e.lang_len.numel() * 8 +
                          e.point_min.numel() * e.point_min.element_size() + n0 * (lid.F.shape[1] + 4) * 4 + 4)
        return pack

    def _new_entry(self, data_dict, pack, lmax, key):
        m = self.model
        dev = pack.points.device
        e = _Entry()
        e.device = dev
        B = data_dict['lang_feat'].shape[0]
        lid = data_dict['lidar']
        rows = ops.round_rows(lid.F.shape[0])
        e.pack_static = {k: torch.empty_like(v) for k, v in pack.static_buffers().items()}
        e.lang_feat = torch.empty(data_dict['lang_feat'].shape, dtype=torch.float32, device=dev)
        e.lang_len = torch.empty(B, dtype=torch.int64, device=dev)
        e.point_min = torch.empty(tuple(data_dict['point_min'].shape), dtype=data_dict['point_min'].dtype, device=dev)
        e.lidar_F = torch.zeros(rows, lid.F.shape[1], dtype=torch.float32, device=dev)
        e.lidar_C = torch.zeros(rows, 4, dtype=torch.int32, device=dev)
        e.n0_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        e.n0_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        e.lmax = lmax
        return e

    def _static_dict(self, e, pack):
        d = dict(lang_feat=e.lang_feat, lang_len=e.lang_len, point_min=e.point_min,
                 lidar=SparseTensor(e.lidar_F, e.lidar_C), _ir_lidar_rows=e.n0_dev, _ir_lang_len_max=e.lmax)
        pack.resident = True
        d[KEY] = pack
        return d

    def __call__(self, data_dict):
        m = self.model
        a = m.args
        if not a.use_gt_lang:
            raise NotImplementedError("graph replay needs use_gt_lang: True (the class filter runs on the host "
                                      "before the language branch)")
        # the two tiny D2H reads the host-side filter / shapes need (object_cat, max token count)
        tgt = data_dict['object_cat']
        ln = data_dict['lang_len']
        data_dict['_ir_target_host'] = tgt.detach().to('cpu') if tgt.is_cuda else tgt
        lmax = int((ln.detach().to('cpu') if ln.is_cuda else ln).max())
        lid = data_dict['lidar']
        B = data_dict['lang_feat'].shape[0]
        # peek at the candidate structure for the signature (cheap: class lists only)
        pred = data_dict['_ir_target_host'].tolist()
        n_c = [sum(1 for c in cl if int(c) == int(pred[i])) for i, cl in enumerate(data_dict['instance_class'])]
        act = [i for i, n in enumerate(n_c) if n >= 2]
        key = (B, lmax, sum(len(data_dict['instance_class'][i]) for i in act), sum(n_c[i] for i in act), len(act),
               ops.round_rows(lid.F.shape[0]), torch.cuda.current_device())
        e = self.cache.get(key)
        dev = torch.device('cuda', torch.cuda.current_device())
        if e is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            pack0 = CandidatePack(data_dict, data_dict['_ir_target_host'], dev)
            e = self._new_entry(data_dict, pack0, lmax, key)
            pack = self._stage(e, data_dict, lmax)
            # warm-up on a side stream (allocations, lazy attribute sets), then capture
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):
                    m(self._static_dict(e, pack))
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            e.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e.graph):
                e.out = m(self._static_dict(e, pack))
            self.cache[key] = e
        else:
            pack = self._stage(e, data_dict, lmax)
        e.graph.replay()
        for k in OUT_KEYS:
            if k in e.out:
                data_dict[k] = e.out[k]
        data_dict['num_filtered_objs'] = pack.num_filtered
        data_dict['pred_obb_batch'] = pack.pred_obb_batch
        return data_dict
