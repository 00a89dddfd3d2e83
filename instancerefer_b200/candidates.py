"""Host-side candidate bookkeeping shared by the attribute / relation / scene modules.

The reference filters instances by class in Python loops that compare host ints against a CUDA
tensor element by element (one D2H sync per instance: models/attribute_module.py:60,
models/relation_module.py:74).  Here the target classes are read back once, the per-scene candidate
lists are built from the host lists, and every instance of the participating scenes is packed into
ONE pinned buffer and copied to the device once per forward; all three modules index into it."""
import numpy as np
import torch

KEY = '_ir_candidates'

_pinned = {}
_last_h2d = {}              # tag -> event recorded after the latest copies out of that tag's staging buffers


def _pinned_buf(name, shape, dtype):
    """Reusable pinned staging buffers (pinning per call would cost more than the copy)."""
    key = (name, tuple(shape), dtype)
    buf = _pinned.get(key)
    if buf is None:
        if len(_pinned) > 64:
            _pinned.clear()
        buf = torch.empty(shape, dtype=dtype).pin_memory()
        _pinned[key] = buf
    return buf


class CandidatePack:
    """cands[i]      : instance ids of scene i with class == target (kept even if < 2)
    active        : scenes with >= 2 candidates, in order (others are skipped in the score vectors)
    points        : (n_inst, ppi, fdim) fp32, every instance of the active scenes (device)
    centres_cls   : (n_inst, 4) fp32 [obb centre xyz, class id]
    inst_ofs      : (len(active)+1,) row offsets of each active scene's instances in `points`
    cand_rows     : (M,) rows of the candidates in `points`, scene-major then instance order
    cand_scene    : (M,) original scene index of each candidate;  cand_seg: index into `active`
    cand_ofs      : (len(active)+1,) candidate offsets per active scene
    ``static``: dict of preallocated device buffers (CUDA-graph replay) to copy into instead of
    allocating; ``tag`` selects a private set of pinned staging buffers (one per in-flight slot)."""

    def __init__(self, data_dict, lang_cls_pred, device, static=None, tag=''):
        classes = data_dict['instance_class']
        pred = lang_cls_pred.detach().to('cpu').tolist() if torch.is_tensor(lang_cls_pred) else list(lang_cls_pred)
        self.cands, self.pred_obb_batch, self.num_filtered = [], [], []
        for i, cl in enumerate(classes):
            tgt = int(pred[i])
            ids = [j for j, c in enumerate(cl) if int(c) == tgt]
            self.cands.append(ids)
            self.num_filtered.append(len(ids))
            self.pred_obb_batch.append(np.asarray([data_dict['instance_obbs'][i][j] for j in ids]))
        self.active = [i for i, ids in enumerate(self.cands) if len(ids) >= 2]
        if not self.active:
            raise ValueError("no scene in the batch has >= 2 candidate instances (the reference forward "
                             "fails on such a batch too: models/attribute_module.py:101,125)")
        pts, centres, cls = [], [], []
        inst_ofs, cand_rows, cand_scene, cand_seg = [0], [], [], []
        for seg, i in enumerate(self.active):
            base = inst_ofs[-1]
            pts += list(data_dict['instance_points'][i])
            centres += [o[:3] for o in data_dict['instance_obbs'][i]]
            cls += [int(c) for c in classes[i]]
            inst_ofs.append(base + len(classes[i]))
            cand_rows += [base + j for j in self.cands[i]]
            cand_scene += [i] * len(self.cands[i])
            cand_seg += [seg] * len(self.cands[i])
        self.n_inst, self.M = inst_ofs[-1], len(cand_rows)
        ppi, fdim = pts[0].shape
        cand_ofs = np.concatenate([[0], np.cumsum([len(self.cands[i]) for i in self.active])])
        if _last_h2d.get(tag) is not None:
            _last_h2d[tag].synchronize()               # previous copies out of these staging buffers are done
        packed = data_dict.get('_ir_packed')
        if packed is not None and len(self.active) == len(classes) and packed['points'].shape[0] == self.n_inst:
            h_points = packed['points']             # loader.collate_packed: every instance already in one pinned buffer
        else:
            h_points = _pinned_buf('points' + tag, (self.n_inst, ppi, fdim), torch.float32)
            np.stack(pts, 0, out=h_points.numpy())
        h_meta = _pinned_buf('meta' + tag, (self.n_inst, 4), torch.float32)
        m = h_meta.numpy()
        m[:, :3] = np.asarray(centres, np.float64)
        m[:, 3] = cls
        a = len(inst_ofs)
        h_ints = _pinned_buf('ints' + tag, (2 * a + 3 * self.M,), torch.int32)
        h_ints.numpy()[:] = np.concatenate([inst_ofs, cand_rows, cand_scene, cand_seg, cand_ofs])
        if static is None:
            self.points = h_points.to(device, non_blocking=True)                # one H2D for all instances
            self.centres_cls = h_meta.to(device, non_blocking=True)
            ints = h_ints.to(device, non_blocking=True)
        else:
            self.points, self.centres_cls, ints = static['points'], static['centres_cls'], static['ints']
            self.points.copy_(h_points, non_blocking=True)
            self.centres_cls.copy_(h_meta, non_blocking=True)
            ints.copy_(h_ints, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _last_h2d[tag] = ev
        self.ints = ints
        self.inst_ofs = ints[:a]
        self.cand_rows = ints[a:a + self.M]
        self.cand_scene = ints[a + self.M:a + 2 * self.M]
        self.cand_seg = ints[a + 2 * self.M:a + 3 * self.M]
        self.cand_ofs = ints[a + 3 * self.M:]
        self.h2d_bytes = h_points.numel() * 4 + h_meta.numel() * 4 + h_ints.numel() * 4
        self.resident = False      # True: inputs pre-staged in HBM, reuse across forwards (bench `value`)

    def scene_ofs(self):
        """(B+1,) int32 device offsets of every scene's candidates in the score vectors (scenes with
        < 2 candidates own an empty range) — the contiguous row range of each partner row in the
        matching heads' backward."""
        if getattr(self, '_scene_ofs', None) is None:
            counts = [len(ids) if len(ids) >= 2 else 0 for ids in self.cands]
            self._scene_ofs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32,
                                           device=self.points.device)
        return self._scene_ofs

    def signature(self):
        return (self.n_inst, self.M, len(self.active), tuple(self.points.shape[1:]))

    def static_buffers(self):
        return dict(points=self.points, centres_cls=self.centres_cls, ints=self.ints)


def target_classes(data_dict, args):
    if not args.use_gt_lang:
        return torch.argmax(data_dict['lang_scores'], dim=1)      # (models/attribute_module.py:93-97)
    return data_dict['object_cat']


def get_pack(data_dict, args, device, rebuild=False):
    """Built (always fresh) by the attribute module, reused by relation / scene in the same forward."""
    pack = data_dict.get(KEY)
    if pack is not None and rebuild and not pack.resident:
        pack = None
    if pack is None:
        pack = CandidatePack(data_dict, target_classes(data_dict, args), device)
        data_dict[KEY] = pack
    return pack
