"""Host-side candidate bookkeeping shared by the attribute / relation / scene modules.

The reference filters instances by class in Python loops that compare host ints against a CUDA
tensor element by element (one D2H sync per instance: models/attribute_module.py:60,
models/relation_module.py:74).  Here the target classes are read back once, the per-scene candidate
lists are built from the host lists, and every instance of the participating scenes is packed into
ONE pinned buffer and copied to the device once per forward; both modules index into it."""
import numpy as np
import torch

KEY = '_ir_candidates'


class CandidatePack:
    """cands[i]      : instance ids of scene i with class == target (kept even if < 2)
    active        : scenes with >= 2 candidates, in order (others are skipped in the score vectors)
    inst_ofs      : (len(active)+1,) row offsets of each active scene's instances in the packed buffer
    cand_rows     : (M,) packed-buffer rows of the candidates, scene-major then instance order
    cand_scene    : (M,) original scene index of each candidate;  cand_seg: index into `active`
    """

    def __init__(self, data_dict, lang_cls_pred, device):
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
            centres += [np.asarray(o[:3], np.float64) for o in data_dict['instance_obbs'][i]]
            cls += [int(c) for c in classes[i]]
            inst_ofs.append(base + len(classes[i]))
            cand_rows += [base + j for j in self.cands[i]]
            cand_scene += [i] * len(self.cands[i])
            cand_seg += [seg] * len(self.cands[i])
        self.n_inst, self.M = inst_ofs[-1], len(cand_rows)
        ppi, fdim = pts[0].shape
        host = torch.empty((self.n_inst, ppi, fdim), dtype=torch.float32, pin_memory=True)
        np.stack(pts, 0, out=host.numpy())
        self.points = host.to(device, non_blocking=True)                       # one H2D for all instances
        meta = np.zeros((self.n_inst, 4), np.float32)
        meta[:, :3] = np.asarray(centres, np.float64).astype(np.float32)
        meta[:, 3] = np.asarray(cls, np.float32)
        self.centres_cls = torch.from_numpy(meta).pin_memory().to(device, non_blocking=True)
        ints = np.concatenate([np.asarray(inst_ofs, np.int32), np.asarray(cand_rows, np.int32),
                               np.asarray(cand_scene, np.int32), np.asarray(cand_seg, np.int32)])
        ints = torch.from_numpy(ints).pin_memory().to(device, non_blocking=True)
        a = len(inst_ofs)
        self.inst_ofs = ints[:a]
        self.cand_rows = ints[a:a + self.M]
        self.cand_scene = ints[a + self.M:a + 2 * self.M]
        self.cand_seg = ints[a + 2 * self.M:a + 3 * self.M]
        self.h2d_bytes = host.numel() * 4 + meta.nbytes + ints.numel() * 4


def get_pack(data_dict, args, device, rebuild=False):
    """Built (always fresh) by the attribute module, reused by relation / scene in the same forward."""
    pack = None if rebuild else data_dict.get(KEY)
    if pack is None:
        if not args.use_gt_lang:
            pred = torch.argmax(data_dict['lang_scores'], dim=1)
        else:
            pred = data_dict['object_cat']
        pack = CandidatePack(data_dict, pred, device)
        data_dict[KEY] = pack
    return pack
