"""Device-side ``get_eval``, drop-in for the reference's ``lib/eval_helper.get_eval(data_dict,
config)`` (lib/eval_helper.py:11-114): same signature, same dict keys written (``lang_acc``,
``ref_acc``, ``ref_iou``, ``ref_iou_rate_0.25/0.5``, ``ref_multiple_mask``, ``ref_others_mask``,
``pred_bboxes``, ``gt_bboxes``).  Call it after ``get_loss`` (it needs ``cluster_label``), like the
reference's solver does (lib/solver.py:207-216).

The reference loops over the scenes on the host with a D2H sync per scene (``torch.argmax`` → numpy
indexing) ; here one kernel (``ir_ref_eval``) scores every scene and the results come back in ONE
copy."""
import torch

from . import ops
from .loss_helper import BOXES, pack_boxes


def get_eval(data_dict, config):
    dev = data_dict['lang_scores'].device
    lang_pred = torch.argmax(data_dict['lang_scores'], dim=1)
    data_dict['lang_acc'] = (lang_pred == data_dict['object_cat'].to(dev)).float().mean()
    bx = data_dict.get(BOXES) or pack_boxes(data_dict, config, dev)
    label = data_dict.get('_ir_cluster_label_flat')
    if label is None:
        cl = data_dict['cluster_label']
        label = torch.cat([torch.as_tensor(c, dtype=torch.float32, device=dev).reshape(-1) for c in cl] or
                          [torch.zeros(1, device=dev)])
    B = len(bx['counts'])
    pred_idx, ref_acc, iou, pc, gc = ops.ref_eval(bx['pred'], bx['obb_ofs'], bx['gt'], bx['score_ofs'],
                                                  data_dict['attribute_scores'].detach(), data_dict['relation_scores'].detach(),
                                                  data_dict['scene_scores'].detach(), label)
    host = torch.cat([iou, ref_acc.double(), pc.reshape(-1), gc.reshape(-1)]).cpu().numpy()      # one D2H
    ious = host[:B]
    data_dict['ref_acc'] = [float(v) for v in host[B:2 * B]]
    data_dict['ref_iou'] = [float(v) for v in ious]
    data_dict['ref_iou_rate_0.25'] = float((ious >= 0.25).sum()) / B
    data_dict['ref_iou_rate_0.5'] = float((ious >= 0.5).sum()) / B
    data_dict['pred_bboxes'] = list(host[2 * B:2 * B + 24 * B].reshape(B, 8, 3))
    data_dict['gt_bboxes'] = list(host[2 * B + 24 * B:].reshape(B, 8, 3))
    data_dict['ref_pred_index'] = pred_idx
    if 'unique_multiple' in data_dict:
        data_dict['ref_multiple_mask'] = [int(v) for v in data_dict['unique_multiple'].detach().cpu().tolist()]
    data_dict['ref_others_mask'] = [1 if int(c) == 17 else 0 for c in data_dict['object_cat'].detach().cpu().tolist()]
    return data_dict
