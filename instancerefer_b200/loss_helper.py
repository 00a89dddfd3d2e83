"""Device-side ``get_loss``, drop-in for the reference's ``lib/loss_helper.get_loss(data_dict,
config)`` (lib/loss_helper.py:196-269): same signature, same dict keys written
(``loss``, ``ref_loss``, ``lang_loss``, ``seg_loss``, ``seg_acc``, ``cluster_label``).

The reference loops over scenes on the host (numpy corner boxes, IoU, argmax, per-scene
ContrastiveLoss) with five ``.cpu()`` round trips; here the candidate boxes go to the device in one
copy and three kernels (``ir_cross_entropy`` x2, ``ir_region_label``, ``ir_ref_loss``) produce every
loss term together with its gradient, so the backward of the loss is a multiply by precomputed
gradients (no graph of small torch ops)."""
import numpy as np
import torch
from torch.autograd import Function

from . import ops


class _CrossEntropy(Function):
    @staticmethod
    def forward(ctx, logits, labels):
        loss, dl = ops.cross_entropy(logits.contiguous(), labels.contiguous())
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None


class _RefLoss(Function):
    """sum_b ContrastiveLoss_b / B with the one-hot labels from the IoU arg-max."""

    @staticmethod
    def forward(ctx, sa, sr, ss, pred_obb, obb_ofs, gt_obb, score_ofs):
        B = gt_obb.shape[0]
        label, loss_scene, dscore, iou_max = ops.ref_loss(pred_obb, obb_ofs, gt_obb, score_ofs, sa.contiguous(),
                                                          sr.contiguous(), ss.contiguous())
        ctx.save_for_backward(dscore)
        ctx.B = B
        ctx.mark_non_differentiable(label, iou_max)
        return loss_scene.sum().reshape(1) / B, label, iou_max

    @staticmethod
    def backward(ctx, g, _gl, _gi):
        (dscore,) = ctx.saved_tensors
        d = dscore * (g / ctx.B)
        return d, d, d, None, None, None, None


HOST_LABELS = '_ir_host_labels'
LABEL_KEYS = ('ref_center_label', 'ref_heading_class_label', 'ref_heading_residual_label', 'ref_size_class_label',
              'ref_size_residual_label')


def stash_host_labels(data_dict):
    """Call BEFORE moving a batch to the GPU: keeps the host copies of the five box-label tensors, so that
    get_loss / get_eval can build the ground-truth boxes (a host numpy function of the dataset config) without
    reading them back — each read-back would stall the host until the whole forward has executed."""
    data_dict[HOST_LABELS] = {k: data_dict[k] for k in LABEL_KEYS + ('lang_len',) if k in data_dict and not data_dict[k].is_cuda}
    return data_dict


def _gt_obb(data_dict, config):
    """config.param2obb_batch on the label tensors (lib/loss_helper.py:212-219), host numpy like the
    reference; labels that only exist on the device are read back (D2H, synchronising)."""
    host = data_dict.get(HOST_LABELS, {})
    t = lambda k: (host[k] if k in host else data_dict[k].detach().cpu()).numpy()
    return np.asarray(config.param2obb_batch(t('ref_center_label'), t('ref_heading_class_label'),
                                             t('ref_heading_residual_label'), t('ref_size_class_label'),
                                             t('ref_size_residual_label')), np.float64)


BOXES = '_ir_boxes'       # packed candidate / ground-truth boxes on the device, shared by get_loss and get_eval


def boxes_host(pred, gt):
    """pred: host list of (c_b,7) float64 candidate boxes per scene, gt (B,7) -> (f64 [all pred | gt] flat,
    i32 [obb_ofs (B+1) | score_ofs (B)], counts, obb_ofs): what goes to the device, as two host arrays."""
    counts = [int(np.asarray(p).reshape(-1, 7).shape[0]) if len(p) else 0 for p in pred]
    obb_ofs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    score_ofs, s = [], 0
    for c in counts:
        score_ofs.append(s if c >= 2 else -1)
        s += c if c >= 2 else 0
    allobb = np.concatenate([np.asarray(p, np.float64).reshape(-1, 7) for p in pred if len(p)] or [np.zeros((1, 7))], 0)
    f64 = np.concatenate([allobb.reshape(-1), np.asarray(gt, np.float64).reshape(-1)])
    i32 = np.concatenate([obb_ofs, np.asarray(score_ofs, np.int32)])
    return f64, i32, counts, obb_ofs


def boxes_views(devbuf, ints, counts, obb_ofs):
    B, n = len(counts), devbuf.numel() - 7 * len(counts)
    return dict(pred=devbuf[:n].view(-1, 7), gt=devbuf[n:].view(B, 7), obb_ofs=ints[:B + 1], score_ofs=ints[B + 1:],
                counts=counts, host_ofs=obb_ofs)


def pack_boxes(data_dict, config, dev):
    """pred_obb_batch (host list of (c_b,7) float64) + the GT boxes -> one H2D copy.
    -> dict(pred (Nc,7) f64, gt (B,7) f64, obb_ofs (B+1) i32, score_ofs (B) i32 [-1: scene not scored], counts).
    A dict already staged by the caller (``static``: train_graph.GraphedTrainStep) is used as it is."""
    bx = data_dict.get(BOXES)
    if bx is not None and bx.get('static'):
        return bx
    f64, i32, counts, obb_ofs = boxes_host(data_dict['pred_obb_batch'], _gt_obb(data_dict, config))
    return boxes_views(torch.from_numpy(f64).to(dev), torch.from_numpy(i32).to(dev), counts, obb_ofs)


def get_loss(data_dict, config):
    dev = data_dict['lang_scores'].device
    lang_loss = _CrossEntropy.apply(data_dict['lang_scores'], data_dict['object_cat'].to(dev).long())[0]
    data_dict['lang_loss'] = lang_loss
    seg_label = ops.region_label(data_dict['ref_center_label'].to(dev), data_dict['point_min'].to(dev),
                                 data_dict['point_max'].to(dev))
    seg_scores = data_dict['seg_scores']
    seg_loss = _CrossEntropy.apply(seg_scores, seg_label)[0]
    seg_acc = (seg_scores.detach().argmax(1) == seg_label).sum() / float(seg_label.numel())
    bx = pack_boxes(data_dict, config, dev)
    data_dict[BOXES] = bx
    B, obb_ofs = len(bx['counts']), bx['host_ofs']
    ref_loss, label, iou_max = _RefLoss.apply(data_dict['attribute_scores'], data_dict['relation_scores'],
                                              data_dict['scene_scores'], bx['pred'], bx['obb_ofs'], bx['gt'], bx['score_ofs'])
    data_dict['_ir_cluster_label_flat'] = label
    data_dict['ref_loss'] = ref_loss
    data_dict['loss'] = 10 * ref_loss + lang_loss + seg_loss                    # (lib/loss_helper.py:263)
    data_dict['seg_loss'] = seg_loss
    data_dict['seg_acc'] = seg_acc
    data_dict['cluster_label'] = [label[obb_ofs[i]:obb_ofs[i + 1]] for i in range(B)]
    data_dict['ref_iou_max'] = iou_max
    return data_dict
