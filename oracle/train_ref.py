"""ORACLE (test infrastructure, never shipped): CPU restatement of one training iteration of the
reference — ``InstanceRefer.forward`` in train mode, ``get_loss`` and the backward pass
(lib/solver.py:196-205, lib/loss_helper.py:131-161,189-269, utils/box_util.py:154-198,310-333) —
plus the Adam update (torch.optim.Adam as built in scripts/train.py:93).

Forward = oracle/model_ref.py with ``train=True`` (batch-statistics BatchNorm); gradients come from
torch autograd over that CPU graph.  Dropout layers are taken with p = 0 (both sides of a parity
test use p = 0: the reference's masks come from torch's RNG stream and cannot be reproduced).

The ground-truth box is derived from the label tensors the way ``get_loss`` does through
``config.param2obb_batch`` (data/scannet/model_util_scannet.py:174-181) with ONE mean-size class of
zeros and ONE heading bin (class2size = residual, heading angle = residual): that is the
``SyntheticConfig`` below, which tests also hand to the reference's own ``get_loss`` when it is run
verbatim.  PARITY STATUS: pinned by tests/test_oracle.py::test_train_ref_matches_reference_live
(reference models + lib/loss_helper.get_loss executed verbatim over the shim, dev container only).
"""
import numpy as np
import torch
import torch.nn.functional as Fn

try:
    from . import model_ref
except ImportError:
    import model_ref


try:                                       # one definition, shared with the product's bench (which may not import oracle/)
    from instancerefer_b200.synthetic import SyntheticConfig
except ImportError:                        # oracle/ used standalone
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
    from instancerefer_b200.synthetic import SyntheticConfig


def box_min_max(obb):
    """get_3d_box_batch + get_box3d_min_max_batch (utils/box_util.py:181-198,310-333):
    8 corners (+-l/2, +-w/2, +-h/2) @ roty(heading)^T + centre -> per-axis min / max.  float64."""
    obb = np.asarray(obb, np.float64)
    l, w, h = obb[:, 3:4], obb[:, 4:5], obb[:, 5:6]
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]) * 0.5
    sy = np.array([1, -1, -1, 1, 1, -1, -1, 1]) * 0.5
    sz = np.array([1, 1, 1, 1, -1, -1, -1, -1]) * 0.5
    x, y, z = l * sx, w * sy, h * sz
    c, s = np.cos(obb[:, 6:7]), np.sin(obb[:, 6:7])
    xr, yr, zr = c * x + s * z, y, -s * x + c * z           # row vector times R^T, R = roty(t)
    corners = np.stack([xr, yr, zr], -1) + obb[:, None, 0:3]
    return corners.min(1), corners.max(1)


def iou_batch(pred_obb, gt_obb_row):
    """box3d_iou_batch of every predicted box against one GT box (utils/box_util.py:154-179)."""
    mn1, mx1 = box_min_max(pred_obb)
    mn2, mx2 = box_min_max(np.tile(np.asarray(gt_obb_row)[None], (pred_obb.shape[0], 1)))
    a = np.maximum(mn1, mn2)
    b = np.minimum(mx1, mx2)
    inter = np.prod(np.maximum(b - a, 0), axis=1)
    v1 = np.prod(mx1 - mn1, axis=1)
    v2 = np.prod(mx2 - mn2, axis=1)
    return inter / (v1 + v2 - inter + 1e-8)


def scene_region_label(ref_center, point_min, point_max):
    """9-way region label of compute_scene_mask_loss (lib/loss_helper.py:131-153)."""
    first = point_min + (point_max - point_min) / 3
    second = point_min + (point_max - point_min) / 3 * 2
    f = ref_center <= first
    s = ref_center <= second
    B = ref_center.shape[0]
    ones = torch.ones(B, dtype=torch.long)
    label = torch.where(f[:, 0] & f[:, 1], ones * 0, ones * 4)
    label = torch.where(~f[:, 0] & s[:, 0] & f[:, 1], ones, label)
    label = torch.where(~s[:, 0] & f[:, 1], ones * 2, label)
    label = torch.where(f[:, 0] & ~f[:, 1] & s[:, 0], ones * 3, label)
    label = torch.where(~s[:, 0] & ~f[:, 1] & s[:, 1], ones * 5, label)
    label = torch.where(f[:, 0] & ~s[:, 1], ones * 6, label)
    label = torch.where(~f[:, 0] & s[:, 0] & ~s[:, 1], ones * 7, label)
    label = torch.where(~s[:, 0] & ~s[:, 1], ones * 8, label)
    return label


def contrastive(score, label, margin=0.2, gamma=5):
    """ContrastiveLoss.forward (lib/loss_helper.py:93-107); the positive's slot contributes exp(0)
    to the logsumexp exactly as in the reference."""
    score = score * gamma
    sim = (score * label).sum()
    neg = torch.logsumexp(score * (1 - label), dim=0)
    return torch.clamp(neg - sim + margin, min=0).sum()


def get_loss(out, data, config=None):
    """lib/loss_helper.py:196-269 on the forward's outputs.  ``data`` carries the label tensors
    (ref_center_label, ref_*_label, point_min/max, object_cat).  -> dict of losses + labels."""
    config = config or SyntheticConfig()
    t = lambda k: torch.as_tensor(np.asarray(data[k]))
    lang_loss = Fn.cross_entropy(out['lang_scores'], t('object_cat').long())               # :189-193
    seg_label = scene_region_label(t('ref_center_label').float(), t('point_min').float(), t('point_max').float())
    seg_loss = Fn.cross_entropy(out['seg_scores'], seg_label)                               # :155
    gt = config.param2obb_batch(t('ref_center_label').numpy(), t('ref_heading_class_label').numpy(),
                                t('ref_heading_residual_label').numpy(), t('ref_size_class_label').numpy(),
                                t('ref_size_residual_label').numpy())                      # :219
    B = len(out['pred_obb_batch'])
    ref_loss = torch.zeros(1)
    start = 0
    cluster_label = []
    for i in range(B):
        pred = np.asarray(out['pred_obb_batch'][i])
        n = pred.shape[0]
        if n == 0:
            cluster_label.append(np.zeros(0))
            continue
        ious = iou_batch(pred, gt[i])
        label = np.zeros(n)
        label[ious.argmax()] = 1                                                            # :246
        cluster_label.append(label)
        if n == 1:
            continue
        score = (out['attribute_scores'][start:start + n] + out['relation_scores'][start:start + n]
                 + out['scene_scores'][start:start + n])
        start += n
        if ious.max() < 0.2:
            continue
        ref_loss = ref_loss + contrastive(score, torch.tensor(label, dtype=torch.float32))
    ref_loss = ref_loss / B                                                                 # :260
    loss = 10 * ref_loss + lang_loss + seg_loss                                             # :263
    return dict(loss=loss, ref_loss=ref_loss, lang_loss=lang_loss, seg_loss=seg_loss,
                seg_label=seg_label, cluster_label=cluster_label)


def train_step(sd, data, args, config=None):
    """One forward + loss + backward on CPU.  -> dict(loss terms, outputs, grads {key: tensor},
    bn_stats {bn prefix: (batch mean, unbiased batch var)})."""
    leaf = {k: (v.detach().clone().float().requires_grad_(True) if v.is_floating_point() and
                not k.endswith(('running_mean', 'running_var')) else v) for k, v in sd.items()}
    model_ref.BN_RECORD = {}
    try:
        out = model_ref.forward(leaf, data, args, train=True, keep_grad=True)
        stats = model_ref.BN_RECORD
    finally:
        model_ref.BN_RECORD = None
    L = get_loss(out, data, config)
    L['loss'].backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v))
             for k, v in leaf.items() if torch.is_tensor(v) and v.requires_grad}
    res = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in L.items()}
    res.update(outputs={k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()},
               grads=grads, bn_stats=stats)
    return res


def updated_running_stats(sd, bn_stats, momentum=0.1):
    """running = (1-m) running + m batch (unbiased var), num_batches_tracked += 1."""
    new = {}
    for p, (mean, var) in bn_stats.items():
        new[p + '.running_mean'] = (1 - momentum) * sd[p + '.running_mean'] + momentum * mean
        new[p + '.running_var'] = (1 - momentum) * sd[p + '.running_var'] + momentum * var
        new[p + '.num_batches_tracked'] = sd[p + '.num_batches_tracked'] + 1
    return new


def adam_update(params, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single step (scripts/train.py:93: Adam(lr, weight_decay)); L2 decay is added
    to the gradient, bias-corrected moments.  ``state`` = {key: (m, v)} and 'step'."""
    step = state.get('step', 0) + 1
    state['step'] = step
    out = {}
    for k, p in params.items():
        g = grads[k] + weight_decay * p
        m, v = state.get(k, (torch.zeros_like(p), torch.zeros_like(p)))
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        state[k] = (m, v)
        denom = v.sqrt() / (1 - betas[1] ** step) ** 0.5 + eps
        out[k] = p - lr / (1 - betas[0] ** step) * m / denom
    return out


def get_eval(out, data, cluster_label, config=None):
    """lib/eval_helper.py:11-114 on the forward's outputs and get_loss's cluster labels: language
    accuracy, per-scene referred box (arg-max of the summed scores), its IoU with the ground truth
    (utils/box_util.py:95-133: the same axis-aligned min/max IoU as the batch form), ref_acc and the
    IoU rates.  -> dict with the reference's keys."""
    config = config or SyntheticConfig()
    t = lambda k: np.asarray(data[k])
    lang_pred = out['lang_scores'].argmax(1)
    res = dict(lang_acc=(lang_pred == torch.as_tensor(t('object_cat'))).float().mean())
    gt = config.param2obb_batch(t('ref_center_label'), t('ref_heading_class_label'), t('ref_heading_residual_label'),
                                t('ref_size_class_label'), t('ref_size_residual_label'))
    ious, ref_acc, pred_boxes, gt_boxes = [], [], [], []
    start = 0
    for i, pred in enumerate(out['pred_obb_batch']):
        pred = np.asarray(pred).reshape(-1, 7)
        n = pred.shape[0]
        if n == 0:
            obb = np.zeros(7)
        elif n == 1:
            obb = pred[0]
        else:
            score = (out['attribute_scores'][start:start + n] + out['relation_scores'][start:start + n]
                     + out['scene_scores'][start:start + n])
            start += n
            cp = int(torch.argmax(score))
            ref_acc.append(1. if int(np.argmax(cluster_label[i])) == cp else 0.)
            obb = pred[cp]
        iou = float(iou_batch(obb[None], gt[i])[0])
        ious.append(iou)
        if n <= 1:
            ref_acc.append(1. if iou > 0.25 else 0.)
        corners = lambda o: box_min_max(np.concatenate([o[:6], [0.0]])[None])       # un-rotated (utils/util.py:21-32)
        pred_boxes.append(np.stack(corners(obb), 0)[:, 0])
        gt_boxes.append(np.stack(corners(gt[i]), 0)[:, 0])
    a = np.asarray(ious)
    res.update(ref_acc=ref_acc, ref_iou=ious, pred_box_min_max=pred_boxes, gt_box_min_max=gt_boxes)
    res['ref_iou_rate_0.25'] = float((a >= 0.25).sum()) / a.shape[0]
    res['ref_iou_rate_0.5'] = float((a >= 0.5).sum()) / a.shape[0]
    res['ref_others_mask'] = [1 if int(c) == 17 else 0 for c in t('object_cat')]
    return res
