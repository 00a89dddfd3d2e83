"""Training / validation loop around the B200 step, drop-in for the reference's ``lib/solver.Solver``
(lib/solver.py:63-342,369-391): same constructor arguments and ``solver(epoch, verbose)`` call, the same
per-iteration order (forward -> get_loss -> backward + step -> get_eval), the same "best model" criterion
(validation ``iou_rate_0.25``) and the same files under ``<output_root>/<stamp>/``:

    model.pth          best validation model          (lib/solver.py:342)
    model_last.pth     after every epoch / at the end (lib/solver.py:156,386)
    checkpoint.tar     {epoch, model_state_dict, optimizer_state_dict}   (lib/solver.py:376-381)
    log.txt

One process per GPU: under ``torchrun`` every rank runs the loop on its own shard of the loader (give the
loader a ``DistributedSampler``), ``FlatAdam.step`` does the one gradient all-reduce, validation metrics are
combined over ranks with one small all-reduce of their sums and counts, and only rank 0 writes files.  TensorBoard export and the
report templates of the reference are host-side cosmetics and are not reproduced."""
import os
import time

import numpy as np
import torch

from .eval_helper import get_eval
from .loss_helper import get_loss, stash_host_labels

DEVICE_KEYS = ('lang_feat', 'lang_len', 'object_cat', 'lidar', 'point_min', 'point_max', 'ref_center_label',
               'ref_size_residual_label')                     # lib/solver.py:242-245
BN_MOMENTUM_INIT, BN_MOMENTUM_MAX = 0.5, 0.001                 # lib/solver.py:131-132
METRICS = ('loss', 'ref_loss', 'lang_loss', 'seg_loss', 'lang_acc', 'ref_acc', 'seg_acc')


class Solver:
    def __init__(self, model, config, dataloader, optimizer, stamp, val_step=10, lr_decay_step=None,
                 lr_decay_rate=None, bn_decay_step=None, bn_decay_rate=None, output_root='outputs', graph_train=False):
        self.model, self.config, self.dataloader, self.optimizer = model, config, dataloader, optimizer
        self.stamp, self.val_step = stamp, val_step
        self.lr_decay_step, self.lr_decay_rate = lr_decay_step, lr_decay_rate
        self.bn_decay_step, self.bn_decay_rate = bn_decay_step, bn_decay_rate
        self.base_lr = optimizer.param_groups[0]['lr']
        self.best = {'epoch': 0, 'iou_rate_0.25': -float('inf')}
        dist = torch.distributed
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.root = os.path.join(output_root, stamp)
        if self.rank == 0:
            os.makedirs(self.root, exist_ok=True)
        self.history = {'train': [], 'val': []}
        # graph_train: training iterations whose batch shape repeats are replayed from one CUDA graph
        # (train_graph.GraphedTrainStep; FlatAdam only); others run the ordinary eager iteration below
        self._graph_step = None
        if graph_train:
            from .train_graph import GraphedTrainStep
            self._graph_step = GraphedTrainStep(model, optimizer, config)
        self._global_iter_id = 0
        self.start_epoch = 0

    # ------------------------------------------------------------------ schedules (lib/solver.py:119-137)
    def _lr_at(self, epoch_id):
        if not (self.lr_decay_step and self.lr_decay_rate):
            return self.base_lr
        if isinstance(self.lr_decay_step, (list, tuple)):                       # MultiStepLR
            k = sum(1 for m in self.lr_decay_step if epoch_id >= m)
        else:                                                                   # StepLR
            k = epoch_id // self.lr_decay_step
        return self.base_lr * self.lr_decay_rate ** k

    def _bn_momentum_at(self, epoch_id):
        return max(BN_MOMENTUM_INIT * self.bn_decay_rate ** int(epoch_id / self.bn_decay_step), BN_MOMENTUM_MAX)

    def _apply_schedules(self, epoch_id):
        lr = self._lr_at(epoch_id)
        if hasattr(self.optimizer, 'lr'):
            self.optimizer.lr = lr
        else:
            for g in self.optimizer.param_groups:
                g['lr'] = lr
        if self.bn_decay_step and self.bn_decay_rate:
            mom = self._bn_momentum_at(epoch_id)
            for m in self.model.modules():
                if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                    m.momentum = mom

    # ------------------------------------------------------------------ one iteration
    def _log(self, msg):
        if self.rank == 0:
            with open(os.path.join(self.root, 'log.txt'), 'a') as f:
                f.write(msg + '\n')
            print(msg, flush=True)

    def _to_device(self, data_dict):
        stash_host_labels(data_dict)
        for k in DEVICE_KEYS:
            if k in data_dict and hasattr(data_dict[k], 'cuda'):
                data_dict[k] = data_dict[k].cuda()
        return data_dict

    def _step(self, data_dict, phase):
        train = phase == 'train'
        if train and self._graph_step is not None:
            # _forward, _compute_loss, _backward (lib/solver.py:195-205) as one replayed graph over the HOST batch
            data_dict = self._graph_step(data_dict)
            with torch.no_grad():
                data_dict = get_eval(data_dict, self.config)
            rec = {k: float(data_dict[k].detach()) for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss', 'lang_acc', 'seg_acc')}
            rec['ref_acc'] = float(np.mean(data_dict['ref_acc']))
            rec['ref_iou'] = list(data_dict['ref_iou'])
            return rec
        data_dict = self._to_device(data_dict)
        with torch.set_grad_enabled(train):
            if train:
                self.optimizer.zero_grad()
            data_dict = self.model(data_dict)                                    # _forward   (lib/solver.py:195)
            data_dict = get_loss(data_dict, self.config)                         # _compute_loss (:207)
            if train:
                data_dict['loss'].backward()                                     # _backward  (:200-205)
                self.optimizer.step()
            data_dict = get_eval(data_dict, self.config)                         # _eval      (:219)
        rec = {k: float(data_dict[k].detach()) for k in ('loss', 'ref_loss', 'lang_loss', 'seg_loss', 'lang_acc', 'seg_acc')}
        rec['ref_acc'] = float(np.mean(data_dict['ref_acc']))
        rec['ref_iou'] = list(data_dict['ref_iou'])
        return rec

    def _feed(self, loader, phase, epoch_id, verbose):
        self.model.train(phase == 'train')
        recs, t0 = [], time.time()
        for data_dict in loader:
            recs.append(self._step(data_dict, phase))
            if phase == 'train':
                self._global_iter_id += 1
                if verbose and self._global_iter_id % verbose == 0:
                    last = recs[-verbose:]
                    self._log('[train] epoch {} iter {}: '.format(epoch_id + 1, self._global_iter_id) +
                              ', '.join('{} {:.4f}'.format(k, np.mean([r[k] for r in last])) for k in METRICS) +
                              ', {:.1f} ms/iter'.format((time.time() - t0) / len(recs) * 1e3))
        ious = np.asarray([v for r in recs for v in r['ref_iou']])
        # sums and counts first, ratios last: under torchrun the ranks may hold different numbers of iterations / samples
        # (last partial batch), so numerators and denominators are all-reduced and divided once
        sums = [float(np.sum([r[k] for r in recs])) for k in METRICS] + \
               [float((ious >= 0.25).sum()), float((ious >= 0.5).sum()), float(len(recs)), float(ious.size)]
        if self.world > 1:
            t = torch.tensor(sums, dtype=torch.float64, device='cuda')
            torch.distributed.all_reduce(t)
            sums = t.tolist()
        n_it, n_iou = max(sums[-2], 1.0), max(sums[-1], 1.0)
        out = {k: sums[i] / n_it for i, k in enumerate(METRICS)}
        out['iou_rate_0.25'] = sums[len(METRICS)] / n_iou
        out['iou_rate_0.5'] = sums[len(METRICS) + 1] / n_iou
        self.history[phase].append(out)
        return out

    # ------------------------------------------------------------------ checkpoints
    def _save_model(self, name):
        if self.rank == 0:
            torch.save(self.model.state_dict(), os.path.join(self.root, name))

    def _finish(self, epoch_id):
        if self.rank == 0:
            torch.save({'epoch': epoch_id, 'model_state_dict': self.model.state_dict(),
                        'optimizer_state_dict': self.optimizer.state_dict()}, os.path.join(self.root, 'checkpoint.tar'))
        self._save_model('model_last.pth')

    def resume(self, path=None):
        """Continue from ``checkpoint.tar`` (what scripts/train.py:110-118 does with --use_checkpoint)."""
        ck = torch.load(path or os.path.join(self.root, 'checkpoint.tar'), map_location='cpu', weights_only=False)
        self.model.load_state_dict(ck['model_state_dict'])
        self.optimizer.load_state_dict(ck['optimizer_state_dict'])
        self.start_epoch = int(ck['epoch']) + 1
        return self.start_epoch

    def __call__(self, epoch, verbose=10):
        epoch_id = self.start_epoch - 1
        for epoch_id in range(self.start_epoch, epoch):
            self._apply_schedules(epoch_id)
            self._log('epoch {} starting...'.format(epoch_id + 1))
            tr = self._feed(self.dataloader['train'], 'train', epoch_id, verbose)
            self._save_model('model_last.pth')
            va = self._feed(self.dataloader['val'], 'val', epoch_id, 0)
            self._log('[epoch {}] train loss {:.4f} | val loss {:.4f} ref_acc {:.4f} iou@0.25 {:.4f} iou@0.5 {:.4f}'.format(
                epoch_id + 1, tr['loss'], va['loss'], va['ref_acc'], va['iou_rate_0.25'], va['iou_rate_0.5']))
            if va['iou_rate_0.25'] > self.best['iou_rate_0.25']:
                self.best = dict(va, epoch=epoch_id + 1)
                self._log('best iou_rate_0.25 achieved: {}'.format(va['iou_rate_0.25']))
                self._save_model('model.pth')
        self._finish(epoch_id)
        return self.best
