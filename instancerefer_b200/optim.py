"""Optimiser side of the training step (scripts/train.py:93, lib/solver.py:200-205) for one process
per GPU: all parameters live in ONE flat fp32 buffer (each tensor a 256-byte-aligned view), so

  * ``zero_grad``      drops the per-parameter gradients (the next backward hands new ones over without an add),
  * the data-parallel gradient reduction is a handful of NCCL all-reduces over contiguous slices of ONE buffer of
                       8.02 M floats,
  * ``step``           is one fused Adam kernel (``ir_adam_step`` / ``ir_adam_step_dev``) with the 1/world_size
                       average folded into its gradient read.

Data parallel (world > 1): the flat gradient buffer is cut into contiguous buckets along gradient availability (per
sub-module; each sparse encoder's deep and shallow stages apart); a post-accumulate hook on every parameter counts the
gradients of its bucket and, when the bucket is complete, packs it and launches its all-reduce asynchronously (NCCL's own
stream) while the rest of the backward is still running — only the bucket that finishes last is exposed (SURVEY §5 /
§8(e); lib/solver.py:200-205 is the step this serves).  ``early_grads`` lets a backward node hand over gradients that
are final before it returns; ``stage_step`` / ``after_replay`` are the host halves of a step replayed from a CUDA graph
(train_graph.GraphedTrainStep).

``torch.distributed`` is plumbing only (process group + all-reduce call)."""
import contextlib

import torch

from . import ops

ALIGN = 64          # floats (256 B): keeps every parameter 16-byte aligned for TMA bulk copies
BUCKET_FLOATS = 1 << 21   # >= 8 MB of gradients per all-reduce bucket when only sizes are known (a plain parameter list)


def _availability_key(name):
    """Parameters whose gradients become final together during the backward share a key: a sub-module of the model, with
    each sparse encoder (``<module>.net``) cut into its deep stages (stage3/4: 2/3 of its parameters, finished first)
    and its shallow ones.  Buckets cut along these keys are complete — and their all-reduce in flight — as early as the
    backward allows, instead of waiting for the last gradient of a size-based range that straddles two branches."""
    parts = name.split('.')
    if len(parts) > 2 and parts[1] == 'net':
        return parts[0], 'net.hi' if parts[2] in ('stage3', 'stage4') else 'net.lo'
    return parts[0], ''


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (amsgrad off) on a flat buffer.

    A real ``torch.optim.Optimizer``: ONE mutable param group whose ``lr`` / ``betas`` / ``eps`` /
    ``weight_decay`` the Adam kernel reads at every step, so the reference's own solver works unchanged on it
    (lib/solver.py:119-125 builds ``MultiStepLR(optimizer, ...)``, scripts/train.py:112-119 loads its
    ``state_dict``).  ``FlatAdam(model, ...)`` and ``FlatAdam(model.parameters(), ...)`` are both accepted.
    Like torch's Adam, parameters that received no gradient in a step are left untouched (no weight decay, no
    moment decay) and own no optimizer state."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        names = None
        if isinstance(params, torch.nn.Module):
            names = [n for n, p in params.named_parameters() if p.requires_grad]
            params = params.parameters()
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("FlatAdam keeps one flat buffer: pass a module or a flat iterable of parameters "
                             "(one param group, as scripts/train.py:93 does)")
        params = [p for p in params if p.requires_grad]
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.params = self.param_groups[0]['params']
        assert all(p.dtype == torch.float32 for p in self.params)
        dev = self.params[0].device
        ofs, total = [], 0
        for p in self.params:
            ofs.append(total)
            total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_views = []
        for p, o in zip(self.params, ofs):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view_as(p)              # parameters become views of the flat buffer
            self.grad_views.append(self.flat_grad[o:o + n].view_as(p))
            p.grad = None
        self.offsets, self.numel = ofs, total
        self.step_count = 0
        self.has_state = [False] * len(self.params)             # parameter saw a gradient at least once
        self._skip_key, self._skip_mask = None, None            # per-64-float block: 1 = no gradient this step
        self.group = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        # --- gradient buckets (contiguous parameter ranges of >= BUCKET_FLOATS) for the overlapped all-reduce
        self.buckets, lo = [], 0
        end_of = lambda i: ofs[i + 1] if i + 1 < len(ofs) else total
        if names is not None and len(names) == len(self.params):
            # cut where the availability key changes; runs of small groups (< 1/8 bucket) are merged
            small = BUCKET_FLOATS // 8
            for i in range(len(self.params)):
                if i + 1 == len(self.params) or _availability_key(names[i + 1]) != _availability_key(names[i]):
                    new = (lo, i + 1, ofs[lo], end_of(i))
                    if self.buckets and new[3] - new[2] < small and self.buckets[-1][3] - self.buckets[-1][2] < small:
                        new = (self.buckets[-1][0], i + 1, self.buckets[-1][2], end_of(i))
                        self.buckets.pop()
                    self.buckets.append(new)                                # params [lo, i+1), floats [ofs[lo], end)
                    lo = i + 1
        else:
            for i in range(len(self.params)):
                if end_of(i) - ofs[lo] >= BUCKET_FLOATS or i + 1 == len(self.params):
                    self.buckets.append((lo, i + 1, ofs[lo], end_of(i)))
                    lo = i + 1
        self.n_buckets = len(self.buckets)
        self._bucket_of = [b for b, (l, h, _, _) in enumerate(self.buckets) for _ in range(h - l)]
        self._fired = [0] * self.n_buckets          # gradients seen in this backward, per bucket
        self._expect = [0] * self.n_buckets         # ... in the previous step (0 = unknown: no early launch)
        self._work = [None] * self.n_buckets        # in-flight all-reduce of the bucket (launched from the hook)
        self._packed = [False] * self.n_buckets
        self._early = set()                         # parameters whose gradient was handed over by early_grads()
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self.overlap = self.world > 1               # launch bucket all-reduces from the backward hooks
        self._sync = True
        self._hyper_dev = self._hyper_host = None   # per-step scalars on the device (graph-replayed steps)
        self._staged = False
        if self.overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))
            from . import training
            training.EARLY_GRAD_HOOK = self.early_grads

    # hyper-parameters live in the param group (what a scheduler mutates); attribute access for convenience
    lr = property(lambda self: self.param_groups[0]['lr'], lambda self, v: self.param_groups[0].__setitem__('lr', v))
    betas = property(lambda self: self.param_groups[0]['betas'], lambda self, v: self.param_groups[0].__setitem__('betas', tuple(v)))
    eps = property(lambda self: self.param_groups[0]['eps'], lambda self, v: self.param_groups[0].__setitem__('eps', v))
    weight_decay = property(lambda self: self.param_groups[0]['weight_decay'],
                            lambda self, v: self.param_groups[0].__setitem__('weight_decay', v))

    def zero_grad(self, set_to_none=True):
        """.grad = None: backward then hands over each gradient tensor without an accumulation kernel."""
        for p in self.params:
            p.grad = None
        self._fired = [0] * self.n_buckets
        self._packed = [False] * self.n_buckets
        self._early = set()

    # --- overlapped bucket all-reduce -------------------------------------------------------------------
    def _make_hook(self, i):
        b = self._bucket_of[i]

        def hook(_p):
            if i in self._early:                      # already counted (and copied) by early_grads
                return
            self._fired[b] += 1
            if self._sync and self._fired[b] == self._expect[b] and self._work[b] is None and not self._packed[b]:
                self._launch(b)
        return hook

    def early_grads(self, params, grads):
        """Gradients that are final before the autograd node computing them returns (the deep stages of an encoder,
        training.EncoderTrainGraphed): copied into the flat buffer now, counted towards their buckets, and a bucket
        they complete starts its all-reduce at once.  The node still returns the same gradients to autograd later;
        the hook and the packing then skip these parameters.  -> False when nothing is reduced early (no_sync,
        single process): the caller has nothing else to do either way."""
        if not (self.overlap and self._sync):
            return False
        idx = [self._index[id(p)] for p in params]
        torch._foreach_copy_([self.grad_views[i] for i in idx], list(grads))
        touched = []
        for i in idx:
            self._early.add(i)
            b = self._bucket_of[i]
            self._fired[b] += 1
            if b not in touched:
                touched.append(b)
        for b in touched:
            if self._fired[b] == self._expect[b] and self._work[b] is None and not self._packed[b]:
                self._launch(b)
        return True

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside this context only accumulate (gradient accumulation over several backwards):
        no bucket is reduced before ``step``."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def _collect(self, b, src, dst, zero):
        """Bucket b's share of a pack: appends (gradient -> flat view) copy pairs and the views of parameters without a
        gradient (packed as zeros), points .grad at the views.  -> indices of its parameters without a gradient."""
        lo, hi, _, _ = self.buckets[b]
        missing = []
        for i in range(lo, hi):
            v, p = self.grad_views[i], self.params[i]
            if i in self._early:                      # already in the flat buffer; autograd still owes p.grad (the same
                continue                              # values): pointing it at the view now would make that an in-place add
            if p.grad is None:
                zero.append(v)
                missing.append(i)
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
            p.grad = v
        self._packed[b] = True
        return missing

    @staticmethod
    def _flush(src, dst, zero):
        if zero:
            torch._foreach_zero_(zero)
        if src:
            torch._foreach_copy_(dst, src)

    def _pack(self, b):
        """Copy bucket b's gradients of this backward into the flat buffer (missing ones as zeros), point .grad at
        the views.  -> indices of its parameters without a gradient."""
        src, dst, zero = [], [], []
        missing = self._collect(b, src, dst, zero)
        self._flush(src, dst, zero)
        return missing

    def _launch(self, b):
        """Bucket complete (called from the last hook of the bucket, inside backward): pack it and start its
        all-reduce; NCCL orders it after the copies on the current stream and runs it beside the backward."""
        self._pack(b)
        _, _, a, e = self.buckets[b]
        self._work[b] = torch.distributed.all_reduce(self.flat_grad[a:e], group=self.group, async_op=True)

    def gather_grads(self):
        """Pack the per-parameter gradients of this backward into the flat buffer (one multi-tensor
        copy; parameters that received no gradient contribute zeros and are masked out of the update)
        and point .grad at the views.  -> indices of the parameters without a gradient."""
        missing, src, dst, zero = [], [], [], []
        for b in range(self.n_buckets):
            if self._work[b] is None:                 # buckets already in flight were packed by their hook
                missing += self._collect(b, src, dst, zero)
            else:
                lo, hi, _, _ = self.buckets[b]
                missing += [i for i in range(lo, hi) if self.params[i].grad is None and i not in self._early]
        self._flush(src, dst, zero)                   # one multi-tensor copy for everything still unpacked
        for i in self._early:
            self.params[i].grad = self.grad_views[i]
        return missing

    def allreduce(self):
        """Sum over ranks (the mean is folded into the Adam kernel): buckets whose all-reduce was launched from the
        backward hooks are only waited for (stream-side), the others are reduced here."""
        if self.world > 1:
            late = [b for b in range(self.n_buckets) if self._work[b] is None]
            if len(late) == self.n_buckets:
                torch.distributed.all_reduce(self.flat_grad, group=self.group)
            else:
                for b in late:
                    _, _, a, e = self.buckets[b]
                    self._work[b] = torch.distributed.all_reduce(self.flat_grad[a:e], group=self.group, async_op=True)
            for b in range(self.n_buckets):
                if self._work[b] is not None:
                    self._work[b].wait()              # the current stream waits; the host does not
                    self._work[b] = None
        self._expect = list(self._fired)              # what a complete bucket looks like, for the next backward
        self._fired = [0] * self.n_buckets

    def _block_skip(self, missing):
        """uint8 per 64-float block (every parameter is block-aligned): 1 where the update is skipped."""
        if not missing:
            return None
        key = tuple(missing)
        if key != self._skip_key:
            m = torch.zeros(self.numel // ALIGN, dtype=torch.uint8)
            for i in missing:
                o, n = self.offsets[i], self.params[i].numel()
                m[o // ALIGN:(o + n + ALIGN - 1) // ALIGN] = 1
            self._skip_key, self._skip_mask = key, m.to(self.flat.device)
        return self._skip_mask

    # --- steps replayed from a CUDA graph (train_graph.GraphedTrainStep) ----------------------------------------------
    def stage_step(self):
        """Host half of a step whose device half is (or is being) captured: advance the step count and put the scalars
        that change per step — lr as the scheduler left it, the two bias corrections — on the device (async, current
        stream).  The next ``step()`` call, or the replay of a captured one, reads them there."""
        if self._hyper_dev is None:
            self._hyper_dev = torch.zeros(4, dtype=torch.float32, device=self.flat.device)
            self._hyper_host = torch.zeros(16, 4, dtype=torch.float32).pin_memory()     # ring: copies in flight keep theirs
        self.step_count += 1
        g = self.param_groups[0]
        h = self._hyper_host[self.step_count % 16]
        ops.adam_hyper(g['lr'], g['betas'][0], g['betas'][1], self.step_count, h)
        self._hyper_dev.copy_(h, non_blocking=True)
        self._staged = True

    def after_replay(self):
        """Host bookkeeping of a replayed step (what ``step()`` does beside launching kernels)."""
        self._staged = False
        from .basic_blocks import bump_weights_epoch
        bump_weights_epoch()

    def graph_key(self):
        """Everything a captured step bakes in as launch arguments."""
        g = self.param_groups[0]
        return (tuple(g['betas']), g['eps'], g['weight_decay'], self.world, self._sync)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        missing = self.gather_grads()
        self.allreduce()
        if not all(self.has_state):
            gone = set(missing)
            self.has_state = [h or (i not in gone) for i, h in enumerate(self.has_state)]
        g = self.param_groups[0]
        if self._staged:                               # scalars already on the device (stage_step)
            self._staged = False
            ops.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._hyper_dev, g['betas'][0],
                              g['betas'][1], g['eps'], g['weight_decay'], 1.0 / self.world, self._block_skip(missing))
        else:
            self.step_count += 1
            ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, g['lr'], g['betas'][0],
                          g['betas'][1], g['eps'], g['weight_decay'], self.step_count, 1.0 / self.world,
                          self._block_skip(missing))
        from .basic_blocks import bump_weights_epoch
        bump_weights_epoch()           # the kernel wrote the parameters through raw pointers: tensor versions did not move
        return loss

    # --- checkpoint format of torch.optim.Adam (lib/solver.py:376-381 stores optimizer.state_dict()) ------
    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``: per-parameter step / exp_avg / exp_avg_sq (copies
        of the flat moments) and one param group, so ``checkpoint.tar`` is interchangeable with the reference's."""
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            if not self.has_state[i]:
                continue
            n = p.numel()
            state[i] = dict(step=torch.tensor(float(self.step_count)),
                            exp_avg=self.exp_avg[o:o + n].view_as(p).clone(),
                            exp_avg_sq=self.exp_avg_sq[o:o + n].view_as(p).clone())
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(self.params)))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        g = sd['param_groups'][0]
        for k in ('lr', 'eps', 'weight_decay', 'initial_lr'):
            if k in g:
                self.param_groups[0][k] = g[k]
        self.param_groups[0]['betas'] = tuple(g['betas'])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        self.has_state = [False] * len(self.params)
        for i, st in sd['state'].items():
            o, p = self.offsets[int(i)], self.params[int(i)]
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            self.step_count = max(self.step_count, int(float(st['step'])))
            self.has_state[int(i)] = True
