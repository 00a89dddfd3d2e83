"""Optimiser side of the training step (scripts/train.py:93, lib/solver.py:200-205) for one process
per GPU: all parameters live in ONE flat fp32 buffer (each tensor a 256-byte-aligned view), so

  * ``zero_grad``      is one memset,
  * the data-parallel gradient reduction is ONE NCCL all-reduce over 8.02 M floats,
  * ``step``           is one fused Adam kernel (``ir_adam_step``) with the 1/world_size average
                       folded into its gradient read.

``torch.distributed`` is plumbing only (process group + all-reduce call)."""
import torch

from . import ops

ALIGN = 64          # floats (256 B): keeps every parameter 16-byte aligned for TMA bulk copies


class FlatAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (amsgrad off) on a flat buffer."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.params = [p for p in model.parameters() if p.requires_grad]
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
        self.group = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)

    def zero_grad(self):
        """.grad = None: backward then hands over each gradient tensor without an accumulation kernel."""
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """Pack the per-parameter gradients of this backward into the flat buffer (one multi-tensor
        copy; parameters that received no gradient contribute zeros) and point .grad at the views."""
        src, dst, zero = [], [], []
        for v, p in zip(self.grad_views, self.params):
            if p.grad is None:
                zero.append(v)
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
            p.grad = v
        if zero:
            torch._foreach_zero_(zero)
        if src:
            torch._foreach_copy_(dst, src)

    def allreduce(self):
        """Sum over ranks (the mean is folded into the Adam kernel)."""
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_grad, group=self.group)

    def step(self):
        self.gather_grads()
        self.allreduce()
        self.step_count += 1
        ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0],
                      self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0 / self.world)

    # --- checkpoint format of torch.optim.Adam (lib/solver.py:376-381 stores optimizer.state_dict()) ------
    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``: per-parameter step / exp_avg / exp_avg_sq (copies
        of the flat moments) and one param group, so ``checkpoint.tar`` is interchangeable with the reference's."""
        state = {}
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = dict(step=torch.tensor(float(self.step_count)),
                                exp_avg=self.exp_avg[o:o + n].view_as(p).clone(),
                                exp_avg_sq=self.exp_avg_sq[o:o + n].view_as(p).clone())
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     decoupled_weight_decay=False, params=list(range(len(self.params))))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        g = sd['param_groups'][0]
        self.lr, self.betas, self.eps, self.weight_decay = g['lr'], tuple(g['betas']), g['eps'], g['weight_decay']
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        for i, st in sd['state'].items():
            o, p = self.offsets[int(i)], self.params[int(i)]
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            self.step_count = max(self.step_count, int(float(st['step'])))

    @property
    def param_groups(self):
        """Read-only view for code that prints / schedules ``optimizer.param_groups[0]['lr']``."""
        return [dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, params=self.params)]
