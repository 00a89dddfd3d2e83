"""world_size-2 gloo test (CPU) of the scene-sharding host logic: each rank runs the forward on its
block of scenes (the CPU oracle stands in for the CUDA kernels here) and the gathered score vectors must
equal the single-process forward of the whole batch."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ('lang_scores', 'seg_scores', 'attribute_scores', 'relation_scores', 'scene_scores', 'obj_feats')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    import model_ref
    import weights
    from conftest import make_args
    from instancerefer_b200 import SparseTensor, parallel, synthetic
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    args = make_args()
    sd = weights.make_state_dict(123)
    b = synthetic.make_batch(55, batch_size=3, num_points=3000, n_inst=6, n_cand=[3, 2, 4], n_tokens=[4, 9, 6])
    full = synthetic.to_data_dict(b, SparseTensor, 'cpu')
    mine = parallel.shard_data_dict(full, rank, world)

    def run(d):
        data = dict(d)
        data['lidar_F'], data['lidar_C'] = d['lidar'].F, d['lidar'].C
        return model_ref.forward(sd, data, args)

    got = parallel.gather_outputs(run(mine))
    if rank == 0:
        want = run(full)
        q.put({k: float((got[k] - want[k]).abs().max()) for k in KEYS} | {'n': int(got['attribute_scores'].shape[0])})
    dist.barrier()
    dist.destroy_process_group()


def test_scene_range_partition():
    sys.path.insert(0, ROOT)
    from instancerefer_b200.parallel import scene_range
    for n in (1, 2, 7, 16):
        for w in (1, 2, 3, 8):
            blocks = [scene_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1


def test_sharded_forward_matches_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res.pop('n') == 9
    assert max(res.values()) < 1e-6, res


# ----------------------------------------------------------------------------- training step, data parallel

def _grad_worker(rank, world, port, q):
    """Flat-buffer optimiser plumbing on CPU: parameters become views of one buffer, the gradients of a
    backward are packed with one multi-tensor copy, ONE all-reduce sums them over ranks;
    the Adam kernel itself is CUDA-only and must refuse to run here."""
    sys.path.insert(0, ROOT)
    from instancerefer_b200 import _lib
    from instancerefer_b200 import optim
    from instancerefer_b200.optim import ALIGN, FlatAdam
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    optim.BUCKET_FLOATS = 64                                        # several buckets even for this toy model
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    ref = [p.detach().clone() for p in model.parameters()]
    opt = FlatAdam(model, lr=1e-3)
    ok = all(torch.equal(p.detach(), r) for p, r in zip(model.parameters(), ref))
    ok &= all(p.data_ptr() == opt.flat.data_ptr() + 4 * o and o % ALIGN == 0 for p, o in zip(opt.params, opt.offsets))
    x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 + rank))
    opt.zero_grad()
    model(x).square().sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    model[2].bias.grad = None                                      # a parameter without gradient packs as zeros
    opt.gather_grads()
    opt.allreduce()
    summed = []
    for i, g in enumerate(local):
        t = g.clone() if i != len(local) - 1 else torch.zeros_like(g)
        dist.all_reduce(t)
        summed.append(t)
    ok &= all(torch.allclose(p.grad, s) for p, s in zip(model.parameters(), summed))
    ok &= opt.world == world and opt.n_buckets >= 2 and all(w is None for w in opt._work)
    # second backward: every bucket now knows how many gradients complete it, so its all-reduce is launched from the
    # parameter hooks DURING backward; step-time packing / reduction only waits for them
    x2 = torch.randn(4, 7, generator=torch.Generator().manual_seed(20 + rank))
    opt.zero_grad()
    model(x2).square().sum().backward()
    early = sum(w is not None for w in opt._work)
    local2 = []
    for r in range(world):                                         # every rank's local gradients, recomputed
        m2 = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
        m2.load_state_dict(model.state_dict())
        m2(torch.randn(4, 7, generator=torch.Generator().manual_seed(20 + r))).square().sum().backward()
        local2.append([p.grad for p in m2.parameters()])
    miss = opt.gather_grads()
    opt.allreduce()
    want = [sum(local2[r][i] for r in range(world)) for i in range(len(local2[0]))]
    ok &= early == opt.n_buckets and miss == [] and all(w is None for w in opt._work)
    ok &= all(torch.allclose(p.grad, w, rtol=1e-5, atol=1e-6) for p, w in zip(model.parameters(), want))
    # third backward: the last layer's gradients are handed over EARLY (as the encoder's backward node does for its deep
    # stages): their bucket is in flight before backward() is even called, the later hooks skip them, the sums agree
    opt.zero_grad()
    out = model(x2).square().sum()
    last = [model[2].weight, model[2].bias]
    ok &= opt.early_grads(last, list(torch.autograd.grad(out, last, retain_graph=True)))
    ok &= opt._work[opt._bucket_of[opt._index[id(last[0])]]] is not None and sum(w is not None for w in opt._work) == 1
    out.backward()
    ok &= sum(w is not None for w in opt._work) == opt.n_buckets
    miss = opt.gather_grads()
    opt.allreduce()
    ok &= miss == [] and all(torch.allclose(p.grad, w, rtol=1e-5, atol=1e-6) for p, w in zip(model.parameters(), want))
    ok &= all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(opt.params, opt.grad_views))
    with opt.no_sync():                                            # accumulation: nothing is reduced before step
        opt.zero_grad()
        model(x2).square().sum().backward()
        ok &= all(w is None for w in opt._work)
        ok &= opt.early_grads(last, [torch.zeros_like(t) for t in last]) is False
    try:
        opt.step()
        refused = False
    except _lib.IrError:
        refused = True                                             # no CPU fallback for the Adam kernel
    if rank == 0:
        q.put(dict(ok=bool(ok), refused=refused))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == dict(ok=True, refused=True), res
