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
