#!/usr/bin/env python
"""bench.py — referrals/sec of the InstanceRefer hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one InstanceRefer.forward (eval) over one batch of the BASELINE.json `configs[1]` workload:
1 synthetic scene, 40k points (~28k voxels @5 cm), 32 candidate instances x 1024 points, 20-token
utterance.  One step = one referral per GPU; scenes shard across ranks with no data-path collective
(weak scaling).  Prints ONE JSON line on rank 0:
  value  referrals/s with inputs resident in HBM (CUDA events around each step, L2 flushed between
         steps outside the timed spans, max over ranks);
  e2e    the same forward through the public module API from HOST buffers: pinned H2D of the step's
         inputs + host packing of the instance lists + D2H of the scores, wall clock with syncs;
  roofline  dominant kernel (tcgen05 pair-GEMM) algorithmic bytes / event-timed duration vs measured
         HBM peak;  cpu_baseline  the oracle port (oracle/model_ref.py) timed on the host cores.
--impl reference: the reference's CPU path.  The reference is pure Python over un-vendored torchsparse /
PyG and cannot travel to the GPU box, so this arm times the oracle port of it (kind "port").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = dict(num_points=40000, n_inst=32, n_cand=32, n_tokens=20)
WORKLOAD_NAME = ('configs[1]: full InstanceRefer.forward, 1 synthetic scene (40k points -> ~28k voxels @5cm, '
                 '32 candidate instances x 1024 pts @2cm, 20-token utterance), eval mode, fp32')
METRIC = 'referrals/sec'


def forward_config():
    """`config` of the forward line — identical on both arms (--impl b200 / reference): it names the workload and the
    cache state of the timed region; how each arm executes it is in `impl_detail`."""
    return dict(workload=WORKLOAD_NAME, referrals_per_step_per_gpu=1,
                l2='cold between steps: a 256 MiB write flushes the 126 MB L2 outside the timed spans (GPU arm); the CPU arm '
                   'cycles the same 4 scenes')


def make_args():
    return types.SimpleNamespace(language_module='lang_module', attribute_module='attribute_module',
                                 relation_module='relation_module', scene_module='scene_module',
                                 num_classes=18, use_bidir=True, voxel_size_ap=0.02, voxel_size_glp=0.05,
                                 k=8, use_gt_lang=True)


def cpu_reference_leg(steps, warmup):
    """Oracle port of the reference forward on the host cores (bounded sample)."""
    import model_ref
    import weights
    from instancerefer_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict(123)
    args = make_args()
    data = [model_ref.data_from_batch(synthetic.make_batch(1000 + 7 * i, batch_size=1, **WORKLOAD)) for i in range(4)]
    for i in range(warmup):
        model_ref.forward(sd, data[i % 4], args)
    t0 = time.perf_counter()
    for i in range(steps):
        model_ref.forward(sd, data[i % 4], args)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=1.0 / dt, unit=METRIC, cores=torch.get_num_threads(), kind='port',
                sample=f'{steps} referrals of the bench workload after {warmup} warm-up; oracle/model_ref.py = the CPU port of the '
                       f'reference forward (torch CPU fp32, {torch.get_num_threads()} threads; GRU token loop and kNN query loop in '
                       f'Python — NOT the reference files verbatim over the shim, which cannot travel to the GPU box), {dt * 1e3:.0f} ms each'), dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '10'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == 'Active'})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def load_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


TRAIN_WORKLOAD_NAME = ('configs[2]: ScanRefer-train-shaped training step — 2 scenes per GPU (16 on 8 GPUs), 40k points, '
                       '32 instances x 1024 pts, 2..8 target-class candidates per scene, 20-token utterances; train-mode '
                       'forward + get_loss + backward + flat gradient all-reduce + Adam, fp32')


def train_batches(rank, n=4, per_gpu=2):
    from instancerefer_b200 import synthetic
    out = []
    for i in range(n):
        rng = np.random.default_rng(5000 + 131 * rank + i)
        out.append(synthetic.make_batch(3000 + 977 * rank + 13 * i, batch_size=per_gpu, num_points=40000, n_inst=32,
                                        n_cand=[int(c) for c in rng.integers(2, 9, per_gpu)], n_tokens=20))
    return out


def cpu_train_leg(steps, warmup, per_gpu=2):
    """Oracle port of one training iteration (forward + loss + backward, torch CPU autograd)."""
    import model_ref
    import train_ref
    import weights
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_state_dict(123)
    data = model_ref.data_from_batch(train_batches(0, 1, per_gpu)[0])
    for _ in range(warmup):
        train_ref.train_step(sd, data, make_args())
    t0 = time.perf_counter()
    for _ in range(steps):
        train_ref.train_step(sd, data, make_args())
    dt = (time.perf_counter() - t0) / steps
    return dict(value=per_gpu / dt, unit=METRIC, cores=torch.get_num_threads(), kind='port',
                sample=f'{steps} training iterations of {per_gpu} scenes after {warmup} warm-up, oracle/train_ref.py '
                       f'(torch CPU fp32 autograd, {torch.get_num_threads()} threads, no optimiser step), {dt * 1e3:.0f} ms each'), dt


def _pin_rank_to_cores(world, local):
    """N ranks on one box: give every rank its own slice of the host cores.  The host side of a step (class filter and
    packing of the instance lists in the forward's e2e path; ~45 library calls and ~20 autograd nodes in the training
    step) is a sizeable part of it, and N interpreters migrating onto each other's cores cost 6 % at N = 8."""
    ncpu = os.cpu_count() or 1
    if world > 1 and hasattr(os, 'sched_setaffinity') and os.environ.get('IR_AFFINITY', '1') == '1':
        per = max(1, ncpu // world)
        try:
            os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
        except OSError:
            pass
    return ncpu


def _init_dist(world, dev):
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=dev)


def train_measure(steps, warmup, rank, world, local, cpu_baseline=True):
    """BASELINE.json configs[2]: one step = train-mode forward + get_loss + backward + gradient all-reduce + fused Adam,
    inputs from HOST buffers every step (so `value` is already end to end); phases are CUDA-event timed.  -> dict."""
    import __graft_entry__ as g
    g.build()
    from instancerefer_b200 import SparseTensor, _lib, ops, synthetic
    from instancerefer_b200.instancerefer import InstanceRefer
    from instancerefer_b200.loss_helper import get_loss, stash_host_labels
    from instancerefer_b200.optim import FlatAdam
    per_gpu = 2
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ops.check_device(local)
    ncpu = _pin_rank_to_cores(world, local)
    _init_dist(world, dev)
    if world > 1:
        import torch.distributed as dist
    lib = _lib.load()
    model = InstanceRefer(7, make_args())
    model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)
    model = model.to(dev).train()
    opt = FlatAdam(model, lr=1e-3, weight_decay=1e-5)          # config/InstanceRefer.yaml:48,53
    if os.environ.get('IR_OVERLAP') == '0':
        opt._sync = False                                      # reduce everything at step() (no bucket launches from the hooks)
    cfg = synthetic.SyntheticConfig()                          # dataset-config stand-in (labels -> GT box)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hosts = []
    for b in train_batches(int(os.environ.get('IR_BATCH_RANK', rank))):     # IR_BATCH_RANK: another rank's batches on one GPU
        h = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        hosts.append(h)
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].values() if torch.is_tensor(v))
    h2d += sum(p.nbytes for sc in hosts[0]['instance_points'] for p in sc)
    ph = {k: 0.0 for k in ('forward', 'loss', 'backward', 'allreduce+adam')}

    def step(i, timed):
        h = hosts[i % len(hosts)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        d = stash_host_labels(dict(h))                          # like the solver: host label copies survive the move
        d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
        d['lidar'] = SparseTensor(d.pop('lidar_feats'), d.pop('lidar_coords'))
        opt.zero_grad()
        d = model(d)
        ev[1].record()
        d = get_loss(d, cfg)
        ev[2].record()
        d['loss'].backward()                                    # bucket all-reduces start from the parameter hooks
        ev[3].record()
        opt.step()
        ev[4].record()
        loss = float(d['loss'].detach())                       # D2H of the step's result
        if timed:
            for j, k in enumerate(ph):
                ph[k] += ev[j].elapsed_time(ev[j + 1])
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # --- eager leg (short): the four phases of the step by CUDA events, and the eager step time for the record
    for i in range(max(warmup, 2 * len(hosts) + 2)):        # every batch shape seen twice: lazy state and the encoder graphs exist
        step(i, False)
    barrier()
    n_eager = min(steps, 8)
    te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = lib.ir_launch_count()
    te0.record()
    losses = [step(i, True) for i in range(n_eager)]
    te1.record()
    barrier()
    eager_ms = te0.elapsed_time(te1) / n_eager
    launches_eager = (lib.ir_launch_count() - c0) // n_eager
    phases = {k: v / n_eager for k, v in ph.items()}
    # --- the measured step: the whole iteration replayed from one CUDA graph per batch signature (train_graph.py);
    #     IR_TRAIN_STEP=eager times the eager iteration instead
    mode = os.environ.get('IR_TRAIN_STEP', 'graph')
    stepper = None
    if mode == 'graph':
        from instancerefer_b200.train_graph import GraphedTrainStep
        stepper = GraphedTrainStep(model, opt, cfg)
        for i in range(3 * len(hosts) + warmup):             # per signature: eager, two slots captured; then replays
            stepper(hosts[i % len(hosts)])['result'].loss()
        stepper.profile = []
        assert stepper.replays >= warmup and len(stepper.cache) >= 1

        pending = []

        def run(i):
            # every step's scalars come back through their own async D2H copy (train_graph.StepResult); the host reads
            # step i-1 after it has queued step i, like a training loop that logs one iteration behind
            pending.append(stepper(hosts[i % len(hosts)])['result'])
            return pending.pop(0).loss() if len(pending) > 1 else None
    else:
        def run(i):
            return step(i, False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    c0 = lib.ir_launch_count()
    r0 = stepper.launches_replayed if stepper else 0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = [run(i) for i in range(steps)]
    if stepper:
        losses = [v for v in losses if v is not None] + [r.loss() for r in pending]      # the last step's read-back
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0]), float(t[1])
    launches = lib.ir_launch_count() - c0 + ((stepper.launches_replayed - r0) if stepper else 0)
    replay = None
    if stepper and stepper.profile:
        pr = stepper.profile[-steps:]
        replay = dict(graph_ms=sum(a.elapsed_time(b) for _, a, b in pr) / len(pr), host_stage_ms=sum(t for t, _, _ in pr) / len(pr) * 1e3,
                      note='graph_ms: CUDA events around the replay of the captured iteration; host_stage_ms: class filter + '
                           'packing + queueing the H2D copies, on the host before each replay')
    clocks = sampler.stop() if sampler else None
    # data-parallel sanity: every rank applied the same averaged gradients, so the parameter buffers must be
    # bitwise identical across ranks (BatchNorm running statistics are per rank and live outside the flat buffer)
    in_sync = True
    if world > 1:
        chk = opt.flat.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs().item() == 0.0)
    cb = None
    if rank == 0 and world == 1 and cpu_baseline:
        cb, _ = cpu_train_leg(2, 1, per_gpu)
    value = world * per_gpu * steps / (dev_ms * 1e-3)
    return dict(
        metric=METRIC, value=value, unit=METRIC, n_gpus=world, steps=steps, warmup=warmup,
        ms_per_step=dev_ms / steps, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f32 (rule GEMM, dgrad and wgrad on tcgen05 as split-fp16 hi/lo with fp32 accumulation)', data='synthetic',
        config=dict(workload=TRAIN_WORKLOAD_NAME, scenes_per_gpu=per_gpu, global_batch=world * per_gpu,
                    l2='inputs (2.6 MB) and activations change every step; working set > L2 over a step',
                    parallelism=f'dp{world}: scenes sharded, gradient all-reduce in {opt.n_buckets} availability buckets launched on a side '
                                f'stream as each branch\'s backward finishes, Adam replicated'),
        e2e=dict(value=world * per_gpu * steps / wall, unit=METRIC, ms_per_step=wall / steps * 1e3,
                 h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=20 if stepper else 4),
        step_mode=('one CUDA graph per batch signature: forward + get_loss + backward + bucketed all-reduce + Adam '
                   f'({len(stepper.cache)} signatures x {stepper.depth} slots, {stepper.replays} replays; per-step scalars read '
                   'back one step behind)') if stepper else 'eager',
        eager=dict(ms_per_step=eager_ms, phases_ms=phases, allreduce_exposed_ms=phases['allreduce+adam'],
                   launches_per_step=int(launches_eager), steps=n_eager),
        phases_ms=phases, allreduce_exposed_ms=phases['allreduce+adam'], replay=replay,
        gpu_launches=int(launches), clocks=clocks, host_cores=ncpu,
        final_loss=losses[-1], first_loss=losses[0], ranks_in_sync=in_sync, cpu_baseline=cb)


def main_train(a):
    """--workload train: BASELINE.json configs[2] as its own line."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(a.warmup, 3)
    per_gpu = 2
    if a.impl == 'reference':
        if rank != 0:
            return
        cb, dt = cpu_train_leg(a.steps, warmup, per_gpu)
        print(json.dumps(dict(metric=METRIC, value=cb['value'], unit=METRIC, impl='reference', n_gpus=a.gpus, steps=a.steps,
                              warmup=warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                              dtype='f32', data='synthetic', config=dict(workload=TRAIN_WORKLOAD_NAME), cpu_baseline=cb,
                              e2e=dict(value=cb['value'], unit=METRIC, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    line = train_measure(a.steps, warmup, rank, world, local, not a.no_cpu_baseline)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def _forward_setup(local):
    import __graft_entry__ as g
    g.build()
    from instancerefer_b200 import ops, synthetic
    from instancerefer_b200.instancerefer import InstanceRefer
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ops.check_device(local)
    model = InstanceRefer(7, make_args())
    model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)
    return model.to(dev).eval(), dev


def _resident_dict(b, dev):
    """Inputs of one batch staged in HBM (loader tensors + the packed instance buffer)."""
    from instancerefer_b200 import SparseTensor, synthetic
    from instancerefer_b200.candidates import KEY, CandidatePack
    d = synthetic.to_data_dict(b, SparseTensor, dev)
    pack = CandidatePack(d, d['object_cat'], dev)
    pack.resident = True
    d[KEY] = pack
    d['_ir_lang_len_max'] = int(np.max(b['lang_len']))
    return d


def _time_graph(model, d, steps, warmup, flush):
    """CUDA-graph replay of model(d): per-step CUDA events, L2 flushed between steps -> ms/step."""
    for _ in range(2):
        model(dict(d))
    torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        model(dict(d))
    for _ in range(warmup):
        g_.replay()
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g_.replay()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / steps


def main_sweep(a):
    """--workload sweep: BASELINE.json configs[4] — forward throughput over 10k-200k points x 8-128 instances
    (n = c), one scene per step per GPU, inputs resident, CUDA-graph replay; --workload relation:
    configs[3] — relation module alone (kNN + EdgeConv + match) on 32 scenes x 64 instances."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    from instancerefer_b200 import synthetic
    model, dev = _forward_setup(local)
    if os.environ.get('IR_PAIR') == '1':
        model.pair_encoders = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps, warmup = max(a.steps, 5), max(a.warmup, 3)
    rows = []
    if a.workload == 'relation':
        b = synthetic.make_batch(7000 + rank, batch_size=32, num_points=40000, n_inst=64, n_cand=64, n_tokens=20)
        d = _resident_dict(b, dev)
        d['lang_rel_feats'] = torch.randn(32, 256, device=dev)
        rel = model.relation
        ms = _time_graph(lambda x: rel(x), d, steps, warmup, flush)
        rows.append(dict(scenes=32, instances_per_scene=64, edges=32 * 64 * 8, ms_per_step=ms, referrals_per_sec=world * 32 / (ms * 1e-3)))
        name = 'configs[3]: relation_module kNN + EdgeConv + cosine match, 32 scenes x 64 instances (64 candidates each), k=8, eval'
    else:
        pts = [10000, 20000, 40000, 80000, 120000, 200000] if not a.quick else [10000, 200000]
        inst = [8, 16, 32, 64, 128] if not a.quick else [8, 128]
        for npts in pts:
            for n in inst:
                room = tuple(np.array([8.0, 6.0, 3.0]) * np.array([(npts / 40000) ** 0.5, (npts / 40000) ** 0.5, 1.0]))
                b = synthetic.make_batch(9000 + rank, batch_size=1, num_points=npts, n_inst=n, n_cand=n, n_tokens=20,
                                         room=tuple(min(v, m) for v, m in zip(room, (11.9, 19.9, 3.9))))
                d = _resident_dict(b, dev)
                ms = _time_graph(model, d, steps, warmup, flush)
                rows.append(dict(points=npts, scene_voxels=int(b['lidar_coords'].shape[0]), instances=n, ms_per_step=ms,
                                 referrals_per_sec=world * 1.0 / (ms * 1e-3)))
                del d
                torch.cuda.empty_cache()
        name = 'configs[4]: forward sweep, points x instances (n = c), 1 scene per step per GPU, eval, CUDA-graph replay'
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        t = torch.tensor([r['ms_per_step'] for r in rows], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for r, v in zip(rows, t.tolist()):
            r['ms_per_step'] = v
            r['referrals_per_sec'] = world * (32 if a.workload == 'relation' else 1) / (v * 1e-3)
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(dict(metric=METRIC, unit=METRIC, n_gpus=world, steps=steps, warmup=warmup, higher_is_better=True,
                              scaling='weak', dtype='f32', data='synthetic', config=dict(workload=name), table=rows,
                              value=rows[0]['referrals_per_sec'], ms_per_step=rows[0]['ms_per_step'])))


def main():
    if '--workload' in sys.argv and sys.argv[sys.argv.index('--workload') + 1] in ('sweep', 'relation'):
        ap = argparse.ArgumentParser()
        ap.add_argument('--workload')
        ap.add_argument('--gpus', type=int, default=1)
        ap.add_argument('--steps', type=int, default=10)
        ap.add_argument('--warmup', type=int, default=3)
        ap.add_argument('--quick', action='store_true')
        return main_sweep(ap.parse_args())
    if '--workload' in sys.argv and sys.argv[sys.argv.index('--workload') + 1] == 'train':
        ap = argparse.ArgumentParser()
        ap.add_argument('--workload')
        ap.add_argument('--gpus', type=int, default=1)
        ap.add_argument('--steps', type=int, default=20)
        ap.add_argument('--warmup', type=int, default=10)      # each of the 4 cycled batches once eagerly, once capturing its CUDA graphs
        ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
        ap.add_argument('--no-cpu-baseline', action='store_true')
        return main_train(ap.parse_args())
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='forward', choices=['forward', 'train', 'sweep', 'relation'])
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager', action='store_true', help='no CUDA-graph replay (launch every kernel from Python)')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step block of the line')
    ap.add_argument('--train-steps', type=int, default=20)
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(a.warmup, 3)

    if a.impl == 'reference':
        if rank != 0:
            return
        # same --steps / --warmup as the GPU arm; one step = one referral of the same workload on the host cores
        # (~0.4 s each, so the default 50 + 5 steps end within half a minute)
        cb, dt = cpu_reference_leg(a.steps, warmup)
        print(json.dumps(dict(metric=METRIC, value=cb['value'], unit=METRIC, impl='reference', n_gpus=a.gpus,
                              steps=a.steps, warmup=warmup, ms_per_step=dt * 1e3, higher_is_better=True,
                              scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                              config=forward_config(), impl_detail=cb['sample'], cpu_baseline=cb,
                              e2e=dict(value=cb['value'], unit=METRIC, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    import __graft_entry__ as g
    g.build()
    from instancerefer_b200 import SparseTensor, _lib, ops, synthetic
    from instancerefer_b200.candidates import KEY, CandidatePack
    from instancerefer_b200.instancerefer import InstanceRefer

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    ops.check_device(local)
    _pin_rank_to_cores(world, local)
    _init_dist(world, dev)
    if world > 1:
        import torch.distributed as dist
    lib = _lib.load()
    args = make_args()
    model = InstanceRefer(7, args)
    model.load_state_dict(synthetic.make_state_dict(123, model=model), strict=True)      # random-init weights (no oracle import on this arm)
    model = model.to(dev).eval()
    if os.environ.get('IR_PAIR') == '1':
        model.pair_encoders = True

    # ---- this rank's stream of scenes (a few distinct scenes, cycled)
    n_scenes = 4
    batches = [synthetic.make_batch(1000 + 97 * rank + 7 * i, batch_size=1, **WORKLOAD) for i in range(n_scenes)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def host_dict(b):
        """step inputs in pinned host memory, in the layout the reference's collate_fn produces"""
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        return dict(lidar_F=pin(b['lidar_feats']), lidar_C=pin(b['lidar_coords']), lang_feat=pin(b['lang_feat']),
                    lang_len=pin(b['lang_len']), object_cat=pin(b['object_cat']), point_min=pin(b['point_min']),
                    instance_points=b['instance_points'], instance_obbs=b['instance_obbs'],
                    instance_class=b['instance_class'])

    def to_device(h):
        d = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in h.items()}
        d['lidar'] = SparseTensor(d.pop('lidar_F'), d.pop('lidar_C'))
        return d

    hosts = [host_dict(b) for b in batches]
    h2d_tensor_bytes = sum(v.numel() * v.element_size() for v in hosts[0].values() if torch.is_tensor(v))

    # ---- resident inputs for `value`
    resident = []
    for h in hosts:
        d = to_device(h)
        pack = CandidatePack(d, d['object_cat'], dev)
        pack.resident = True
        d[KEY] = pack
        d['_ir_lang_len_max'] = int(h['lang_len'].max())
        resident.append(d)
    torch.cuda.synchronize()

    out_keys = ('attribute_scores', 'relation_scores', 'scene_scores', 'lang_scores', 'seg_scores', 'ref_pred')
    from instancerefer_b200.graphed import GraphedInstanceRefer
    runner = GraphedInstanceRefer(model)
    use_graph = not a.eager

    # `value`: inputs resident in HBM -> one captured graph per resident scene, replay only
    res_graphs = []
    stamps = torch.zeros(256, 4, dtype=torch.int64, device=dev)    # live per-launch spans of the conv kernels (scene 0's graph)
    if use_graph:
        for i, d in enumerate(resident):
            for _ in range(2):
                model(dict(d))
            torch.cuda.synchronize()
            if i == 0 and rank == 0:
                lib.ir_conv_stamps_set(ctypes.c_void_p(stamps.data_ptr()))
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                o_ = model(dict(d))
            lib.ir_conv_stamps_set(None)
            res_graphs.append((g_, o_))

    def step_resident(i):
        if use_graph:
            g_, o_ = res_graphs[i % n_scenes]
            g_.replay()
            return o_
        d = dict(resident[i % n_scenes])
        return model(d)

    d2h_bytes = [0]
    pinned_out = {}

    def host_step_dict(i):
        h = hosts[i % n_scenes]
        d = dict(h)
        d['lidar'] = SparseTensor(d.pop('lidar_F'), d.pop('lidar_C'))
        return d

    def run_e2e(n):
        """n steps from HOST buffers to HOST scores.  Graph mode: double-buffered submit()/result() — step
        i+1 is filtered, packed and uploaded while the GPU replays step i; every step still uploads its own
        inputs and reads its own scores back."""
        if use_graph:
            prev = None
            for i in range(n):
                h = runner.submit(host_step_dict(i))
                if prev is not None:
                    prev.result()
                prev = h
            prev.result()
            d2h_bytes[0] = runner.d2h_bytes
            return
        for i in range(n):
            d = to_device(hosts[i % n_scenes])
            out = model(d)
            nb = 0
            for k in out_keys:
                t = out[k]
                if k not in pinned_out or pinned_out[k].shape != t.shape:
                    pinned_out[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                pinned_out[k].copy_(t, non_blocking=True)
                nb += t.numel() * t.element_size()
            torch.cuda.synchronize()
            d2h_bytes[0] = nb

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: resident inputs, per-step CUDA events, L2 flush between steps
    sampler = ClockSampler(local) if rank == 0 else None        # 10 ms samples from before the warm-up to the end of e2e
    for i in range(warmup):
        step_resident(i)
    barrier()
    launches0 = lib.ir_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()
        evs[i][0].record()
        step_resident(i)
        evs[i][1].record()
    barrier()
    launches = lib.ir_launch_count() - launches0
    if use_graph:                           # replayed kernels are not re-counted by the library: count one step
        c0 = lib.ir_launch_count()
        model(dict(resident[0]))
        launches = (lib.ir_launch_count() - c0) * a.steps
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    dev_ms = max_over_ranks(dev_ms)
    value = world * a.steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers in, scores out
    run_e2e(warmup)
    barrier()
    t0 = time.perf_counter()
    run_e2e(a.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = dict(value=world * a.steps / e2e_s, unit=METRIC, ms_per_step=e2e_s / a.steps * 1e3,
               h2d_bytes_per_step=int(runner.h2d_bytes if use_graph else h2d_tensor_bytes + resident[0][KEY].h2d_bytes),
               d2h_bytes_per_step=int(d2h_bytes[0]))
    clocks = sampler.stop() if sampler else None

    # ---- roofline pass: per-launch spans of the conv kernels measured INSIDE the replayed 4-stream graph (in-kernel GPU
    # timer: begin = first CTA past its dependency wait, end = last CTA done), L2 flushed before every replay
    roofline, detail, layers = None, None, None
    if rank == 0:
        nprof = min(a.steps, 20)
        I64MAX = (1 << 63) - 1
        cap = 256
        meta, nout = (ctypes.c_int32 * (4 * cap))(), ctypes.c_int32(0)
        maps = [0, 5, 1, 1, 6, 2, 2, 7, 3, 3, 8, 4, 4]                  # layer -> kernel map id
        acc = None
        if use_graph:
            _lib.call('ir_conv_stamps_meta', meta, cap, ctypes.byref(nout))
            n_l = nout.value
            acc = np.zeros((n_l, 2))
            for _ in range(nprof):
                stamps[:, 0] = I64MAX; stamps[:, 2] = I64MAX; stamps[:, 1] = 0; stamps[:, 3] = 0
                flush.zero_()
                res_graphs[0][0].replay()
                torch.cuda.synchronize()
                st = stamps[:n_l].cpu().numpy()
                acc[:, 0] += np.where(st[:, 1] > 0, st[:, 1] - st[:, 0], 0) / 1e3        # pair-GEMM us
                acc[:, 1] += np.where(st[:, 3] > 0, st[:, 3] - st[:, 2], 0) / 1e3        # reduce / stem us
            acc /= nprof
            kc = [model.attribute.net._last_ws.kcount().cpu().numpy(), model.scene.net._last_ws.kcount().cpu().numpy()]
            nl = [model.attribute.net._last_ws.nlvl().cpu().numpy(), model.scene.net._last_ws.nlvl().cpu().numpy()]
            lvl_of = [0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4]
            layers = []
            for j in range(n_l):
                enc, layer = divmod(j, 13)
                cin, cout, K, tc = (meta[4 * j + q] for q in range(4))
                P = int(kc[enc][maps[layer]][:K].sum())
                layers.append(dict(encoder=('instance', 'scene')[enc], layer=layer, cin=cin, cout=cout, K=K, tcgen05=bool(tc),
                                   rows_out=int(nl[enc][lvl_of[layer]]), pairs=P, gemm_us=round(float(acc[j, 0]), 2),
                                   reduce_us=round(float(acc[j, 1]), 2),
                                   bytes_gemm=P * (4 * cin + 4) + P * 4 * cout + 4 * K * cin * cout,          # gather + T write + W
                                   bytes_conv=P * (4 * cin + 4) + P * (4 * cout + 4) + 4 * K * cin * cout))   # G + S + W
        peak, peak_src = load_peak()
        traffic = ncu = None
        tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.isfile(tp):
            tj = json.load(open(tp))
            traffic, ncu = tj.get('dram_bytes_per_launch'), tj.get('ncu')
        if layers:
            tcl = [l for l in layers if l['tcgen05']]
            g_us, g_by = sum(l['gemm_us'] for l in tcl), sum(l['bytes_gemm'] for l in tcl)
            c_us = sum(l['gemm_us'] + l['reduce_us'] for l in layers)
            c_by = sum(l['bytes_conv'] for l in layers)
            ach, cach = g_by / (g_us * 1e-6) / 1e9, c_by / (c_us * 1e-6) / 1e9
            roofline = dict(bound='hbm', kernel=f'k_pairgemm_tc (tcgen05 split-fp16 pair-GEMM, {len(tcl)} launches/step)',
                            achieved=ach, peak=peak, unit='GB/s', frac=ach / peak, traffic=traffic, peak_source=peak_src,
                            timing='in-kernel GPU-timer span of every launch (first CTA past its dependency wait -> last CTA '
                                   'done) inside the replayed 4-stream CUDA graph, L2 flushed before each replay, mean of '
                                   f'{nprof} replays; both encoders run concurrently, so a span includes the time its CTAs '
                                   'waited for SMs held by the other encoder',
                            bytes_per_launch=g_by / len(tcl), us_per_launch=g_us / len(tcl),
                            conv=dict(what='whole sparse conv = pair-GEMM + reduce/epilogue (+ the two fused stems), SURVEY 8(d) '
                                           'bytes G+S+W over the summed spans of both kernels', achieved=cach, frac=cach / peak,
                                      us_per_step=c_us, bytes_per_step=c_by),
                            ncu=ncu)
            groups = {}
            for l in layers:
                e = groups.setdefault(f"{'tcgen05' if l['tcgen05'] else 'direct'}_{l['cin']}x{l['cout']}_k{l['K']}", [0, 0.0, 0.0, 0, 0])
                e[0] += 1; e[1] += l['gemm_us']; e[2] += l['reduce_us']; e[3] += l['pairs']; e[4] += l['bytes_conv']
            detail = {k: dict(launches_per_step=v[0], gemm_us=v[1] / v[0], reduce_us=v[2] / v[0], pairs=v[3] / v[0],
                              conv_GBs=v[4] / ((v[1] + v[2]) * 1e-6) / 1e9) for k, v in groups.items()}

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb, _ = cpu_reference_leg(5, 1)

    # ---- the training step (BASELINE configs[2]) at the same N: the path's one collective, under the driver's eyes
    train = None
    if not a.no_train:
        del res_graphs, runner, resident
        torch.cuda.empty_cache()
        t = train_measure(a.train_steps, 10, rank, world, local, cpu_baseline=False)
        train = {k: t[k] for k in ('value', 'unit', 'ms_per_step', 'steps', 'warmup', 'phases_ms', 'allreduce_exposed_ms',
                                   'ranks_in_sync', 'e2e', 'gpu_launches', 'host_cores', 'config', 'first_loss', 'final_loss', 'step_mode',
                                   'eager', 'replay')}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=METRIC, n_gpus=world, steps=a.steps, warmup=warmup,
                    ms_per_step=dev_ms / a.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype='f32 (rule GEMM on tcgen05 as split-fp16 hi/lo, fp32 accumulate)', data='synthetic',
                    config=forward_config(),
                    impl_detail=dict(parallelism=f'scenes sharded over {world} GPU(s), no data-path collective',
                                     launch='eager' if not use_graph else 'CUDA-graph replay, 4 streams; e2e double-buffered (host of step i+1 overlaps GPU of step i)'),
                    e2e=e2e, gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                    cpu_baseline=cb, spconv_detail=detail, conv_layers=layers, train=train)
        print(json.dumps(line))


if __name__ == '__main__':
    main()
