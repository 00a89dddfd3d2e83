"""Timing of the scan preprocessing on one ScanNet-sized synthetic scan: the device array work
alone (CUDA events around the seven library calls, inputs resident), the whole `export_arrays` + `filter_scan` call from
host arrays to host arrays, and the numpy oracle on the same scan.  Prints one JSON line.
    python tools/prepare_bench.py [--verts 150000 --faces 300000 --objects 60 --props 80 --reps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instancerefer_b200 import _lib, prepare_data as P          # noqa: E402
from oracle import prepare_ref as PR                            # noqa: E402  (CPU baseline leg only)

NYU = {'chair': 5, 'table': 7, 'wall': 1, 'floor': 2, 'cabinet': 3, 'sofa': 6, 'door': 8, 'window': 9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--verts', type=int, default=150000)
    ap.add_argument('--faces', type=int, default=300000)
    ap.add_argument('--objects', type=int, default=60)
    ap.add_argument('--props', type=int, default=80)
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    from instancerefer_b200 import synthetic
    s = synthetic.synth_scan(seed=41, n_verts=a.verts, n_faces=a.faces, n_objects=a.objects, n_props=a.props)
    vertex = np.concatenate([s['xyz'], s['rgb'].astype(np.float32)], 1)
    n, nf, no, ni = a.verts, a.faces, a.objects, a.props
    ch = np.random.RandomState(0).choice(n, min(n, P.MAX_NUM_POINT), replace=False)

    def whole():
        A = P.export_arrays(vertex, s['faces'], s['matrix'], s['seg_indices'], s['seg_groups'], NYU, s['masks'], s['cls'])
        return P.filter_scan(A, choices=ch)
    whole()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        whole()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3

    # device work alone: the same seven calls on resident inputs
    dev = torch.device('cuda')
    n_tab = int(s['seg_indices'].max()) + 1
    seg_label, seg_object, obj_label = P.segment_tables(s['seg_groups'], NYU, n_tab)
    v9 = torch.zeros(n, 9, device=dev)
    v9[:, :6] = torch.from_numpy(vertex).to(dev)
    faces = torch.from_numpy(s['faces']).to(dev)
    seg, sl, so, ol = (torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev) for x in (s['seg_indices'], seg_label, seg_object, obj_label))
    masks, cls = torch.from_numpy(s['masks']).to(dev), torch.from_numpy(s['cls'].astype(np.int32)).to(dev)
    chd = torch.from_numpy(ch.astype(np.int64)).to(dev)
    scratch = torch.empty(_lib.load().ir_prepare_scratch_bytes(n, nf, no), dtype=torch.uint8, device=dev)
    al, sem, ins, spg, ipg = torch.empty_like(v9), *(torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4))
    bb, bba = (torch.empty(no, 8, dtype=torch.float64, device=dev) for _ in range(2))
    m = np.ascontiguousarray(s['matrix'].reshape(16))
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()

    def device():
        _lib.call('ir_mesh_normals', p(v9), n, p(faces), nf, p(scratch), st)
        _lib.call('ir_align_vertices', p(v9), n, m.ctypes.data, p(al), st)
        _lib.call('ir_vertex_labels', p(seg), n, p(sl), p(so), n_tab, p(sem), p(ins), st)
        _lib.call('ir_instance_boxes', p(v9), p(ins), n, no, p(ol), p(scratch), p(bb), st)
        _lib.call('ir_instance_boxes', p(al), p(ins), n, no, p(ol), p(scratch), p(bba), st)
        _lib.call('ir_pointgroup_labels', p(masks), p(cls), ni, n, p(spg), p(ipg), st)
        for t in (v9, al, sem, ins, spg, ipg):
            P.select_rows(t, chd)
    for _ in range(3):
        device()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        device()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / a.reps
    # algorithmic bytes: vertex rows read+written (normals, align, 2 boxes, gathers), faces, face normals, masks, labels
    m_sel = len(ch)
    bytes_ = (n * 36 * 2 + nf * 12 + nf * 12 * 2 + n * 12 * 2        # normals: rows r/w, faces, face normals w+r, last table
              + n * 36 * 2 + n * 4 * 3 + 2 * n * 16 + ni * n + n * 8  # align, labels, boxes (xyz+id), masks, pg labels
              + m_sel * (36 * 2 + 4 * 4) * 2)                         # gathers
    t0 = time.perf_counter()
    PR.export_one_scan(PR.export(vertex, s['faces'], s['matrix'], s['seg_indices'], s['seg_groups'], NYU, s['masks'], s['cls']), choices=ch)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({'metric': 'scan_preprocess', 'config': {'verts': n, 'faces': nf, 'objects': no, 'proposals': ni},
                      'device_ms': dev_ms, 'device_gbps_algorithmic': bytes_ / dev_ms / 1e6, 'e2e_ms_host_to_host': e2e_ms,
                      'cpu_oracle_ms': cpu_ms, 'cpu_kind': 'port (numpy, 1 process)', 'launches_per_scan': 17}))


if __name__ == '__main__':
    main()
