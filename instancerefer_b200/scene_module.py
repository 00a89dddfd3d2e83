"""Scene (GLP) module, drop-in for the reference's ``models/scene_module.py``: whole-scene sparse
encoder at 5 cm -> crop -> dense BEV 15x25 (z-slice matvec + scatter-sum, BN, ReLU) -> 2x Conv2d
3x3 -> language-guided attention over 11x21 cells -> 9-way region classifier and cosine(obj, scene).
Reference lines: models/scene_module.py:10-58 (ctor), :60-108 (forward)."""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .basic_blocks import (BEVEncoder, PrepCache, ReLU, SparseCrop, ToDenseBEVConvolution, fold_bn,
                           require_eval)
from .candidates import get_pack


class GlobalMaxPooling(nn.Module):
    pass


class SceneModule(nn.Module, PrepCache):
    def __init__(self, input_feature_dim, args, v_dim=128, h_dim=128, l_dim=256, dropout_rate=0.15):
        super().__init__()
        self.args = args
        self.input_feature_dim = input_feature_dim
        self.voxel_size = np.array([args.voxel_size_glp] * 3)
        self.net = BEVEncoder(self.input_feature_dim)
        self.pooling = GlobalMaxPooling()
        self.to_bev = nn.Sequential(SparseCrop((0, 0, 0), (240, 400, 80)),
                                    ToDenseBEVConvolution(128, 128, shape=(15, 25, 5), z_dim=2),
                                    nn.BatchNorm2d(128), ReLU(True))
        self.h_dim = h_dim
        self.vis_emb_fc = nn.Sequential(nn.Conv2d(v_dim, h_dim, 3), nn.BatchNorm2d(h_dim), nn.ReLU(),
                                        nn.Dropout(dropout_rate), nn.Conv2d(h_dim, h_dim, 3))
        self.vis_emb_fc1 = nn.Sequential(nn.Linear(128, h_dim), nn.LayerNorm(h_dim), nn.ReLU(),
                                         nn.Dropout(dropout_rate), nn.Linear(h_dim, h_dim))
        self.lang_emb_fc = nn.Sequential(nn.Linear(l_dim, h_dim), nn.LayerNorm(h_dim), nn.ReLU(),
                                         nn.Dropout(dropout_rate), nn.Linear(h_dim, h_dim))
        self.cls = nn.Sequential(nn.Linear(h_dim, h_dim), nn.BatchNorm1d(h_dim), nn.ReLU(),
                                 nn.Linear(h_dim, 9))

    def _prep_tensors(self):
        mods = (self.to_bev, self.vis_emb_fc, self.vis_emb_fc1, self.lang_emb_fc, self.cls)
        return [t for m in mods for t in list(m.parameters()) + list(m.buffers())]

    def _prepare(self):
        f = lambda t: t.detach().float().contiguous()
        ft = lambda t: t.detach().float().t().contiguous()         # (in,out) layout for ir_mlp_head
        pk = lambda conv: conv.weight.detach().float().permute(2, 3, 1, 0).contiguous()   # [ky][kx][Cin][Cout]
        bs, bb = fold_bn(self.to_bev[2])
        c1s, c1b = fold_bn(self.vis_emb_fc[1])
        cs, cb = fold_bn(self.cls[1])
        v1, l = self.vis_emb_fc1, self.lang_emb_fc
        c1bias, c2bias = f(self.vis_emb_fc[0].bias), f(self.vis_emb_fc[4].bias)
        return dict(bev_kernel=f(self.to_bev[1].kernel), bev_s=bs, bev_b=bb,
                    c1w=pk(self.vis_emb_fc[0]), c1bias=c1bias, c1s=c1s, c1b=c1b,
                    c2w=pk(self.vis_emb_fc[4]), c2bias=c2bias,
                    # tcgen05 form: weights as (9, Cin, Cout) rule-GEMM operands, conv bias folded into the shift
                    c1w9=pk(self.vis_emb_fc[0]).view(9, 128, 128), c1shift=(c1s * c1bias + c1b).contiguous(),
                    c2w9=pk(self.vis_emb_fc[4]).view(9, 128, 128), c2shift=c2bias,
                    ow1=ft(v1[0].weight), ob1=f(v1[0].bias), og=f(v1[1].weight), obeta=f(v1[1].bias),
                    ow2=ft(v1[4].weight), ob2=f(v1[4].bias),
                    lw1=ft(l[0].weight), lb1=f(l[0].bias), lg=f(l[1].weight), lbeta=f(l[1].bias),
                    lw2=ft(l[4].weight), lb2=f(l[4].bias),
                    kw1=ft(self.cls[0].weight), kb1=f(self.cls[0].bias), kg=cs, kbeta=cb,
                    kw2=ft(self.cls[3].weight), kb2=f(self.cls[3].bias))

    conv2d_tc = os.environ.get('IR_CONV2D', 'tc') != 'simt'      # BEV Conv2d on the tcgen05 rule GEMM (default) or the SIMT kernel

    def _bev_tail(self, p, f4, c4, n4, n_max, B):
        """crop + dense BEV + BN + ReLU (:70), Conv2d-BN-ReLU-Conv2d (:71) -> (B,11,21,128) NHWC."""
        if not self.conv2d_tc:
            bev = ops.bev(f4, c4, n4, n_max, 16, p['bev_kernel'], p['bev_s'], p['bev_b'], B)
            ops.stamp('scene:bev')
            x = ops.conv2d_3x3(bev, p['c1w'], p['c1bias'], p['c1s'], p['c1b'], True)
            return ops.conv2d_3x3(x, p['c2w'], p['c2bias'], None, None, False)
        amax = torch.zeros(2, dtype=torch.float32, device=f4.device)       # max|BEV map|, max|conv1 out|: range scales
        bev = ops.bev(f4, c4, n4, n_max, 16, p['bev_kernel'], p['bev_s'], p['bev_b'], B, absmax=amax[0:1])
        ops.stamp('scene:bev')
        x = ops.conv2d_3x3_tc(bev, p['c1w9'], p['c1s'], p['c1shift'], True, in_absmax=amax[0:1], out_absmax=amax[1:2])
        return ops.conv2d_3x3_tc(x, p['c2w9'], None, p['c2shift'], False, in_absmax=amax[1:2])

    def encode_scene(self, data_dict, device):
        """Phase A (no language dependency): whole-scene sparse encoder (:69), crop + dense BEV + BN +
        ReLU (:70), Conv2d-BN-ReLU-Conv2d (:71) -> (B,11,21,128) NHWC."""
        require_eval(self)
        p = self.prepared()
        lidar = data_dict['lidar']
        B = data_dict['point_min'].shape[0]                                     # (:62-63)
        F0 = lidar.F.to(device, torch.float32).contiguous()
        C0 = lidar.C.to(device, torch.int32).contiguous()
        ws = self.net.workspace(F0.shape[0], device)
        ops.stamp('scene:start')
        f4, c4, n4 = self.net.encode(ws, F0, C0, data_dict.get('_ir_lidar_rows'))
        ops.stamp('scene:features')
        data_dict['_ir_bev_feats'] = self._bev_tail(p, f4, c4, n4, ws.n_max, B)
        ops.stamp('scene:conv2d')
        return data_dict

    def prepare_maps(self, data_dict, device):
        """Phase A1: hash the loader's 5 cm voxels, levels + kernel maps."""
        require_eval(self)
        lidar = data_dict['lidar']
        F0 = lidar.F.to(device, torch.float32).contiguous()
        C0 = lidar.C.to(device, torch.int32).contiguous()
        ws = self.net.workspace(F0.shape[0], device)
        self.net.build_maps(ws, C0, data_dict.get('_ir_lidar_rows'))
        return ws, F0

    def bev_convs(self, data_dict, ws, f4):
        """Phase A3: crop + dense BEV + BN + ReLU (:70), Conv2d-BN-ReLU-Conv2d (:71)."""
        p = self.prepared()
        B = data_dict['point_min'].shape[0]
        data_dict['_ir_bev_feats'] = self._bev_tail(p, f4, ws.coords(4), ws.nlvl()[4:5], ws.n_max, B)
        return data_dict

    def embed_language(self, data_dict):
        p = self.prepared()
        q, _ = ops.mlp_head(data_dict['lang_scene_feats'].float().contiguous(), p['lw1'], p['lb1'], ops.NORM_LAYER,
                            p['lg'], p['lbeta'], p['lw2'], p['lb2'], ops.MODE_RAW)
        return q

    def match(self, data_dict):
        """Phase B: language-guided attention over the 11x21 cells (:73-83), region classifier (:84),
        cosine(vis_emb_fc1(obj_feats), scene_feat) (:89-104)."""
        p = self.prepared()
        x = data_dict['_ir_bev_feats']
        B, h, w = x.shape[0], x.shape[1], x.shape[2]
        q = data_dict.pop('_ir_scene_lang', None)
        if q is None:
            q = self.embed_language(data_dict)
        atten, scene_feats = ops.scene_attention(x.view(B, h * w, -1), q)
        data_dict['vis_atten'] = atten.view(B, h, w)
        # the region classifier and the candidate head both hang off scene_feats and are independent of each other:
        # the classifier runs on a side stream (fork / join by events, capturable), the candidate head — the one the
        # step's final softmax waits for — stays on this stream
        cur = torch.cuda.current_stream(x.device)
        side = self.__dict__.get('_seg_stream')
        if side is None or side.device != x.device:
            side = self.__dict__['_seg_stream'] = torch.cuda.Stream(device=x.device)
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            seg, _ = ops.mlp_head(scene_feats, p['kw1'], p['kb1'], ops.NORM_AFFINE, p['kg'], p['kbeta'],
                                  p['kw2'], p['kb2'], ops.MODE_RAW)
            join.record(side)
        data_dict['seg_scores'] = seg
        pack = get_pack(data_dict, self.args, x.device)
        _, scores = ops.mlp_head(data_dict['obj_feats'], p['ow1'], p['ob1'], ops.NORM_LAYER, p['og'],
                                 p['obeta'], p['ow2'], p['ob2'], ops.MODE_COS, partner=scene_feats,
                                 seg=pack.cand_scene)
        cur.wait_event(join)
        data_dict['scene_scores'] = scores
        ops.stamp('scene:matched')
        return data_dict

    def forward(self, data_dict):
        ops.check_device()
        if self.training:
            from . import training
            pack = get_pack(data_dict, self.args, data_dict['lang_scene_feats'].device)
            return training.scene_forward_train(self, data_dict, pack)
        data_dict = self.encode_scene(data_dict, data_dict['lang_scene_feats'].device)
        return self.match(data_dict)
