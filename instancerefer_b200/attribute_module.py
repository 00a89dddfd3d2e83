"""Attribute (AP) module, drop-in for the reference's ``models/attribute_module.py``: class-filtered
candidates are voxelised at 2 cm ON THE GPU (first-point-wins), encoded by the sparse backbone,
max-pooled per candidate and matched against the language feature.
Reference lines: models/attribute_module.py:12-40 (ctor), :42-81 (filter_candidates), :83-131."""
import numpy as np
import torch.nn as nn

from . import ops
from .basic_blocks import PrepCache, SparseConvEncoder, fold_bn, require_eval
from .candidates import KEY as _KEY
from .candidates import get_pack


class GlobalMaxPooling(nn.Module):
    """Parameter-free holder (spnn.GlobalMaxPooling); the pooling runs in ops.segmax."""


class AttributeModule(nn.Module, PrepCache):
    def __init__(self, input_feature_dim, args, v_dim=128, h_dim=256, l_dim=256):
        super().__init__()
        self.args = args
        self.input_feature_dim = input_feature_dim
        self.voxel_size = np.array([args.voxel_size_ap] * 3)
        self.net = SparseConvEncoder(self.input_feature_dim)
        self.pooling = GlobalMaxPooling()
        self.vis_emb_fc = nn.Sequential(nn.Linear(v_dim, h_dim), nn.LayerNorm(h_dim), nn.ReLU(),
                                        nn.Linear(h_dim, h_dim))
        self.lang_emb_fc = nn.Sequential(nn.Linear(l_dim, h_dim), nn.BatchNorm1d(h_dim), nn.ReLU(),
                                         nn.Linear(h_dim, h_dim))
        self.weight_initialization()

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _prepare(self):
        f = lambda t: t.detach().float().contiguous()
        ft = lambda t: t.detach().float().t().contiguous()         # (in,out) layout for ir_mlp_head
        v, l = self.vis_emb_fc, self.lang_emb_fc
        s, b = fold_bn(l[1])
        return dict(vw1=ft(v[0].weight), vb1=f(v[0].bias), vg=f(v[1].weight), vbeta=f(v[1].bias),
                    vw2=ft(v[3].weight), vb2=f(v[3].bias),
                    lw1=ft(l[0].weight), lb1=f(l[0].bias), lg=s, lbeta=b, lw2=ft(l[3].weight), lb2=f(l[3].bias))

    def _prep_tensors(self):        # heads only; the encoder caches its own copies
        return list(self.vis_emb_fc.parameters()) + list(self.lang_emb_fc.parameters()) + list(self.lang_emb_fc.buffers())

    def encode_candidates(self, data_dict, device, pack=None):
        """Phase A (no language dependency): host class filter + one packed H2D (:42-81,101), GPU
        voxelisation @2 cm, sparse encoder (:104), per-candidate max-pool (:105)."""
        require_eval(self)
        pack = pack or get_pack(data_dict, self.args, device, rebuild=True)
        data_dict['num_filtered_objs'] = pack.num_filtered
        ppi = pack.points.shape[1]
        ws = self.net.workspace(pack.M * ppi, device)
        ops.stamp('attr:start')
        ops.encoder_reset(ws)
        ops.voxelize(pack.points, pack.cand_rows, float(self.voxel_size[0]), ws)
        ops.stamp('attr:voxelized')
        f4, c4, n4 = self.net.encode(ws)
        ops.stamp('attr:features')
        data_dict['obj_feats'] = ops.segmax(f4, c4, n4, ws.n_max, pack.M)
        ops.stamp('attr:pooled')
        data_dict['pred_obb_batch'] = pack.pred_obb_batch
        return data_dict

    def embed_language(self, data_dict):
        """Language side only (needs nothing from the encoder): Linear-BN-ReLU-Linear + L2 norm."""
        p = self.prepared()
        lang_emb, _ = ops.mlp_head(data_dict['lang_attr_feats'].float().contiguous(), p['lw1'], p['lb1'],
                                   ops.NORM_AFFINE, p['lg'], p['lbeta'], p['lw2'], p['lb2'], ops.MODE_L2)
        return lang_emb

    def prepare_maps(self, data_dict, device, pack=None):
        """Phase A1: host class filter + packed H2D, GPU voxelisation, levels + kernel maps."""
        require_eval(self)
        pack = pack or get_pack(data_dict, self.args, device, rebuild=True)
        data_dict['num_filtered_objs'] = pack.num_filtered
        data_dict['pred_obb_batch'] = pack.pred_obb_batch
        ws = self.net.workspace(pack.M * pack.points.shape[1], device)
        ops.encoder_reset(ws)
        ops.voxelize(pack.points, pack.cand_rows, float(self.voxel_size[0]), ws)
        self.net.build_maps(ws)
        return ws, pack

    def pool(self, data_dict, ws, f4, pack):
        """Phase A3: per-candidate max-pool of the stride-16 features (:105)."""
        data_dict['obj_feats'] = ops.segmax(f4, ws.coords(4), ws.nlvl()[4:5], ws.n_max, pack.M)
        return data_dict

    def match(self, data_dict):
        """Phase B: language side Linear-BN-ReLU-Linear + L2 normalise (:88-90); visual side
        Linear-LN-ReLU-Linear + L2 normalise, dot with the scene's language vector (:108-126)."""
        p = self.prepared()
        pack = data_dict[_KEY]
        lang_emb = data_dict.pop('_ir_attr_lang', None)
        if lang_emb is None:
            lang_emb = self.embed_language(data_dict)
        _, scores = ops.mlp_head(data_dict['obj_feats'], p['vw1'], p['vb1'], ops.NORM_LAYER, p['vg'], p['vbeta'],
                                 p['vw2'], p['vb2'], ops.MODE_DOT, partner=lang_emb, seg=pack.cand_scene)
        data_dict['attribute_scores'] = scores
        ops.stamp('attr:matched')
        return data_dict

    def forward(self, data_dict):
        ops.check_device()
        if self.training:
            from . import training
            pack = get_pack(data_dict, self.args, data_dict['lang_attr_feats'].device, rebuild=True)
            return training.attribute_forward_train(self, data_dict, pack)
        data_dict = self.encode_candidates(data_dict, data_dict['lang_attr_feats'].device)
        return self.match(data_dict)
