"""Relation (RP) module, drop-in for the reference's ``models/relation_module.py``: per-instance
25-d features (OBB centre, mean colour/height, class one-hot), per-scene kNN graph from candidates
to all instances, fused EdgeConv with max aggregation, cosine match with the language feature.
Reference lines: models/relation_module.py:8-36 (ctor), :38-78 (filter_candidates), :80-107."""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .basic_blocks import DynamicEdgeConv, PrepCache, fold_bn, require_eval
from .candidates import KEY as _KEY
from .candidates import get_pack


class RelationModule(nn.Module, PrepCache):
    def __init__(self, input_feature_dim, args, v_dim=128, h_dim=128, l_dim=256, dropout_rate=0.15):
        super().__init__()
        self.args = args
        self.input_feature_dim = input_feature_dim
        self.vis_emb_fc = nn.Sequential(nn.Linear(v_dim, h_dim), nn.LayerNorm(h_dim), nn.ReLU(),
                                        nn.Dropout(dropout_rate), nn.Linear(h_dim, h_dim))
        self.lang_emb_fc = nn.Sequential(nn.Linear(l_dim, h_dim), nn.BatchNorm1d(h_dim), nn.ReLU(),
                                         nn.Dropout(dropout_rate), nn.Linear(h_dim, h_dim))
        self.gcn = DynamicEdgeConv(input_feature_dim + args.num_classes, 128, k=args.k,
                                   num_classes=args.num_classes)
        self.one_hot_array = np.eye(args.num_classes)
        self.weight_initialization()

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _prep_tensors(self):
        return list(self.vis_emb_fc.parameters()) + list(self.lang_emb_fc.parameters()) + list(self.lang_emb_fc.buffers())

    def _prepare(self):
        f = lambda t: t.detach().float().contiguous()
        ft = lambda t: t.detach().float().t().contiguous()         # (in,out) layout for ir_mlp_head
        v, l = self.vis_emb_fc, self.lang_emb_fc
        s, b = fold_bn(l[1])
        return dict(vw1=ft(v[0].weight), vb1=f(v[0].bias), vg=f(v[1].weight), vbeta=f(v[1].bias),
                    vw2=ft(v[4].weight), vb2=f(v[4].bias),
                    lw1=ft(l[0].weight), lb1=f(l[0].bias), lg=s, lbeta=b, lw2=ft(l[4].weight), lb2=f(l[4].bias))

    def encode_graph(self, data_dict, device):
        """Phase A (no language dependency): per-instance 25-d rows (:66-73: mean of the points with
        xyz := OBB centre, + class one-hot), per-scene kNN, fused EdgeConv (:100)."""
        require_eval(self)
        pack = get_pack(data_dict, self.args, device)
        mean = ops.instance_mean(pack.points)
        ncls = self.args.num_classes
        cls_ids = torch.arange(ncls, device=device, dtype=torch.float32)
        onehot = (pack.centres_cls[:, 3:4] == cls_ids[None, :]).float()          # capture-safe one-hot
        xyz = pack.centres_cls[:, :3].contiguous()
        feats = torch.cat([xyz, mean[:, 3:], onehot], 1).contiguous()
        g, nbr = self.gcn(xyz, pack.inst_ofs, pack.cand_rows, pack.cand_seg, feats)
        data_dict['_ir_gcn'], data_dict['_ir_knn'] = g, nbr
        return data_dict

    def embed_language(self, data_dict):
        p = self.prepared()
        lang_emb, _ = ops.mlp_head(data_dict['lang_rel_feats'].float().contiguous(), p['lw1'], p['lb1'],
                                   ops.NORM_AFFINE, p['lg'], p['lbeta'], p['lw2'], p['lb2'], ops.MODE_RAW)
        return lang_emb

    def match(self, data_dict):
        """Phase B: language MLP (:82), visual MLP, cosine (:101-103)."""
        p = self.prepared()
        pack = data_dict[_KEY]
        lang_emb = data_dict.pop('_ir_rel_lang', None)
        if lang_emb is None:
            lang_emb = self.embed_language(data_dict)
        _, scores = ops.mlp_head(data_dict['_ir_gcn'], p['vw1'], p['vb1'], ops.NORM_LAYER, p['vg'], p['vbeta'],
                                 p['vw2'], p['vb2'], ops.MODE_COS, partner=lang_emb, seg=pack.cand_scene)
        data_dict['relation_scores'] = scores
        return data_dict

    def forward(self, data_dict):
        ops.check_device()
        if self.training:
            from . import training
            pack = get_pack(data_dict, self.args, data_dict['lang_rel_feats'].device)
            return training.relation_forward_train(self, data_dict, pack)
        data_dict = self.encode_graph(data_dict, data_dict['lang_rel_feats'].device)
        return self.match(data_dict)
