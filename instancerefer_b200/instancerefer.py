"""Model shell, drop-in for the reference's ``models/instancerefer.py``: same constructor
(``InstanceRefer(input_feature_dim, args)``), same YAML-driven plugin loading by module name
(models/instancerefer.py:20-34) and the same ``forward(data_dict) -> data_dict`` chain (:56-70).
Module names given in the YAML are resolved inside this package first (``lang_module`` ->
``instancerefer_b200.lang_module``), then as plain import names, so the reference's
config/InstanceRefer.yaml works unchanged."""
import importlib
import os

import torch
import torch.nn as nn

from . import ops
from .candidates import KEY as _PACK_KEY
from .candidates import CandidatePack, target_classes


def _load(name):
    try:
        return importlib.import_module(f'{__package__}.{name}')
    except ModuleNotFoundError:
        return importlib.import_module(name)


class InstanceRefer(nn.Module):
    def __init__(self, input_feature_dim=0, args=None):
        super().__init__()
        self.args = args
        self.lang = _load(args.language_module).LangModule(args.num_classes, True, args.use_bidir, 300, 128)
        if args.attribute_module:
            self.attribute = _load(args.attribute_module).AttributeModule(input_feature_dim, args)
        if args.relation_module:
            self.relation = _load(args.relation_module).RelationModule(input_feature_dim, args)
        if args.scene_module:
            self.scene = _load(args.scene_module).SceneModule(input_feature_dim, args)

    concurrent = True      # run the three language-independent encoders on side streams
    pair_encoders = False  # run both sparse encoders through shared launches (ir_encoder_features_pair)

    def _streams(self, device):
        st = self.__dict__.get('_side_streams')
        if st is None or st[0].device != device:
            st = [torch.cuda.Stream(device=device) for _ in range(3)]
            self.__dict__['_side_streams'] = st
        return st

    def forward_train(self, data_dict):
        """Train mode (lib/solver.py:196): batch-statistics BatchNorm, Dropout, outputs with autograd
        history; same chain and dict keys as the reference (models/instancerefer.py:56-70)."""
        from . import training as T
        a = self.args
        pack = data_dict.pop(_PACK_KEY, None)
        if not getattr(pack, 'resident', False):               # a pre-staged pack (graph replay) is reused
            pack = None
        full = bool(a.attribute_module and a.relation_module and a.scene_module and a.use_gt_lang)
        prep_a = prep_s = None
        if full:
            # host class filter + packed H2D, then the coordinate phase of both encoders and ONE read-back of
            # their level sizes, before any feature kernel is queued: the rest of the step is issued without
            # a host synchronisation
            dev = data_dict['lang_feat'].device
            if pack is None:
                pack = CandidatePack(data_dict, target_classes(data_dict, a), dev)
            if self.concurrent and data_dict.get('_ir_capacity') and os.environ.get('IR_TRAIN_STREAMS', '1') != '0':
                return self._forward_train_streams(data_dict, pack, dev)
            prep_a, prep_s = T.prepare_encoder_maps(self, data_dict, pack)
        data_dict = T.lang_forward_train(self.lang, data_dict)
        if not full and pack is None:
            dev = data_dict['lang_feat'].device
            pack = CandidatePack(data_dict, target_classes(data_dict, a), dev)
        data_dict[_PACK_KEY] = pack
        if a.attribute_module:
            data_dict = T.attribute_forward_train(self.attribute, data_dict, pack, prep_a)
        if a.relation_module:
            data_dict = T.relation_forward_train(self.relation, data_dict, pack)
        if a.scene_module:
            data_dict = T.scene_forward_train(self.scene, data_dict, pack, prep_s)
        return data_dict

    def _forward_train_streams(self, data_dict, pack, dev):
        """The train-mode chain on four streams, for the captured iteration (train_graph.GraphedTrainStep; capacity
        mode, so nothing below touches the host): instance encoder, scene encoder and relation graph start at once on
        three side streams while the language branch runs on the caller's; each branch's matching head joins the
        language event, the scene head also the pooled instance features.  Autograd replays every node on the stream
        its forward ran on, so the backward of the three branches forms parallel chains of the same CUDA graph —
        their latency-bound launches (a few thousand rows each) overlap instead of queueing behind each other.
        Cross-stream tensors (language features, obj_feats, scores) stay referenced by ``data_dict`` until the step
        returns, so no block is recycled while another stream still reads it; gradients crossing streams are
        synchronised and recorded by the autograd engine itself."""
        from . import training as T
        main = torch.cuda.current_stream(dev)
        sa, ss, sr = self._streams(dev)
        data_dict[_PACK_KEY] = pack
        for s_ in (sa, ss, sr):
            s_.wait_stream(main)
        ev_obj, ev_lang = torch.cuda.Event(), torch.cuda.Event()
        ops.stamp('fwd:start')
        with torch.cuda.stream(sa):
            ws_a = T.prepare_attribute_maps(self, pack)
            ops.stamp('fwd:attr:maps')
            T.attribute_encode_train(self.attribute, data_dict, pack, (ws_a, T.EncoderGraph(ws_a, [ws_a.n_max] * 5)))
            ev_obj.record(sa)
            ops.stamp('fwd:attr:encoded')
        with torch.cuda.stream(ss):
            ws_s, F0, C0 = T.prepare_scene_maps(self, data_dict, dev)
            ops.stamp('fwd:scene:maps')
            T.scene_encode_train(self.scene, data_dict, pack, (ws_s, T.EncoderGraph(ws_s, [ws_s.n_max] * 5), F0, C0))
            ops.stamp('fwd:scene:encoded')
        with torch.cuda.stream(sr):
            T.relation_encode_train(self.relation, data_dict, pack)
            ops.stamp('fwd:rel:encoded')
        data_dict = T.lang_forward_train(self.lang, data_dict)
        ev_lang.record(main)
        ops.stamp('fwd:lang:done')
        with torch.cuda.stream(sa):
            sa.wait_event(ev_lang)
            T.attribute_match_train(self.attribute, data_dict, pack)
            ops.stamp('fwd:attr:matched')
        with torch.cuda.stream(sr):
            sr.wait_event(ev_lang)
            T.relation_match_train(self.relation, data_dict, pack)
            ops.stamp('fwd:rel:matched')
        with torch.cuda.stream(ss):
            ss.wait_event(ev_lang)
            ss.wait_event(ev_obj)
            T.scene_match_train(self.scene, data_dict, pack)
            ops.stamp('fwd:scene:matched')
        for s_ in (sa, ss, sr):
            main.wait_stream(s_)
        ops.stamp('fwd:joined')
        return data_dict

    def forward(self, data_dict):
        ops.check_device()
        if self.training:
            return self.forward_train(data_dict)
        if not getattr(data_dict.get(_PACK_KEY), 'resident', False):
            data_dict.pop(_PACK_KEY, None)
        a = self.args
        full = bool(a.attribute_module and a.relation_module and a.scene_module)
        with torch.no_grad():
            if not (full and self.concurrent and a.use_gt_lang):
                # plain chain, exactly the reference order (models/instancerefer.py:56-70)
                data_dict = self.lang(data_dict)
                if a.attribute_module:
                    data_dict = self.attribute(data_dict)
                if a.relation_module:
                    data_dict = self.relation(data_dict)
                if a.scene_module:
                    data_dict = self.scene(data_dict)
            else:
                # same arithmetic, B200 schedule: the instance encoder, the scene encoder and the
                # relation graph do not depend on the language branch, so they run on three side
                # streams while the GRU runs on the caller's stream; the matching heads join them.
                dev = data_dict['lang_feat'].device
                main = torch.cuda.current_stream(dev)
                sa, ss, sr = self._streams(dev)
                pack = data_dict.get(_PACK_KEY)
                if pack is None:                       # host filter + packed H2D on the caller's stream
                    pack = CandidatePack(data_dict, target_classes(data_dict, a), dev)
                    data_dict[_PACK_KEY] = pack
                for s_ in (sa, ss, sr):
                    s_.wait_stream(main)
                ev_obj = torch.cuda.Event()
                if self.pair_encoders:
                    # both encoders through shared launches (measured slower than two overlapping chains
                    # at the bench size: kept as an option for small scenes)
                    ev_maps, ev_feat = torch.cuda.Event(), torch.cuda.Event()
                    with torch.cuda.stream(ss):
                        ws_s, F0 = self.scene.prepare_maps(data_dict, dev)
                        ev_maps.record(ss)
                    with torch.cuda.stream(sa):
                        ws_a, _ = self.attribute.prepare_maps(data_dict, dev, pack)
                        sa.wait_event(ev_maps)
                        pa, ps_ = self.attribute.net.prepared(), self.scene.net.prepared()
                        f4a = torch.empty(ws_a.n_max, 128, dtype=torch.float32, device=dev)
                        f4s = torch.empty(ws_s.n_max, 128, dtype=torch.float32, device=dev)
                        ops.encoder_features_pair(pa['params'], ws_a, None, f4a, ps_['params'], ws_s, F0, f4s)
                        ev_feat.record(sa)
                        self.attribute.pool(data_dict, ws_a, f4a, pack)
                        ev_obj.record(sa)
                    with torch.cuda.stream(ss):
                        ss.wait_event(ev_feat)
                        self.scene.bev_convs(data_dict, ws_s, f4s)
                else:
                    with torch.cuda.stream(sa):
                        self.attribute.encode_candidates(data_dict, dev, pack)
                        ev_obj.record(sa)                              # obj_feats ready (scene head needs it)
                    with torch.cuda.stream(ss):
                        self.scene.encode_scene(data_dict, dev)
                with torch.cuda.stream(sr):
                    ops.stamp('rel:start')
                    self.relation.encode_graph(data_dict, dev)
                    ops.stamp('rel:graph')
                ops.stamp('lang:start')
                data_dict = self.lang(data_dict)
                ops.stamp('lang:encoded')
                # language-side embeddings need nothing from the encoders
                data_dict['_ir_attr_lang'] = self.attribute.embed_language(data_dict)
                data_dict['_ir_rel_lang'] = self.relation.embed_language(data_dict)
                data_dict['_ir_scene_lang'] = self.scene.embed_language(data_dict)
                ops.stamp('lang:embedded')
                ev_lang = torch.cuda.Event()
                ev_lang.record(main)
                # each branch finishes with its own matching head on its own stream
                with torch.cuda.stream(sa):
                    sa.wait_event(ev_lang)
                    self.attribute.match(data_dict)
                with torch.cuda.stream(sr):
                    sr.wait_event(ev_lang)
                    self.relation.match(data_dict)
                    ops.stamp('rel:matched')
                with torch.cuda.stream(ss):
                    ss.wait_event(ev_lang)
                    ss.wait_event(ev_obj)
                    self.scene.match(data_dict)
                for s_ in (sa, ss, sr):
                    main.wait_stream(s_)
            if full:
                # extra fused output (not in the reference dict): per-scene softmax / argmax over
                # candidates of the summed score (the reference does this on the host,
                # lib/eval_helper.py:61-67)
                pack = data_dict[_PACK_KEY]
                prob, arg = ops.candidate_softmax(data_dict['attribute_scores'], data_dict['relation_scores'],
                                                  data_dict['scene_scores'], pack.cand_ofs)
                data_dict['ref_probs'], data_dict['ref_pred'] = prob, arg
                ops.stamp('end')
        return data_dict
