"""Model shell, drop-in for the reference's ``models/instancerefer.py``: same constructor
(``InstanceRefer(input_feature_dim, args)``), same YAML-driven plugin loading by module name
(models/instancerefer.py:20-34) and the same ``forward(data_dict) -> data_dict`` chain (:56-70).
Module names given in the YAML are resolved inside this package first (``lang_module`` ->
``instancerefer_b200.lang_module``), then as plain import names, so the reference's
config/InstanceRefer.yaml works unchanged."""
import importlib

import torch
import torch.nn as nn

from . import ops
from .candidates import KEY as _PACK_KEY


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

    def forward(self, data_dict):
        data_dict.pop(_PACK_KEY, None)
        with torch.no_grad():
            data_dict = self.lang(data_dict)
            if self.args.attribute_module:
                data_dict = self.attribute(data_dict)
            if self.args.relation_module:
                data_dict = self.relation(data_dict)
            if self.args.scene_module:
                data_dict = self.scene(data_dict)
            if self.args.attribute_module and self.args.relation_module and self.args.scene_module:
                # extra fused output (not in the reference dict): per-scene softmax / argmax over
                # candidates of the summed score (the reference does this on the host,
                # lib/eval_helper.py:61-67)
                pack = data_dict[_PACK_KEY]
                nf = [n for n in pack.num_filtered if n >= 2]
                ofs = torch.tensor([0] + list(torch.tensor(nf).cumsum(0).tolist()), dtype=torch.int32)
                ofs = ofs.pin_memory().to(data_dict['attribute_scores'].device, non_blocking=True)
                prob, arg = ops.candidate_softmax(data_dict['attribute_scores'], data_dict['relation_scores'],
                                                  data_dict['scene_scores'], ofs)
                data_dict['ref_probs'], data_dict['ref_pred'] = prob, arg
        return data_dict
