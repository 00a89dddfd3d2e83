"""Language encoder, drop-in for the reference's ``models/lang_module.py`` (same class name,
constructor signature, data_dict keys and state_dict layout) running on the CUDA library:
word MLP -> 2-layer packed biGRU (recurrent matrix resident in shared memory) -> four masked
attention poolings -> 18-way classifier.  Reference lines: models/lang_module.py:8-49 (ctor),
:51-93 (rnn_encoding), :95-108 (forward)."""
import torch
import torch.nn as nn

from . import ops
from .basic_blocks import PrepCache, require_eval


class LangModule(nn.Module, PrepCache):
    def __init__(self, num_text_classes, use_lang_classifier=True, use_bidir=False, emb_size=300,
                 hidden_size=256):
        super().__init__()
        self.num_text_classes = num_text_classes
        self.use_lang_classifier = use_lang_classifier
        self.use_bidir = use_bidir
        self.hidden_size = hidden_size
        self.gru = nn.GRU(input_size=256, hidden_size=hidden_size, num_layers=2, batch_first=True,
                          bidirectional=self.use_bidir)
        h_dim = 256
        self.word_projection = nn.Sequential(nn.Linear(emb_size, h_dim), nn.ReLU(), nn.Dropout(0.1),
                                             nn.Linear(h_dim, h_dim), nn.ReLU())
        o_dim = 128 * (1 + self.use_bidir)
        self.fc_a = nn.Linear(o_dim, 1)
        self.fc_cls = nn.Linear(o_dim, 1)
        self.fc_rel = nn.Linear(o_dim, 1)
        self.fc_scene = nn.Linear(o_dim, 1)
        if use_lang_classifier:
            self.lang_cls = nn.Sequential(nn.Linear(256, num_text_classes))

    def _prepare(self):
        if not self.use_bidir or self.hidden_size != 128:
            raise NotImplementedError("CUDA GRU path is built for the reference configuration: "
                                      "bidirectional, hidden 128 (models/instancerefer.py:21)")
        f = lambda t: t.detach().float().contiguous()
        g = self.gru
        prep = dict(w0=f(self.word_projection[0].weight), b0=f(self.word_projection[0].bias),
                    w3=f(self.word_projection[3].weight), b3=f(self.word_projection[3].bias))
        for l in (0, 1):
            prep[f'wih{l}'] = f(torch.cat([getattr(g, f'weight_ih_l{l}'), getattr(g, f'weight_ih_l{l}_reverse')], 0))
            prep[f'bih{l}'] = f(torch.cat([getattr(g, f'bias_ih_l{l}'), getattr(g, f'bias_ih_l{l}_reverse')], 0))
            prep[f'whh{l}'] = f(torch.stack([getattr(g, f'weight_hh_l{l}'), getattr(g, f'weight_hh_l{l}_reverse')], 0))
            prep[f'bhh{l}'] = f(torch.stack([getattr(g, f'bias_hh_l{l}'), getattr(g, f'bias_hh_l{l}_reverse')], 0))
        fcs = (self.fc_a, self.fc_cls, self.fc_rel, self.fc_scene)
        prep['fcw'] = f(torch.cat([m.weight for m in fcs], 0))
        prep['fcb'] = f(torch.cat([m.bias for m in fcs], 0))
        if self.use_lang_classifier:
            prep['wc'], prep['bc'] = f(self.lang_cls[0].weight), f(self.lang_cls[0].bias)
        return prep

    def rnn_encoding(self, embed_in, length, data_dict):
        require_eval(self)
        p = self.prepared()
        dev = embed_in.device
        len_dev = length.to(dev, torch.int64).contiguous()
        B = embed_in.shape[0]
        if data_dict.get('_ir_lang_len_max') is not None:                      # caller already knows max(len)
            L = int(data_dict['_ir_lang_len_max'])
        else:                                                                  # one small D2H (ref: :60)
            L = int((length.detach().to('cpu') if length.is_cuda else length).max())
        x = embed_in[:, :L].float().contiguous().view(B * L, -1)               # only the L live tokens
        e = ops.linear(ops.linear(x, p['w0'], p['b0'], relu=True), p['w3'], p['b3'], relu=True)
        h = e
        for l in (0, 1):
            xp = ops.linear(h, p[f'wih{l}'], p[f'bih{l}'])                      # (B*L, 2*3H) hoisted input GEMM
            h = ops.gru_layer(xp, p[f'whh{l}'], p[f'bhh{l}'], len_dev, B, L).view(B * L, -1)
        feats = h.view(B, L, -1)
        embed = e.view(B, L, -1)
        data_dict['lang_feat'] = feats                                         # overwritten, as in the reference (:58)
        atten, pooled = ops.token_attention(feats, embed, len_dev, p['fcw'], p['fcb'])
        data_dict['atten_attr'] = atten[0]
        data_dict['atten_rel'] = atten[2]
        data_dict['atten_scene'] = atten[3]
        data_dict['lang_attr_feats'] = pooled[0]
        data_dict['lang_cls_feats'] = pooled[1]
        data_dict['lang_rel_feats'] = pooled[2]
        data_dict['lang_scene_feats'] = pooled[3]
        return data_dict

    def forward(self, data_dict):
        ops.check_device()
        if self.training:                                   # batch-statistics / Dropout / autograd path
            from . import training
            return training.lang_forward_train(self, data_dict)
        data_dict = self.rnn_encoding(data_dict['lang_feat'], data_dict['lang_len'], data_dict)
        if self.use_lang_classifier:
            p = self.prepared()
            data_dict['lang_scores'] = ops.linear(data_dict['lang_cls_feats'], p['wc'], p['bc'])
        return data_dict

    def length_to_mask(self, length, max_len=None, dtype=None):
        """(B,) -> (B,max_len) bool mask, arange < length (models/lang_module.py:127-139)."""
        assert len(length.shape) == 1, "Length shape should be 1 dimensional."
        max_len = max_len or length.max().item()
        mask = torch.arange(max_len, device=length.device, dtype=length.dtype).expand(len(length), max_len) \
            < length.unsqueeze(1)
        return mask if dtype is None else mask.to(dtype)
