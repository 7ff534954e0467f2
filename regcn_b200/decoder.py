"""ConvTransE / ConvTransR decoders (src/decoder.py:10-100) on the sm_100a kernels.

Same constructor / forward signatures, same parameter and buffer names (conv1, bn0, bn1, bn2, bn3, bn_init,
fc, b) so reference checkpoints load.  Inference path: BatchNorm uses running statistics (eval mode),
dropout is the identity; training mode raises until the backward kernels exist.
"""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops


def _fold_bn(bn):
    """Eval-mode BatchNorm1d as y = x*scale + shift; cached on the module until one of its tensors changes."""
    ts = [t for t in (bn.running_var, bn.running_mean, bn.weight, bn.bias) if t is not None]
    stamp = tuple((t._version, t.data_ptr()) for t in ts)
    hit = bn.__dict__.get("_regcn_fold")
    if hit is not None and hit[0] == stamp:
        return hit[1]
    with torch.no_grad():
        inv = torch.rsqrt(bn.running_var + bn.eps)
        w = bn.weight if bn.weight is not None else torch.ones_like(inv)
        b = bn.bias if bn.bias is not None else torch.zeros_like(inv)
        scale = (w * inv).contiguous()
        shift = (b - bn.running_mean * w * inv).contiguous()
    bn.__dict__["_regcn_fold"] = (stamp, (scale, shift))
    return scale, shift


class _ConvTransBase(nn.Module):
    def __init__(self, n_out_bias, embedding_dim, input_dropout, hidden_dropout, feature_map_dropout, channels,
                 kernel_size):
        super().__init__()
        self.inp_drop = nn.Dropout(input_dropout)
        self.hidden_drop = nn.Dropout(hidden_dropout)
        self.feature_map_drop = nn.Dropout(feature_map_dropout)
        self.loss = nn.BCELoss()
        self.conv1 = nn.Conv1d(2, channels, kernel_size, stride=1, padding=int(math.floor(kernel_size / 2)))
        self.bn0 = nn.BatchNorm1d(2)
        self.bn1 = nn.BatchNorm1d(channels)
        self.bn2 = nn.BatchNorm1d(embedding_dim)
        self.register_parameter('b', Parameter(torch.zeros(n_out_bias)))
        self.fc = nn.Linear(embedding_dim * channels, embedding_dim)
        self.bn3 = nn.BatchNorm1d(embedding_dim)
        self.bn_init = nn.BatchNorm1d(embedding_dim)

    def _check_eval(self):
        if self.training:
            raise NotImplementedError("regcn_b200 decoders: the standalone forward() is the inference path (folded BatchNorm, no dropout); batch-statistics BatchNorm, dropout and gradients run through the model's get_loss() (regcn_b200/train.py, train_hyp.py)")

    def _tower(self, ent_act, second, triplets, col0, col1, always_bn2, batch_total=None):
        """K10: bn0 -> conv1d(2->C,k) -> bn1 -> relu -> fc -> bn2 -> relu, returns the (B,d) query matrix.
        batch_total: size of the whole query batch when `triplets` is one rank's slice of it (query-sharded tower): the
        split-K factor of the FC and the B == 1 rule of bn2 then follow the whole batch, so a row's value does not depend
        on how the batch was cut."""
        B = len(triplets) if batch_total is None else int(batch_total)
        if ops.gemm_impl() == "tc" and ops.convtrans_fc_ok(ent_act.shape[1], self.conv1.weight, self.fc.out_features):
            # the feature map is computed inside the FC GEMM's operand ring (regcn_convtrans_fc): never written or read
            x = ops.convtrans_fc(ent_act, second, triplets, col0, col1, _fold_bn(self.bn0), self.conv1.weight.detach(),
                                 self.conv1.bias.detach(), _fold_bn(self.bn1), self.fc.weight,
                                 self.fc.bias.detach(), batch_total=B,
                                 bn2=_fold_bn(self.bn2) if (always_bn2 or B > 1) else None, relu=True)
            return x
        feats = ops.convtranse_features(ent_act, second, triplets, col0, col1, _fold_bn(self.bn0),
                                        self.conv1.weight.detach(), self.conv1.bias.detach(), _fold_bn(self.bn1),
                                        split=False)
        K = feats.shape[1]
        split_k = max(1, min(16, (148 * 2) // max(1, ((B + 127) // 128) * ((self.fc.out_features + 127) // 128))))
        split_k = min(split_k, max(1, K // 512))
        # the feature map stays one fp32 matrix (B, 50 d): the FC GEMM splits it to TF32 on chip
        x = ops.gemm(feats, self.fc.weight.detach(), trans_b=True, bias=self.fc.bias.detach(), split_k=split_k,
                     b_key=(self.fc.weight, "w"), split_a_on_chip=True)
        if always_bn2 or B > 1:
            s, t = _fold_bn(self.bn2)
            ops.affine_relu_(x, s, t, relu=True)
        else:
            ops.affine_relu_(x, None, None, relu=True)
        return x


class ConvTransE(_ConvTransBase):
    """src/decoder.py:55-100.  score = relu(bn2(fc(relu(bn1(conv(bn0([tanh(E)[s]; rel[r]]))))))) . tanh(E)^T"""

    def __init__(self, num_entities, embedding_dim, input_dropout=0, hidden_dropout=0, feature_map_dropout=0,
                 channels=50, kernel_size=3, use_bias=True):
        super().__init__(num_entities, embedding_dim, input_dropout, hidden_dropout, feature_map_dropout, channels,
                         kernel_size)

    @torch.no_grad()
    def query(self, embedding, emb_rel, triplets, batch_total=None, normalize=False):
        """Returns (tanh(E), Q): the activated entity table and the (B,d) query matrix of the dot scoring.
        normalize: `embedding` is the evolved table BEFORE the predict-time F.normalize (src/rrgcn.py:190), applied here in
        the same pass as the tanh."""
        self._check_eval()
        e_all = ops.row_map(embedding, ops.ROW_NORMALIZE_TANH if normalize else ops.ROW_TANH,
                            split=ops.gemm_impl() == "tc" and ops.score_dtype() == "fp32")
        q = self._tower(e_all, emb_rel.contiguous(), triplets, 0, 1, always_bn2=False, batch_total=batch_total)
        return e_all, q

    def _forward_train(self, embedding, emb_rel, triplets, partial_embeding=None):
        """train() mode (batch-statistics BatchNorm, dropout, running-stat updates) with gradients: the kernel-backed
        autograd nodes of regcn_b200.train; the (B,N) scores are materialised because the caller asked for them."""
        from . import train as T
        with torch.enable_grad():
            t = torch.as_tensor(triplets).to(embedding.device).contiguous()
            e_all = T.tanh(embedding)
            q = T.conv_tower(self, e_all, emb_rel, t, 0, 1)
            return T.linear(q, e_all if partial_embeding is None else partial_embeding)

    @torch.no_grad()
    def forward(self, embedding, emb_rel, triplets, nodes_id=None, mode="train", negative_rate=0,
                partial_embeding=None):
        if self.training:
            return self._forward_train(embedding, emb_rel, triplets, partial_embeding)
        e_all, q = self.query(embedding, emb_rel, triplets)
        cand = e_all if partial_embeding is None else partial_embeding.contiguous()
        return ops.gemm(q, cand, trans_b=True)   # K11 (note: the reference registers `b` but never adds it, :72,:96-99)


class ConvTransR(_ConvTransBase):
    """src/decoder.py:10-52.  Same tower on [tanh(E)[s]; tanh(E)[o]], scored against emb_rel."""

    def __init__(self, num_relations, embedding_dim, input_dropout=0, hidden_dropout=0, feature_map_dropout=0,
                 channels=50, kernel_size=3, use_bias=True):
        super().__init__(num_relations * 2, embedding_dim, input_dropout, hidden_dropout, feature_map_dropout,
                         channels, kernel_size)

    @torch.no_grad()
    def forward(self, embedding, emb_rel, triplets, nodes_id=None, mode="train", negative_rate=0):
        if self.training:
            from . import train as T
            with torch.enable_grad():
                t = torch.as_tensor(triplets).to(embedding.device).contiguous()
                e_all = T.tanh(embedding)
                return T.linear(T.conv_tower(self, e_all, e_all, t, 0, 2), emb_rel)
        e_all = ops.row_map(embedding, ops.ROW_TANH)
        q = self._tower(e_all, e_all, triplets, 0, 2, always_bn2=True)
        return ops.gemm(q, emb_rel.contiguous(), trans_b=True)
