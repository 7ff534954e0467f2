"""R-GCN layers with the reference's constructor / forward signatures and parameter names
(rgcn/layers.py:7-91 RGCNLayer, :147-179 RGCNBlockLayer, :182-279 UnionRGCNLayer), computing on the
sm_100a kernels.  eval(): the inference kernels under no_grad (dropout = identity).  train(): UnionRGCNLayer runs on the
kernel-backed autograd nodes of regcn_b200/train.py (dropout, gradients, the live skip gate when prev_h is given); the
layers without backward kernels of their own (RGCNBlockLayer outside the static-graph path) raise instead of returning
values without gradients.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

_RELU_GAIN = nn.init.calculate_gain('relu')


def _act_code(activation):
    if activation is None:
        return 0
    if activation is F.rrelu:
        return 1
    raise NotImplementedError("regcn_b200 layers support activation=None or F.rrelu (the reference's only choice)")


def _no_train_dropout(mod):
    if mod.training and getattr(mod, "dropout", None) is not None and mod.dropout.p > 0:
        raise NotImplementedError("regcn_b200: a layer's standalone forward() is the inference path (kernels under no_grad); training-mode dropout and gradients run through the model's get_loss() (regcn_b200/train.py, train_hyp.py)")


class UnionRGCNLayer(nn.Module):
    """rgcn/layers.py:182-279.  out = act( norm * sum_in (h[src]+rel[type]) W_n + where(indeg>0, h W_loop, h W_evolve) )."""

    def __init__(self, in_feat, out_feat, num_rels, num_bases=-1, bias=None, activation=None, self_loop=False,
                 dropout=0.0, skip_connect=False, rel_emb=None):
        super().__init__()
        self.in_feat, self.out_feat = in_feat, out_feat
        self.bias = bias
        self.activation = activation
        self.self_loop = self_loop
        self.num_rels = num_rels
        self.rel_emb = None
        self.skip_connect = skip_connect
        self.weight_neighbor = nn.Parameter(torch.Tensor(in_feat, out_feat))
        nn.init.xavier_uniform_(self.weight_neighbor, gain=_RELU_GAIN)
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_RELU_GAIN)
            self.evolve_loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.evolve_loop_weight, gain=_RELU_GAIN)
        if self.skip_connect:
            self.skip_connect_weight = nn.Parameter(torch.Tensor(out_feat, out_feat))
            nn.init.xavier_uniform_(self.skip_connect_weight, gain=_RELU_GAIN)
            self.skip_connect_bias = nn.Parameter(torch.Tensor(out_feat))
            nn.init.zeros_(self.skip_connect_bias)
        self.dropout = nn.Dropout(dropout) if dropout else None

    def _loop_cat(self):
        # [W_loop | W_evolve] (d, 2d): one GEMM yields both self-loop candidates, the combine kernel picks per row.
        # Cached per parameter version so the concatenation is not redone every snapshot.
        key = (self.loop_weight._version, self.evolve_loop_weight._version, self.loop_weight.data_ptr())
        if getattr(self, "_loop_cat_key", None) != key:
            self._loop_cat_val = torch.cat((self.loop_weight.detach(), self.evolve_loop_weight.detach()), dim=1).contiguous()
            self._loop_cat_key = key
        return self._loop_cat_val

    def _forward_train(self, g, emb_rel, prev_h=None):
        """train() mode with gradients and dropout: the kernel-backed autograd nodes RecurrentRGCN's training path uses
        (regcn_b200/train.py), for a caller that drives the layer itself (rgcn/layers.py:222-255).  prev_h: the LIVE skip
        gate (rgcn/layers.py:234-245): sigmoid(prev_h W_s + b_s) mixes the node representation with prev_h in front of
        the activation -- the time-gate node without its normalisation."""
        from . import train as T
        with torch.enable_grad():
            h = g.ndata['h']
            p = float(self.dropout.p) if self.dropout is not None else 0.0
            P = T.linear(T.union_aggregate(h, emb_rel, g), self.weight_neighbor, None, True)
            L = T.linear(h, torch.cat((self.loop_weight, self.evolve_loop_weight), dim=1), None, True)
            if prev_h is None:
                out = T.union_combine(P, L, g, p)
            else:
                S = T.linear(prev_h, self.skip_connect_weight, None, True)
                out = T.rrelu_drop(T.time_gate(S, self.skip_connect_bias, T.union_sum(P, L, g), prev_h, False), p)
        g.ndata['h'] = out
        return out

    @torch.no_grad()
    def forward(self, g, prev_h, emb_rel):
        if self.training and self.self_loop and self.activation is F.rrelu:
            self.rel_emb = emb_rel
            live = len(prev_h) != 0 and self.skip_connect
            return self._forward_train(g, emb_rel, prev_h if live else None)
        _no_train_dropout(self)
        self.rel_emb = emb_rel
        h = g.ndata['h']
        agg = ops.union_aggregate(h, emb_rel, g)                      # K4: (h[src]+rel[type]) summed, norm applied
        P = ops.gemm(agg, self.weight_neighbor, b_key=(self.weight_neighbor, "w"))   # aggregate-then-transform
        lc = self._loop_cat() if self.self_loop else None
        L = ops.gemm(h, lc, b_key=(lc, "w")) if self.self_loop else None
        S = sb = prev = None
        if len(prev_h) != 0 and self.skip_connect:
            S = ops.gemm(prev_h, self.skip_connect_weight, b_key=(self.skip_connect_weight, "w"))
            sb, prev = self.skip_connect_bias, prev_h
        out, _, _ = ops.union_combine(P, L, g.indeg, act=_act_code(self.activation), skip=S, skip_bias=sb, prev=prev)
        g.ndata['h'] = out
        return out


class RGCNLayer(nn.Module):
    """rgcn/layers.py:7-91 base layer (propagate() supplied by the subclass)."""

    def __init__(self, in_feat, out_feat, bias=None, activation=None, self_loop=False, skip_connect=False,
                 dropout=0.0, layer_norm=False):
        super().__init__()
        self.bias = bias
        self.activation = activation
        self.self_loop = self_loop
        self.skip_connect = skip_connect
        self.layer_norm = layer_norm
        if self.bias:
            self.bias = nn.Parameter(torch.Tensor(out_feat))
            nn.init.zeros_(self.bias)
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_RELU_GAIN)
        if self.skip_connect:
            self.skip_connect_weight = nn.Parameter(torch.Tensor(out_feat, out_feat))
            nn.init.xavier_uniform_(self.skip_connect_weight, gain=_RELU_GAIN)
            self.skip_connect_bias = nn.Parameter(torch.Tensor(out_feat))
            nn.init.zeros_(self.skip_connect_bias)
        self.dropout = nn.Dropout(dropout) if dropout else None
        if self.layer_norm:
            self.normalization_layer = nn.LayerNorm(out_feat, elementwise_affine=False)

    def propagate(self, g):
        raise NotImplementedError

    @torch.no_grad()
    def forward(self, g, prev_h=[]):
        _no_train_dropout(self)
        if self.self_loop or self.bias or self.layer_norm or (len(prev_h) != 0 and self.skip_connect):
            # The reference only instantiates this base through RGCNBlockLayer(self_loop=False, skip_connect=False,
            # bias=None) for the static graph (src/rrgcn.py:104-105); other combinations are out of scope.
            raise NotImplementedError("regcn_b200.RGCNLayer: only the static-graph configuration is implemented")
        agg = self.propagate(g)
        out, _, _ = ops.union_combine(agg, None, None, act=_act_code(self.activation))
        g.ndata['h'] = out
        return out


class RGCNBlockLayer(RGCNLayer):
    """rgcn/layers.py:147-179: per-edge block-diagonal transform, sum, degree norm."""

    def __init__(self, in_feat, out_feat, num_rels, num_bases, bias=None, activation=None, self_loop=False,
                 dropout=0.0, skip_connect=False, layer_norm=False):
        super().__init__(in_feat, out_feat, bias, activation, self_loop=self_loop, skip_connect=skip_connect,
                         dropout=dropout)
        self.num_rels = num_rels
        self.num_bases = num_bases
        assert self.num_bases > 0
        self.out_feat = out_feat
        self.submat_in = in_feat // self.num_bases
        self.submat_out = out_feat // self.num_bases
        self.weight = nn.Parameter(torch.Tensor(self.num_rels, self.num_bases * self.submat_in * self.submat_out))
        nn.init.xavier_uniform_(self.weight, gain=_RELU_GAIN)

    def propagate(self, g):
        return ops.block_aggregate(g.ndata['h'], self.weight, g, self.num_bases, self.out_feat)
