"""Hyperbolic encoder layers with the reference's signatures and parameter names
(hyperbolic_src/hyperbolic_layers.py:21-161 HyperbolicRGCNLayer, :164-323 HyperbolicUnionRGCNLayer, :524-743 LorentzRGCNLayer/Cell;
hyperbolic_src/hyperbolic_model.py:114-154 HyperbolicRGCNCell) on the sm_100a kernels.

Layer inputs/outputs are points on the Poincare ball like the reference.  Internally each layer works on the
tangent-space copy log_0(h) and the radius |h| of its input; when a layer is called by the cells below those are
handed over from the producer kernel (the previous layer's combine epilogue already computed them), otherwise
they are computed on entry.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import _act_code, _no_train_dropout

_RELU_GAIN = nn.init.calculate_gain('relu')


class _LoopMixin:
    def _loop_cat(self):
        key = (self.loop_weight._version, self.evolve_loop_weight._version, self.loop_weight.data_ptr())
        if getattr(self, "_loop_cat_key", None) != key:
            self._loop_cat_val = torch.cat((self.loop_weight.detach(), self.evolve_loop_weight.detach()), dim=1).contiguous()
            self._loop_cat_key = key
        return self._loop_cat_val


class HyperbolicRGCNLayer(nn.Module):
    """hyperbolic_layers.py:21-161: RGCN message passing in tangent space with block-diagonal relation transforms.

        t = log_0(h);  agg[v] = norm[v] * sum_{(u,r)->v} exp(-gamma |rad_u - rad_v|) * blockdiag(W[r]) t[u]     (K6)
        x = agg (+ t W_loop) -> optional skip gate with log_0(prev_h) -> activation -> exp_0

    Same constructor, parameter names (`weight`, `loop_weight`, `skip_weight`, `skip_bias`) and forward signature as the
    reference.  Unlike the union layers it has no evolve-loop weight and no +-10 tangent clamps."""

    def __init__(self, in_feat, out_feat, num_rels, num_bases=-1, c=0.01, activation=None, self_loop=False,
                 dropout=0.0, skip_connect=False, radius_msg_gamma=1.0):
        super().__init__()
        self.in_feat, self.out_feat, self.num_rels = in_feat, out_feat, num_rels
        self.num_bases = num_bases if num_bases > 0 else num_rels
        if self.num_bases > self.num_rels:
            self.num_bases = self.num_rels
        self.c = c
        self.activation = activation
        self.self_loop = self_loop
        self.skip_connect = skip_connect
        self.radius_msg_gamma = radius_msg_gamma
        self.submat_in = in_feat // self.num_bases
        self.submat_out = out_feat // self.num_bases
        self.weight = nn.Parameter(torch.Tensor(self.num_rels, self.num_bases * self.submat_in * self.submat_out))
        nn.init.xavier_uniform_(self.weight, gain=_RELU_GAIN)
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_RELU_GAIN)
        if self.skip_connect:
            self.skip_weight = nn.Parameter(torch.Tensor(out_feat, out_feat))
            nn.init.xavier_uniform_(self.skip_weight, gain=_RELU_GAIN)
            self.skip_bias = nn.Parameter(torch.zeros(out_feat))
        self.dropout = nn.Dropout(dropout) if dropout > 0 else None

    @torch.no_grad()
    def forward(self, g, h_hyper, rel_emb=None, prev_h=None):
        _no_train_dropout(self)
        if self.in_feat % self.num_bases or self.out_feat % self.num_bases:
            raise NotImplementedError("HyperbolicRGCNLayer: num_bases must divide in_feat and out_feat")
        act = _act_code(self.activation)
        ht, _, rad = ops.hyp_tangent(h_hyper, self.c, want_clamped=False, want_radius=True)
        x = ops.block_aggregate(ht, self.weight.detach(), g, self.num_bases, self.out_feat, radius=rad,
                                gamma=self.radius_msg_gamma)
        if self.self_loop:
            ops.gemm(ht, self.loop_weight, b_key=(self.loop_weight, "w"), out=x, accumulate=True)
        if self.skip_connect and prev_h is not None:
            prev_t, _, _ = ops.hyp_tangent(prev_h, self.c, want_clamped=False, want_radius=False)
            G = ops.gemm(prev_t, self.skip_weight, b_key=(self.skip_weight, "w"))
            x = ops.time_gate(G, self.skip_bias.detach(), x, prev_t, False)      # sigmoid(G + b) * x + (1 - sigmoid) * prev_t
        return ops.row_map(x, ops.ROW_RRELU_EXP0 if act else ops.ROW_EXP0, c=self.c)


class HyperbolicUnionRGCNLayer(nn.Module, _LoopMixin):
    """hyperbolic_layers.py:164-323."""

    def __init__(self, in_feat, out_feat, num_rels, num_bases=-1, c=0.01, activation=None, self_loop=False,
                 dropout=0.0, skip_connect=False, radius_msg_gamma=1.0):
        super().__init__()
        self.in_feat, self.out_feat, self.num_rels = in_feat, out_feat, num_rels
        self.c = c
        self.activation = activation
        self.self_loop = self_loop
        self.skip_connect = skip_connect
        self.rel_emb = None
        self.radius_msg_gamma = radius_msg_gamma
        self.weight_neighbor = nn.Parameter(torch.Tensor(in_feat, out_feat))
        nn.init.xavier_uniform_(self.weight_neighbor, gain=_RELU_GAIN)
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_RELU_GAIN)
            self.evolve_loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.evolve_loop_weight, gain=_RELU_GAIN)
        if self.skip_connect:
            self.skip_weight = nn.Parameter(torch.Tensor(out_feat, out_feat))
            nn.init.xavier_uniform_(self.skip_weight, gain=_RELU_GAIN)
            self.skip_bias = nn.Parameter(torch.zeros(out_feat))
        self.dropout = nn.Dropout(dropout) if dropout > 0 else None

    @torch.no_grad()
    def forward(self, g, h_hyper, rel_emb, prev_h=None, _tangent=None, _radius=None, _want_next=False):
        _no_train_dropout(self)
        self.rel_emb = rel_emb
        if _tangent is None or _radius is None:
            _tangent, _, _radius = ops.hyp_tangent(h_hyper, self.c, want_clamped=False, want_radius=True)
        agg = ops.union_aggregate(_tangent, rel_emb, g, radius=_radius, gamma=self.radius_msg_gamma)
        P = ops.gemm(agg, self.weight_neighbor, b_key=(self.weight_neighbor, "w"))
        lc = self._loop_cat() if self.self_loop else None
        L = ops.gemm(_tangent, lc, b_key=(lc, "w")) if self.self_loop else None
        S = sb = prev_t = None
        if self.skip_connect and prev_h is not None:
            prev_t, _, _ = ops.hyp_tangent(prev_h, self.c, want_clamped=False, want_radius=False)
            S, sb = ops.gemm(prev_t, self.skip_weight, b_key=(self.skip_weight, "w")), self.skip_bias
        out, ht_next, rad_next = ops.union_combine(P, L, g.indeg, act=_act_code(self.activation), hyper=True, c=self.c,
                                                   skip=S, skip_bias=sb, prev=prev_t, want_tangent=_want_next,
                                                   want_radius=_want_next)
        if _want_next:
            return out, ht_next, rad_next
        return out


class LorentzRGCNLayer(nn.Module, _LoopMixin):
    """hyperbolic_layers.py:524-694."""

    def __init__(self, in_feat, out_feat, num_rels, num_bases=-1, c=0.01, activation=None, self_loop=False,
                 dropout=0.0, skip_connect=False):
        super().__init__()
        self.in_feat, self.out_feat, self.num_rels = in_feat, out_feat, num_rels
        self.num_bases = num_bases if num_bases > 0 else num_rels
        if self.num_bases > self.num_rels:
            self.num_bases = self.num_rels
        self.c = c
        self.activation = activation
        self.self_loop = self_loop
        self.skip_connect = skip_connect
        self.submat_in = in_feat // self.num_bases
        self.submat_out = out_feat // self.num_bases
        self.weight = nn.Parameter(torch.Tensor(self.num_rels, self.num_bases * self.submat_in * self.submat_out))
        nn.init.xavier_uniform_(self.weight, gain=_RELU_GAIN)
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_RELU_GAIN)
            self.evolve_loop_weight = nn.Parameter(torch.Tensor(in_feat, out_feat))
            nn.init.xavier_uniform_(self.evolve_loop_weight, gain=_RELU_GAIN)
        if self.skip_connect:
            self.skip_weight = nn.Parameter(torch.Tensor(out_feat, out_feat))
            nn.init.xavier_uniform_(self.skip_weight, gain=_RELU_GAIN)
            self.skip_bias = nn.Parameter(torch.zeros(out_feat))
        self.dropout = nn.Dropout(dropout) if dropout > 0 else None
        self.rel_emb = None

    @torch.no_grad()
    def forward(self, g, h_hyper, rel_emb=None, prev_h=None, _tangent=None, _want_next=False):
        _no_train_dropout(self)
        if self.in_feat != self.out_feat or self.in_feat % self.num_bases:
            raise NotImplementedError("LorentzRGCNLayer: needs in_feat == out_feat divisible by num_bases")
        self.rel_emb = rel_emb
        if rel_emb is not None and rel_emb.shape[-1] != self.out_feat:
            rel_emb = rel_emb[:, :self.out_feat].contiguous()
        if _tangent is None:
            _tangent, _, _ = ops.hyp_tangent(h_hyper, self.c, want_clamped=False, want_radius=False)
        agg = ops.lorentz_aggregate(_tangent, self.weight, rel_emb, g, self.num_bases, self.c)
        lc = self._loop_cat() if self.self_loop else None
        L = ops.gemm(_tangent, lc, b_key=(lc, "w")) if self.self_loop else None
        S = sb = prev_t = None
        if self.skip_connect and prev_h is not None:
            prev_t, _, _ = ops.hyp_tangent(prev_h, self.c, want_clamped=False, want_radius=False)
            S, sb = ops.gemm(prev_t, self.skip_weight, b_key=(self.skip_weight, "w")), self.skip_bias
        out, ht_next, _ = ops.union_combine(agg, L, g.indeg, act=_act_code(self.activation), hyper=True, c=self.c,
                                            skip=S, skip_bias=sb, prev=prev_t, want_tangent=_want_next)
        if _want_next:
            return out, ht_next
        return out


class LorentzRGCNCell(nn.Module):
    """hyperbolic_layers.py:697-743."""

    def __init__(self, num_nodes, h_dim, out_dim, num_rels, num_bases=-1, num_hidden_layers=1, dropout=0.0, c=0.01,
                 self_loop=False, skip_connect=False, encoder_name="lgcn", rel_emb=None, use_cuda=False,
                 analysis=False):
        super().__init__()
        self.h_dim = h_dim
        self.c = c
        self.layers = nn.ModuleList()
        for idx in range(num_hidden_layers):
            sc = False if idx == 0 or not skip_connect else True
            self.layers.append(LorentzRGCNLayer(h_dim, h_dim, num_rels, num_bases, c=c, activation=F.rrelu,
                                                self_loop=self_loop, dropout=dropout, skip_connect=sc))

    @torch.no_grad()
    def forward(self, g, init_ent_emb, init_rel_emb, _tangent=None, _radius=None):
        h = init_ent_emb  # ndata['id'] is arange(N): identity gather
        rel_embs = init_rel_emb if isinstance(init_rel_emb, list) else [init_rel_emb] * len(self.layers)
        prev_h, ht = None, _tangent
        for i, layer in enumerate(self.layers):
            last = i == len(self.layers) - 1
            res = layer(g, h, rel_embs[i], prev_h=prev_h, _tangent=ht, _want_next=not last)
            prev_h = h
            if last:
                h = res
            else:
                h, ht = res
        return h


class HyperbolicBaseRGCN(nn.Module):
    """hyperbolic_model.py:67-111."""

    def __init__(self, num_nodes, h_dim, out_dim, num_rels, num_bases=-1, num_hidden_layers=1, dropout=0, c=0.01,
                 self_loop=False, skip_connect=False, encoder_name="", rel_emb=None, use_cuda=False, analysis=False,
                 radius_msg_gamma=1.0):
        super().__init__()
        self.num_nodes, self.h_dim, self.out_dim, self.num_rels = num_nodes, h_dim, out_dim, num_rels
        self.num_bases = num_bases
        self.num_hidden_layers = num_hidden_layers
        self.dropout = dropout
        self.c = c
        self.skip_connect = skip_connect
        self.self_loop = self_loop
        self.encoder_name = encoder_name
        self.use_cuda = use_cuda
        self.run_analysis = analysis
        self.rel_emb = rel_emb       # registers `rgcn.rel_emb` like the reference
        self.radius_msg_gamma = radius_msg_gamma
        self.build_model()

    def build_model(self):
        self.layers = nn.ModuleList()
        for idx in range(self.num_hidden_layers):
            self.layers.append(self.build_hidden_layer(idx))

    def build_hidden_layer(self, idx):
        raise NotImplementedError


class HyperbolicRGCNCell(HyperbolicBaseRGCN):
    """hyperbolic_model.py:114-154."""

    def build_hidden_layer(self, idx):
        sc = False if idx == 0 or not self.skip_connect else True
        return HyperbolicUnionRGCNLayer(self.h_dim, self.h_dim, self.num_rels, self.num_bases, c=self.c,
                                        activation=F.rrelu, self_loop=self.self_loop, dropout=self.dropout,
                                        skip_connect=sc, radius_msg_gamma=self.radius_msg_gamma)

    @torch.no_grad()
    def forward(self, g, init_ent_emb, init_rel_emb, _tangent=None, _radius=None):
        h = init_ent_emb
        rel_embs = init_rel_emb if isinstance(init_rel_emb, list) else [init_rel_emb] * len(self.layers)
        ht, rad = _tangent, _radius
        for i, layer in enumerate(self.layers):
            last = i == len(self.layers) - 1
            res = layer(g, h, rel_embs[i], _tangent=ht, _radius=rad, _want_next=not last)
            if last:
                h = res
            else:
                h, ht, rad = res
        return h
