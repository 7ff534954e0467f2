"""Hyperbolic decoders with the reference's signatures and parameter names on the sm_100a kernels:
HyperbolicConvTransE / ConvTransR (hyperbolic_src/hyperbolic_decoder.py:310-510), HyperbolicMuRP (:647-817),
HyperbolicRotH (:931-1138) and HyperbolicRotHRel (:1141-1280).

The all-entity hyperbolic score (:89-179, proxy-distance branch) is evaluated in norm/dot form: one dense
contraction <q_b, e_n> followed by an epilogue that needs only |q_b|^2 and |e_n|^2, instead of
the reference's B*N*d Mobius-add temporaries.
Entity Euclidean bias (candidate bias + entity_bias[subject], :1097-1098) and relation-specific curvature (per-query
c_q, true-distance artanh branch, :145-163) are served by the same kernels (`col_bias` / `qbias` / `row_c`).
Not implemented in this round (raise): the streaming-CE `loss` heads (training), AttH.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import ops
from ._lib import call as _call, ptr as _ptr
from .decoder import _fold_bn

SCORE_SCALE_EPSILON = 1e-6
REL_CURVATURE_EPSILON = 1e-5
REL_CURVATURE_INIT_RATIO = 0.95


def _relation_curvature_theta_init(global_c):
    """hyperbolic_decoder.py:38-63: softplus(theta) = 0.95 * c."""
    target = max(float(global_c) * REL_CURVATURE_INIT_RATIO, REL_CURVATURE_EPSILON)
    return math.log(max(math.expm1(target), 1e-12))


class _HypConvBase(nn.Module):
    def __init__(self, n_bias, embedding_dim, c, input_dropout, hidden_dropout, feature_map_dropout, channels,
                 kernel_size):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.c = c
        self.inp_drop = nn.Dropout(input_dropout)
        self.hidden_drop = nn.Dropout(hidden_dropout)
        self.feature_map_drop = nn.Dropout(feature_map_dropout)
        self.conv1 = nn.Conv1d(2, channels, kernel_size, stride=1, padding=int(math.floor(kernel_size / 2)))
        self.bn0 = nn.BatchNorm1d(2)
        self.bn1 = nn.BatchNorm1d(channels)
        self.bn2 = nn.BatchNorm1d(embedding_dim)
        self.fc = nn.Linear(embedding_dim * channels, embedding_dim)
        self.register_parameter('b', Parameter(torch.zeros(n_bias)))

    def _tower(self, ent_act, second, triplets, col0, col1, always_bn2):
        if self.training:
            raise NotImplementedError("regcn_b200 decoders: the standalone forward() is the inference path (folded BatchNorm, no dropout); batch-statistics BatchNorm, dropout and gradients run through the model's get_loss() (regcn_b200/train.py, train_hyp.py)")
        B = len(triplets)
        if ops.gemm_impl() == "tc" and ops.convtrans_fc_ok(ent_act.shape[1], self.conv1.weight, self.fc.out_features):
            # one GEMM whose A operand (the conv feature map) is computed on chip; bn2 + relu in its split-K reduction
            return ops.convtrans_fc(ent_act, second, triplets, col0, col1, _fold_bn(self.bn0), self.conv1.weight.detach(),
                                    self.conv1.bias.detach(), _fold_bn(self.bn1), self.fc.weight, self.fc.bias.detach(),
                                    bn2=_fold_bn(self.bn2) if (always_bn2 or B > 1) else None, relu=True)
        feats = ops.convtranse_features(ent_act, second, triplets, col0, col1, _fold_bn(self.bn0),
                                        self.conv1.weight.detach(), self.conv1.bias.detach(), _fold_bn(self.bn1),
                                        split=False)
        split_k = max(1, min(16, (148 * 2) // max(1, ((B + 127) // 128) * ((self.fc.out_features + 127) // 128))))
        # the feature map stays one fp32 matrix (B, 50 d): the FC GEMM splits it to TF32 on chip
        x = ops.gemm(feats, self.fc.weight.detach(), trans_b=True, bias=self.fc.bias.detach(), split_k=split_k,
                     b_key=(self.fc.weight, "w"), split_a_on_chip=True)
        if always_bn2 or B > 1:
            s, t = _fold_bn(self.bn2)
            ops.affine_relu_(x, s, t, relu=True)
        else:
            ops.affine_relu_(x, None, None, relu=True)
        return x


def _hyp_conv_forward_train(self, entity_embedding, rel_embedding, triplets, relation_head):
    """train() mode of the two ConvTrans heads with gradients (kernel-backed autograd nodes of regcn_b200.train /
    train_hyp); the scores are materialised because the caller asked for them."""
    from . import train as T, train_hyp as TH
    with torch.enable_grad():
        t = torch.as_tensor(triplets).to(entity_embedding.device).contiguous()
        et = TH.eltwise(TH.radial(entity_embedding, TH.LOG0, float(self.c)), 1, 0.0)      # 0.9 tanh(log_0 E) + 0.1 log_0 E
        if relation_head:
            return T.linear(T.conv_tower(self, et, et, t, 0, 2), rel_embedding, self.b)
        return T.linear(T.conv_tower(self, et, rel_embedding, t, 0, 1), et, self.b)


_HypConvBase._forward_train = _hyp_conv_forward_train


class HyperbolicConvTransE(_HypConvBase):
    """hyperbolic_decoder.py:310-413."""

    def __init__(self, num_entities, embedding_dim, c=0.01, input_dropout=0.0, hidden_dropout=0.0,
                 feature_map_dropout=0.0, channels=50, kernel_size=3):
        super().__init__(num_entities, embedding_dim, c, input_dropout, hidden_dropout, feature_map_dropout, channels,
                         kernel_size)
        self.num_entities = num_entities

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        if self.training:
            return self._forward_train(entity_embedding, rel_embedding, triplets, False)
        et = ops.row_map(entity_embedding, ops.ROW_LEAKY_TANH_LOG0, c=self.c)
        q = self._tower(et, rel_embedding.contiguous(), triplets, 0, 1, always_bn2=False)
        return ops.gemm(q, et, trans_b=True, bias=self.b.detach())


class HyperbolicConvTransR(_HypConvBase):
    """hyperbolic_decoder.py:416-510."""

    def __init__(self, num_relations, embedding_dim, c=0.01, input_dropout=0.0, hidden_dropout=0.0,
                 feature_map_dropout=0.0, channels=50, kernel_size=3):
        super().__init__(num_relations * 2, embedding_dim, c, input_dropout, hidden_dropout, feature_map_dropout,
                         channels, kernel_size)
        self.num_relations = num_relations

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        if self.training:
            return self._forward_train(entity_embedding, rel_embedding, triplets, True)
        et = ops.row_map(entity_embedding, ops.ROW_LEAKY_TANH_LOG0, c=self.c)
        q = self._tower(et, et, triplets, 0, 2, always_bn2=True)
        return ops.gemm(q, rel_embedding.contiguous(), trans_b=True, bias=self.b.detach())


class _HypDistBase(nn.Module):
    """Shared scoring tail of the distance decoders: scale*(margin - |(-q)(+)e|^2) + bias  (:164-172)."""

    def _train_loss(self, builder, entity_embedding, rel_embedding, triplets):
        """The decoder's own training head: regcn_b200.train_hyp builds it from kernel-backed autograd nodes."""
        from . import train_hyp
        with torch.enable_grad():
            t = torch.as_tensor(triplets).to(entity_embedding.device).contiguous()
            return getattr(train_hyp, builder)(self, entity_embedding, rel_embedding, t, float(self.c),
                                               self.training).reshape(())

    def _score_scale(self):
        return F.softplus(self.score_scale_raw) + SCORE_SCALE_EPSILON

    def _scale_margin(self):
        return torch.stack((self._score_scale().detach(), self.score_margin.detach())).float().contiguous()

    def _unsupported_flags(self):
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("regcn_b200 decoders: the standalone forward() is the inference path (folded BatchNorm, no dropout); batch-statistics BatchNorm, dropout and gradients run through the model's get_loss() (regcn_b200/train.py, train_hyp.py)")

    def _relation_curvature(self, triplets):
        """(B,) per-query curvature or None (hyperbolic_decoder.py:1020-1026); inverse relations share c_r."""
        if getattr(self, "rel_curvature_raw", None) is None:
            return None
        return ops.rel_curvature(self.rel_curvature_raw, triplets, self.num_relations, self.c, self.rel_curvature_max)

    def set_relation_curvature_bounds(self, curvature_max=None):
        if curvature_max is not None:
            self.rel_curvature_max = float(curvature_max)

    def _entity_bias(self, triplets):
        """(candidate bias (N,), per-query bias entity_bias[subject] (B,)) or (None, None)."""
        eb = getattr(self, "entity_bias", None)
        if eb is None:
            return None, None
        eb = eb.detach().contiguous()
        return eb, eb[triplets[:, 0]].contiguous()

    def hyp_operands(self, q_sumsq, cand, triplets):
        """The `hyp` tuple of ops.fused_rank_counts for this decoder's score."""
        rc = self._relation_curvature(triplets)
        base = (self.c, q_sumsq, ops.row_sumsq(cand), self._scale_margin())
        return base if rc is None else base + (rc,)

    def _dist_scores(self, query, q_sumsq, cand, bias, triplets=None):
        e_sumsq = ops.row_sumsq(cand)
        S = ops.gemm(query, cand, trans_b=True)                                  # <q_b, e_n>
        qbias = row_c = None
        if triplets is not None:
            cb, qbias = self._entity_bias(triplets)
            bias = cb if cb is not None else bias
            row_c = self._relation_curvature(triplets)
        return ops.hyp_score_epilogue_(S, q_sumsq, e_sumsq, bias, qbias, self.c, self._scale_margin(), row_c)


class HyperbolicRotH(_HypDistBase):
    """hyperbolic_decoder.py:931-1138."""

    def __init__(self, num_entities, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256, init_scale=1e-3, score_scale_init=1.0, score_margin_init=1.0,
                 use_entity_euclidean_bias=False, use_relation_specific_curvature=False):
        super().__init__()
        assert embedding_dim % 2 == 0, "embedding_dim must be even (required for Givens rotation)"
        self.num_entities = num_entities
        self.embedding_dim = embedding_dim
        self.half_dim = embedding_dim // 2
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.num_relations = num_relations
        self.use_entity_euclidean_bias = use_entity_euclidean_bias
        self.use_relation_specific_curvature = use_relation_specific_curvature
        self.rot_proj = nn.Linear(embedding_dim, self.half_dim)
        self.trans_proj = nn.Linear(embedding_dim, embedding_dim)
        self.reshape_fc1 = nn.Linear(embedding_dim, embedding_dim)
        self.reshape_fc2 = nn.Linear(embedding_dim, embedding_dim)
        for lin in (self.rot_proj, self.trans_proj, self.reshape_fc1, self.reshape_fc2):
            nn.init.uniform_(lin.weight, -init_scale, init_scale)
            nn.init.zeros_(lin.bias)
        if use_entity_euclidean_bias:
            self.entity_bias = nn.Parameter(torch.zeros(num_entities))
        else:
            self.register_parameter("entity_bias", None)
        if use_relation_specific_curvature:
            self.rel_curvature_raw = nn.Parameter(torch.full((num_relations,), _relation_curvature_theta_init(c)))
        else:
            self.register_parameter("rel_curvature_raw", None)
        self.rel_curvature_max = float(c) if use_relation_specific_curvature else None
        self.score_scale_raw = nn.Parameter(torch.tensor(float(score_scale_init)))
        self.score_margin = nn.Parameter(torch.tensor(float(score_margin_init)))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def query(self, entity_embedding, rel_embedding, triplets):
        """K12: (B,d) query points and their squared norms."""
        self._unsupported_flags()
        s_tan = ops.gather_log0(entity_embedding, triplets, 0, True, self.c)                     # :1066-1070
        h1 = ops.gemm(s_tan, self.reshape_fc1.weight.detach(), trans_b=True, bias=self.reshape_fc1.bias.detach(), b_key=(self.reshape_fc1.weight, "w"))
        ops.affine_relu_(h1, None, None, relu=True)
        ops.gemm(h1, self.reshape_fc2.weight.detach(), trans_b=True, bias=self.reshape_fc2.bias.detach(), b_key=(self.reshape_fc2.weight, "w"),
                 out=s_tan, accumulate=True)                                                      # x + fc2(relu(fc1 x))
        rel = rel_embedding.contiguous()
        ang = ops.gemm(rel, self.rot_proj.weight.detach(), trans_b=True, bias=self.rot_proj.bias.detach(), b_key=(self.rot_proj.weight, "w"))      # (2R, d/2)
        trans = ops.gemm(rel, self.trans_proj.weight.detach(), trans_b=True, bias=self.trans_proj.bias.detach(), b_key=(self.trans_proj.weight, "w"))  # (2R, d)
        return ops.hyp_query(s_tan, ang, trans, None, triplets, 0, self.c)

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        q, qss = self.query(entity_embedding, rel_embedding, triplets)
        return self._dist_scores(q, qss, entity_embedding.contiguous(), None, triplets)

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:1101-1138: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("roth_ent_loss", entity_embedding, rel_embedding, triplets)


class HyperbolicMuRP(_HypDistBase):
    """hyperbolic_decoder.py:647-817."""

    def __init__(self, num_entities, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256, init_scale=1e-3, score_scale_init=1.0, score_margin_init=1.0,
                 use_entity_euclidean_bias=False, use_relation_specific_curvature=False):
        super().__init__()
        self.num_entities = num_entities
        self.embedding_dim = embedding_dim
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.num_relations = num_relations
        self.use_entity_euclidean_bias = use_entity_euclidean_bias
        self.use_relation_specific_curvature = use_relation_specific_curvature
        self.rot_proj = nn.Linear(embedding_dim, embedding_dim)
        self.trans_proj = nn.Linear(embedding_dim, embedding_dim)
        for lin in (self.rot_proj, self.trans_proj):
            nn.init.uniform_(lin.weight, -init_scale, init_scale)
            nn.init.zeros_(lin.bias)
        if use_entity_euclidean_bias:
            self.entity_bias = nn.Parameter(torch.zeros(num_entities))
        else:
            self.register_parameter("entity_bias", None)
        if use_relation_specific_curvature:
            self.rel_curvature_raw = nn.Parameter(torch.full((num_relations,), _relation_curvature_theta_init(c)))
        else:
            self.register_parameter("rel_curvature_raw", None)
        self.rel_curvature_max = float(c) if use_relation_specific_curvature else None
        self.score_scale_raw = nn.Parameter(torch.tensor(float(score_scale_init)))
        self.score_margin = nn.Parameter(torch.tensor(float(score_margin_init)))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def query(self, entity_embedding, rel_embedding, triplets):
        self._unsupported_flags()
        s_tan = ops.gather_log0(entity_embedding, triplets, 0, True, self.c)
        rel = rel_embedding.contiguous()
        diag = ops.gemm(rel, self.rot_proj.weight.detach(), trans_b=True, bias=self.rot_proj.bias.detach(), b_key=(self.rot_proj.weight, "w"))
        trans = ops.gemm(rel, self.trans_proj.weight.detach(), trans_b=True, bias=self.trans_proj.bias.detach(), b_key=(self.trans_proj.weight, "w"))
        return ops.hyp_query(s_tan, diag, trans, None, triplets, 1, self.c)

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        q, qss = self.query(entity_embedding, rel_embedding, triplets)
        return self._dist_scores(q, qss, entity_embedding.contiguous(), None, triplets)

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:781-817: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("murp_ent_loss", entity_embedding, rel_embedding, triplets)


class HyperbolicRotHRel(_HypDistBase):
    """hyperbolic_decoder.py:1141-1280."""

    def __init__(self, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256, init_scale=1e-3, score_scale_init=1.0, score_margin_init=1.0):
        super().__init__()
        assert embedding_dim % 2 == 0, "embedding_dim must be even (required for Givens rotation)"
        self.num_relations = num_relations
        self.embedding_dim = embedding_dim
        self.half_dim = embedding_dim // 2
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.global_rot = nn.Parameter(torch.Tensor(self.half_dim))
        nn.init.uniform_(self.global_rot, -math.pi, math.pi)
        self.reshape_fc1 = nn.Linear(embedding_dim, embedding_dim)
        self.reshape_fc2 = nn.Linear(embedding_dim, embedding_dim)
        for lin in (self.reshape_fc1, self.reshape_fc2):
            nn.init.uniform_(lin.weight, -init_scale, init_scale)
            nn.init.zeros_(lin.bias)
        self.rel_bias = nn.Parameter(torch.zeros(num_relations * 2))
        self.score_scale_raw = nn.Parameter(torch.tensor(float(score_scale_init)))
        self.score_margin = nn.Parameter(torch.tensor(float(score_margin_init)))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        self._unsupported_flags()
        E = entity_embedding.contiguous()
        s_tan = ops.gather_log0(E, triplets, 0, False, self.c)
        h1 = ops.gemm(s_tan, self.reshape_fc1.weight.detach(), trans_b=True, bias=self.reshape_fc1.bias.detach(), b_key=(self.reshape_fc1.weight, "w"))
        ops.affine_relu_(h1, None, None, relu=True)
        ops.gemm(h1, self.reshape_fc2.weight.detach(), trans_b=True, bias=self.reshape_fc2.bias.detach(), b_key=(self.reshape_fc2.weight, "w"), out=s_tan,
                 accumulate=True)
        q, qss = ops.hyp_query(s_tan, self.global_rot.detach(), None, E, triplets, 2, self.c)
        rel_hyp, rss = ops.row_map(rel_embedding.contiguous(), ops.ROW_EXP0, c=self.c, want_sumsq=True)
        S = ops.gemm(q, rel_hyp, trans_b=True)
        return ops.hyp_score_epilogue_(S, qss, rss, self.rel_bias.detach(), None, self.c, self._scale_margin())

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:1249-1280: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("roth_rel_loss", entity_embedding, rel_embedding, triplets)


class HyperbolicMuRPRel(_HypDistBase):
    """hyperbolic_decoder.py:820-928: query = exp_0(log_0(h_s) W_s + log_0(h_o) W_o), scored against exp_0(rel)
    with no scale / margin (score = -|(-q)(+)r|^2 + rel_bias)."""

    def __init__(self, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256):
        super().__init__()
        self.num_relations = num_relations
        self.embedding_dim = embedding_dim
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.W_s = nn.Parameter(torch.Tensor(embedding_dim, embedding_dim))
        nn.init.xavier_uniform_(self.W_s)
        self.W_o = nn.Parameter(torch.Tensor(embedding_dim, embedding_dim))
        nn.init.xavier_uniform_(self.W_o)
        self.rel_bias = nn.Parameter(torch.zeros(num_relations * 2))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        self._unsupported_flags()
        E = entity_embedding.contiguous()
        s_tan = ops.gather_log0(E, triplets, 0, False, self.c)
        o_tan = ops.gather_log0(E, triplets, 2, False, self.c)
        q_tan = ops.gemm(s_tan, self.W_s.detach(), b_key=(self.W_s, "w"))
        ops.gemm(o_tan, self.W_o.detach(), out=q_tan, accumulate=True, b_key=(self.W_o, "w"))
        q, qss = ops.row_map(q_tan, ops.ROW_EXP0, c=self.c, want_sumsq=True)
        rel_hyp, rss = ops.row_map(rel_embedding.contiguous(), ops.ROW_EXP0, c=self.c, want_sumsq=True)
        S = ops.gemm(q, rel_hyp, trans_b=True)
        sm = torch.tensor([1.0, 0.0], device=S.device, dtype=torch.float32)
        return ops.hyp_score_epilogue_(S, qss, rss, self.rel_bias.detach(), None, self.c, sm)

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:897-928: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("murp_rel_loss", entity_embedding, rel_embedding, triplets)


class HyperbolicAttH(_HypDistBase):
    """hyperbolic_decoder.py:1283-1512: query = exp_0(a Rot_r(log_0 s) + (1-a) Ref_r(log_0 s)) (+)_c exp_0(trans_proj(rel_r)),
    a = sigmoid(<attn_proj(rel_r), [log_0 s ; rel_r]>); scored like RotH."""

    def __init__(self, num_entities, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256, init_scale=1e-3, score_scale_init=1.0, score_margin_init=1.0,
                 use_entity_euclidean_bias=False, use_relation_specific_curvature=False):
        super().__init__()
        assert embedding_dim % 2 == 0, "embedding_dim must be even"
        self.num_entities = num_entities
        self.embedding_dim = embedding_dim
        self.half_dim = embedding_dim // 2
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.num_relations = num_relations
        self.use_entity_euclidean_bias = use_entity_euclidean_bias
        self.use_relation_specific_curvature = use_relation_specific_curvature
        self.rot_proj = nn.Linear(embedding_dim, self.half_dim)
        self.ref_proj = nn.Linear(embedding_dim, self.half_dim)
        self.trans_proj = nn.Linear(embedding_dim, embedding_dim)
        self.attn_proj = nn.Linear(embedding_dim, 2 * embedding_dim)
        for lin in (self.rot_proj, self.ref_proj, self.trans_proj, self.attn_proj):
            nn.init.uniform_(lin.weight, -init_scale, init_scale)
            nn.init.zeros_(lin.bias)
        if use_entity_euclidean_bias:
            self.entity_bias = nn.Parameter(torch.zeros(num_entities))
        else:
            self.register_parameter("entity_bias", None)
        if use_relation_specific_curvature:
            self.rel_curvature_raw = nn.Parameter(torch.full((num_relations,), _relation_curvature_theta_init(c)))
        else:
            self.register_parameter("rel_curvature_raw", None)
        self.rel_curvature_max = float(c) if use_relation_specific_curvature else None
        self.score_scale_raw = nn.Parameter(torch.tensor(float(score_scale_init)))
        self.score_margin = nn.Parameter(torch.tensor(float(score_margin_init)))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def query(self, entity_embedding, rel_embedding, triplets):
        self._unsupported_flags()
        s_tan = ops.gather_log0(entity_embedding, triplets, 0, True, self.c)                     # :1417-1421
        rel = rel_embedding.contiguous()

        def table(lin):                                                                           # per-relation projections
            return ops.gemm(rel, lin.weight.detach(), trans_b=True, bias=lin.bias.detach(), b_key=(lin.weight, "w"))

        rot, ref, trans, attn = table(self.rot_proj), table(self.ref_proj), table(self.trans_proj), table(self.attn_proj)
        B, d = s_tan.shape
        Q = torch.empty((B, d), device=s_tan.device, dtype=torch.float32)
        qss = torch.empty(B, device=s_tan.device, dtype=torch.float32)
        _call("regcn_atth_query", _ptr(s_tan), _ptr(rot.contiguous()), _ptr(ref.contiguous()), _ptr(attn.contiguous()),
              _ptr(rel), _ptr(trans.contiguous()), None, _ptr(triplets.contiguous()), B, d, 0, float(self.c), _ptr(Q),
              _ptr(qss))
        return Q, qss

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        q, qss = self.query(entity_embedding, rel_embedding, triplets)
        return self._dist_scores(q, qss, entity_embedding.contiguous(), None, triplets)

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:1464-1512: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("atth_ent_loss", entity_embedding, rel_embedding, triplets)


class HyperbolicAttHRel(_HypDistBase):
    """hyperbolic_decoder.py:1515-1700: global rotation / reflection mixed by a(s, o), query = (-exp_0(mix)) (+)_c h_o,
    scored against exp_0(rel)."""

    def __init__(self, num_relations, embedding_dim, c=0.01, dropout=0.0, query_chunk_size=128,
                 candidate_chunk_size=256, init_scale=1e-3, score_scale_init=1.0, score_margin_init=1.0):
        super().__init__()
        assert embedding_dim % 2 == 0, "embedding_dim must be even"
        self.num_relations = num_relations
        self.embedding_dim = embedding_dim
        self.half_dim = embedding_dim // 2
        self.c = c
        self.query_chunk_size = query_chunk_size
        self.candidate_chunk_size = candidate_chunk_size
        self.global_rot = nn.Parameter(torch.Tensor(self.half_dim))
        nn.init.uniform_(self.global_rot, -math.pi, math.pi)
        self.global_ref = nn.Parameter(torch.Tensor(self.half_dim))
        nn.init.uniform_(self.global_ref, -math.pi, math.pi)
        self.attn_weight = nn.Parameter(torch.Tensor(2 * embedding_dim))
        nn.init.uniform_(self.attn_weight, -init_scale, init_scale)
        self.rel_bias = nn.Parameter(torch.zeros(num_relations * 2))
        self.score_scale_raw = nn.Parameter(torch.tensor(float(score_scale_init)))
        self.score_margin = nn.Parameter(torch.tensor(float(score_margin_init)))
        self.dropout = nn.Dropout(dropout)

    @torch.no_grad()
    def forward(self, entity_embedding, rel_embedding, triplets, mode="train"):
        self._unsupported_flags()
        E = entity_embedding.contiguous()
        s_tan = ops.gather_log0(E, triplets, 0, False, self.c)
        B, d = s_tan.shape
        q = torch.empty((B, d), device=E.device, dtype=torch.float32)
        qss = torch.empty(B, device=E.device, dtype=torch.float32)
        _call("regcn_atth_query", _ptr(s_tan), _ptr(self.global_rot.detach().contiguous()),
              _ptr(self.global_ref.detach().contiguous()), _ptr(self.attn_weight.detach().contiguous()), None, None,
              _ptr(E), _ptr(triplets.contiguous()), B, d, 1, float(self.c), _ptr(q), _ptr(qss))
        rel_hyp, rss = ops.row_map(rel_embedding.contiguous(), ops.ROW_EXP0, c=self.c, want_sumsq=True)
        S = ops.gemm(q, rel_hyp, trans_b=True)
        return ops.hyp_score_epilogue_(S, qss, rss, self.rel_bias.detach(), None, self.c, self._scale_margin())

    def loss(self, entity_embedding, rel_embedding, triplets):
        """hyperbolic_decoder.py:1641-1700: scalar cross entropy over all candidates, with gradients (no (B,N) logits)."""
        return self._train_loss("atth_rel_loss", entity_embedding, rel_embedding, triplets)
