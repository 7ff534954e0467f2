"""HyperbolicRecurrentRGCN with the reference's constructor / forward / predict signatures and state-dict names
(hyperbolic_src/hyperbolic_model.py:157-1088), evolving snapshots on the sm_100a kernels.

Per history snapshot (hyperbolic_model.py:797-888):
  tangent prep (log_0 h, |h|)  -> relation mean-pool (K2) -> relation GRU (K3)
  -> 2 x {HyperbolicUnionRGCNLayer | LorentzRGCNLayer}  (K4/K7 aggregate + node GEMMs + K5 combine with exp_0)
  -> tangent-space time gate + projection + residual radius evolution (K9 + K8, one fused row kernel).
The five unconditional `.item()` host syncs of the reference's TemporalRadiusEvolution (hyperbolic_ops.py:426-434)
do not exist here.  Not in this round: EST add-ons, FHNN/HGAT encoders, AttH decoders, geoopt ManifoldParameter,
learnable curvature, the training losses (all default-off or SURVEY.md 8f "next").
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .hyperbolic_decoder import (HyperbolicAttH, HyperbolicAttHRel, HyperbolicConvTransE, HyperbolicConvTransR,
                                 HyperbolicMuRP, HyperbolicMuRPRel, HyperbolicRotH, HyperbolicRotHRel)
from .hyperbolic_layers import HyperbolicRGCNCell, LorentzRGCNCell
from .layers import RGCNBlockLayer


class TemporalRadiusEvolution(nn.Module):
    """Parameter holder for hyperbolic_ops.py:364-435 (the arithmetic is fused into the time-gate kernel)."""

    def __init__(self, dim, c=0.01, epsilon=0.1, anchor_beta=1.0):
        super().__init__()
        if anchor_beta < 0.0 or anchor_beta > 1.0:
            raise ValueError("anchor_beta must be in [0, 1]")
        self.dim, self.c, self.epsilon, self.anchor_beta = dim, c, epsilon, float(anchor_beta)
        self.radius_mlp = nn.Linear(dim, 1)
        nn.init.xavier_uniform_(self.radius_mlp.weight, gain=0.1)
        nn.init.zeros_(self.radius_mlp.bias)
        self.last_evolution_stats = None

    def get_evolution_stats(self):
        return self.last_evolution_stats


class HyperbolicRecurrentRGCN(nn.Module):
    def __init__(self, decoder_name, encoder_name, num_ents, num_rels, num_static_rels, num_words, h_dim, opn,
                 sequence_len, num_bases=-1, num_hidden_layers=1, dropout=0, c=0.01, self_loop=False,
                 skip_connect=False, layer_norm=False, input_dropout=0, hidden_dropout=0, feat_dropout=0, weight=1,
                 discount=0, angle=0, use_static=False, entity_prediction=False, relation_prediction=False,
                 use_cuda=False, gpu=0, analysis=False, learn_curvature=False, use_residual_evolution=True,
                 radius_target=None, radius_lambda=0.02, radius_min=0.5, radius_max=3.0, radius_epsilon=0.1,
                 radius_anchor_beta=1.0, curvature_min=1e-4, curvature_max=1e-1, num_heads=4, query_chunk_size=128,
                 candidate_chunk_size=256, hyp_init_scale=1e-3, hyp_score_scale_init=1.0, hyp_score_margin_init=1.0,
                 use_entity_euclidean_bias=False, use_relation_specific_curvature=False, use_est=False,
                 est_state_alpha=0.2, est_encoder="gru", use_time_aware_negative=False, radius_msg_gamma=1.0):
        super().__init__()
        if learn_curvature or use_est or use_time_aware_negative:
            raise NotImplementedError("learnable curvature / EST add-ons are out of this round's scope (default-off flags)")
        self.decoder_name, self.encoder_name = decoder_name, encoder_name
        self.num_rels, self.num_ents = num_rels, num_ents
        self.opn = opn
        self.num_words, self.num_static_rels = num_words, num_static_rels
        self.sequence_len = sequence_len
        self.h_dim = h_dim
        self.layer_norm = layer_norm
        self.h = None
        self.run_analysis = analysis
        self.weight, self.discount, self.use_static, self.angle = weight, discount, use_static, angle
        self.relation_prediction, self.entity_prediction = relation_prediction, entity_prediction
        self.gpu = gpu
        self.learn_curvature = learn_curvature
        self.use_residual_evolution = use_residual_evolution
        self.radius_lambda = radius_lambda
        self.radius_min, self.radius_max = radius_min, radius_max
        self.radius_anchor_beta = radius_anchor_beta
        self.curvature_min, self.curvature_max = curvature_min, curvature_max
        self.num_heads = num_heads
        self.query_chunk_size, self.candidate_chunk_size = query_chunk_size, candidate_chunk_size
        self.use_entity_euclidean_bias = use_entity_euclidean_bias
        self.use_relation_specific_curvature = use_relation_specific_curvature
        self.radius_msg_gamma = radius_msg_gamma
        self.use_est = use_est
        self._c_float = float(c)

        self.register_buffer('c', torch.tensor(c))
        self.dynamic_emb = nn.Parameter(torch.Tensor(num_ents, h_dim))
        nn.init.normal_(self.dynamic_emb, std=1.0)
        self.emb_rel = nn.Parameter(torch.Tensor(num_rels * 2, h_dim))
        nn.init.xavier_normal_(self.emb_rel)
        self.temporal_radius_evolution = TemporalRadiusEvolution(h_dim, c=c, epsilon=radius_epsilon,
                                                                 anchor_beta=radius_anchor_beta)
        self.w1 = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_normal_(self.w1)
        self.w2 = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_normal_(self.w2)
        if self.use_static:
            self.words_emb = nn.Parameter(torch.Tensor(num_words, h_dim))
            nn.init.xavier_normal_(self.words_emb)
            self.static_rgcn_layer = RGCNBlockLayer(h_dim, h_dim, num_static_rels * 2, num_bases, activation=F.rrelu,
                                                    dropout=dropout, self_loop=False, skip_connect=False)
            self.static_loss = nn.MSELoss()
        self.loss_r = nn.CrossEntropyLoss()
        self.loss_e = nn.CrossEntropyLoss()

        if encoder_name == "hyperbolic_uvrgcn":
            self.rgcn = HyperbolicRGCNCell(num_ents, h_dim, h_dim, num_rels * 2, num_bases, num_hidden_layers, dropout,
                                           c=c, self_loop=self_loop, skip_connect=skip_connect,
                                           encoder_name=encoder_name, rel_emb=self.emb_rel, use_cuda=use_cuda,
                                           analysis=analysis, radius_msg_gamma=radius_msg_gamma)
        elif encoder_name == "lgcn":
            self.rgcn = LorentzRGCNCell(num_ents, h_dim, h_dim, num_rels * 2, num_bases, num_hidden_layers, dropout,
                                        c=c, self_loop=self_loop, skip_connect=skip_connect, encoder_name=encoder_name,
                                        rel_emb=self.emb_rel, use_cuda=use_cuda, analysis=analysis)
        else:
            raise NotImplementedError(f"Encoder '{encoder_name}' not implemented here (hyperbolic_uvrgcn, lgcn)")

        self.time_gate_weight = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_uniform_(self.time_gate_weight, gain=nn.init.calculate_gain('relu'))
        self.time_gate_bias = nn.Parameter(torch.zeros(h_dim))
        self.relation_gru = nn.GRUCell(h_dim * 2, h_dim)

        dist_kw = dict(c=c, dropout=input_dropout, query_chunk_size=query_chunk_size,
                       candidate_chunk_size=candidate_chunk_size)
        ent_kw = dict(init_scale=hyp_init_scale, score_scale_init=hyp_score_scale_init,
                      score_margin_init=hyp_score_margin_init, use_entity_euclidean_bias=use_entity_euclidean_bias,
                      use_relation_specific_curvature=use_relation_specific_curvature)
        if decoder_name == "hyperbolic_convtranse":
            self.decoder_ob = HyperbolicConvTransE(num_ents, h_dim, c=c, input_dropout=input_dropout,
                                                   hidden_dropout=hidden_dropout, feature_map_dropout=feat_dropout)
            self.rdecoder = HyperbolicConvTransR(num_rels, h_dim, c=c, input_dropout=input_dropout,
                                                 hidden_dropout=hidden_dropout, feature_map_dropout=feat_dropout)
        elif decoder_name == "murp":
            self.decoder_ob = HyperbolicMuRP(num_ents, num_rels * 2, h_dim, **dist_kw, **ent_kw)
            self.rdecoder = HyperbolicMuRPRel(num_rels, h_dim, **dist_kw)
        elif decoder_name == "roth":
            self.decoder_ob = HyperbolicRotH(num_ents, num_rels * 2, h_dim, **dist_kw, **ent_kw)
            self.rdecoder = HyperbolicRotHRel(num_rels, h_dim, **dist_kw, init_scale=hyp_init_scale,
                                              score_scale_init=hyp_score_scale_init,
                                              score_margin_init=hyp_score_margin_init)
        elif decoder_name == "atth":
            self.decoder_ob = HyperbolicAttH(num_ents, num_rels * 2, h_dim, **dist_kw, **ent_kw)
            self.rdecoder = HyperbolicAttHRel(num_rels, h_dim, **dist_kw, init_scale=hyp_init_scale,
                                              score_scale_init=hyp_score_scale_init,
                                              score_margin_init=hyp_score_margin_init)
        else:
            raise NotImplementedError(f"Decoder '{decoder_name}' not implemented. Choose from: hyperbolic_convtranse, "
                                      "murp, roth, atth")

        if radius_target is None:
            target = torch.full((num_ents,), 0.5 * (radius_min + radius_max))
        else:
            target = torch.as_tensor(radius_target, dtype=torch.float)
        self.register_buffer("radius_target", target)
        self.radius_static = nn.Parameter(self.radius_target.clone())

    def get_curvature(self):
        return self.c

    def _relation_step(self, g, ht, h0_prev):
        cell, d = self.relation_gru, self.h_dim
        x_mean = ops.rel_mean_pool(ht, g)
        w_ih = cell.weight_ih.detach()
        gi = ops.gemm(self.emb_rel.detach(), w_ih[:, :d], trans_b=True, bias=cell.bias_ih.detach(),
                      b_key=(cell.weight_ih, "left"))
        ops.gemm(x_mean, w_ih[:, d:], trans_b=True, out=gi, accumulate=True, b_key=(cell.weight_ih, "right"))
        gh = ops.gemm(h0_prev, cell.weight_hh.detach(), trans_b=True, bias=cell.bias_hh.detach(),
                      b_key=(cell.weight_hh, "w"))
        return ops.gru_gate(gi, gh, h0_prev, self.layer_norm)

    # ------------------------------------------------------------------ whole-recurrence fast path
    def _engine_ok(self):
        first = self.rgcn.layers[0]
        # --skip-connect is inert for hyperbolic_uvrgcn (its cell never passes prev_h, hyperbolic_src/hyperbolic_model.py:152)
        # and live from the second Lorentz layer on (hyperbolic_src/hyperbolic_layers.py:737-740): that one takes the
        # per-layer path
        live_skip = self.encoder_name == "lgcn" and any(l.skip_connect for l in self.rgcn.layers)
        return (ops.gemm_impl() == "tc" and first.self_loop and not live_skip
                and self.encoder_name in ("hyperbolic_uvrgcn", "lgcn") and self.h_dim % 4 == 0 and self.h_dim <= 256)

    def _engine_tables(self, h_init=None):
        """Pointer / int / double tables of regcn_hyp_evolve (include/regcn_b200.h HM_* / HMI_* / HMD_*).
        h_init: persistent buffer holding the initial tangent table instead of dynamic_emb (use_static)."""
        import numpy as np
        cell = self.relation_gru
        tre = self.temporal_radius_evolution
        params = [self.dynamic_emb if h_init is None else h_init, self.emb_rel, self.radius_static, cell.weight_ih, cell.weight_hh, cell.bias_ih,
                  cell.bias_hh, self.time_gate_weight, self.time_gate_bias, tre.radius_mlp.weight, tre.radius_mlp.bias]
        for layer in self.rgcn.layers:
            params += [getattr(layer, "weight_neighbor", None) if self.encoder_name != "lgcn" else layer.weight,
                       layer.loop_weight, layer.evolve_loop_weight]
        # the static buffer is rewritten in place before every call: its version is not part of the stamp
        stamp = tuple((p._version if (i or h_init is None) else -1, p.data_ptr()) for i, p in enumerate(params))
        if getattr(self, "_engine_stamp", None) == stamp:
            return self._engine_tab
        d = self.h_dim
        keep = []

        def split(m):
            hi, lo = ops.split_tf32(m.detach().contiguous())
            keep.extend((hi, lo))
            return hi, lo

        emb_rel = self.emb_rel.detach().contiguous()
        er_hi, er_lo = split(emb_rel)
        w_ih = cell.weight_ih.detach()
        gi_static = ops.gemm(emb_rel, w_ih[:, :d], trans_b=True, bias=cell.bias_ih.detach())
        wr_hi, wr_lo = split(w_ih[:, d:])
        wh_hi, wh_lo = split(cell.weight_hh)
        gw_hi, gw_lo = split(self.time_gate_weight.detach().t())
        b_hh = cell.bias_hh.detach().contiguous()
        gate_b = self.time_gate_bias.detach().contiguous()
        dyn = self.dynamic_emb.detach().contiguous() if h_init is None else h_init
        rs = self.radius_static.detach().contiguous()
        rw = tre.radius_mlp.weight.detach().view(-1).contiguous()
        keep += [emb_rel, gi_static, b_hh, gate_b, dyn, rs, rw]
        ptrs = [dyn, rs, emb_rel, er_hi, er_lo, gi_static, wr_hi, wr_lo, wh_hi, wh_lo, b_hh, gw_hi, gw_lo, gate_b, rw]
        lgcn = self.encoder_name == "lgcn"
        for layer in self.rgcn.layers:
            if lgcn:
                w = layer.weight.detach().contiguous()
                keep.append(w)
                first = [w, w]
            else:
                first = list(split(layer.weight_neighbor.detach().t()))
            wl_hi, wl_lo = split(torch.cat([layer.loop_weight.detach(), layer.evolve_loop_weight.detach()], dim=1).t())
            ptrs += first + [wl_hi, wl_lo]
        ptab = np.array([t.data_ptr() for t in ptrs], dtype=np.uint64)
        nb = self.rgcn.layers[0].num_bases if lgcn else 0
        itab = np.array([self.num_ents, 2 * self.num_rels, d, len(self.rgcn.layers), int(bool(self.layer_norm)), 1,
                         1 if lgcn else 0, nb, int(bool(self.use_residual_evolution))], dtype=np.int32)
        dtab = np.array([self._c_float, float(self.radius_msg_gamma), float(self.radius_min), float(self.radius_max),
                         float(tre.anchor_beta), float(tre.epsilon), float(tre.radius_mlp.bias.detach().item())],
                        dtype=np.float64)
        self._engine_tab = (ptab, itab, dtab, keep)
        self._engine_ptr_tensors = ptrs
        self._engine_stamp = stamp
        self._engine_tab_batch = {}
        return self._engine_tab

    def _engine_tables_batch(self, G):
        """Tables of the recurrence over G independent history windows at once (see RecurrentRGCN._engine_tables_batch):
        per-entity tables (tangent table, static radii) tiled G times, per-relation tables (relation embeddings, the
        Lorentz layers' relation blocks) tiled in the numbering of graph.concat_graphs, dense weights shared."""
        ptab, itab, dtab, _ = self._engine_tables()
        hit = self._engine_tab_batch.get(G)
        if hit is not None:
            return hit
        R = self.num_rels
        t = self._engine_ptr_tensors

        def tile_rel(x):
            return torch.cat([x[:R]] * G + [x[R:]] * G).contiguous()

        heads = {0: t[0].repeat(G, 1).contiguous(), 1: t[1].repeat(G, *([1] * (t[1].dim() - 1))).contiguous()}
        for i in (2, 3, 4, 5):
            heads[i] = tile_rel(t[i])
        if self.encoder_name == "lgcn":
            for l in range(len(self.rgcn.layers)):
                w = tile_rel(t[15 + 4 * l])
                heads[15 + 4 * l] = heads[16 + 4 * l] = w
        ptab_g, itab_g = ptab.copy(), itab.copy()
        for i, x in heads.items():
            ptab_g[i] = x.data_ptr()
        itab_g[0], itab_g[1] = G * self.num_ents, 2 * G * R
        self._engine_tab_batch[G] = (ptab_g, itab_g, dtab, heads)
        return self._engine_tab_batch[G]

    def _forward_engine(self, g_list, h_init=None, members=1):
        import numpy as np
        from . import _lib
        ptab, itab, dtab, _ = self._engine_tables(h_init) if members == 1 else self._engine_tables_batch(members)
        L = len(g_list)
        N, R2, d = members * self.num_ents, 2 * members * self.num_rels, self.h_dim
        dev = self.dynamic_emb.device
        gp = np.concatenate([g.ptr_table for g in g_list])
        gi = np.concatenate([g.int_table for g in g_list])
        max_split = max(g.n_split_chunks for g in g_list)
        rel_nsplit = max(max(1, min(64, g.n_rel_ents // (max(1, R2 // 2) * 512))) for g in g_list)
        need = _lib.load().regcn_hyp_evolve_workspace_bytes(N, R2, d, max_split, rel_nsplit)
        ws = getattr(self, "_engine_ws", None)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(need, device=dev, dtype=torch.uint8)
            self._engine_ws = ws
        hist = torch.empty((L, N, d), device=dev, dtype=torch.float32)
        h0 = torch.empty((R2, d), device=dev, dtype=torch.float32)
        _lib.call("regcn_hyp_evolve", ptab.ctypes.data, itab.ctypes.data, dtab.ctypes.data, gp.ctypes.data,
                  gi.ctypes.data, L, hist.data_ptr(), h0.data_ptr(), rel_nsplit, ws.data_ptr(), ws.numel())
        return [hist[i] for i in range(L)], h0

    def batch_ok(self):
        """True when forward_batch can evolve several history windows at once (whole-recurrence engine, no static graph)."""
        return self._engine_ok() and not self.use_static

    @torch.no_grad()
    def forward_batch(self, windows):
        """G independent history windows (lists of L SnapshotGraphs, same L) as ONE recurrence over the block-diagonal union
        graphs -- RecurrentRGCN.forward_batch for the hyperbolic model (hyperbolic_main.py:100-113 evaluates the test
        timestamps one after the other, each over its own window).  Returns [(h_g (N,d) on the ball, r_emb_g (2R,d))]."""
        from .graph import concat_graphs
        G = len(windows)
        L = len(windows[0])
        if not self.batch_ok() or L == 0 or any(len(w) != L for w in windows):
            raise RuntimeError("forward_batch: needs the recurrence engine, no static graph and windows of one length > 0")
        if G == 1:
            embs, _, r_emb, _, _ = self.forward(windows[0], None, True)
            return [(embs[-1], r_emb)]
        pool = self.__dict__.setdefault("_batch_graphs", {})
        key = (G, torch.cuda.current_stream().cuda_stream)
        old = pool.get(key, [])
        comb = [concat_graphs([w[i] for w in windows], old[i] if i < len(old) else None) for i in range(L)]
        pool[key] = comb
        hist, h0 = self._forward_engine(comb, None, members=G)
        N, R, d = self.num_ents, self.num_rels, self.h_dim
        rel = h0.view(2, G, R, d).transpose(0, 1).contiguous().view(G, 2 * R, d)
        return [(hist[-1][g * N:(g + 1) * N], rel[g]) for g in range(G)]

    def forward(self, g_list, static_graph, use_cuda):
        """hyperbolic_model.py:722-890.  In train() mode with autograd enabled the outputs carry gradients (the
        kernel-backed autograd nodes of regcn_b200.train_hyp, the ones get_loss() trains through); otherwise the
        one-call inference engine."""
        if self.training and torch.is_grad_enabled():
            from . import train, train_hyp
            train.begin_step()
            with torch.enable_grad():
                hist, h0, static_emb = train_hyp.hyp_evolve(self, g_list, static_graph)
            self.h_0 = h0
            self.h = hist[-1] if hist else None
            return hist, static_emb, h0, [], []
        return self._forward_eval(g_list, static_graph, use_cuda)

    @torch.no_grad()
    def _forward_eval(self, g_list, static_graph, use_cuda):
        gate_list, degree_list = [], []
        static_emb = None
        if self.use_static and static_graph is not None:
            static_graph = static_graph.to(self.gpu)
            static_graph.ndata['h'] = torch.cat((self.dynamic_emb, self.words_emb), dim=0).detach()
            self.static_rgcn_layer(static_graph, [])
            static_emb = static_graph.ndata.pop('h')[:self.num_ents, :].contiguous()
            static_emb = ops.row_map(static_emb, ops.ROW_NORMALIZE) if self.layer_norm else static_emb
        if self._engine_ok() and len(g_list) > 0:
            h_init = None
            if static_emb is not None:
                # fixed-pointer initial table, as in RecurrentRGCN.forward: the engine's row normalisation of the already
                # normalised static embedding is idempotent (hyperbolic_src/hyperbolic_model.py:769-771)
                buf = getattr(self, "_static_h", None)
                if buf is None or buf.shape != static_emb.shape or buf.device != static_emb.device:
                    buf = self._static_h = torch.empty_like(static_emb)
                buf.copy_(static_emb)
                h_init = buf
            history_embs, self.h_0 = self._forward_engine(g_list, h_init)
            self.h = history_embs[-1]
            return history_embs, static_emb, self.h_0, gate_list, degree_list
        c = self._c_float
        rs_raw = self.radius_static.detach()
        if static_emb is not None:
            self.h = ops.hyp_init(static_emb, rs_raw, False, False, c, self.radius_min, self.radius_max)
        else:
            self.h = ops.hyp_init(self.dynamic_emb.detach(), rs_raw, self.layer_norm, False, c, self.radius_min,
                                  self.radius_max)

        tre = self.temporal_radius_evolution
        rw = tre.radius_mlp.weight.detach().view(-1).contiguous()
        bias = tre.radius_mlp.bias
        if getattr(self, "_rb_cache", (None, None))[0] != bias._version:    # one D2H sync per optimiser step, not per call
            self._rb_cache = (bias._version, float(bias.detach().item()))
        rb = self._rb_cache[1]
        history_embs = []
        for i, g in enumerate(g_list):
            g = g.to(self.gpu)
            ht, pt, radius = ops.hyp_tangent(self.h, c, want_clamped=True, want_radius=True)
            h0_prev = self.emb_rel.detach() if i == 0 else self.h_0
            self.h_0 = self._relation_step(g, ht, h0_prev)
            current_h = self.rgcn.forward(g, self.h, [self.h_0, self.h_0], _tangent=ht, _radius=radius)
            G = ops.gemm(pt, self.time_gate_weight.detach(), b_key=(self.time_gate_weight, "w"))
            self.h = ops.hyp_time_gate(current_h, pt, G, self.time_gate_bias.detach(), rs_raw, rw, rb, self.layer_norm,
                                       self.use_residual_evolution, c, self.radius_min, self.radius_max,
                                       tre.anchor_beta, tre.epsilon)
            history_embs.append(self.h)
        return history_embs, static_emb, self.h_0, gate_list, degree_list

    @torch.no_grad()
    def predict(self, test_graph, num_rels, static_graph, test_triplets, use_cuda):
        inverse_test_triplets = test_triplets.flip(1)
        inverse_test_triplets[:, 1] = inverse_test_triplets[:, 1] + num_rels
        all_triples = torch.cat((test_triplets, inverse_test_triplets)).contiguous()
        evolve_embs, _, r_emb, _, _ = self.forward(test_graph, static_graph, use_cuda)
        embedding = evolve_embs[-1]
        if self.layer_norm:
            embedding = ops.row_map(embedding, ops.ROW_TANGENT_NORMALIZE, c=self._c_float)
        score = self.decoder_ob.forward(embedding, r_emb, all_triples, mode="test")
        score_rel = self.rdecoder.forward(embedding, r_emb, all_triples, mode="test")
        return all_triples, score, score_rel

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._c_float = float(self.c.item())
        if hasattr(self, "_rb_cache"):
            del self._rb_cache
        return out

    def get_loss(self, glist, triples, static_graph, use_cuda, query_time=None):
        """hyperbolic_model.py:941-1088: (loss_ent, loss_rel, loss_static, loss_radius).

        Training mode (`model.train()`): losses with gradients for the hyperbolic_uvrgcn encoder + hyperbolic_convtranse
        decoder (regcn_b200/train_hyp.py: kernels as autograd nodes); other encoders / decoders raise there.
        Evaluation mode: forward values.  Entity head: the decoders' streaming CE (hyperbolic_decoder.py:182-307) as the
        scoring GEMM's log-sum-exp epilogue; relation head on its (B,2R) score matrix; radius supervision :1066-1073."""
        if self.training:
            from . import train_hyp
            with torch.enable_grad():
                return train_hyp.hyp_get_loss(self, glist, triples, static_graph)
        with torch.no_grad():
            return self._get_loss_eval(glist, triples, static_graph, use_cuda)

    def _get_loss_eval(self, glist, triples, static_graph, use_cuda):
        from . import evaluate
        dev = self.dynamic_emb.device
        triples = torch.as_tensor(triples).to(dev)
        inverse_triples = triples.flip(1)
        inverse_triples[:, 1] = inverse_triples[:, 1] + self.num_rels
        all_triples = torch.cat([triples, inverse_triples]).contiguous()
        evolve_embs, static_emb, r_emb, _, _ = self.forward(glist, static_graph, use_cuda)
        pre_emb = evolve_embs[-1]
        if self.layer_norm:
            pre_emb = ops.row_map(pre_emb, ops.ROW_TANGENT_NORMALIZE, c=self._c_float)
        loss_ent = torch.zeros(1, device=dev)
        loss_rel = torch.zeros(1, device=dev)
        loss_static = torch.zeros(1, device=dev)
        if getattr(self, "use_static", False) and static_emb is not None and self.discount in (0, 1):
            # hyperbolic_src/hyperbolic_model.py:1039-1064: angle loss against the tangent vectors log_0(evolve_emb)
            from . import train
            tangent = [ops.row_map(e, ops.ROW_LOG0, c=self._c_float) for e in evolve_embs]
            loss_static, _ = train.static_angle_terms(static_emb.contiguous(), tangent, self.layer_norm, self.angle,
                                                      self.discount, self.weight)
        if self.entity_prediction:
            q, cand, hyp, col_bias = evaluate._scoring_operands(self, pre_emb, r_emb, all_triples)
            _, loss_ent = ops.fused_ce(q, cand, all_triples[:, 2], hyp=hyp, col_bias=col_bias)
        if self.relation_prediction:
            score_rel = self.rdecoder.forward(pre_emb, r_emb, all_triples, mode="train")
            _, loss_rel = ops.ce_dense(score_rel, all_triples, 1)
        # radius supervision on the entities of the batch (tiny: a few thousand scalars)
        ids = torch.unique(all_triples[:, [0, 2]].reshape(-1))
        rs = torch.clamp(self.radius_static.detach(), min=self.radius_min, max=self.radius_max)
        rs = torch.clamp(rs, max=1.0 / (self._c_float ** 0.5) - 1e-6).index_select(0, ids)
        rt = self.radius_target.index_select(0, ids)
        loss_radius = (self.radius_lambda * torch.mean((rs - rt) ** 2)).reshape(1)
        return loss_ent, loss_rel, loss_static, loss_radius
