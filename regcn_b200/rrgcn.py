"""RecurrentRGCN with the reference's constructor / forward / predict / get_loss signatures
(src/rrgcn.py:14-248) and state-dict names, evolving snapshots on the sm_100a kernels.

Per history snapshot (src/rrgcn.py:159-179): relation mean-pool (K2) -> relation GRU (K3) ->
2 x UnionRGCNLayer (K4 aggregate + node GEMMs + K5 combine) -> time gate (K9).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .decoder import ConvTransE, ConvTransR
from .layers import RGCNBlockLayer, UnionRGCNLayer


class RGCNCell(nn.Module):
    """The entity encoder `rgcn` of RecurrentRGCN (src/rrgcn.py:14-54): a ModuleList `layers` of UnionRGCNLayer and the
    `rel_emb` handle (the reference registers emb_rel a second time as `rgcn.rel_emb`; state dicts carry both names).
    Constructor arguments as the reference passes them (src/rrgcn.py:108-121)."""

    def __init__(self, num_nodes, h_dim, out_dim, num_rels, num_bases=-1, num_basis=-1, num_hidden_layers=1, dropout=0,
                 self_loop=False, skip_connect=False, encoder_name="", opn="sub", rel_emb=None, use_cuda=False,
                 analysis=False):
        super().__init__()
        if encoder_name != "uvrgcn":
            raise NotImplementedError(f"encoder {encoder_name!r}: RecurrentRGCN evolves with 'uvrgcn' (src/rrgcn.py:17-27)")
        self.num_nodes, self.h_dim, self.out_dim, self.num_rels = num_nodes, h_dim, out_dim, num_rels
        self.num_bases, self.num_basis, self.num_hidden_layers = num_bases, num_basis, num_hidden_layers
        self.dropout, self.self_loop, self.skip_connect = dropout, self_loop, skip_connect
        self.encoder_name, self.opn, self.use_cuda, self.run_analysis = encoder_name, opn, use_cuda, analysis
        self.rel_emb = rel_emb
        self.layers = nn.ModuleList(
            UnionRGCNLayer(h_dim, h_dim, num_rels, num_bases, activation=F.rrelu, dropout=dropout, self_loop=self_loop,
                           skip_connect=bool(skip_connect and idx != 0), rel_emb=rel_emb)
            for idx in range(num_hidden_layers))

    def forward(self, g, init_ent_emb, init_rel_emb):
        # ndata['id'] is arange(N) (rgcn/utils.py:122), so init_ent_emb[node_id] is the identity gather; the cell never
        # passes prev_h (src/rrgcn.py:37-38), so a configured skip connection stays inert exactly like the reference's
        g.ndata['h'] = init_ent_emb
        for i, layer in enumerate(self.layers):
            layer(g, [], init_rel_emb[i])
        return g.ndata.pop('h')


class RecurrentRGCN(nn.Module):
    def __init__(self, decoder_name, encoder_name, num_ents, num_rels, num_static_rels, num_words, h_dim, opn,
                 sequence_len, num_bases=-1, num_basis=-1, num_hidden_layers=1, dropout=0, self_loop=False,
                 skip_connect=False, layer_norm=False, input_dropout=0, hidden_dropout=0, feat_dropout=0,
                 aggregation='cat', weight=1, discount=0, angle=0, use_static=False, entity_prediction=False,
                 relation_prediction=False, use_cuda=False, gpu=0, analysis=False):
        super().__init__()
        self.decoder_name = decoder_name
        self.encoder_name = encoder_name
        self.num_rels = num_rels
        self.num_ents = num_ents
        self.opn = opn
        self.num_words = num_words
        self.num_static_rels = num_static_rels
        self.sequence_len = sequence_len
        self.h_dim = h_dim
        self.layer_norm = layer_norm
        self.h = None
        self.run_analysis = analysis
        self.aggregation = aggregation
        self.relation_evolve = False
        self.weight = weight
        self.discount = discount
        self.use_static = use_static
        self.angle = angle
        self.relation_prediction = relation_prediction
        self.entity_prediction = entity_prediction
        self.emb_rel = None
        self.gpu = gpu

        self.w1 = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_normal_(self.w1)
        self.w2 = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_normal_(self.w2)
        self.emb_rel = nn.Parameter(torch.Tensor(num_rels * 2, h_dim))
        nn.init.xavier_normal_(self.emb_rel)
        self.dynamic_emb = nn.Parameter(torch.Tensor(num_ents, h_dim))
        nn.init.normal_(self.dynamic_emb)

        if self.use_static:
            self.words_emb = nn.Parameter(torch.Tensor(num_words, h_dim))
            nn.init.xavier_normal_(self.words_emb)
            self.statci_rgcn_layer = RGCNBlockLayer(h_dim, h_dim, num_static_rels * 2, num_bases, activation=F.rrelu,
                                                    dropout=dropout, self_loop=False, skip_connect=False)
            self.static_loss = nn.MSELoss()

        self.loss_r = nn.CrossEntropyLoss()
        self.loss_e = nn.CrossEntropyLoss()

        self.rgcn = RGCNCell(num_ents, h_dim, h_dim, num_rels * 2, num_bases, num_basis, num_hidden_layers, dropout,
                             self_loop, skip_connect, encoder_name, self.opn, self.emb_rel, use_cuda, analysis)

        self.time_gate_weight = nn.Parameter(torch.Tensor(h_dim, h_dim))
        nn.init.xavier_uniform_(self.time_gate_weight, gain=nn.init.calculate_gain('relu'))
        self.time_gate_bias = nn.Parameter(torch.Tensor(h_dim))
        nn.init.zeros_(self.time_gate_bias)

        self.relation_cell_1 = nn.GRUCell(h_dim * 2, h_dim)

        if decoder_name == "convtranse":
            self.decoder_ob = ConvTransE(num_ents, h_dim, input_dropout, hidden_dropout, feat_dropout)
            self.rdecoder = ConvTransR(num_rels, h_dim, input_dropout, hidden_dropout, feat_dropout)
        else:
            raise NotImplementedError

    # ------------------------------------------------------------------ relation GRU (K2 + K3)
    def _relation_step(self, g, h, h0_prev):
        cell = self.relation_cell_1
        d = self.h_dim
        x_mean = ops.rel_mean_pool(h, g)                                       # (2R, d)
        w_ih = cell.weight_ih.detach()                                         # (3d, 2d): [emb_rel | x_mean] halves
        gi = ops.gemm(self.emb_rel.detach(), w_ih[:, :d], trans_b=True, bias=cell.bias_ih.detach(),
                      b_key=(cell.weight_ih, "left"))
        ops.gemm(x_mean, w_ih[:, d:], trans_b=True, out=gi, accumulate=True, b_key=(cell.weight_ih, "right"))
        gh = ops.gemm(h0_prev, cell.weight_hh.detach(), trans_b=True, bias=cell.bias_hh.detach(),
                      b_key=(cell.weight_hh, "w"))
        return ops.gru_gate(gi, gh, h0_prev, self.layer_norm)

    # ------------------------------------------------------------------ whole-recurrence fast path
    def _engine_ok(self):
        # skip_connect is not a condition: the uvrgcn cell passes prev_h=[] to every layer (src/rrgcn.py:37-38)
        return (ops.gemm_impl() == "tc" and self.rgcn.self_loop
                and self.encoder_name == "uvrgcn" and self.h_dim % 4 == 0 and self.h_dim <= 256)

    def _engine_tables(self, h_init=None):
        """Pointer / int tables of regcn_regcn_evolve (include/regcn_b200.h RM_* / RMI_*), rebuilt when any parameter
        changes.  GEMM weights are stored K-major and TF32-split once here, never inside the recurrence.
        h_init: persistent buffer holding the initial entity table instead of dynamic_emb (use_static)."""
        cell = self.relation_cell_1
        params = [self.dynamic_emb if h_init is None else h_init, self.emb_rel, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh,
                  self.time_gate_weight, self.time_gate_bias]
        for layer in self.rgcn.layers:
            params += [layer.weight_neighbor, layer.loop_weight, layer.evolve_loop_weight]
        # (the static buffer is rewritten before every call: only its address identifies it)
        stamp = tuple((p._version if (i or h_init is None) else -1, p.data_ptr()) for i, p in enumerate(params))
        if getattr(self, "_engine_stamp", None) == stamp:
            return self._engine_tab
        import numpy as np
        d = self.h_dim
        keep = []

        def split(m):
            hi, lo = ops.split_tf32(m.detach().contiguous())
            keep.extend((hi, lo))
            return hi, lo

        emb_rel = self.emb_rel.detach().contiguous()
        er_hi, er_lo = split(emb_rel)
        w_ih = cell.weight_ih.detach()
        prev = ops.gemm_impl()
        gi_static = ops.gemm(emb_rel, w_ih[:, :d], trans_b=True, bias=cell.bias_ih.detach())
        wr_hi, wr_lo = split(w_ih[:, d:])
        wh_hi, wh_lo = split(cell.weight_hh)
        b_hh = cell.bias_hh.detach().contiguous()
        gate_b = self.time_gate_bias.detach().contiguous()
        dyn = self.dynamic_emb.detach().contiguous() if h_init is None else h_init
        keep += [emb_rel, gi_static, b_hh, gate_b, dyn]
        ptrs = [dyn, emb_rel, er_hi, er_lo, gi_static, wr_hi, wr_lo, wh_hi, wh_lo, b_hh, gate_b]
        for l, layer in enumerate(self.rgcn.layers):
            wn, wl, we = (layer.weight_neighbor.detach(), layer.loop_weight.detach(),
                          layer.evolve_loop_weight.detach())
            gate = [self.time_gate_weight.detach()] if l == 0 else []
            wn_hi, wn_lo = split(wn.t())                                        # dense path: agg . W_n
            wl_hi, wl_lo = split(torch.cat([wl, we] + gate, dim=1).t())         #             x . [W_loop|W_evolve(|W_t)]
            wc_hi, wc_lo = split(torch.cat([wn, wl], dim=0).t())                # sparse path: [agg|x] . [W_n;W_loop]
            we_hi, we_lo = split(torch.cat([we] + gate, dim=1).t())             #              x . [W_evolve(|W_t)]
            ptrs += [wn_hi, wn_lo, wl_hi, wl_lo, wc_hi, wc_lo, we_hi, we_lo]
        ptab = np.array([t.data_ptr() for t in ptrs], dtype=np.uint64)
        itab = np.array([self.num_ents, 2 * self.num_rels, d, len(self.rgcn.layers), int(bool(self.layer_norm)), 1],
                        dtype=np.int32)
        self._engine_tab = (ptab, itab, keep)
        self._engine_heads = ptrs[:5]           # the tables with one row per entity / relation (see _engine_tables_batch)
        self._engine_stamp = stamp
        self._engine_tab_batch = {}
        return self._engine_tab

    def _engine_tables_batch(self, G, tile_entities=True):
        """Tables of the recurrence over G independent history windows at once: the per-entity / per-relation tables are
        tiled G times in the numbering of graph.concat_graphs (entity g*N + v; relation g*R + r, inverse G*R + g*R + r),
        the weights are shared.  tile_entities=False (the shared-trajectory engine reads the first N rows only): the
        entity table is passed as it is instead of G copies of it."""
        ptab, itab, _ = self._engine_tables()
        hit = self._engine_tab_batch.get((G, tile_entities))
        if hit is not None:
            return hit
        R = self.num_rels
        dyn, emb_rel, er_hi, er_lo, gi_static = self._engine_heads

        def tile_rel(t):
            return torch.cat([t[:R]] * G + [t[R:]] * G).contiguous()

        other = self._engine_tab_batch.get((G, not tile_entities))
        rel_heads = other[2][1:] if other is not None else [tile_rel(emb_rel), tile_rel(er_hi), tile_rel(er_lo), tile_rel(gi_static)]
        heads = [dyn.repeat(G, 1).contiguous() if tile_entities else dyn] + list(rel_heads)
        ptab_g, itab_g = ptab.copy(), itab.copy()
        for i, t in enumerate(heads):
            ptab_g[i] = t.data_ptr()
        itab_g[0], itab_g[1] = G * self.num_ents, 2 * G * R
        self._engine_tab_batch[(G, tile_entities)] = (ptab_g, itab_g, heads)
        return self._engine_tab_batch[(G, tile_entities)]

    def _forward_engine(self, g_list, h_init=None, members=1):
        import numpy as np
        from . import _lib
        ptab, itab, _ = self._engine_tables(h_init) if members == 1 else self._engine_tables_batch(members)
        L = len(g_list)
        N, R2, d = members * self.num_ents, 2 * members * self.num_rels, self.h_dim
        dev = self.dynamic_emb.device
        gp = np.concatenate([g.ptr_table for g in g_list]) if L else np.zeros(1, dtype=np.uint64)
        gi = np.concatenate([g.int_table for g in g_list]) if L else np.zeros(1, dtype=np.int32)
        max_split = max([g.n_split_chunks for g in g_list], default=0)
        rel_nsplit = max([max(1, min(64, g.n_rel_ents // (max(1, R2 // 2) * 512))) for g in g_list], default=1)
        need = _lib.load().regcn_regcn_evolve_workspace_bytes(N, R2, d, max_split, rel_nsplit)
        ws = getattr(self, "_engine_ws", None)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(need, device=dev, dtype=torch.uint8)
            self._engine_ws = ws
        hist = torch.empty((max(L, 1), N, d), device=dev, dtype=torch.float32)
        h0 = torch.empty((R2, d), device=dev, dtype=torch.float32)
        _lib.call("regcn_regcn_evolve", ptab.ctypes.data, itab.ctypes.data, gp.ctypes.data, gi.ctypes.data, L,
                  hist.data_ptr(), h0.data_ptr(), rel_nsplit, ws.data_ptr(), ws.numel())
        return [hist[i] for i in range(L)], h0

    def _forward_engine_shared(self, g_list, members):
        """regcn_regcn_evolve_shared: the recurrence over `members` windows with the entity state kept compact (one shared
        row per entity until it is first active in its window).  Returns (h_final (G N, d), h0) or None when the
        preconditions do not hold (a dense snapshot, one layer) -- the caller then runs the plain engine."""
        import numpy as np
        from . import _lib
        G = members
        n_act = [int(g.n_active) for g in g_list]
        N, R2, d = G * self.num_ents, 2 * G * self.num_rels, self.h_dim
        if len(self.rgcn.layers) < 2 or any(2 * a > N for a in n_act) or os.environ.get("REGCN_SHARED_ROWS", "1") == "0":
            return None
        ptab, itab, _ = self._engine_tables_batch(G, tile_entities=False)
        dev = self.dynamic_emb.device
        gp = np.concatenate([g.ptr_table for g in g_list])
        gi = np.concatenate([g.int_table for g in g_list])
        max_split = max([g.n_split_chunks for g in g_list], default=0)
        rel_nsplit = max([max(1, min(64, g.n_rel_ents // (max(1, R2 // 2) * 512))) for g in g_list], default=1)
        need = _lib.load().regcn_regcn_evolve_shared_workspace_bytes(self.num_ents, G, R2, d, max_split, rel_nsplit,
                                                                     sum(n_act), max(n_act))
        ws = getattr(self, "_engine_ws_shared", None)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(int(need * 1.25), device=dev, dtype=torch.uint8)
            self._engine_ws_shared = ws
        h_final = torch.empty((N, d), device=dev, dtype=torch.float32)
        h0 = torch.empty((R2, d), device=dev, dtype=torch.float32)
        _lib.call("regcn_regcn_evolve_shared", ptab.ctypes.data, itab.ctypes.data, gp.ctypes.data, gi.ctypes.data,
                  len(g_list), G, h_final.data_ptr(), h0.data_ptr(), rel_nsplit, ws.data_ptr(), ws.numel())
        return h_final, h0

    def batch_ok(self):
        """True when forward_batch can evolve several history windows at once (the whole-recurrence engine, no static
        graph: its initial table is rebuilt per call)."""
        return self._engine_ok() and not self.use_static

    @torch.no_grad()
    def forward_batch(self, windows):
        """Evolve G independent history windows (lists of L SnapshotGraphs each, same L) in ONE recurrence over the
        block-diagonal union graphs (graph.concat_graphs).  The reference evaluates test timestamps one after the other,
        each over its own window (src/main.py:60-90); the windows do not depend on each other, so this runs the same
        kernels at G times the rows per launch.  Returns [(h_g (N,d), r_emb_g (2R,d))] -- row for row what
        forward(windows[g]) returns as (history_embs[-1], h_0)."""
        from .graph import concat_graphs
        G = len(windows)
        L = len(windows[0])
        if not self.batch_ok() or L == 0 or any(len(w) != L for w in windows):
            raise RuntimeError("forward_batch: needs the recurrence engine, no static graph and windows of one length > 0")
        if G == 1:
            embs, _, r_emb, _, _ = self.forward(windows[0], None, True)
            return [(embs[-1], r_emb)]
        # the union graphs of the previous call with this G are overwritten in place: every kernel that read them was
        # enqueued on this stream before (the engine joins its side streams before it returns)
        pool = self.__dict__.setdefault("_batch_graphs", {})
        key = (G, torch.cuda.current_stream().cuda_stream)
        old = pool.get(key, [])
        comb = [concat_graphs([w[i] for w in windows], old[i] if i < len(old) else None) for i in range(L)]
        pool[key] = comb
        shared = self._forward_engine_shared(comb, G)
        if shared is not None:
            h_last, h0 = shared
        else:
            hist, h0 = self._forward_engine(comb, None, members=G)
            h_last = hist[-1]
        N, R, d = self.num_ents, self.num_rels, self.h_dim
        rel = h0.view(2, G, R, d).transpose(0, 1).contiguous().view(G, 2 * R, d)
        return [(h_last[g * N:(g + 1) * N], rel[g]) for g in range(G)]

    def forward(self, g_list, static_graph, use_cuda):
        """src/rrgcn.py:142-180: (history_embs, static_emb, h_0, gate_list, degree_list).  In train() mode with autograd
        enabled the outputs carry gradients (layer dropout active), as the reference's do: the recurrence runs on the
        kernel-backed autograd nodes of regcn_b200.train (the same ones get_loss() trains through).  Otherwise: the
        one-call inference engine."""
        if self.training and torch.is_grad_enabled():
            from . import train
            train.begin_step()
            with torch.enable_grad():
                hist, h0, static_emb = train.regcn_evolve(self, g_list, static_graph)
            self.h_0 = h0
            self.h = hist[-1] if hist else None
            return hist, static_emb, h0, [], []
        return self._forward_eval(g_list, static_graph, use_cuda)

    @torch.no_grad()
    def _forward_eval(self, g_list, static_graph, use_cuda):
        gate_list, degree_list = [], []
        static_emb = None
        if self.use_static:
            static_graph = static_graph.to(self.gpu)
            static_graph.ndata['h'] = torch.cat((self.dynamic_emb, self.words_emb), dim=0).detach()
            self.statci_rgcn_layer(static_graph, [])
            static_emb = static_graph.ndata.pop('h')[:self.num_ents, :].contiguous()
            static_emb = ops.row_map(static_emb, ops.ROW_NORMALIZE) if self.layer_norm else static_emb
        if self._engine_ok() and len(g_list) > 0:
            h_init = None
            if static_emb is not None:
                # the engine reads its initial table through a fixed pointer: keep the static embedding in a persistent
                # buffer (its row normalisation inside the engine is then idempotent, src/rrgcn.py:150-152)
                buf = getattr(self, "_static_h", None)
                if buf is None or buf.shape != static_emb.shape or buf.device != static_emb.device:
                    buf = self._static_h = torch.empty_like(static_emb)
                buf.copy_(static_emb)
                h_init = buf
            history_embs, self.h_0 = self._forward_engine(g_list, h_init)
            self.h = history_embs[-1]
            return history_embs, static_emb, self.h_0, gate_list, degree_list
        if self.use_static:
            self.h = static_emb
        else:
            dyn = self.dynamic_emb.detach()
            self.h = ops.row_map(dyn, ops.ROW_NORMALIZE) if self.layer_norm else dyn
            static_emb = None

        history_embs = []
        for i, g in enumerate(g_list):
            g = g.to(self.gpu)
            h0_prev = self.emb_rel.detach() if i == 0 else self.h_0
            self.h_0 = self._relation_step(g, self.h, h0_prev)
            current_h = self.rgcn.forward(g, self.h, [self.h_0, self.h_0])
            G = ops.gemm(self.h, self.time_gate_weight.detach(), b_key=(self.time_gate_weight, "w"))
            self.h = ops.time_gate(G, self.time_gate_bias.detach(), current_h, self.h, self.layer_norm)
            history_embs.append(self.h)
        return history_embs, static_emb, self.h_0, gate_list, degree_list

    @torch.no_grad()
    def predict(self, test_graph, num_rels, static_graph, test_triplets, use_cuda):
        inverse_test_triplets = test_triplets.flip(1)
        inverse_test_triplets[:, 1] = inverse_test_triplets[:, 1] + num_rels
        all_triples = torch.cat((test_triplets, inverse_test_triplets)).contiguous()

        evolve_embs, _, r_emb, _, _ = self.forward(test_graph, static_graph, use_cuda)
        embedding = ops.row_map(evolve_embs[-1], ops.ROW_NORMALIZE) if self.layer_norm else evolve_embs[-1]

        score = self.decoder_ob.forward(embedding, r_emb, all_triples, mode="test")
        score_rel = self.rdecoder.forward(embedding, r_emb, all_triples, mode="test")
        return all_triples, score, score_rel

    def get_loss(self, glist, triples, static_graph, use_cuda):
        """src/rrgcn.py:197-248: (loss_ent, loss_rel, loss_static), each of shape (1,).

        Training mode (`model.train()`): the losses carry gradients -- the evolution, the train-mode ConvTransE/R
        towers (batch-statistics BatchNorm, dropout) and the cross entropies run as autograd nodes whose forward and
        backward are kernels (regcn_b200/train.py); `loss.backward()` then `regcn_b200.optim.Adam.step()` is the
        reference's optimisation step (src/main.py:235-246).

        Evaluation mode: forward values only.  The entity head is CrossEntropy over all entities computed by the
        scoring GEMM's streaming log-sum-exp epilogue (no (B,N) logits); the relation head materialises its (B,2R)
        scores (dropout off, BatchNorm running statistics)."""
        from . import evaluate
        if not use_cuda:
            raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
        if self.training:
            from . import train
            with torch.enable_grad():
                return train.regcn_get_loss(self, glist, triples, static_graph)
        with torch.no_grad():
            dev = self.dynamic_emb.device
            triples = torch.as_tensor(triples).to(dev)
            inverse_triples = triples.flip(1)
            inverse_triples[:, 1] = inverse_triples[:, 1] + self.num_rels
            all_triples = torch.cat([triples, inverse_triples]).contiguous()
            evolve_embs, static_emb, r_emb, _, _ = self.forward(glist, static_graph, use_cuda)
            pre_emb = ops.row_map(evolve_embs[-1], ops.ROW_NORMALIZE) if self.layer_norm else evolve_embs[-1]
            loss_ent = torch.zeros(1, device=dev)
            loss_rel = torch.zeros(1, device=dev)
            loss_static = torch.zeros(1, device=dev)
            if self.entity_prediction:
                q, cand, hyp, col_bias = evaluate._scoring_operands(self, pre_emb, r_emb, all_triples)
                _, loss_ent = ops.fused_ce(q, cand, all_triples[:, 2], hyp=hyp, col_bias=col_bias)
            if self.relation_prediction:
                score_rel = self.rdecoder.forward(pre_emb, r_emb, all_triples, mode="train")
                _, loss_rel = ops.ce_dense(score_rel, all_triples, 1)
            if self.use_static and self.discount in (0, 1):
                from . import train
                loss_static, _ = train.static_angle_terms(static_emb, list(evolve_embs), self.layer_norm, self.angle,
                                                          self.discount, self.weight)
            return loss_ent, loss_rel, loss_static
