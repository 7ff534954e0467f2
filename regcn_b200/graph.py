"""SnapshotGraph: the device-resident edge index of one history snapshot.

Replaces the DGL graph object produced by the reference's `build_sub_graph`
(rgcn/utils.py:100-134) and the python `r2e` loop (rgcn/utils.py:78-97).  The index is built by one
C-ABI call (`regcn_csr_build`) and kept in HBM as CSR-by-destination plus a relation->entity CSR; the
attributes the reference's modules read from `g` (`ndata['id'|'norm']`, `edata['type']`, `in_degrees`,
`number_of_nodes`, `uniq_r`, `r_to_e`, `r_len`, `to`) are provided as lazily materialised views so
code written against the reference keeps working.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

I32 = torch.int32
AGG_CHUNK = 32  # must match csrc/graph_build.cu kAggChunk


class _Frame(dict):
    pass


# arrays of one snapshot's index, in arena order (norm, float32, follows them)
_VIEWS = ("src", "dst", "etype", "indeg", "rowptr", "src_sorted", "etype_sorted", "eperm", "vptr", "sptr", "vrow_row",
          "active_pos", "rel_rowptr", "rel_ents", "_counts", "active_rows")


class SnapshotGraph:
    @staticmethod
    def _arena_sizes(N, R, E):
        return [E, E, E, N, N + 1, E, E, E, N + 1, N + 1, min(N, E) + E // AGG_CHUNK + 1, N, R + 1, E, 8, min(N, E)]

    @staticmethod
    def arena_ints(N, R, T):
        """int32 words of one snapshot's index arena (what __init__ allocates)."""
        return sum((max(int(n), 1) + 3) // 4 * 4 for n in SnapshotGraph._arena_sizes(int(N), int(R), 2 * int(T))) + int(N)

    def __init__(self, num_nodes, num_rels, triples_dev, _defer_counts=False, _defer_build=False, _shell=None, _arena=None):
        self.num_nodes = int(num_nodes)
        self.num_rels = int(num_rels)
        self.triples = triples_dev  # (T,3) int64 on the device, reference layout (None: index made by concat_graphs)
        if triples_dev is None:
            T, self.device = int(_shell[0]), _shell[1]
            _defer_build = True
        else:
            self.device = triples_dev.device
            T = int(triples_dev.shape[0])
        N, R, E = self.num_nodes, self.num_rels, 2 * T
        self.num_edges = E
        dev = self.device
        # one arena for every int32 array of the index (a single allocation per snapshot); the per-array tensor views are
        # made on first access (_VIEWS / __getattr__): the evaluation loop only ever needs their addresses
        sizes = self._arena_sizes(N, R, E)
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (max(int(n), 1) + 3) // 4 * 4          # keep every view 16-byte aligned
        # _arena: a slice of a buffer shared by the snapshots of one batched build (build_sub_graphs)
        if _arena is not None and (_arena.numel() < tot + N or _arena.dtype != I32 or _arena.data_ptr() % 16):
            raise ValueError("SnapshotGraph: the shared arena slice is too small or misaligned")
        arena = _arena if _arena is not None else torch.empty(tot + N, device=dev, dtype=I32)
        self._arena = arena
        self._layout = {name: (o, max(int(n), 1)) for name, o, n in zip(_VIEWS, offs, sizes)}
        self._layout["norm"] = (tot, N)
        base = arena.data_ptr()
        self._addr = {name: base + 4 * o for name, (o, _) in self._layout.items()}
        if not _defer_build:
            ws_bytes = _lib.load().regcn_csr_build_workspace_bytes(T, N, R)
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            a = self._addr
            call("regcn_csr_build", ptr(triples_dev), T, N, R, a["src"], a["dst"], a["etype"], a["indeg"], a["norm"], a["rowptr"],
                 a["src_sorted"], a["etype_sorted"], a["eperm"], a["vptr"], a["sptr"], a["vrow_row"], a["active_pos"],
                 a["active_rows"], a["rel_rowptr"], a["rel_ents"], a["_counts"], ptr(ws), ws_bytes)
        self._ndata = None
        self._edata = None
        self._r2e = None
        self.ptr_table = np.array([self._addr[k] for k in ("rowptr", "src_sorted", "etype_sorted", "indeg", "norm", "vptr", "sptr",
                                                          "vrow_row", "rel_rowptr", "rel_ents", "active_pos", "active_rows")],
                                  dtype=np.uint64)
        if not _defer_counts and not _defer_build:
            self._set_counts(self._counts.tolist())          # the one host sync of graph construction

    def __getattr__(self, name):
        # tensor view of one array of the index, created on first access (only reached when the attribute is missing)
        lay = self.__dict__.get("_layout")
        if lay is not None and name in lay:
            o, n = lay[name]
            v = self._arena[o:o + n]
            if name == "norm":
                v = v.view(torch.float32)
            self.__dict__[name] = v
            return v
        raise AttributeError(name)

    def _descriptor(self, desc):
        """Fill a struct regcn_csr_arrays with this snapshot's pointers (batched build)."""
        desc.triples = self.triples.data_ptr() if self.triples is not None else None
        desc.T = self.num_edges // 2
        for name in ("src", "dst", "etype", "indeg", "norm", "rowptr", "src_sorted", "etype_sorted", "eperm", "vptr",
                     "sptr", "vrow_row", "active_pos", "active_rows", "rel_rowptr", "rel_ents"):
            setattr(desc, name, self._addr[name])
        desc.counts = self._addr["_counts"]

    def _set_counts(self, c):
        self.n_vrows, self.n_split_chunks, self.n_rel_ents, self.max_hub_degree, self.n_active = c[:5]
        # pointer / int tables consumed by the whole-recurrence entry points (include/regcn_b200.h RG_* / RGI_*)
        max_chunks = (self.max_hub_degree + AGG_CHUNK - 1) // AGG_CHUNK
        self.int_table = np.array([self.num_edges, self.n_vrows, self.n_split_chunks, self.n_rel_ents, self.n_active,
                                   max_chunks], dtype=np.int32)

    # ---- the slice of the DGL surface the reference modules use -----------------
    def number_of_nodes(self):
        return self.num_nodes

    def number_of_edges(self):
        return self.num_edges

    def in_degrees(self, v=None):
        deg = self.indeg[: self.num_nodes].long()
        if v is None or isinstance(v, range):
            return deg
        return deg[torch.as_tensor(v, device=self.device, dtype=torch.long)]

    def to(self, device):
        return self

    @property
    def ndata(self):
        if self._ndata is None:
            f = _Frame()
            f["id"] = torch.arange(self.num_nodes, device=self.device, dtype=torch.long).view(-1, 1)
            f["norm"] = self.norm.view(-1, 1)
            self._ndata = f
        return self._ndata

    @property
    def edata(self):
        if self._edata is None:
            f = _Frame()
            f["type"] = self.etype[: self.num_edges].long()
            self._edata = f
        return self._edata

    def edges(self):
        return self.src[: self.num_edges].long(), self.dst[: self.num_edges].long()

    def _host_r2e(self):
        if self._r2e is None:
            R = self.num_rels
            rp = self.rel_rowptr.cpu().numpy().astype(np.int64)
            ents = self.rel_ents[: self.n_rel_ents].cpu().numpy().astype(np.int64)
            present = np.nonzero(rp[1:] > rp[:-1])[0]
            uniq_r = np.concatenate((present, present + R))
            r_len, e_idx, idx = [], [], 0
            for r in uniq_r:
                b, e = rp[r % R], rp[r % R + 1]
                r_len.append((idx, idx + int(e - b)))
                e_idx.extend(ents[b:e].tolist())
                idx += int(e - b)
            self._r2e = (uniq_r, r_len, torch.as_tensor(e_idx, dtype=torch.long, device=self.device))
        return self._r2e

    @property
    def uniq_r(self):
        return self._host_r2e()[0]

    @property
    def r_len(self):
        return self._host_r2e()[1]

    @property
    def r_to_e(self):
        return self._host_r2e()[2]


def build_sub_graph(num_nodes, num_rels, triples, use_cuda=True, gpu=0):
    """Drop-in for rgcn/utils.py:100 `build_sub_graph(num_nodes, num_rels, triples, use_cuda, gpu)`.

    `triples` is the reference's (T,3) int numpy array (or a tensor).  The index is always built on
    the device (`use_cuda=False` is refused: the product path has no CPU implementation)."""
    if not use_cuda:
        raise RuntimeError("regcn_b200.build_sub_graph: use_cuda=False is not supported (no CPU path)")
    _lib.require_device()
    dev = torch.device("cuda", gpu) if isinstance(gpu, int) else torch.device(gpu)
    if isinstance(triples, torch.Tensor):
        t = triples.to(device=dev, dtype=torch.int64)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(triples, dtype=np.int64)).reshape(-1, 3)).to(dev)
    return SnapshotGraph(num_nodes, num_rels, t.contiguous())


def build_sub_graphs(num_nodes, num_rels, triples_list, device, sync=True):
    """Build the edge index of several snapshots with ONE C-ABI call (`regcn_csr_build_batch`: one CTA per small
    snapshot, all of them in one launch) and ONE host synchronisation for the per-snapshot size counters.
    `triples_list`: (T_i,3) int64 tensors (pinned host or device).  With sync=False the counters are left on the
    device: call `finish_sub_graphs(gs)` (or pass them through `pending_counts`) before using the graphs."""
    import ctypes
    _lib.require_device()
    dev_t = [None] * len(triples_list)
    host = [i for i, t in enumerate(triples_list) if not t.is_cuda and t.dtype == torch.int64 and t.dim() == 2 and t.shape[0] > 0]
    if len(host) > 2:
        # several host snapshots (the start of a test() call uploads the whole history window): ONE staged copy instead of
        # one ~25 us host-side copy call each; the pinned staging block comes from torch's caching host allocator
        sizes = [int(triples_list[i].numel()) for i in host]
        stage = torch.empty(sum(sizes), dtype=torch.int64, pin_memory=True)
        torch.cat([triples_list[i].reshape(-1) for i in host], out=stage)
        on_dev = stage.to(device, non_blocking=True)
        o = 0
        for i, n in zip(host, sizes):
            dev_t[i] = on_dev[o:o + n].view(-1, 3)
            o += n
    arenas = [None] * len(triples_list)
    if len(triples_list) > 2:
        # one allocation for the arenas of the batch, in a size class that repeats from call to call (a fresh size per
        # snapshot keeps the caching allocator growing its pool -- a device allocation while kernels are queued stalls the
        # host -- through the first calls of a process)
        need = [(SnapshotGraph.arena_ints(num_nodes, num_rels, t.shape[0]) + 63) // 64 * 64 for t in triples_list]
        chunk = 1 << 23                               # 32 MB steps: a 24-snapshot ICEWS18-shaped batch is ~18 MB
        big = torch.empty((sum(need) + chunk - 1) // chunk * chunk, device=device, dtype=I32)
        o = 0
        for i, n_ in enumerate(need):
            arenas[i] = big[o:o + n_]
            o += n_
    gs = [SnapshotGraph(num_nodes, num_rels, (t.to(device, non_blocking=True) if d is None else d).contiguous(), _defer_build=True,
                        _arena=a) for t, d, a in zip(triples_list, dev_t, arenas)]
    if gs:
        L = len(gs)
        descs = (_lib.CsrArrays * L)()
        for g, dsc in zip(gs, descs):
            g._descriptor(dsc)
        Ts = (ctypes.c_int32 * L)(*[g.num_edges // 2 for g in gs])
        lib = _lib.load()
        ws_bytes = lib.regcn_csr_build_batch_workspace_bytes(Ts, L, int(num_nodes), int(num_rels))
        ws = torch.empty((max(ws_bytes, 1) + (1 << 22) - 1) >> 22 << 22, device=gs[0].device, dtype=torch.uint8)
        call("regcn_csr_build_batch", ctypes.cast(descs, ctypes.c_void_p), L, int(num_nodes), int(num_rels), ptr(ws),
             ws_bytes)
        if sync:
            finish_sub_graphs(gs)
    return gs


def pending_counts(gs):
    """Device tensor (L, 8) of the size counters of graphs built with sync=False."""
    return torch.stack([g._counts for g in gs])


def finish_sub_graphs(gs, counts=None):
    """Read the size counters back (one device->host copy) and finalise the graphs."""
    if gs:
        if counts is None:
            counts = pending_counts(gs).tolist()
        for g, c in zip(gs, counts):
            g._set_counts(c)
    return gs


def concat_graphs(graphs, out=None):
    """Block-diagonal union of G finished SnapshotGraphs over the same (N, R) as ONE SnapshotGraph over G*N entities and
    G*R relations (`regcn_csr_concat`, one launch of offset copies): member g's entity v is g*N + v, its relation r < R
    is g*R + r, the inverse relation R + r is G*R + g*R + r.  This is the graph one recurrence step runs on when G
    independent history windows are evolved together (evaluate.test: consecutive test timestamps, src/main.py:60-90).
    No host synchronisation: the members' size counters are already on the host.
    out: a union graph returned earlier for the same G whose arrays are reused when they are large enough (the caller
    guarantees that no kernel still reads it on another stream)."""
    import ctypes
    G = len(graphs)
    g0 = graphs[0]
    N, R = g0.num_nodes, g0.num_rels
    if any(g.num_nodes != N or g.num_rels != R or g.device != g0.device for g in graphs):
        raise ValueError("concat_graphs: members must share num_nodes, num_rels and device")
    T = sum(g.num_edges // 2 for g in graphs)
    if (out is None or getattr(out, "_cap_T", -1) < T or out.num_nodes != G * N or out.num_rels != G * R
            or out.device != g0.device):
        cap = T + T // 4 + 16
        out = SnapshotGraph(G * N, G * R, None, _shell=(cap, g0.device))
        out._cap_T = cap
        out._cdesc = _lib.CsrArrays()
        out._descriptor(out._cdesc)
    out.num_edges = 2 * T
    for g in graphs:                                   # per-member descriptor and size row, made once per snapshot
        if "_cdesc" not in g.__dict__:
            g._cdesc = _lib.CsrArrays()
            g._descriptor(g._cdesc)
            g._sz5 = np.array([g.n_vrows, g.n_split_chunks, g.n_rel_ents, g.n_active, g.max_hub_degree], dtype=np.int32)
    descs = (_lib.CsrArrays * G)(*[g._cdesc for g in graphs])
    sz = np.stack([g._sz5 for g in graphs])
    sizes = np.ascontiguousarray(sz[:, :4])
    call("regcn_csr_concat", ctypes.addressof(descs), sizes.ctypes.data, G, N, R, ctypes.addressof(out._cdesc))
    tot = sz.sum(axis=0)
    out._set_counts([int(tot[0]), int(tot[1]), int(tot[2]), int(sz[:, 4].max()), int(tot[3])])
    out._ndata = out._edata = out._r2e = None
    out.members = G
    return out


class SnapshotCache:
    """Device-resident index cache: the reference rebuilds the DGL graph of every history snapshot for every evaluated
    timestamp (src/main.py:68, hyperbolic_main.py:100-113 -- L graphs per step, of which L-1 were built the step before);
    here a snapshot's index is built once, kept in HBM (a few hundred KB per snapshot) and reused while it stays in the
    history window.  Keyed by the identity of the triple array (the reference's sliding `input_list` holds the same
    array objects from step to step); entries leave in insertion order once `capacity` is exceeded."""

    def __init__(self, num_nodes, num_rels, device, capacity=64):
        self.num_nodes, self.num_rels, self.device, self.capacity = int(num_nodes), int(num_rels), device, int(capacity)
        self._graphs = {}          # id(array) -> (array (kept alive), SnapshotGraph)

    def ensure(self, snapshots, device_copies=None):
        """Graphs of `snapshots` (list of (T,3) int64 numpy arrays / tensors), building the missing ones in one batched
        call WITHOUT reading their size counters back.  Returns (graphs, new_graphs): call finish_sub_graphs(new_graphs,
        counts) (or let `finish` do it) before the graphs are used.  device_copies: {id(snapshot): (T,3) int64 device
        tensor} for snapshots whose triples are already in HBM (a test snapshot was uploaded as the queries of its own
        timestamp before it slides into a history window): no second upload."""
        missing = [s for s in snapshots if id(s) not in self._graphs]
        new = []
        if missing:
            uniq, seen = [], set()
            for s in missing:
                if id(s) not in seen:
                    seen.add(id(s))
                    uniq.append(s)
            have = device_copies or {}
            tens = [have[id(s)] if id(s) in have else
                    (s if isinstance(s, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(s, dtype=np.int64)))
                    for s in uniq]
            new = build_sub_graphs(self.num_nodes, self.num_rels, tens, self.device, sync=False)
            for s, g in zip(uniq, new):
                self._graphs[id(s)] = (s, g)
            while len(self._graphs) > self.capacity:
                self._graphs.pop(next(iter(self._graphs)))
        return [self._graphs[id(s)][1] for s in snapshots], new

    def __len__(self):
        return len(self._graphs)
