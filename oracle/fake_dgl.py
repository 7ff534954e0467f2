"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (regcn_b200/).

Pure-torch stand-in for the small slice of the DGL 0.5.2 API that the reference's
hot path touches.  It exists so that the *unmodified*
reference modules under /root/reference can be imported and executed on CPU in
this container (DGL is not installable: no network) to
  (a) generate the golden fixtures committed under tests/golden/ and
  (b) validate the independent restatement in oracle/restate.py.

Semantics honoured (call sites in the reference):
  dgl.graph((src, dst), num_nodes=N)      rgcn/utils.py:120   multigraph, edge id = position
  g.in_degrees(range(N))                   rgcn/utils.py:111, rgcn/layers.py:231
  g.ndata / g.edata dict frames            rgcn/utils.py:123-125
  g.apply_edges(udf)                       rgcn/utils.py:124
  g.update_all(msg, fn.sum(...), apply)    rgcn/layers.py:144,175,220; hyperbolic_layers.py:134,290
  g.update_all(msg, reduce_udf)            hyperbolic_layers.py:453,665,892 (degree-bucketed mailbox)
  g.to(device), g.number_of_nodes()
Zero-in-degree nodes receive exact zeros from both reduce flavours (DGL >= 0.5
behaviour, recalled in SURVEY.md section 5.1; not verifiable here).

The scatter-sum is torch.index_add_ (sequential edge order on CPU), not DGL's
C++ gspmm kernel -- this caveat is printed next to every CPU-baseline number.
"""
import sys
import types

import torch


class _Frame(dict):
    """dict with the .update/.pop/.get the reference uses on ndata/edata."""


class _EdgeBatch:
    def __init__(self, g):
        self._g = g
        self.src = _Lazy(lambda k: g.ndata[k][g._src])
        self.dst = _Lazy(lambda k: g.ndata[k][g._dst])
        self.data = g.edata


class _Lazy:
    def __init__(self, fn):
        self._fn = fn

    def __getitem__(self, k):
        return self._fn(k)


class _NodeBatch:
    def __init__(self, data, mailbox=None):
        self.data = data
        self.mailbox = mailbox


class _SumReduce:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


class Graph:
    def __init__(self, edges, num_nodes):
        src, dst = edges
        self._src = torch.as_tensor(src, dtype=torch.long)
        self._dst = torch.as_tensor(dst, dtype=torch.long)
        self._n = int(num_nodes)
        self.ndata = _Frame()
        self.edata = _Frame()

    # -- structure ---------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.numel())

    def in_degrees(self, v=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        if v is None:
            return deg
        idx = torch.as_tensor(list(v) if isinstance(v, range) else v, dtype=torch.long)
        return deg[idx]

    def edges(self):
        return self._src, self._dst

    def to(self, device):
        # DGL 0.5 returns a shallow copy that keeps python attributes;
        # on CPU this is the identity.
        return self

    # -- message passing ---------------------------------------------------
    def apply_edges(self, udf):
        out = udf(_EdgeBatch(self))
        for k, v in out.items():
            self.edata[k] = v

    def update_all(self, message_func, reduce_func, apply_node_func=None):
        msgs = message_func(_EdgeBatch(self))
        if isinstance(reduce_func, _SumReduce):
            m = msgs[reduce_func.msg]
            out = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype)
            out.index_add_(0, self._dst, m)
            self.ndata[reduce_func.out] = out
        else:
            self._udf_reduce(msgs, reduce_func)
        if apply_node_func is not None:
            res = apply_node_func(_NodeBatch(self.ndata))
            for k, v in res.items():
                self.ndata[k] = v

    def _udf_reduce(self, msgs, reduce_func):
        deg = torch.bincount(self._dst, minlength=self._n)
        order = torch.argsort(self._dst, stable=True)  # mailbox keeps edge order per node
        rowptr = torch.zeros(self._n + 1, dtype=torch.long)
        rowptr[1:] = torch.cumsum(deg, 0)
        results = {}
        for k in torch.unique(deg).tolist():
            if k == 0:
                continue
            nodes = torch.nonzero(deg == k, as_tuple=False).view(-1)
            eidx = (rowptr[nodes].view(-1, 1) + torch.arange(k).view(1, -1)).view(-1)
            eids = order[eidx]
            mailbox = {name: m[eids].view((nodes.numel(), k) + tuple(m.shape[1:]))
                       for name, m in msgs.items()}
            data = {name: v[nodes] for name, v in self.ndata.items()}
            out = reduce_func(_NodeBatch(data, mailbox))
            for name, v in out.items():
                if name not in results:
                    results[name] = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype)
                results[name][nodes] = v
        for name, v in results.items():
            self.ndata[name] = v


def graph(edges, num_nodes=None):
    return Graph(edges, num_nodes)


def install():
    """Register stub `dgl`, `dgl.function`, `dgl.data.utils`, `rdflib` modules and
    neutralise the reference's hard-coded `.cuda()` (rgcn/layers.py:230)."""
    if "dgl" in sys.modules and getattr(sys.modules["dgl"], "_regcn_fake", False):
        return
    dgl = types.ModuleType("dgl")
    dgl._regcn_fake = True
    dgl.graph = graph
    dgl.DGLGraph = Graph
    fn = types.ModuleType("dgl.function")
    fn.sum = lambda msg, out: _SumReduce(msg, out)
    dgl.function = fn
    data = types.ModuleType("dgl.data")
    data_utils = types.ModuleType("dgl.data.utils")
    for name in ("download", "extract_archive", "get_download_dir", "_get_dgl_url"):
        setattr(data_utils, name, lambda *a, **k: None)
    data.utils = data_utils
    dgl.data = data
    rdflib = types.ModuleType("rdflib")
    rdflib.Graph = object
    rdflib.URIRef = object
    sys.modules.update({"dgl": dgl, "dgl.function": fn, "dgl.data": data,
                        "dgl.data.utils": data_utils, "rdflib": rdflib})
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
