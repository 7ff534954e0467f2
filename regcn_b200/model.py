"""BaseRGCN layer-stack scaffold (src/model.py:4-71), host glue only."""
import torch.nn as nn


class BaseRGCN(nn.Module):
    def __init__(self, num_nodes, h_dim, out_dim, num_rels, num_bases=-1, num_basis=-1, num_hidden_layers=1,
                 dropout=0, self_loop=False, skip_connect=False, encoder_name="", opn="sub", rel_emb=None,
                 use_cuda=False, analysis=False):
        super().__init__()
        self.num_nodes = num_nodes
        self.h_dim = h_dim
        self.out_dim = out_dim
        self.num_rels = num_rels
        self.num_bases = num_bases
        self.num_basis = num_basis
        self.num_hidden_layers = num_hidden_layers
        self.dropout = dropout
        self.skip_connect = skip_connect
        self.self_loop = self_loop
        self.encoder_name = encoder_name
        self.use_cuda = use_cuda
        self.run_analysis = analysis
        self.rel_emb = rel_emb          # registers `rgcn.rel_emb` in the state dict exactly like the reference
        self.opn = opn
        self.build_model()
        self.features = self.create_features()

    def build_model(self):
        self.layers = nn.ModuleList()
        i2h = self.build_input_layer()
        if i2h is not None:
            self.layers.append(i2h)
        for idx in range(self.num_hidden_layers):
            self.layers.append(self.build_hidden_layer(idx))
        h2o = self.build_output_layer()
        if h2o is not None:
            self.layers.append(h2o)

    def create_features(self):
        return None

    def build_input_layer(self):
        return None

    def build_hidden_layer(self, idx):
        raise NotImplementedError

    def build_output_layer(self):
        return None

    def forward(self, g):
        if self.features is not None:
            g.ndata['id'] = self.features
        for layer in self.layers:
            layer(g)
        return g.ndata.pop('h')
