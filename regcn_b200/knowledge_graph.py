"""On-disk dataset format of the reference (rgcn/knowledge_graph.py:173-228, 526-555; rgcn/utils.py:356-365):

    <dir>/<name>/entity2id.txt, relation2id.txt     one `name<TAB>id` per line
    <dir>/<name>/train.txt, valid.txt, test.txt     one `s<TAB>r<TAB>o<TAB>t[<TAB>...]` quadruple per line (ids)

Host-side input parsing only: the arrays it returns are what `utils.split_by_time` cuts into the
snapshots the device path consumes.  Same class / attribute names as the reference so `src/main.py`-style callers work.
"""
import os

import numpy as np

TKG_DATASETS = ('ICEWS18', 'ICEWS14', "GDELT", "SMALL", "ICEWS14s", "ICEWS05-15", "YAGO", "WIKI")


def _read_dictionary(filename):
    """rgcn/knowledge_graph.py:526-532: {id: name}."""
    d = {}
    with open(filename, 'r') as f:
        for line in f:
            line = line.strip().split('\t')
            if len(line) < 2:
                continue
            d[int(line[1])] = line[0]
    return d


def _read_triplets_as_array(filename, load_time):
    """rgcn/knowledge_graph.py:542-555: the first 3 (or 4 with the timestamp) integer columns, int64."""
    width = 4 if load_time else 3
    rows = []
    with open(filename, 'r') as f:
        for line in f:
            parts = line.strip().split('\t')
            if len(parts) < width:
                if not line.strip():
                    continue
                raise ValueError(f"{filename}: expected {width} tab-separated integer columns, got {line!r}")
            rows.append([int(parts[i]) for i in range(width)])
    return np.asarray(rows, dtype=np.int64).reshape(-1, width)


class RGCNLinkDataset(object):
    """rgcn/knowledge_graph.py:137-206 (local-directory form; the download branch needs a network and is not kept)."""

    def __init__(self, name, dir=None):
        self.name = name
        if not dir:
            raise ValueError("regcn_b200.RGCNLinkDataset reads local datasets only: pass dir=<data root>")
        self.dir = os.path.join(dir, self.name)

    def load(self, load_time=True):
        entity_dict = _read_dictionary(os.path.join(self.dir, 'entity2id.txt'))
        relation_dict = _read_dictionary(os.path.join(self.dir, 'relation2id.txt'))
        self.train = _read_triplets_as_array(os.path.join(self.dir, 'train.txt'), load_time)
        self.valid = _read_triplets_as_array(os.path.join(self.dir, 'valid.txt'), load_time)
        self.test = _read_triplets_as_array(os.path.join(self.dir, 'test.txt'), load_time)
        self.num_nodes = len(entity_dict)
        self.num_rels = len(relation_dict)
        self.relation_dict = relation_dict
        self.entity_dict = entity_dict
        return self


def load_from_local(dir, dataset):
    """rgcn/knowledge_graph.py:221-228."""
    data = RGCNLinkDataset(dataset, dir)
    data.load()
    return data


def load_data(dataset, data_dir="../data"):
    """rgcn/utils.py:356-365 for the temporal datasets (the RDF entity-classification and FB15k branches are out of
    scope: SURVEY.md section 2, row 6)."""
    if dataset in TKG_DATASETS:
        return load_from_local(data_dir, dataset)
    raise ValueError('Unknown dataset: {}'.format(dataset))
