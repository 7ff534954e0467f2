"""CPU: the C-ABI library loads and exports exactly what include/regcn_b200.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import restate, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "regcn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"REGCN_API\s+[\w\s\*]+?\b(regcn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        protos[name] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_library_exports_every_declared_symbol():
    from regcn_b200 import _lib
    protos = _header_prototypes()
    assert len(protos) >= 27
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/regcn_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(protos), set(_lib.SIGNATURES) ^ set(protos)
    for name, n_args in protos.items():
        assert len(_lib.SIGNATURES[name][1]) == n_args, f"{name}: ctypes binding has the wrong arity"
    assert _lib.load().regcn_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_product_path_refuses_cpu():
    """No CPU fallback: without a B200 the public API raises instead of computing somewhere else."""
    import regcn_b200 as R
    from regcn_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no sm_100"):
        R.build_sub_graph(10, 2, np.array([[0, 0, 1]]), True, 0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.gemm(torch.randn(4, 4), torch.randn(4, 4))
    with pytest.raises(RuntimeError):
        R.build_sub_graph(10, 2, np.array([[0, 0, 1]]), False, 0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "regcn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("oracle/", ""), fn


def test_answer_dicts_and_split_by_time():
    from regcn_b200 import utils
    rng = np.random.default_rng(0)
    quads = np.stack([rng.integers(0, 20, 60), rng.integers(0, 4, 60), rng.integers(0, 20, 60),
                      np.sort(rng.integers(0, 5, 60)) * 24], 1)
    snaps = utils.split_by_time(quads)
    assert sum(len(s) for s in snaps) == 60 and len(snaps) == len(np.unique(quads[:, 3]))
    for s, t in zip(snaps, np.unique(quads[:, 3])):
        assert np.array_equal(s, quads[quads[:, 3] == t][:, :3])
    d = utils.load_all_answers_for_filter(snaps[0], 4, False)
    assert d == synth.answers_of(snaps[0], 4, False)
    assert utils.load_all_answers_for_filter(snaps[0], 4, True) == synth.answers_of(snaps[0], 4, True)
    assert len(utils.load_all_answers_for_time_filter(quads, 4, 20)) == len(snaps)


def test_filter_csr_builders_agree_on_cpu():
    from regcn_b200 import utils
    case = synth.make_case("small", 8)
    r = case["num_rels"]
    all_t = torch.as_tensor(restate.add_inverse(case["test"], r))
    for rel_p, nk in ((0, 2 * r), (1, case["num_ents"])):
        d = synth.answers_of(case["test"], r, bool(rel_p))
        a = utils.filter_csr_from_dict(all_t, d, rel_predict=rel_p)
        b = utils.filter_csr_from_snapshot(all_t, nk, rel_predict=rel_p)
        assert a.lists() == b.lists()
        # every query's own target is in its list; lists are sorted and unique
        al = a.lists()
        for q in range(0, len(all_t), 17):
            lst = al[q]
            assert lst == sorted(set(lst)) and int(all_t[q, 1 if rel_p else 2]) in lst


def test_state_dict_names_match_reference_listing():
    """The names the reference's checkpoints carry exist with the right shapes."""
    import regcn_b200 as R
    m = R.RecurrentRGCN("convtranse", "uvrgcn", 50, 4, 0, 0, 200, "sub", 3, num_bases=100, num_hidden_layers=2,
                        dropout=0.2, self_loop=True, skip_connect=True, layer_norm=True)
    sd = m.state_dict()
    for k in ("w1", "w2", "emb_rel", "dynamic_emb", "time_gate_weight", "time_gate_bias", "rgcn.rel_emb",
              "relation_cell_1.weight_ih", "relation_cell_1.bias_hh", "rgcn.layers.0.weight_neighbor",
              "rgcn.layers.1.loop_weight", "rgcn.layers.1.evolve_loop_weight", "rgcn.layers.1.skip_connect_weight",
              "rgcn.layers.1.skip_connect_bias", "decoder_ob.conv1.weight", "decoder_ob.bn0.running_mean",
              "decoder_ob.bn3.weight", "decoder_ob.bn_init.bias", "decoder_ob.fc.weight", "decoder_ob.b", "rdecoder.b"):
        assert k in sd, k
    assert "rgcn.layers.0.skip_connect_weight" not in sd
    assert tuple(sd["decoder_ob.fc.weight"].shape) == (200, 10000) and tuple(sd["rdecoder.b"].shape) == (8,)
    h = R.HyperbolicRecurrentRGCN("roth", "lgcn", 50, 50, 0, 0, 200, "sub", 3, num_bases=100, num_hidden_layers=2,
                                  dropout=0.2, self_loop=True)
    sd = h.state_dict()
    for k in ("c", "radius_target", "radius_static", "temporal_radius_evolution.radius_mlp.weight",
              "relation_gru.weight_ih", "rgcn.layers.0.weight", "rgcn.layers.1.evolve_loop_weight",
              "decoder_ob.rot_proj.weight", "decoder_ob.trans_proj.bias", "decoder_ob.reshape_fc2.weight",
              "decoder_ob.score_scale_raw", "decoder_ob.score_margin", "rdecoder.global_rot", "rdecoder.rel_bias"):
        assert k in sd, k
    assert tuple(sd["rgcn.layers.0.weight"].shape) == (100, 400)
    assert tuple(sd["decoder_ob.rot_proj.weight"].shape) == (100, 200)


def test_dataset_reader_round_trip(tmp_path):
    """The reference's on-disk format (rgcn/knowledge_graph.py:189-206,526-555): write a tiny dataset, read it back
    through load_data, cut it with split_by_time; compared with the reference's own reader when it is present."""
    from regcn_b200 import knowledge_graph, utils
    rng = np.random.default_rng(3)
    root = tmp_path / "data"
    d = root / "SMALL"
    d.mkdir(parents=True)
    (d / "entity2id.txt").write_text("".join(f"ent {i}\t{i}\n" for i in range(30)))
    (d / "relation2id.txt").write_text("".join(f"rel_{i}\t{i}\n" for i in range(5)))
    quads = {}
    t0 = 0
    for split, n in (("train", 80), ("valid", 20), ("test", 25)):
        q = np.stack([rng.integers(0, 30, n), rng.integers(0, 5, n), rng.integers(0, 30, n),
                      t0 + np.sort(rng.integers(0, 4, n)) * 24, np.zeros(n, dtype=np.int64)], 1)
        t0 = int(q[-1, 3]) + 24
        quads[split] = q
        (d / f"{split}.txt").write_text("".join("\t".join(str(int(v)) for v in row) + "\n" for row in q))
    data = utils.load_data("SMALL", str(root))
    assert (data.num_nodes, data.num_rels) == (30, 5)
    assert data.entity_dict[7] == "ent 7" and data.relation_dict[4] == "rel_4"
    for split in ("train", "valid", "test"):
        got = getattr(data, split)
        assert got.dtype == np.int64 and np.array_equal(got, quads[split][:, :4])
    snaps = utils.split_by_time(data.train)
    assert sum(len(s) for s in snaps) == 80
    with pytest.raises(ValueError):
        utils.load_data("FB15k", str(root))
    with pytest.raises(ValueError):
        knowledge_graph.RGCNLinkDataset("SMALL")
    ref_root = os.environ.get("REGCN_REFERENCE", "/root/reference")
    if os.path.isdir(ref_root):                      # build container only: the reference's own reader on the same files
        import sys
        from oracle import fake_dgl
        fake_dgl.install()
        sys.path.insert(0, ref_root)
        try:
            from rgcn import knowledge_graph as ref_kg
            ref = ref_kg.load_from_local(str(root), "SMALL")
        finally:
            sys.path.remove(ref_root)
        assert (ref.num_nodes, ref.num_rels) == (data.num_nodes, data.num_rels)
        for split in ("train", "valid", "test"):
            assert np.array_equal(np.asarray(getattr(ref, split)), getattr(data, split))


def test_group_sizes_of_the_evaluation_loop():
    """evaluate.group_sizes: every timestamp in exactly one group, no group above G, ramp sizes only while below G."""
    from regcn_b200.evaluate import group_sizes
    assert group_sizes(32, 32) == [8, 24]
    assert group_sizes(32, 32, "4,8") == [4, 8, 20]
    assert group_sizes(5, 32) == [5]
    assert group_sizes(0, 32) == []
    assert group_sizes(7, 1) == [1] * 7
    assert group_sizes(100, 32) == [8, 31, 31, 30]
    assert group_sizes(20, 8) == [7, 7, 6]                # a ramp step that is not below G is skipped
    assert group_sizes(20, 8, "") == [7, 7, 6]
    for K in range(0, 80):
        for G in (1, 2, 8, 24, 32):
            for ramp in ("8", "4,8", "2", ""):
                sz = group_sizes(K, G, ramp)
                assert sum(sz) == K and all(0 < s <= max(G, 1) for s in sz)
