"""CPU: host-side pieces of the operator surface (loaders, synthetic graphs, argument checking)."""
import os

import numpy as np
import pytest
import torch

import oracle
from graph_loader import GraphDataLoader, save_warp4, warp4_path
from synth_graphs import SHAPES, shape_graph, symmetrize, synth_graph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_graph_loader_round_trip_and_reference_edge_values(tmp_path):
    ref = np.load(os.path.join(GOLD, "ref_py.npz"))
    loader = GraphDataLoader(str(tmp_path) + "/")
    loader.save_graph("small", ref["indptr"], ref["indices"])
    assert loader.get_available_graphs() == ["small"]
    g = loader.load_graph("small.dgl")                      # extension is stripped like graph_loader.py:52
    assert np.array_equal(g["indptr"], ref["indptr"]) and np.array_equal(g["indices"], ref["indices"])
    assert g["v_num"] == 200 and g["e_num"] == 3000
    assert np.array_equal(g["values"], ref["loader_values"])   # np.random.seed(123) U[0,1), graph_loader.py:71-72
    with pytest.raises(FileNotFoundError):
        loader.load_graph("missing")


def test_warp4_file_convention(tmp_path):
    p = warp4_path("reddit.dgl", root=str(tmp_path))
    assert p.endswith("w12_nz64_warp_4/reddit.dgl.warp4")
    assert warp4_path("g", csc=True).endswith("w12_nz64_warp_4_csc/g.warp4_csc")
    quads, _ = oracle.warp4(np.array([0, 3, 3, 200], np.int32), 64)
    save_warp4(p, quads)
    assert np.array_equal(np.fromfile(p, dtype=np.int32), quads)
    assert quads.reshape(-1, 4).tolist() == [[0, 0, 3, 0], [2, 3, 64, 0], [2, 67, 64, 0], [2, 131, 64, 0], [2, 195, 5, 0]]


@pytest.mark.parametrize("kind", ["uniform", "powerlaw"])
def test_synthetic_graph_is_valid_csr(kind):
    g = synth_graph(1000, 25000, seed=1, kind=kind)
    ip, ix = g["indptr"].numpy(), g["indices"].numpy()
    assert ip[0] == 0 and ip[-1] == 25000 == ix.size and (np.diff(ip) >= 0).all()
    assert ix.min() >= 0 and ix.max() < 1000
    for r in (0, 10, 999):
        assert (np.diff(ix[ip[r]:ip[r + 1]]) >= 0).all()          # sorted within a row
    g2 = synth_graph(1000, 25000, seed=1, kind=kind)
    assert torch.equal(g["indices"], g2["indices"]) and torch.equal(g["values"], g2["values"])   # seeded
    if kind == "powerlaw":
        assert np.diff(ip).max() > 5 * 25           # heavy tail: hubs well above the mean degree


def test_symmetrize_is_symmetric_with_self_loops():
    s = symmetrize(synth_graph(200, 1500, seed=4))
    import scipy.sparse as sp
    a = sp.csr_matrix((s["values"].numpy(), s["indices"].numpy(), s["indptr"].numpy()), shape=(200, 200))
    assert (a != a.T).nnz == 0 and (a.diagonal() == 1).all()


def test_named_shapes_match_baseline_configs():
    assert SHAPES["reddit"] == (232_965, 114_615_892) and SHAPES["flickr"][0] == 89_250
    g = shape_graph("flickr", scale=0.01)
    assert g["v_num"] == 892 and abs(g["e_num"] / g["v_num"] - SHAPES["flickr"][1] / SHAPES["flickr"][0]) < 1


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the binding raises before reaching the library when a tensor is not on a GPU."""
    import maxk_cuda_kernels as k
    from maxk_models_integrated import MaxK
    from maxk_spgemm_function import maxk_spgemm
    x = torch.rand(4, 256)
    with pytest.raises(RuntimeError):
        k.topk_cbsr(x, 8)
    with pytest.raises(RuntimeError):
        MaxK.apply(x, 8)
    with pytest.raises(RuntimeError):
        maxk_spgemm(torch.zeros(1, dtype=torch.int32), torch.ones(1), x, 8, None, 0, torch.zeros(5, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        k.load_warp4_metadata("definitely_missing_graph")


def test_clean_edges_matches_dataset_gen_steps():
    """dataset_gen.py:44-98: undirected -> self loops -> dedup, checked against a plain Python set."""
    from graph_loader import clean_edges
    rng = np.random.default_rng(5)
    n = 300
    src, dst = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    src[:50], dst[:50] = src[50:100], dst[50:100]           # guaranteed multi-edges
    indptr, indices = clean_edges(torch.from_numpy(src), torch.from_numpy(dst), n)
    want = set(zip(src.tolist(), dst.tolist())) | set(zip(dst.tolist(), src.tolist())) | {(i, i) for i in range(n)}
    rows = np.repeat(np.arange(n), np.diff(indptr.numpy()))
    got = list(zip(rows.tolist(), indices.tolist()))
    assert len(got) == len(set(got)) == len(want) and set(got) == want
    assert indptr.dtype == torch.int32 and indices.dtype == torch.int32 and int(indptr[-1]) == len(want)
    # directed, no loops: only dedup
    ip2, ix2 = clean_edges(src, dst, n, undirected=False, self_loops=False)
    assert int(ip2[-1]) == len(set(zip(src.tolist(), dst.tolist())))
    e_ptr, e_idx = clean_edges([], [], 4)                      # empty edge list -> the identity pattern
    assert e_ptr.tolist() == [0, 1, 2, 3, 4] and e_idx.tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        clean_edges([0, 5], [1, 1], 4)


def test_csr_to_csc_matches_scipy():
    from scipy.sparse import csr_matrix
    from graph_loader import csr_to_csc
    g = synth_graph(400, 9000, seed=3, kind="powerlaw")
    ip, ix, va = g["indptr"].numpy(), g["indices"].numpy(), g["values"].numpy()
    rows = np.repeat(np.arange(400), np.diff(ip))
    t_ptr, t_idx, t_val = csr_to_csc(g["indptr"], g["indices"], g["values"])
    dense = np.zeros((400, 400))
    np.add.at(dense, (rows, ix), va.astype(np.float64))
    t_rows = np.repeat(np.arange(400), np.diff(t_ptr.numpy()))
    dense_t = np.zeros((400, 400))
    np.add.at(dense_t, (t_rows, t_idx.numpy()), t_val.numpy().astype(np.float64))
    assert np.array_equal(dense.T, dense_t)
    ref = csr_matrix((np.ones(len(ix), np.float32), ix, ip), shape=(400, 400)).tocsc()   # generate_meta_csc.py:134-141
    assert np.array_equal(t_ptr.numpy(), ref.indptr)
    # a twice-transposed matrix is the original with sorted columns
    b_ptr, b_idx = csr_to_csc(t_ptr, t_idx)
    assert np.array_equal(b_ptr.numpy(), ip)
    for r in (0, 7, 399):
        assert np.array_equal(b_idx.numpy()[ip[r]:ip[r + 1]], np.sort(ix[ip[r]:ip[r + 1]]))


def test_sharding_of_any_graph_is_a_partition_of_its_edges():
    """hypothesis: for random ragged graphs, world sizes and both partitions, the row slabs hold every edge
    exactly once, padded column positions are a bijection onto the real rows of the padded numbering, and
    re-assembling the slabs gives back A (as a dense matrix in the padded numbering)."""
    from hypothesis import given, settings, strategies as st
    from sharded import balanced_bounds, padded_position, shard_columns, shard_rows, uniform_bounds

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 60), st.integers(1, 8), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(n, world, seed, balanced):
        rng = np.random.default_rng(seed)
        deg = rng.integers(0, 12, n)
        deg[rng.random(n) < 0.3] = 0
        if rng.random() < 0.5:
            deg[rng.integers(0, n)] += 200                   # a hub
        indptr = np.zeros(n + 1, np.int64)
        indptr[1:] = np.cumsum(deg)
        e = int(indptr[-1])
        g = {"indptr": torch.from_numpy(indptr.astype(np.int32)), "indices": torch.from_numpy(rng.integers(0, n, e).astype(np.int32)),
             "values": torch.from_numpy(rng.standard_normal(e).astype(np.float32)), "v_num": n, "e_num": e}
        bounds = balanced_bounds(g["indptr"], world) if balanced else uniform_bounds(n, world)
        assert bounds[0] == 0 and bounds[-1] == n and len(bounds) == world + 1 and bounds == sorted(bounds)
        explicit = bounds if balanced else None
        slabs = [shard_rows(g, world, r, explicit) for r in range(world)]
        m = slabs[0]["v_num"]
        assert all(s["v_num"] == m for s in slabs) and sum(s["e_num"] for s in slabs) == e
        pos = padded_position(torch.arange(n), bounds, m)
        assert len(set(pos.tolist())) == n and int(pos.max()) < world * m
        dense = np.zeros((n, n))
        rows = np.repeat(np.arange(n), deg)
        np.add.at(dense, (rows, g["indices"].numpy()), g["values"].numpy().astype(np.float64))
        padded = np.zeros((world * m, world * m))
        padded[np.ix_(pos.numpy(), pos.numpy())] = dense
        rebuilt, rebuilt_cols = np.zeros_like(padded), np.zeros_like(padded)
        for r, s in enumerate(slabs):
            lrows = np.repeat(np.arange(m), np.diff(s["indptr"].numpy()))
            np.add.at(rebuilt, (r * m + lrows, s["indices"].numpy()), s["values"].numpy().astype(np.float64))
            c = shard_columns(g, world, r, explicit)
            crows = np.repeat(np.arange(world * m), np.diff(c["indptr"].numpy()))
            np.add.at(rebuilt_cols, (crows, r * m + c["indices"].numpy()), c["values"].numpy().astype(np.float64))
        assert np.array_equal(rebuilt, padded) and np.array_equal(rebuilt_cols, padded)

    check()
