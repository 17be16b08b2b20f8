"""The reference's own tests replayed against the CPU oracle (pins the restatement).

test/search.jl:26-49  -- the only known-answer test of the reference (toy 2x13 matrix)
test/utils.jl          -- push!/pushfirst!/pop!/popfirst!/delete_from_index! semantics
test/index.jl:31-42    -- constructor assertions (host mirror; fire before any device use)
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers


@pytest.mark.parametrize("seed", range(8))
def test_search_results_toy_known_answer(seed):
    # IVFADCIndex(data, kc=3, k=8, m=2); expected neighbour sets test/search.jl:34-46
    idx, qz, _, _ = helpers.build_oracle_index(helpers.TOY, kc=3, k=8, m=2, seed=seed)
    for w, expected in ((1, helpers.TOY_W1), (2, helpers.TOY_W2)):
        for point, result in zip(helpers.TOY_POINTS, expected):
            ids, dists, counts = idx.knn_search(np.array(point)[None, :], 5, w=w)
            neighbors = [int(i) + 1 for i in ids[0, :counts[0]]]
            assert set(neighbors) <= set(result), (seed, w, point, neighbors)
            # (fewer than k results when the probed lists are short: test/search.jl:35)
            assert counts[0] == len(result) or w == 2
            assert np.all(np.diff(dists[0, :counts[0]]) >= 0)


def test_search_types_and_asserts():
    # test/search.jl:1-23
    idx, qz, _, _, data = helpers.reference_fixture()
    rng = np.random.default_rng(5)
    q = rng.random(10)
    ids, dists, counts = idx.knn_search(q[None, :], 3, w=2)
    assert ids.dtype == np.uint64 and dists.dtype == np.float64 and counts[0] <= 3
    with pytest.raises(AssertionError):
        idx.knn_search(q[None, :], 0)
    with pytest.raises(AssertionError):
        idx.knn_search(q[None, :], 1, w=0)
    Q = rng.random((10, 10))
    ids, dists, counts = idx.knn_search(Q, 3, w=2)
    assert ids.shape == (10, 3)


def test_push_pushfirst_capacity_and_dimension():
    # test/utils.jl:1-29 with UInt8 ids: 243 -> 256 vectors, the 257th push throws
    idx, qz, _, _, data = helpers.reference_fixture(id_bytes=1)
    rng = np.random.default_rng(1)
    ol = len(idx)
    nnv = 256 - 243
    for _ in range(nnv):
        idx.push(rng.random(10))
    assert len(idx) == ol + nnv
    with pytest.raises(AssertionError):
        idx.push(rng.random(10))
    idx.delete_from_index([1])
    with pytest.raises(AssertionError):
        idx.push(rng.random(11))
    for i in range(1, nnv):
        idx.delete_from_index([i])
    for _ in range(nnv):
        idx.pushfirst(rng.random(10))
    assert len(idx) == ol + nnv
    with pytest.raises(AssertionError):
        idx.pushfirst(rng.random(10))
    # ids stay a permutation of 0..N-1
    all_ids = sorted(i for ids, _ in idx.lists for i in ids)
    assert all_ids == list(range(256))


def test_pop_popfirst():
    # test/utils.jl:32-55
    idx, qz, _, X, data = helpers.reference_fixture(id_bytes=1)
    ol = len(idx)
    v = idx.pop()
    assert v.dtype == np.float64 and v.shape == (10,)
    assert len(idx) == ol - 1
    # the popped vector is the quantised reconstruction of the last data point
    assert np.linalg.norm(v - X[-1]) < np.linalg.norm(X[-1])
    v = idx.popfirst()
    assert v.shape == (10,) and len(idx) == ol - 2
    assert sorted(i for ids, _ in idx.lists for i in ids) == list(range(ol - 2))


def test_delete_from_index_renumbering():
    # test/utils.jl:58-105, same ranges and the same bookkeeping
    idx, qz, _, _, data = helpers.reference_fixture()
    import copy
    before = copy.deepcopy(idx.lists)
    n = len(idx)
    L1s, L1e, L2s, L2e, L3s, L3e = 1, 5, 10, 30, n - 5, n
    to_delete = list(range(L1s, L1e + 1)) + list(range(L2s, L2e + 1)) + list(range(L3s, L3e + 1))
    idx.delete_from_index(to_delete)
    assert len(idx) == n - len(to_delete)
    mismatches = 0
    for cl, (ids_old, codes_old) in enumerate(before):
        ids_new, codes_new = idx.lists[cl]
        found = set(ids_old) & {d - 1 for d in to_delete}
        assert len(ids_old) == len(ids_new) + len(found)
        for i, id0 in enumerate(ids_old):
            id1 = id0 + 1
            if L1e < id1 < L2s:
                shift = L1e - L1s + 1
            elif L2e < id1 < L3s:
                shift = (L1e - L1s + 1) + (L2e - L2s + 1)
            else:
                continue
            newval = id1 - shift - 1
            newpos = ids_new.index(newval)
            if not np.array_equal(codes_old[i], codes_new[newpos]):
                mismatches += 1
    assert mismatches == 0


def test_delete_closed_form():
    """The closed form the CUDA compaction implements: new_id = old_id - |{deleted < old_id}|,
    in-list order preserved -- checked against the literal descending loop."""
    idx, qz, _, _, data = helpers.reference_fixture(seed=3)
    import copy
    before = copy.deepcopy(idx.lists)
    rng = np.random.default_rng(9)
    dele = sorted(set(int(x) for x in rng.integers(1, 244, size=60)))
    idx.delete_from_index(dele + dele[:5] + [10_000])  # duplicates and unknown ids are ignored
    d0 = np.array([d - 1 for d in dele])
    for cl, (ids_old, codes_old) in enumerate(before):
        keep = [j for j, i in enumerate(ids_old) if i not in set(d0)]
        exp_ids = [ids_old[j] - int((d0 < ids_old[j]).sum()) for j in keep]
        assert exp_ids == idx.lists[cl][0]
        for a, j in zip(idx.lists[cl][1], keep):
            assert np.array_equal(a, codes_old[j])
    with pytest.raises(OverflowError):
        idx.delete_from_index([0])  # I.(points .- 1) -> InexactError


def test_oracle_vs_fp64_truth():
    """Error budget: oracle fp32 distances within 1e-5 relative of exact arithmetic, ids equal
    except at near-ties."""
    rng = np.random.default_rng(2)
    data = rng.random((50, 1000)).astype(np.float32)
    idx, qz, _, X = helpers.build_oracle_index(data, kc=20, k=32, m=10, seed=2)
    off, codes, ids = idx.csr()
    Q = rng.random((20, 50)).astype(np.float32)
    oi, od, oc = idx.knn_search(Q, 5, w=3)
    for i in range(20):
        ti, td = orc.truth_search_fp64(qz, off, codes, ids, Q[i], 5, 3)
        assert oc[i] == len(ti)
        np.testing.assert_allclose(od[i, :oc[i]], td, rtol=1e-5)
        diff = oi[i, :oc[i]] != ti
        if diff.any():  # only allowed at near-ties of the exact distances
            assert np.all(np.abs(np.diff(td))[np.flatnonzero(diff)[:-1]] < 1e-5 * td.max())
