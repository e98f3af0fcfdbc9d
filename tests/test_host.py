"""Host logic of the drop-in (no device): Parameters, TrainSet ids, KFold, BaseLine SGD, synth."""
import numpy as np
import pytest

import recommend_sys_b200 as rs
from oracle import binding as ob
from conftest import split


def test_parameters_typed_getters_panic_like_go():
    p = rs.Parameters({"k": 10, "userBased": False, "reg": 0.1, "sim": rs.Pearson, "type": "x"})
    assert p.GetInt("k", 40) == 10 and p.GetInt("mink", 1) == 1
    assert p.GetBool("userBased", True) is False
    assert p.GetFloat64("reg", 0.02) == 0.1 and p.GetFloat64("lr", 0.005) == 0.005
    assert p.GetSim("sim", rs.MSD) is rs.Pearson and p.GetSim("other", rs.MSD) is rs.MSD
    with pytest.raises(TypeError):        # val.(int) on a float64 panics, core/base.go:26
        rs.Parameters({"k": 10.0}).GetInt("k", 40)
    with pytest.raises(TypeError):        # val.(Sim) on anything else panics, core/base.go:47
        rs.Parameters({"sim": "pearson"}).GetSim("sim", rs.MSD)
    with pytest.raises(TypeError):
        rs.Parameters({"reg": 1}).GetFloat64("reg", 0.02)
    q = p.Copy()
    q["k"] = 3
    assert p["k"] == 10


def test_trainset_inner_ids_first_appearance(ml100k):
    u, i, r = split(ml100k["u1_base"])
    ts = rs.NewTrainSet(rs.NewRawSet(u, i, r))
    ots = ob.TrainSet(u, i, r)
    assert ts.UserCount == ots.user_count == 943
    assert ts.ItemCount == ots.item_count
    assert np.array_equal(ts.innerUsers, ots.inner_users())
    assert np.array_equal(ts.innerItems, ots.inner_items())
    assert ts.GlobalMean == ots.global_mean
    assert ts.ConvertUserID(int(u[0])) == 0 and ts.ConvertItemID(int(i[0])) == 0
    assert ts.ConvertUserID(10 ** 9) == rs.core.newID
    got = ts.convert_users(np.array([u[0], 10 ** 9, u[-1]]))
    assert got[0] == 0 and got[1] == -1 and got[2] == ts.ConvertUserID(int(u[-1]))


def test_kfold_partitions(ml100k):
    d = rs.NewRawSet(*split(ml100k["u_data"][:1003]))
    trains, tests = d.KFold(5, 0)
    assert [t.Length() for t in tests] == [201, 201, 201, 200, 200]   # core/data.go:54-60
    assert all(tr.Length() + te.Length() == 1003 for tr, te in zip(trains, tests))
    seen = np.concatenate([np.stack([t.Users, t.Items], 1) for t in tests])
    assert len({tuple(x) for x in seen.tolist()}) == 1003


def test_baseline_sgd_matches_oracle_bits(ml100k):
    u, i, r = split(ml100k["u2_base"])
    ts = rs.NewTrainSet(rs.NewRawSet(u, i, r))
    bl = rs.NewBaseLine(rs.Parameters({"nEpochs": 7}))
    bl.Fit(ts)
    ub, ib, gb = ob.TrainSet(u, i, r).baseline(n_epochs=7)
    assert np.array_equal(bl.userBias, ub) and np.array_equal(bl.itemBias, ib) and bl.globalBias == gb
    assert bl.Predict(int(u[0]), 10 ** 9) == gb + ub[0]


def test_synth_is_deterministic_and_well_formed():
    a = rs.core.synth_ratings(943, 1682, 100000, 0x5EED0001)
    b = rs.core.synth_ratings(943, 1682, 100000, 0x5EED0001)
    assert a.Length() == 100000
    assert np.array_equal(a.Users, b.Users) and np.array_equal(a.Items, b.Items) and np.array_equal(a.Ratings, b.Ratings)
    assert len(set(zip(a.Users.tolist(), a.Items.tolist()))) == 100000      # unique pairs
    assert set(np.unique(a.Ratings)) == {1.0, 2.0, 3.0, 4.0, 5.0}
    assert np.bincount(rs.NewTrainSet(a).innerUsers).min() >= 20            # MovieLens-like floor
    c = rs.core.synth_ratings(943, 1682, 100000, 0x5EED0002)
    assert not np.array_equal(a.Items, c.Items)


def test_metrics_match_oracle():
    rng = np.random.RandomState(0)
    p, t = rng.rand(1000) * 5, rng.randint(1, 6, 1000).astype(float)
    assert abs(rs.RMSE(p, t) - ob.rmse(p, t)) < 1e-12
    assert abs(rs.MAE(p, t) - ob.mae(p, t)) < 1e-12


def test_metric_known_answers():
    """core/eval_test.go:32-46 — TestRMSE / TestMAE, the reference's own vectors (tolerance 1e-5), on the
    host mirror and on the oracle."""
    a, b = [-2.0, 0.0, 2.0], [0.0, 0.0, 0.0]
    for rmse, mae in ((rs.RMSE, rs.MAE), (ob.rmse, ob.mae)):
        assert abs(rmse(np.array(a), np.array(b)) - 1.63299) < 1e-5
        assert abs(mae(np.array(a), np.array(b)) - 1.33333) < 1e-5


def test_slab_owner_deals_every_slab_once_and_evenly():
    from recommend_sys_b200.shard import slab_owner

    for world in (1, 2, 3, 8):
        for n_slabs in (1, 5, 25, 64):
            owners = [slab_owner(s, world) for s in range(n_slabs)]
            assert all(0 <= o < world for o in owners)
            # slab s costs ~ (n_slabs - s): the snake deal keeps the shares within one round of each other
            cost = np.zeros(world)
            for s_, o in enumerate(owners):
                cost[o] += n_slabs - s_
            if n_slabs >= 4 * world:
                assert cost.max() / cost.mean() < 1.12, (world, n_slabs, cost)


def test_shard_rows_cover_everything():
    from recommend_sys_b200.shard import shard_rows

    for n in (1, 7, 128, 943, 26744, 138493):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_rows(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[x][1] == parts[x + 1][0] for x in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 128


# ---- persistence of neighbour lists, and the loaders ----
def test_neighbor_lists_round_trip(tmp_path):
    rng = np.random.RandomState(3)
    idx = rng.randint(-1, 5000, size=(777, 40)).astype(np.int32)
    sim = rng.uniform(-1, 1, size=(777, 40))
    sim[idx < 0] = np.nan
    sim[5, 3] = -0.0
    f = tmp_path / "sub" / "lists.rsknn"          # core/dump.go:25 creates the directory
    rs.SaveNeighbors(f, idx, sim)
    i2, s2 = rs.LoadNeighbors(f)
    assert np.array_equal(i2, idx)
    assert np.array_equal(s2.view(np.uint64), sim.view(np.uint64))     # bit for bit, NaN payloads and -0.0 included
    raw = bytearray(f.read_bytes())
    raw[100] ^= 0x40                                # a flipped payload bit is caught by the checksum
    f.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        rs.LoadNeighbors(f)
    f.write_bytes(b"not a list")
    with pytest.raises(ValueError):
        rs.LoadNeighbors(f)
    with pytest.raises(OSError):
        rs.LoadNeighbors(tmp_path / "missing")


def test_loader_reference_semantics_and_half_stars(tmp_path):
    f = tmp_path / "ratings.csv"
    f.write_text("userId,movieId,rating,timestamp\n1,31,2.5,1260759144\n1,1029,3.0,1260759179\n7,31,4,1\n9,x,5,2\n")
    # the reference's loader (core/data.go:302-304): Atoi -> the header line and "2.5"/"3.0" become 0
    d = rs.LoadDataFromFile(f, sep=",")
    assert d.Users.tolist() == [0, 1, 1, 7, 9] and d.Items.tolist() == [0, 31, 1029, 31, 0]
    assert d.Ratings.tolist() == [0.0, 0.0, 0.0, 4.0, 5.0]
    # the extension keeps half-stars and drops the header
    d = rs.LoadDataFromFile(f, sep=",", floatRatings=True, hasHeader=True)
    assert d.Users.tolist() == [1, 1, 7, 9] and d.Ratings.tolist() == [2.5, 3.0, 4.0, 5.0]
    # tab-separated, CRLF, no trailing newline (the ml-100k u.data form)
    g = tmp_path / "u.data"
    g.write_bytes(b"196\t242\t3\t881250949\r\n186\t302\t3\t891717742")
    d = rs.LoadDataFromFile(g)
    assert d.Users.tolist() == [196, 186] and d.Items.tolist() == [242, 302] and d.Ratings.tolist() == [3.0, 3.0]
    with pytest.raises(OSError):
        rs.LoadDataFromFile(tmp_path / "nope")


def test_threaded_id_conversion_equals_serial():
    """rs_host_convert_dense_mt (the batch form of ConvertUserID / ConvertItemID, core/data.go:110-122): any number of
    host threads gives the serial result, unknown and negative ids become newID = -1, and a caller-provided output
    array (pinned staging memory on the GPU box) is filled in place."""
    rng = np.random.RandomState(5)
    known = rng.permutation(50_000)[:30_000].astype(np.int64)
    d = rs.NewRawSet(known, known[::-1].copy(), np.ones(len(known)))
    ts = rs.NewTrainSet(d)
    raw = rng.randint(-5, 60_000, size=700_000).astype(np.int64)
    want = np.array([ts.ConvertUserID(int(x)) for x in raw[:2000]], dtype=np.int32)
    got = ts.convert_users(raw)
    assert got.dtype == np.int32 and np.array_equal(got[:2000], want)
    assert (got[raw < 0] == -1).all() and (got[raw >= 50_000] == -1).all()
    table = ts._ulook[1]
    L = rs.core.host_lib()
    for threads in (1, 2, 3, 7, 64):
        out = np.full(len(raw), 123, dtype=np.int32)
        L.rs_host_convert_dense_mt(rs.core._ptr(table), len(table) - 1, rs.core._ptr(raw), len(raw), rs.core._ptr(out), threads)
        assert np.array_equal(out, got), threads
    out = np.empty(len(raw), dtype=np.int32)
    assert ts.convert_users(raw, out=out) is out and np.array_equal(out, got)
    small = ts.convert_items(raw[:10])              # below the threading threshold
    assert np.array_equal(small, np.array([ts.ConvertItemID(int(x)) for x in raw[:10]], dtype=np.int32))
    assert rs.core.host_threads() >= 1
