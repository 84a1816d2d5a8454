"""CPU tests of the input pipeline (iins_vae_b200/dataset.py): the zenodo schema loader, the splits, the scaler (against
sklearn's StandardScaler, which the reference uses at dataset.py:73-76) and the pinned batch ring."""
import numpy as np
import pytest
import torch

from iins_vae_b200 import dataset as D


def _frame(n, seed=0):
    import pandas as pd
    rng = np.random.RandomState(seed)
    codes = ["0000000000"] + D._OBSTACLE_FULL
    return pd.DataFrame({"CIR": [rng.rand(157) * (1 + i % 7) for i in range(n)], "Error": rng.rand(n) * 0.6,
                         "Room": rng.randint(0, 5, n), "Obstacles": [codes[i % len(codes)] for i in range(n)]})


def test_load_pkl_room_full_and_obstacle_full(tmp_path):
    df = _frame(200)
    path = tmp_path / "dataset.pkl"
    df.to_pickle(path)
    cir, err, label, room = D.load_pkl_data(str(path), "room_full", np.random.RandomState(1))
    assert cir.shape == (200, 157) and err.shape == (200, 1) and label.shape == (200, 1)
    # a joint shuffle: every (cir row, error, room) triple of the file is still together
    key = {round(float(c[0]), 12): (e, r) for c, e, r in zip(np.vstack(df["CIR"]), df["Error"], df["Room"])}
    for c, e, l in zip(cir, err, label):
        e0, r0 = key[round(float(c[0]), 12)]
        assert abs(e0 - e[0]) < 1e-12 and r0 == l[0]
    cir, err, label, room = D.load_pkl_data(str(path), "obstacle_full", np.random.RandomState(1))
    assert len(cir) == 160 and set(np.unique(label)) == {0.0, 1.0, 2.0, 3.0}          # 4 of the 5 obstacle codes, 40 rows each
    with pytest.raises(NotImplementedError):
        D.load_pkl_data(str(path), "paper")


def test_splits_and_scaler_match_sklearn():
    from sklearn.preprocessing import StandardScaler as SK
    rng = np.random.RandomState(0)
    cir = rng.randn(500, 157) * rng.rand(157) * 3 + rng.randn(157)
    cir[:, 5] = 2.5                                                     # zero-variance feature
    err, label = rng.rand(500, 1), rng.randint(0, 5, (500, 1)).astype(float)
    train, test, _, _ = D.err_mitigation_dataset(None, data=(cir, err, label), split_factor=0.8, scaling=True, mode="full")
    assert train[0].shape == (400, 157) and test[0].shape == (100, 157) and train[1].shape == (400, 1)
    sk = SK().fit(cir[:400])
    np.testing.assert_allclose(train[0], sk.transform(cir[:400]), rtol=0, atol=1e-12)
    np.testing.assert_allclose(test[0], sk.transform(cir[400:]), rtol=0, atol=1e-12)
    train, test, _, _ = D.err_mitigation_dataset(None, data=(cir, err, label), mode="paper")
    assert (test[2] == 2).all() and not (train[2] == 2).any() and len(train[0]) + len(test[0]) == 500


def test_dataset_items_and_ring_cover_every_sample_once():
    rng = np.random.RandomState(3)
    cir = np.arange(1030, dtype=np.float64)[:, None] + np.zeros((1, 157))
    ds = D.UWBDataset((cir, rng.rand(1030, 1), rng.randint(0, 5, (1030, 1))))
    item = ds[7]
    assert item["CIR"].shape == (157,) and item["CIR"].dtype == torch.float32 and item["Err"].shape == (1,) and item["Label"].shape == (1,)
    ring = D.PinnedBatchRing(ds, 256, shuffle=True, seed=5, pin=False)
    assert len(ring) == 5 and ring.batch_size == 256
    for epoch in range(2):
        seen, sizes = [], []
        for batch in ring:
            assert batch["CIR"].dtype == torch.float32 and batch["Err"].shape == (batch["CIR"].shape[0], 1)
            seen.append(batch["CIR"][:, 0].clone())           # clone: the ring recycles its buffers
            sizes.append(batch["CIR"].shape[0])
        assert sizes == [256, 256, 256, 256, 6]
        assert sorted(torch.cat(seen).tolist()) == list(range(1030))
    a = [b["CIR"][:, 0].clone() for b in D.PinnedBatchRing(ds, 256, shuffle=True, seed=5, pin=False)]
    assert not torch.equal(torch.cat(a), torch.arange(1030.0)), "shuffle=True must permute"
