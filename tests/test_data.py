"""svr_b200.data: the reference's on-disk formats and ImplicitDataset (dataset/implicit_dataset.py:10-58,
data_processing/volume_reader.py:36-52) behind the same interface.  The reference modules cannot be imported here
(skimage / trimesh are missing), so the expected values are restated in the test exactly as the reference computes them:
struct-unpacked .df voxels in Fortran order, np.random.randint sub-sampling per noise level in file order."""
import struct

import numpy as np
import pytest
import torch

import svr_b200
from svr_b200 import data as D


def _make_sample(folder, rng, dims=(7, 5, 6), n_samples=300):
    folder.mkdir(parents=True)
    grid = rng.random(dims, dtype=np.float32)
    np.savez(folder / "depth_grid.npz", grid=grid)
    target = rng.random(dims, dtype=np.float32) * 3
    D.write_df(folder / "target.df", target)
    occ = {}
    for sigma in ("0.10", "0.01"):
        occ[sigma] = {"points": rng.random((n_samples, 3), dtype=np.float32) - 0.5,
                      "grid_coords": rng.random((n_samples, 3), dtype=np.float32) * 2 - 1,
                      "occupancies": (rng.random(n_samples) < 0.5).astype(np.float32)}
        np.savez(folder / f"occupancy_{sigma}.npz", **occ[sigma])
    return grid, target, occ


def _read_df_reference_way(path):
    with open(path, "rb") as fh:
        dx, dy, dz = struct.unpack("QQQ", fh.read(24))
        vals = struct.unpack("f" * (dx * dy * dz), fh.read(4 * dx * dy * dz))
    return np.array(vals, dtype=np.float32).reshape([dx, dy, dz], order="F")


def test_read_df_matches_the_struct_reader_and_block_means(tmp_path):
    rng = np.random.default_rng(0)
    df = rng.random((9, 6, 7), dtype=np.float32)
    D.write_df(tmp_path / "a.df", df)
    assert np.array_equal(D.read_df(tmp_path / "a.df"), df)
    assert np.array_equal(D.read_df(tmp_path / "a.df"), _read_df_reference_way(tmp_path / "a.df"))
    half = D.read_df(tmp_path / "a.df", 2)                       # block_reduce(df, (2,2,2), np.mean): zero padding to even sizes
    pad = np.pad(df, [(0, 1), (0, 0), (0, 1)])
    want = pad.reshape(5, 2, 3, 2, 4, 2).mean(axis=(1, 3, 5))
    assert half.shape == (5, 3, 4) and np.allclose(half, want, atol=1e-6)
    (tmp_path / "bad.df").write_bytes(struct.pack("QQQ", 2, 2, 2) + struct.pack("fff", 1.0, 2.0, 3.0))     # 3 of 8 voxels
    with pytest.raises(ValueError):
        D.read_df(tmp_path / "bad.df")


def test_dataset_items_equal_the_reference_recipe(tmp_path):
    rng = np.random.default_rng(1)
    names = ["00000", "00001", "00002"]
    truth = {n: _make_sample(tmp_path / "data" / "processed" / "overfit" / n, rng) for n in names}
    splits = tmp_path / "splits" / "overfit"
    splits.mkdir(parents=True)
    (splits / "train.txt").write_text("\n".join(names) + "\n\n")
    (splits / "val.txt").write_text(names[0] + "\n")
    ds = D.ImplicitDataset("train", tmp_path / "data", 40, "overfit", splits_root=tmp_path / "splits")
    assert len(ds) == 150 and len(D.ImplicitDataset("val", tmp_path / "data", 40, "overfit", splits_root=tmp_path / "splits")) == 1
    np.random.seed(7)
    item = ds[4]                                               # 4 % 3 -> "00001"
    grid, target, occ = truth["00001"]
    np.random.seed(7)
    pts, crd, oc = [], [], []
    for sigma in ("0.10", "0.01"):                             # implicit_dataset.py:36-44
        idx = np.random.randint(0, occ[sigma]["points"].shape[0], 40)
        pts.extend(occ[sigma]["points"][idx])
        crd.extend(occ[sigma]["grid_coords"][idx])
        oc.extend(occ[sigma]["occupancies"][idx])
    assert item["name"] == "00001"
    assert torch.equal(item["points"], torch.from_numpy(np.array(pts, dtype=np.float32)))
    assert torch.equal(item["grid"], torch.from_numpy(np.array(crd, dtype=np.float32)))
    assert torch.equal(item["occupancies"], torch.from_numpy(np.array(oc, dtype=np.float32)))
    assert torch.equal(item["input"], torch.from_numpy(grid)[None]) and torch.equal(item["target"], torch.from_numpy(target)[None])
    assert item["points"].shape == (80, 3) and item["occupancies"].shape == (80,)


def test_collate_and_host_batches(tmp_path):
    rng = np.random.default_rng(2)
    names = [f"{i:05d}" for i in range(5)]
    for n in names:
        _make_sample(tmp_path / "data" / "processed" / "s" / n, rng)
    (tmp_path / "splits" / "s").mkdir(parents=True)
    (tmp_path / "splits" / "s" / "train.txt").write_text("\n".join(names))
    ds = D.ImplicitDataset("train", tmp_path / "data", 16, "s", splits_root=tmp_path / "splits")
    b = D.collate_pinned([ds[0], ds[1]], pin=False)
    assert b["input"].shape == (2, 1, 7, 5, 6) and b["points"].shape == (2, 32, 3) and b["name"] == ["00000", "00001"]
    got = list(D.batches(ds, 2, device=None, shuffle=False))
    assert len(got) == 2 and [g["name"] for g in got] == [["00000", "00001"], ["00002", "00003"]]
    assert len(list(D.batches(ds, 2, device=None, shuffle=True, drop_last=False, seed=3))) == 3


@pytest.mark.gpu
def test_device_batches_arrive_intact(tmp_path):
    rng = np.random.default_rng(3)
    names = [f"{i:05d}" for i in range(6)]
    for n in names:
        _make_sample(tmp_path / "data" / "processed" / "s" / n, rng)
    (tmp_path / "splits" / "s").mkdir(parents=True)
    (tmp_path / "splits" / "s" / "train.txt").write_text("\n".join(names))
    ds = D.ImplicitDataset("train", tmp_path / "data", 16, "s", splits_root=tmp_path / "splits")
    np.random.seed(11)
    host = list(D.batches(ds, 2, device=None, shuffle=False))
    np.random.seed(11)
    n = 0
    for hb, db in zip(host, D.batches(ds, 2, device="cuda", shuffle=False)):
        assert db["input"].is_cuda and db["points"].is_cuda and not db["grid"].is_cuda
        for k in ("input", "points", "occupancies"):
            assert torch.equal(db[k].cpu(), hb[k])              # (staging slots are reused: compare before the next batch)
        n += 1
    assert n == 3
