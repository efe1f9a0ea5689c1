"""Input side of the IF-Net trainer: the reference's on-disk formats and its ``ImplicitDataset``
(dataset/implicit_dataset.py:10-58, data_processing/volume_reader.py:36-52) behind the same interface, written for a
GPU step that takes 8 ms: once the step is that short, ``np.load`` + per-point Python list handling on the trainer's
main process (the reference's ``num_workers=0`` default) is the throughput limit.

* ``read_df``: the ``.df`` distance-field format (3 x uint64 dims, float32 voxels in Fortran order) through one
  ``np.fromfile`` instead of ``struct.unpack`` of dimX*dimY*dimZ Python floats (a 139x104x112 grid: 6.5 MB, 1.6 M
  Python objects in the reference).
* ``ImplicitDataset``: same constructor, same item dict (name / grid / points / input / occupancies / target), the same
  ``np.random.randint`` draws in the same order (so a seeded run picks the same samples); the sub-sampling is one
  fancy-index per array instead of ``list.extend`` of 2 x num_points rows.
* ``collate_pinned`` + ``batches``: batches assembled directly in pinned host memory by a background thread and handed
  to ``HostPrefetcher`` (H2D copy of batch i+1 on a side stream while batch i computes)."""
from __future__ import annotations

import queue
import threading
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

SIGMAS = ("0.10", "0.01")     # implicit_dataset.py:36: boundary samples at two noise levels


def read_df(filename, scale_factor: int = 1) -> np.ndarray:
    """volume_reader.py:36-46: (dimX, dimY, dimZ) float32, file order = Fortran order; ``scale_factor`` > 1 block-averages
    (``skimage.measure.block_reduce(df, (f, f, f), np.mean)``: zero padding up to a multiple of f, mean over f^3)."""
    with open(filename, "rb") as fh:
        dims = np.fromfile(fh, dtype=np.uint64, count=3)
        if dims.size != 3:
            raise ValueError(f"{filename}: truncated .df header")
        n = int(dims[0]) * int(dims[1]) * int(dims[2])
        vox = np.fromfile(fh, dtype=np.float32, count=n)
    if vox.size != n:
        raise ValueError(f"{filename}: expected {n} voxels, found {vox.size}")
    df = vox.reshape([int(dims[0]), int(dims[1]), int(dims[2])], order="F")
    if scale_factor != 1:
        f = int(scale_factor)
        pad = [(0, (-s) % f) for s in df.shape]
        df = np.pad(df, pad, mode="constant", constant_values=0)
        sx, sy, sz = (s // f for s in df.shape)
        df = df.reshape(sx, f, sy, f, sz, f).mean(axis=(1, 3, 5), dtype=np.float64).astype(np.float32)
    return df


def write_df(filename, df: np.ndarray) -> None:
    """Inverse of ``read_df`` (tests, synthetic data)."""
    df = np.asarray(df, dtype=np.float32)
    with open(filename, "wb") as fh:
        np.asarray(df.shape, dtype=np.uint64).tofile(fh)
        df.reshape(-1, order="F").tofile(fh)


class ImplicitDataset(Dataset):
    """dataset/implicit_dataset.py:10-58 with the same arguments and item layout.  ``splits_root`` is the reference's
    hard-coded ``data/splits`` (relative to the working directory) unless given."""

    def __init__(self, split, dataset_path, num_points, splitsdir, splits_root: Optional[str] = None):
        self.dataset_path = Path(dataset_path)
        self.split = split
        self.splitsdir = splitsdir
        root = Path(splits_root) if splits_root is not None else Path("data/splits")
        self.split_shapes = [x.strip() for x in (root / splitsdir / f"{split}.txt").read_text().split("\n") if x.strip() != ""]
        self.data = [x for x in self.split_shapes]
        self.data = self.data * (50 if ("overfit" in splitsdir) and split == "train" else 1)
        self.num_points = num_points

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        item = self.data[idx]
        folder = self.dataset_path / "processed" / self.splitsdir / item
        sample_input = torch.from_numpy(np.load(folder / "depth_grid.npz")["grid"]).float()
        sample_target = torch.from_numpy(read_df(str(folder / "target.df"))).float()
        points, grids, occupancies = [], [], []
        for sigma in SIGMAS:
            with np.load(folder / f"occupancy_{sigma}.npz") as z:
                pts, coords, occ = z["points"], z["grid_coords"], z["occupancies"]
            idx_s = np.random.randint(0, pts.shape[0], self.num_points)       # the reference's draw, same order
            points.append(pts[idx_s])
            grids.append(coords[idx_s])
            occupancies.append(occ[idx_s])
        return {
            "name": item,
            "grid": torch.from_numpy(np.concatenate(grids).astype(np.float32, copy=False)),
            "points": torch.from_numpy(np.concatenate(points).astype(np.float32, copy=False)),
            "input": sample_input.unsqueeze(0),
            "occupancies": torch.from_numpy(np.concatenate(occupancies).astype(np.float32, copy=False)),
            "target": sample_target.unsqueeze(0),
        }


def collate_pinned(items: Sequence[Dict], pin: bool = True) -> Dict:
    """Default-collate semantics (tensors stacked on a new batch axis, names as a list) with the stacked tensors
    allocated in PINNED host memory, so that the H2D copy is asynchronous and runs at link speed."""
    out: Dict = {}
    for k in items[0]:
        v0 = items[0][k]
        if torch.is_tensor(v0):
            buf = torch.empty((len(items),) + tuple(v0.shape), dtype=v0.dtype, pin_memory=pin and torch.cuda.is_available())
            for i, it in enumerate(items):
                buf[i].copy_(it[k])
            out[k] = buf
        else:
            out[k] = [it[k] for it in items]
    return out


def batches(dataset: Dataset, batch_size: int, device=None, keys: Sequence[str] = ("input", "points", "occupancies"),
            shuffle: bool = True, drop_last: bool = True, depth: int = 2, seed: Optional[int] = None) -> Iterator[Dict]:
    """One epoch of batches.  A background thread reads and collates (``collate_pinned``) up to ``depth`` batches ahead;
    with a CUDA ``device`` the tensors named in ``keys`` are staged through ``HostPrefetcher`` (the copy of batch i+1
    overlaps the step on batch i) and yielded on the device, everything else stays on the host."""
    order = np.arange(len(dataset))
    if shuffle:
        np.random.default_rng(seed).shuffle(order)
    chunks: List[np.ndarray] = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    if drop_last and chunks and len(chunks[-1]) < batch_size:
        chunks.pop()
    q: "queue.Queue" = queue.Queue(maxsize=max(depth, 1))

    def reader():
        try:
            for ch in chunks:
                q.put(collate_pinned([dataset[int(i)] for i in ch]))
            q.put(None)
        except BaseException as e:      # noqa: BLE001  (surface loader errors in the consumer)
            q.put(e)

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    use_dev = device is not None and torch.device(device).type == "cuda"
    pf = None
    if use_dev:
        from .prefetch import HostPrefetcher
        pf = HostPrefetcher(torch.device(device))

    def nxt():
        b = q.get()
        if isinstance(b, BaseException):
            raise b
        return b

    cur = nxt()
    handle = pf.issue(tuple(cur[k] for k in keys)) if (pf is not None and cur is not None) else None
    while cur is not None:
        following = nxt()
        out = dict(cur)
        if pf is not None:
            dev_t = pf.wait(handle)
            for k, t in zip(keys, dev_t):
                out[k] = t
            handle = pf.issue(tuple(following[k] for k in keys)) if following is not None else None
        yield out
        cur = following
    th.join()
