"""The lungs mask and the tumour table the environment runs on.

The reference keeps `data/lungs.npy` and 1000 dense `data/tumours/*.npy` volumes
(environment.py:28-29, 90-97; 807 MB).  `data/phantom.npz` is the same information
packed by tools/pack_phantom.py (0.17 MB): a lungs bitmask and, per tumour, the sorted
voxel list plus the per-tumour constants the reference recomputes every step.
"""
import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _native

DEFAULT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "phantom.npz")


class Phantom:
    def __init__(self, grid, lungs_bits, vox_offsets, vox, centroid, tumour_sum, lung_mask_sum,
                 names: Optional[Sequence[str]] = None, meta=None):
        self.grid = np.ascontiguousarray(grid, dtype=np.int32)
        self.nvox = int(np.prod(self.grid))
        self.lungs_bits = np.ascontiguousarray(lungs_bits, dtype=np.uint32)
        self.vox_offsets = np.ascontiguousarray(vox_offsets, dtype=np.int32)
        self.vox = np.ascontiguousarray(vox, dtype=np.int32)
        self.centroid = np.ascontiguousarray(centroid, dtype=np.float64).reshape(-1, 3)
        self.tumour_sum = np.ascontiguousarray(tumour_sum, dtype=np.float32)
        self.lung_mask_sum = np.ascontiguousarray(lung_mask_sum, dtype=np.float32)
        self.n_tumours = len(self.vox_offsets) - 1
        self.names = list(names) if names is not None else [f"tumour_{i}" for i in range(self.n_tumours)]
        self.meta = None if meta is None else np.asarray(meta, dtype=np.float32)
        if self.lungs_bits.size != (self.nvox + 31) // 32:
            raise ValueError("lungs_bits has the wrong length for the grid")
        for arr in (self.centroid, self.tumour_sum, self.lung_mask_sum):
            if arr.shape[0] != self.n_tumours:
                raise ValueError("per-tumour arrays disagree on the number of tumours")

    @classmethod
    def load(cls, path: str = DEFAULT_PATH) -> "Phantom":
        z = np.load(path)
        return cls(z["grid"], z["lungs_bits"], z["vox_offsets"], z["vox"], z["centroid"], z["tumour_sum"],
                   z["lung_mask_sum"], names=[str(x) for x in z["names"]], meta=z["meta"])

    @classmethod
    def from_volumes(cls, lungs: np.ndarray, tumours: Sequence[np.ndarray], names=None) -> "Phantom":
        """Build the table from dense volumes with the reference's own expressions
        (environment.py:145-148, 167, 174-178)."""
        lungs_f = np.asarray(lungs).astype(np.float32)
        grid = np.array(lungs_f.shape, dtype=np.int32)
        nvox = lungs_f.size
        padded = np.zeros(((nvox + 31) // 32) * 32, dtype=np.uint8)
        padded[:nvox] = (lungs_f.reshape(-1) != 0).astype(np.uint8)
        bits = np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
        offsets, vox, cen, tsum, msum = [0], [], [], [], []
        for t in tumours:
            t = np.clip(np.asarray(t).astype(np.float32), 0.0, 1.0)
            if t.shape != lungs_f.shape:
                raise ValueError("tumour volume shape differs from the lungs volume")
            lin = np.flatnonzero(t.reshape(-1) == 1.0).astype(np.int32)
            if lin.size == 0:
                raise ValueError("empty tumour volume")
            vox.append(lin)
            offsets.append(offsets[-1] + lin.size)
            cen.append(np.mean(np.stack(np.where(t == 1.0), axis=-1), axis=0))
            tsum.append(np.sum(t))
            msum.append(np.sum(lungs_f * (1 - t)))
        return cls(grid, bits, offsets, np.concatenate(vox), np.array(cen), tsum, msum, names=names)

    def lungs_volume(self) -> np.ndarray:
        flat = np.unpackbits(self.lungs_bits.view(np.uint8), bitorder="little")[: self.nvox]
        return flat.astype(np.float32).reshape(tuple(int(g) for g in self.grid))

    def tumour_voxels(self, tid: int) -> np.ndarray:
        return self.vox[self.vox_offsets[tid]: self.vox_offsets[tid + 1]]

    def tumour_volume(self, tid: int) -> np.ndarray:
        v = np.zeros(self.nvox, dtype=np.float32)
        v[self.tumour_voxels(tid)] = 1.0
        return v.reshape(tuple(int(g) for g in self.grid))

    def desc(self) -> "_native.PhantomDesc":
        """rt_phantom_desc over this object's arrays (which must outlive the call)."""
        d = _native.PhantomDesc()
        d.grid = (C.c_int32 * 3)(*[int(g) for g in self.grid])
        d.lungs_bits = self.lungs_bits.ctypes.data_as(C.POINTER(C.c_uint32))
        d.n_tumours = self.n_tumours
        d.vox_offsets = self.vox_offsets.ctypes.data_as(C.POINTER(C.c_int32))
        d.vox = self.vox.ctypes.data_as(C.POINTER(C.c_int32))
        d.centroid = self.centroid.ctypes.data_as(C.POINTER(C.c_double))
        d.tumour_sum = self.tumour_sum.ctypes.data_as(C.POINTER(C.c_float))
        d.lung_mask_sum = self.lung_mask_sum.ctypes.data_as(C.POINTER(C.c_float))
        return d


_default = None


def default_phantom() -> Phantom:
    global _default
    if _default is None:
        _default = Phantom.load()
    return _default
