#!/usr/bin/env python
"""Pack the reference's phantom data into the compact table the CUDA path loads.

Input  (reference layout, environment.py:28-29,90-97):
    <data>/lungs.npy            bool  (67,43,70)
    <data>/tumours/X_Y_Z_R.npy  f32   (67,43,70), values in {0,1}, one file per tumour
Output (one .npz, ~0.5 MB instead of 807 MB of dense volumes):
    grid            int32[3]      volume shape G (C order, axis 2 contiguous)
    lungs_bits      uint32[W]     bit i of word i>>5 == lungs.flat[i]
    names           str[K]        tumour file names, SORTED -> tumour id (the reference
                                  uses unsorted os.listdir order, environment.py:28,
                                  which is not reproducible across machines)
    vox_offsets     int32[K+1]    CSR offsets into vox
    vox             int32[sum]    ascending linear voxel indices with tumour == 1.0
    centroid        float64[K,3]  np.mean of the voxel index triples (environment.py:145-148)
    tumour_sum      float32[K]    np.sum(tumours)                    (environment.py:167)
    lung_mask_sum   float32[K]    np.sum(lungs * (1 - tumours))      (environment.py:174,178)
    meta            float32[K,4]  (x, y, z, radius) parsed from the name (environment.py:91-93)

Every derived number is produced by the same NumPy expression the reference
evaluates, so the table holds the reference's own values.
"""
import argparse
import os

import numpy as np


def pack(data_dir: str, out_path: str) -> None:
    lungs_bool = np.load(os.path.join(data_dir, "lungs.npy"))
    lungs = lungs_bool.astype(np.float32)
    grid = np.array(lungs.shape, dtype=np.int32)
    nvox = int(lungs.size)

    flat = lungs_bool.reshape(-1).astype(np.uint8)
    padded = np.zeros(((nvox + 31) // 32) * 32, dtype=np.uint8)
    padded[:nvox] = flat
    lungs_bits = np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little")
    lungs_bits = lungs_bits.view(np.uint32).reshape(-1)

    tdir = os.path.join(data_dir, "tumours")
    names = sorted(x for x in os.listdir(tdir) if x.endswith(".npy"))
    offsets = [0]
    vox, centroid, tsum, msum, meta = [], [], [], [], []
    for name in names:
        t = np.load(os.path.join(tdir, name)).astype(np.float32)
        t = np.clip(t, 0.0, 1.0)
        assert t.shape == lungs.shape
        assert set(np.unique(t).tolist()) <= {0.0, 1.0}, name
        where = np.stack(np.where(t == 1.0), axis=-1)
        centroid.append(np.mean(where, axis=0))
        lin = np.flatnonzero(t.reshape(-1) == 1.0).astype(np.int32)
        vox.append(lin)
        offsets.append(offsets[-1] + lin.size)
        tsum.append(np.sum(t))
        msum.append(np.sum(lungs * (1 - t)))
        attrs = name.split(".npy")[0].split("_")
        meta.append([float(a) for a in attrs[:4]])

    np.savez_compressed(
        out_path,
        grid=grid,
        lungs_bits=lungs_bits,
        names=np.array(names),
        vox_offsets=np.array(offsets, dtype=np.int32),
        vox=np.concatenate(vox).astype(np.int32),
        centroid=np.array(centroid, dtype=np.float64),
        tumour_sum=np.array(tsum, dtype=np.float32),
        lung_mask_sum=np.array(msum, dtype=np.float32),
        meta=np.array(meta, dtype=np.float32),
    )
    print(f"packed {len(names)} tumours, {offsets[-1]} voxels -> {out_path} "
          f"({os.path.getsize(out_path)} bytes)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="/root/reference/data")
    ap.add_argument("--out", default=os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
        "ppo-radiotherapy_b200", "data", "phantom.npz"))
    args = ap.parse_args()
    pack(args.data, args.out)
