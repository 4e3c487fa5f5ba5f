"""Order-independent 64-bit checksum of a sparse float32 volume (test infrastructure).

hash = sum over non-zero voxels of splitmix64((index << 32) | float32_bits)  (mod 2^64).
Used to commit compact golden vectors for whole dose/beam volumes: two volumes hash
equal iff (up to 2^-64 collisions) they have the same non-zero voxels with the same bits.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def volume_hash(indices, values) -> np.uint64:
    idx = np.asarray(indices).astype(np.uint64)
    bits = np.ascontiguousarray(values, dtype=np.float32).view(np.uint32).astype(np.uint64)
    keep = bits != 0
    key = (idx[keep] << np.uint64(32)) | bits[keep]
    with np.errstate(over="ignore"):
        return np.uint64(np.sum(_splitmix64(key), dtype=np.uint64))


def dense_hash(volume) -> np.uint64:
    flat = np.ascontiguousarray(volume, dtype=np.float32).reshape(-1)
    nz = np.flatnonzero(flat.view(np.uint32))
    return volume_hash(nz, flat[nz])


def batch_hash(idx, w, count):
    """Per-row volume_hash of padded sparse traces (idx, w: [m, cap]; count: [m]) -> (nnz, hash) arrays."""
    idx = np.asarray(idx)
    bits = np.ascontiguousarray(w, dtype=np.float32).view(np.uint32).astype(np.uint64)
    valid = (np.arange(idx.shape[1])[None, :] < np.asarray(count)[:, None]) & (bits != 0)
    key = (idx.astype(np.uint64) << np.uint64(32)) | bits
    mixed = np.where(valid, _splitmix64(key), np.uint64(0))
    with np.errstate(over="ignore"):
        return valid.sum(axis=1).astype(np.int32), np.sum(mixed, axis=1, dtype=np.uint64)
