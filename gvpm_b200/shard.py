"""Image-tile sharding of the gather across GPUs.

Every camera ray's gather reads the whole photon set and writes only its own 27 floats
(gvpm.cpp:1008-1069), so the unit of partition is the reference's 32x32 gather block
(gvpm.cpp:271-290): blocks are dealt round-robin to the ranks, the photon set is replicated
(broadcast), and the per-ray results are gathered to rank 0 and scattered back to pixel order.
"""
import numpy as np

BLOCK = 32


def tile_owner(px, py, width, world, block=BLOCK):
    """Rank owning each ray: block index (row-major over the block grid) modulo world size."""
    tiles_x = (width + block - 1) // block
    return ((np.asarray(py) // block) * tiles_x + (np.asarray(px) // block)) % world


def local_indices(px, py, width, world, rank, block=BLOCK):
    return np.nonzero(tile_owner(px, py, width, world, block) == rank)[0]


def assemble(parts, index_lists, n_total, width=27):
    """Scatter per-rank result rows back to the global ray order.  parts[r]: [>=len(idx_r), width]."""
    out = np.zeros((n_total, width), dtype=np.float32)
    for part, idx in zip(parts, index_lists):
        out[idx] = np.asarray(part)[:len(idx)]
    return out


def to_image(ray_out, px, py, w, h):
    """Sum the rays of each pixel (several medium edges may share one), gvpm.cpp:1018-1052."""
    img = np.zeros((h, w, ray_out.shape[1]), dtype=np.float32)
    np.add.at(img, (np.asarray(py), np.asarray(px)), ray_out)
    return img
