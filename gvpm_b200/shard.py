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


def band_owner(px, py, width, height, world, cycles=2, block=BLOCK):
    """Rank owning each ray under the column-band partition: the gather blocks, numbered column-major, are cut into
    cycles*world runs of (almost) equal block count - vertical bands of the image - and run j goes to rank
    j % world.  A rank's rays then cross only `cycles` wedges of the scene, which is what lets
    gvpm_build_points_for_rays leave most of the photon set out of that rank's hierarchy, while the cyclic
    deal pairs a band near the image centre with one near the edge (load balance under a centre-weighted
    photon density).  cycles*world == number of blocks degenerates into tile_owner's round-robin deal."""
    tiles_x = (width + block - 1) // block
    tiles_y = (height + block - 1) // block
    n_tiles = tiles_x * tiles_y
    runs = max(1, min(n_tiles, cycles * world))
    t = (np.asarray(px) // block).astype(np.int64) * tiles_y + (np.asarray(py) // block)
    return (t * runs // n_tiles) % world


def band_indices(px, py, width, height, world, rank, cycles=2, block=BLOCK):
    return np.nonzero(band_owner(px, py, width, height, world, cycles, block) == rank)[0]


def block_index(px, py, height, block=BLOCK):
    """column-major index of each ray's gather block (the order band_owner cuts into runs)"""
    tiles_y = (height + block - 1) // block
    return (np.asarray(px) // block).astype(np.int64) * tiles_y + (np.asarray(py) // block)


def band_owner_weighted(px, py, width, height, world, block_cost, cycles=2, block=BLOCK):
    """band_owner with runs of (almost) equal COST instead of equal block count: block_cost[t] is what gather block t
    (column-major) cost in the previous iteration (contributing pairs + a per-ray term).  The photon density of a
    rendered scene is far from uniform over the image, and a sharded iteration is as slow as its busiest rank."""
    tiles_x = (width + block - 1) // block
    tiles_y = (height + block - 1) // block
    n_tiles = tiles_x * tiles_y
    cost = np.maximum(np.asarray(block_cost, dtype=np.float64).reshape(-1)[:n_tiles], 0.0)
    assert cost.size == n_tiles
    runs = max(1, min(n_tiles, cycles * world))
    total = cost.sum()
    if not total > 0:
        return band_owner(px, py, width, height, world, cycles, block)
    before = np.cumsum(cost) - cost                      # cost in front of each block
    run_of_block = np.minimum((before / total * runs).astype(np.int64), runs - 1)
    return run_of_block[block_index(px, py, height, block)] % world


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
