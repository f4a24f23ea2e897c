"""Numpy models of two pieces of device logic the sharded step relies on, checked independently of the GPU:

* ``slim_topk`` (seesaw_b200/csrc/ssw_scan.cu): the k best of n unique 64-bit keys by an 8-pass MSB radix select
  (``merge_pick_bin`` picks the bin holding the remaining-th largest key), then a rank of the <= k survivors by counting;
  key 0 = empty slot.  Both uses are modelled: a shard's compacted candidates and the world's [world, k] lists.
* ``part_build_kernel``: the image partition of a scan grid other than the database's own — range g starts at the first
  image whose first row is >= n_rows * g / ranges — which must be image-aligned, cover every row once and equal the
  host rule of ``build_layout`` (ssw_capi.cu)."""
import numpy as np
import pytest


def pick_bin(hist, remaining):
    """merge_pick_bin: scanning bins 255 .. 0, the bin where the running count first reaches `remaining`; returns
    (bin, what is still to take inside that bin)."""
    seen = 0
    for b in range(255, -1, -1):
        if seen < remaining <= seen + hist[b]:
            return b, remaining - seen
        seen += hist[b]
    raise AssertionError("fewer keys than remaining")


def slim_topk(keys, k):
    """-> the best min(n_valid, k) keys, best first, padded with 0 (the kernel's S.ok)."""
    keys = np.asarray(keys, dtype=np.uint64)
    valid = keys[keys != 0]
    threshold = np.uint64(1)                       # fewer than k valid keys: keep them all
    if len(valid) > k:
        prefix, remaining = 0, k
        for d in range(7, -1, -1):
            match = valid if d == 7 else valid[(valid >> np.uint64(8 * (d + 1))) == np.uint64(prefix >> (8 * (d + 1)))]
            hist = np.bincount(((match >> np.uint64(8 * d)) & np.uint64(255)).astype(np.int64), minlength=256)
            b, remaining = pick_bin(hist, remaining)
            prefix |= b << (8 * d)
        threshold = np.uint64(prefix)              # the k-th largest key (keys are unique)
    survivors = keys[(keys != 0) & (keys >= threshold)]
    assert len(survivors) <= 64
    out = np.zeros(k, np.uint64)
    for key in survivors:                          # rank by counting the larger ones: the rank is the output slot
        rank = int((survivors > key).sum())
        if rank < k:
            out[rank] = key
    return out


def make_keys(rng, n, n_empty, clustered):
    """unique keys shaped like the scan's: (order-preserving score bits << 32) | ~row"""
    rows = rng.choice(1 << 24, size=n, replace=False).astype(np.uint64)
    score = rng.integers(0x3F000000, 0x3F000040 if clustered else 0x3F800000, size=n).astype(np.uint64)
    keys = (score << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - rows)
    keys[rng.choice(n, size=min(n_empty, n), replace=False)] = 0
    return keys


@pytest.mark.parametrize("clustered", [False, True])
def test_slim_topk_selects_and_orders_like_a_sort(clustered):
    rng = np.random.default_rng(5)
    for n, k, n_empty in [(400, 50, 0), (400, 50, 390), (37, 50, 3), (1, 1, 0), (64, 64, 0), (700, 3, 100), (400, 50, 350),
                          (51, 50, 0), (8 * 50, 17, 40)]:
        keys = make_keys(rng, n, n_empty, clustered)
        want = np.sort(keys[keys != 0])[::-1][:k]
        got = slim_topk(keys, k)
        assert (got[:len(want)] == want).all() and (got[len(want):] == 0).all(), (n, k, n_empty)


def test_world_merge_of_shard_lists_equals_the_global_topk():
    """phase 4 of the exchange: every shard's own top-k (padded with empty slots), merged, is the global top-k"""
    rng = np.random.default_rng(6)
    for world, k, per_shard in [(8, 50, 400), (2, 3, 10), (4, 64, 30), (8, 50, 20)]:
        allkeys = make_keys(rng, world * per_shard, 0, False).reshape(world, per_shard)
        lists = np.stack([slim_topk(allkeys[r], k) for r in range(world)])          # [world, k], what the ranks exchange
        got = slim_topk(lists.reshape(-1), k)
        want = np.sort(allkeys.reshape(-1))[::-1][:k]
        assert (got == want).all()


def test_groups_push_everything_before_they_wait():
    """A group serves queries first, first + stride, ...: it raises the flags of ALL of them (phase 1) before it waits
    for any (phase 2).  With the same assignment on every rank, replaying the ranks in any interleaving completes: a wait
    only ever depends on phase 1 of the same group index on another rank."""
    nq, blocks, groups = 64, 4, 8
    stride = blocks * groups
    served = sorted(q for g in range(stride) for q in range(g, nq, stride))
    assert served == list(range(nq))                                     # every query has exactly one group
    world = 3
    flags = np.zeros((world, world, nq), bool)                           # flags[dst, src, q]
    done = np.zeros((world, stride), bool)
    rng = np.random.default_rng(7)
    pushed = np.zeros((world, stride), bool)
    for _ in range(10 * world * stride):
        r, g = int(rng.integers(world)), int(rng.integers(stride))
        if not pushed[r, g]:
            for q in range(g, nq, stride):
                flags[:, r, q] = True                                    # peer stores + release flags
            pushed[r, g] = True
        elif not done[r, g] and all(flags[r, :, q].all() for q in range(g, nq, stride)):
            done[r, g] = True
        if done.all():
            break
    assert pushed.all() and done.all()


def host_partition(row_ptr, n_rows, ranges):
    """build_layout (ssw_capi.cu): lower_bound over row_ptr[0 .. n_images], clamped, ends pinned"""
    n_images = len(row_ptr) - 1
    part = np.empty(ranges + 1, np.int64)
    for g in range(ranges + 1):
        part[g] = min(int(np.searchsorted(row_ptr, n_rows * g // ranges, side="left")), n_images)
    part[0], part[ranges] = 0, n_images
    return part


def device_partition(row_ptr, n_rows, ranges):
    """part_build_kernel: the same rule as an explicit binary search"""
    n_images = len(row_ptr) - 1
    part = np.empty(ranges + 1, np.int64)
    for g in range(ranges + 1):
        target = n_rows * g // ranges
        lo, hi = 0, n_images + 1
        while lo < hi:
            mid = (lo + hi) >> 1
            if row_ptr[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        lo = min(lo, n_images)
        part[g] = 0 if g == 0 else (n_images if g == ranges else lo)
    return part


@pytest.mark.parametrize("grid", [144, 146, 148, 64])
def test_side_grid_partition_is_image_aligned_and_complete(grid):
    rng = np.random.default_rng(8)
    for counts in (rng.integers(1, 61, size=5000), np.full(31250, 40), rng.integers(1, 4, size=100), np.array([7]),
                   np.zeros(0, np.int64)):
        row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        n_rows, ranges = int(row_ptr[-1]), grid * 8
        part = device_partition(row_ptr, n_rows, ranges)
        assert (part == host_partition(row_ptr, n_rows, ranges)).all()
        assert part[0] == 0 and part[-1] == len(counts) and (np.diff(part) >= 0).all()
        rows_per_cta = np.diff(row_ptr[part[::8]])                       # K2 uses a CTA's 8 ranges as one
        assert rows_per_cta.sum() == n_rows
        if len(counts) > ranges:                                         # balanced to within one image per range end
            assert rows_per_cta.max() - rows_per_cta.min() <= 2 * counts.max() + 8
