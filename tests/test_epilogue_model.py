"""A numpy model of the batched-scan epilogue's segment walk (seesaw_b200/csrc/ssw_scan_tc.cu: scan_tc8_group /
scan_tc8_boundary), checked against a direct per-image argmax.  It pins the bit logic the kernel relies on — the
funnel-shifted boundary words of an unaligned tile, the per-thread column ownership of a 16x256b tensor-memory load
(thread j of a quad holds columns 8i+2j, 8i+2j+1), the segment masks, the argmax-tree index -> column mapping and
the (score desc, column asc) quad reduction — independently of the GPU."""
import numpy as np


def argmax8(x):
    """depth-3 tree with strict '>' : the LOWEST index among equal maxima (mirrors argmax8 in the kernel)"""
    idx = list(range(8))
    vals = list(x)
    while len(vals) > 1:
        nv, ni = [], []
        for a in range(0, len(vals), 2):
            take = vals[a + 1] > vals[a]
            nv.append(vals[a + 1] if take else vals[a])
            ni.append(idx[a + 1] if take else idx[a])
        vals, idx = nv, ni
    return vals[0], idx[0]


def walk_cta(scores, last_bits_words, r_begin, r_end, nt=128):
    """scores: [n_rows_total] one query's accumulator values; returns [(max, device_row)] per image of the range."""
    out = []
    m = [-np.inf] * 4          # the four threads of one quad
    c = [0] * 4
    ntiles = -(-(r_end - r_begin) // nt)
    for t in range(ntiles):
        row0 = r_begin + t * nt
        sh, valid = row0 & 31, min(nt, r_end - row0)
        w = [int(last_bits_words[(row0 >> 5) + i]) for i in range(nt // 32 + 1)]
        for g in range(nt // 32):
            em = ((w[g] | (w[g + 1] << 32)) >> sh) & 0xFFFFFFFF                     # __funnelshift_r
            rem = valid - 32 * g
            em &= 0xFFFFFFFF if rem >= 32 else (0 if rem <= 0 else (1 << rem) - 1)
            colbase = row0 - r_begin + 32 * g
            sa = [[scores[r_begin + colbase + 8 * i + 2 * j + e] if colbase + 8 * i + 2 * j + e < r_end - r_begin + 0 and
                   r_begin + colbase + 8 * i + 2 * j + e < len(scores) else -np.inf
                   for i in range(4) for e in range(2)] for j in range(4)]
            lo = 0
            while True:
                p = (em & -em).bit_length() - 1 if em else 31
                seg = (0xFFFFFFFF >> (31 - p)) & ((0xFFFFFFFF << lo) & 0xFFFFFFFF)
                for j in range(4):
                    mine = seg >> (2 * j)
                    xa = [sa[j][2 * i + e] if (mine >> (8 * i + e)) & 1 else -np.inf for i in range(4) for e in range(2)]
                    ma, ia = argmax8(xa)
                    if ma > m[j]:
                        m[j], c[j] = ma, colbase + 2 * j + ((ia >> 1) << 3) + (ia & 1)
                if em == 0:
                    break
                em &= em - 1
                best = max(range(4), key=lambda j: (m[j], -c[j]))                     # quad reduce: score desc, column asc
                out.append((m[best], r_begin + c[best]))
                m, c = [-np.inf] * 4, [0] * 4
                lo = p + 1
                if lo == 32:
                    break
    return out


def test_segment_walk_matches_per_image_argmax():
    rng = np.random.default_rng(0)
    for trial in range(30):
        counts = rng.integers(1, int(rng.choice([2, 9, 70, 200])), size=int(rng.integers(1, 60)))
        n = int(counts.sum())
        scores = (rng.integers(-6, 7, size=n) / 4.0)                                  # many exact ties
        ends = np.cumsum(counts) - 1
        words = np.zeros(n // 32 + 8, dtype=np.uint64)
        for r in ends:
            words[r >> 5] |= np.uint64(1) << np.uint64(r & 31)
        # a CTA range that starts at an arbitrary image (unaligned to 32 rows and to the tile)
        first = int(rng.integers(0, len(counts)))
        last = int(rng.integers(first + 1, len(counts) + 1))
        starts = np.concatenate([[0], np.cumsum(counts)])
        r_begin, r_end = int(starts[first]), int(starts[last])
        got = walk_cta(scores, words, r_begin, r_end, nt=int(rng.choice([64, 128])))
        want = []
        for im in range(first, last):
            s = scores[starts[im]:starts[im + 1]]
            want.append((s.max(), int(starts[im] + np.flatnonzero(s == s.max())[0])))
        assert got == want, trial


def pooled_group_min(v, ngp, vmax=6):
    """numpy model of pooled_group_min (csrc/ssw_common.cuh): lane l holds CTAs l + 32m; CTA c is in group c mod ngp."""
    g_total = len(v)
    lanes = [[v[l + 32 * m] if l + 32 * m < g_total else 0 for m in range(vmax)] for l in range(32)]
    if ngp >= 32:
        gpl = ngp >> 5
        t = []
        for l in range(32):
            tl = 0xFFFFFFFF
            for j in range(gpl):
                tl = min(tl, max(lanes[l][m] for m in range(vmax) if (m & (gpl - 1)) == j))
            t.append(tl)
    else:
        t = [max(x) for x in lanes]
        sft = 16
        while sft >= ngp:
            t = [max(t[l], t[l ^ sft]) for l in range(32)]
            sft >>= 1
    return min(t)


def test_pooled_bound_is_reached_by_at_least_k_ctas():
    """The invariant the scan kernels rely on: the pooled value never exceeds the k-th largest published best."""
    rng = np.random.default_rng(1)
    for trial in range(300):
        g_total = int(rng.choice([148, 132, 160, 64, 33]))
        k = int(rng.integers(1, 65))
        ngp = 1
        while ngp < k:
            ngp <<= 1
        if g_total < ngp:
            continue
        v = rng.integers(1, 2 ** 31, size=g_total).astype(np.int64)
        v[rng.random(g_total) < rng.choice([0.0, 0.1, 0.9])] = 0            # CTAs that have published nothing yet
        t = pooled_group_min(v.tolist(), ngp)
        if t > 0:
            assert int((v >= t).sum()) >= k, (trial, g_total, k, ngp)
        kth = np.sort(v)[::-1][k - 1]
        assert t <= kth
