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
