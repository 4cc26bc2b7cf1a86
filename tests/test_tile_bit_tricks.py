"""CPU checks of two integer formulas the tile kernel (lyon_b200/csrc/tile_kernels.cuh) relies on, restated in numpy next
to the definitions they replace (the kernel itself is exercised by the -m gpu parity tests)."""
import numpy as np

TF_END, TF_BEGIN, TF_FLAGGED = 0x40, 0x20, 0x02


def run_starts_per_slot(flags: np.ndarray) -> np.ndarray:
    """tile_run_start (stroke_tile.cuh): slot s starts a merge run iff it is FLAGGED, neither a Begin nor an End, not slot
    0, and its predecessor is not FLAGGED."""
    out = np.zeros(len(flags), bool)
    for s in range(1, len(flags)):
        fl = flags[s]
        out[s] = bool(fl & TF_FLAGGED) and not (fl & (TF_BEGIN | TF_END)) and not (flags[s - 1] & TF_FLAGGED)
    return out


def run_starts_per_word(flags: np.ndarray) -> np.ndarray:
    """The kernel's form: four flag bytes as one little-endian 32-bit word (phase 1c of k_stroke_tile)."""
    n = len(flags)
    padded = np.zeros((n + 3) // 4 * 4, np.uint8)
    padded[:n] = flags
    words = padded.view("<u4")
    out = np.zeros(len(padded), bool)
    for i, word in enumerate(words.tolist()):
        fl = (word >> 1) & 0x01010101
        if not fl:
            continue
        be = ((word >> 5) | (word >> 6)) & 0x01010101
        prevf = ((int(padded[4 * i - 1]) >> 1) & 1) if i else 1
        start = fl & ~be & ~(((fl << 8) & 0xFFFFFFFF) | prevf) & 0xFFFFFFFF
        while start:
            b = (start & -start).bit_length() - 1
            start &= start - 1
            s = 4 * i + (b >> 3)
            if s < n:
                out[s] = True
    return out[:n]


def test_run_starts_on_words_match_the_per_slot_definition():
    rng = np.random.default_rng(5)
    pool = np.array([0, 0x80, TF_FLAGGED, 0x80 | TF_FLAGGED, TF_BEGIN, TF_END, TF_END | 0x10, TF_BEGIN | TF_FLAGGED,
                     TF_END | TF_FLAGGED, 0x04, 0x04 | TF_FLAGGED], np.uint8)
    for n in (0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 1000):
        for _ in range(20):
            flags = pool[rng.integers(0, len(pool), n)]
            assert np.array_equal(run_starts_per_word(flags), run_starts_per_slot(flags))
    # every pair of neighbouring bytes at every position inside a word
    for a in pool:
        for b in pool:
            for pos in range(8):
                flags = np.zeros(12, np.uint8)
                flags[pos], flags[pos + 1] = a, b
                assert np.array_equal(run_starts_per_word(flags), run_starts_per_slot(flags))


def test_fast_join_index_stores_cover_the_six_words_once():
    """tile_fast_store: a slot's six index words go out as one 128-bit store of the four words that share a 16-byte line and
    one 64-bit store of the other two, chosen by bit 3 of the address (8-byte aligned addresses only; the others take the
    scalar path).  Whatever the alignment, the two stores write words 0..5 exactly once, in order."""
    words = np.arange(6)  # (p0_neg, p0_pos, p1_pos, p0_neg, p1_pos, p1_neg) by position
    for addr in range(0, 64, 8):
        hi = (addr & 8) != 0
        wide_at, wide = (2, words[2:6]) if hi else (0, words[0:4])
        narrow_at, narrow = (0, words[0:2]) if hi else (4, words[4:6])
        assert (addr + 4 * wide_at) % 16 == 0  # the 128-bit store is 16-byte aligned
        assert (addr + 4 * narrow_at) % 8 == 0
        out = np.full(6, -1)
        out[wide_at:wide_at + 4] = wide
        assert np.all(out[narrow_at:narrow_at + 2] == -1)  # the two stores do not overlap
        out[narrow_at:narrow_at + 2] = narrow
        assert np.array_equal(out, words)


def packed_triangle_stores(start_word: int, batches):
    """EmitSinkPacked::add_triangle / flush (stroke_core.cuh) on a word-addressed buffer: `batches` is a list of triangle
    lists, with a flush() between two batches (what arc() does before it hands the position on) and at the end.  Returns
    the buffer and the list of stores as (word address, n words)."""
    total = sum(len(b) for b in batches) * 3
    buf = np.full(start_word + total + 2, -1, np.int64)
    stores = []
    nt = 0
    pend, has_pend = None, False

    def st(addr, vals):
        assert np.all(buf[addr:addr + len(vals)] == -1), "a word written twice"
        if len(vals) == 2:
            assert addr % 2 == 0, "64-bit store off an 8-byte boundary"
        buf[addr:addr + len(vals)] = vals
        stores.append((addr, len(vals)))

    for batch in batches:
        for (a, b, c) in batch:
            d = start_word + 3 * nt
            if has_pend:
                st(d - 1, [pend, a])
                st(d + 1, [b, c])
                has_pend = False
            elif d % 2 == 0:
                st(d, [a, b])
                pend, has_pend = c, True
            else:
                st(d, [a])
                st(d + 1, [b, c])
            nt += 1
        if has_pend:  # flush()
            st(start_word + 3 * nt - 1, [pend])
            has_pend = False
    return buf, stores


def test_packed_triangle_stores_write_every_index_once_in_order():
    rng = np.random.default_rng(9)
    for start in (0, 1, 2, 3):
        for _ in range(50):
            batches, k = [], 0
            for _b in range(rng.integers(1, 4)):
                n = int(rng.integers(0, 6))
                batches.append([(3 * (k + i), 3 * (k + i) + 1, 3 * (k + i) + 2) for i in range(n)])
                k += n
            buf, stores = packed_triangle_stores(start, batches)
            assert np.array_equal(buf[start:start + 3 * k], np.arange(3 * k))
            assert np.all(buf[:start] == -1) and np.all(buf[start + 3 * k:] == -1)
            # three words in at most two stores per triangle, plus one flush word per batch at most
            assert len(stores) <= 2 * k + len(batches)


def tile_end_bsearch(psb, p, hi, cap):
    limit = psb[p] + cap
    a, b = p + 1, hi + 1
    while a < b:
        mid = a + ((b - a) >> 1)
        if psb[mid] <= limit:
            a = mid + 1
        else:
            b = mid
    return a - 1


def tile_end_walk(psb, p, hi, cap, est):
    """k_tile_pack: start from the previous tile's path count, walk at most eight steps, then the binary search."""
    limit = psb[p] + cap
    a, b = p + 1, hi + 1
    g = min(p + est, hi)
    if psb[g] <= limit:
        a = g + 1
        k = 0
        while k < 8 and a < b and psb[a] <= limit:
            a += 1
            k += 1
        if a < b and psb[a] > limit:
            b = a
    else:
        b = g
        k = 0
        while k < 8 and a < b and psb[b - 1] > limit:
            b -= 1
            k += 1
        if a < b and psb[b - 1] <= limit:
            a = b
    while a < b:
        mid = a + ((b - a) >> 1)
        if psb[mid] <= limit:
            a = mid + 1
        else:
            b = mid
    return a - 1


def test_tile_packer_walk_finds_the_same_tile_ends_as_the_binary_search():
    rng = np.random.default_rng(13)
    for trial in range(200):
        n = int(rng.integers(1, 400))
        kind = trial % 4
        if kind == 0:
            sizes = rng.integers(1, 50, n)
        elif kind == 1:
            sizes = rng.integers(380, 440, n)            # C3-like: ~410 slots per path
        elif kind == 2:
            sizes = np.where(rng.random(n) < 0.05, rng.integers(2000, 5000, n), rng.integers(1, 20, n))
        else:
            sizes = rng.integers(0, 3, n)                # empty paths
        psb = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        cap = int(rng.choice([64, 512, 3520]))
        hi = n
        p, est = 0, int(rng.integers(1, 40))
        while p < hi:
            want = tile_end_bsearch(psb, p, hi, cap)
            got = tile_end_walk(psb, p, hi, cap, max(1, est))
            assert got == want, (trial, p, est)
            q = max(want, p + 1)  # a path longer than the capacity is a tile of its own
            est = q - p
            p = q
