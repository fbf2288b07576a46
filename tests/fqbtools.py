"""Hand-built FQB records for edge-case tests (packing as fq2b.c:33-61 of the reference)."""
import numpy as np

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def pack_read(codes):
    """151 2-bit codes -> 15 words (10 base words, 5 qual words of all-good quality)."""
    codes = np.asarray(codes, dtype=np.uint32)
    assert codes.size == 151
    u = np.zeros(15, np.uint32)
    for i in range(9):
        v = 0
        for c in codes[16 * i:16 * i + 16]:
            v = (v << 2) | int(c)
        u[i] = v
    v = 0
    for c in codes[144:151]:
        v = (v << 2) | int(c)
    u[9] = v                        # last 7 bases right-aligned (fq2b.c:41)
    u[10:14] = 0xFFFFFFFF
    u[14] = 0x7FFFFF
    return u


def barcode_codes(word):
    return [(word >> (2 * (15 - i))) & 3 for i in range(16)]


def make_record(barcode_word, r1_tail, r2):
    """barcode word + 135 codes for read-1 positions 16..150 + 151 codes of read 2."""
    r1 = np.concatenate([barcode_codes(barcode_word), np.asarray(r1_tail)]).astype(np.uint32)
    rec = np.concatenate([pack_read(r1), pack_read(r2)])
    assert rec[0] == barcode_word
    return rec


def random_records(rng, barcode_words, counts):
    """counts[i] random records for barcode_words[i], grouped in that order."""
    out = []
    for w, n in zip(barcode_words, counts):
        for _ in range(n):
            out.append(make_record(int(w), rng.integers(0, 4, 135), rng.integers(0, 4, 151)))
    return np.array(out, dtype=np.uint32).reshape(-1, 30)


def const_records(barcode_word, n, base1, base2):
    """n records whose read-1 tail is all base1 and read 2 all base2 (poly-A gives hash 0 moshes,
    poly-C none: SURVEY.md Appendix E)."""
    rec = make_record(barcode_word, np.full(135, base1), np.full(151, base2))
    return np.tile(rec, (n, 1))


def mixed_records(words, counts, kinds, seed):
    """runs of random / poly-A / poly-C / dinucleotide-repeat reads (used by the property tests)"""
    rng = np.random.default_rng(seed)
    out = []
    for w, n, kind in zip(words, counts, kinds):
        if kind == "rand":
            out.append(random_records(rng, [w], [n]))
        elif kind == "polyA":
            out.append(const_records(w, n, 0, 0))
        elif kind == "polyC":
            out.append(const_records(w, n, 1, 1))
        else:
            r1 = np.tile([0, 1], 68)[:135]
            r2 = np.tile([2, 3], 76)[:151]
            out.append(np.tile(make_record(w, r1, r2), (n, 1)))
    return np.concatenate(out).astype(np.uint32)


def fastq_from_fqb(recs, read_len):
    """the two FASTQ texts fq2b would have packed into these records (fq2b.c:33-61: the last partial word of a line is
    right-aligned; quality bit 1 -> 'I', 0 -> '#'); read_len bases per read"""
    ws, wq = (read_len + 15) // 16, (read_len + 31) // 32
    assert recs.shape[1] == 2 * (ws + wq)

    def unpack(words, bits, n):
        per = 32 // bits
        out = []
        for i, u in enumerate(words):
            k = per if (i + 1 < len(words)) else n - per * i
            out += [(int(u) >> (bits * (k - 1 - j))) & ((1 << bits) - 1) for j in range(k)]
        return out
    f1, f2 = [], []
    for r, rec in enumerate(recs):
        for f, o, tag in ((f1, 0, b"1"), (f2, ws + wq, b"2")):
            s = bytes(b"ACGT"[x] for x in unpack(rec[o:o + ws], 2, read_len))
            q = bytes(b"#I"[x] for x in unpack(rec[o + ws:o + ws + wq], 1, read_len))
            f.append(b"@r%d %s\n" % (r, tag) + s + b"\n+\n" + q + b"\n")
    return b"".join(f1), b"".join(f2)


def abandonment_case(n_groups=300, seed=5):
    """A barcode block with more than 255 sub-clusters (hash10x.c:810-817): block 1 holds n_groups read pairs, and read pair
    g is also the only read pair of its own barcode block 2 + g, so the moshes of pair g are shared by exactly those two
    blocks (depth 2).  With --hashDepthRange 2 3 (min <= depth < max) and -ct 1 the second good hash of every pair founds a cluster in block
    1: the 256th founding abandons the block's clustering.  One more run closes the file (the last run is never hashed)."""
    rng = np.random.default_rng(seed)
    words = rng.choice(np.arange(1, 1 << 31, dtype=np.int64), size=n_groups + 2, replace=False)
    pairs = [(rng.integers(0, 4, 135), rng.integers(0, 4, 151)) for _ in range(n_groups)]
    out = [make_record(int(words[0]), t, r2) for t, r2 in pairs]
    out += [make_record(int(words[1 + g]), t, r2) for g, (t, r2) in enumerate(pairs)]
    out.append(make_record(int(words[n_groups + 1]), rng.integers(0, 4, 135), rng.integers(0, 4, 151)))
    return np.array(out, dtype=np.uint32).reshape(-1, 30)
