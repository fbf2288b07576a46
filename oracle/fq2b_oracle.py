"""CPU restatement of the reference's fq2b (fq2b.c) + the `bsort -k 4 -r <record bytes>` that follows it in the README
pipeline.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and by tests/golden/make_golden_fq2b.py, which pins it to the
output of the unmodified reference binary oracle/_ref/fq2b); never imported by hash10x_b200.

Pure-Python loops: meant for the few hundred entries of the test inputs.
"""
import numpy as np

_SPACK = {ord(c): i % 4 for i, c in enumerate("acgtACGT")}          # fq2b.c:27-28, everything else -> 0


class FastqError(Exception):
    """die() of gzReadFastq (fq2b.c:180-208); .text is the reference's message"""

    def __init__(self, text):
        super().__init__(text)
        self.text = text


def seq_pack(s):
    """seqPack (fq2b.c:33-42): 16 bases per U32, first base on top; what is left after `while (len > 16)` is right-aligned"""
    out, i, n = [], 0, len(s)
    while n > 16:
        u = 0
        for c in s[i:i + 16]:
            u = (u << 2) | _SPACK.get(c, 0)
        out.append(u)
        i += 16
        n -= 16
    u = 0
    for c in s[i:i + n]:
        u = (u << 2) | _SPACK.get(c, 0)
    out.append(u)
    return out


def qual_pack(q):
    """qualPack (fq2b.c:52-61): one bit per base, 1 for q >= '$' + 20 (fq2b.c:29)"""
    out, i, n = [], 0, len(q)
    while n > 32:
        u = 0
        for c in q[i:i + 32]:
            u = (u << 1) | (1 if c >= ord('$') + 20 else 0)
        out.append(u)
        i += 32
        n -= 32
    u = 0
    for c in q[i:i + n]:
        u = (u << 1) | (1 if c >= ord('$') + 20 else 0)
    out.append(u)
    return out


def switch_base(u, code):
    """switchBase (fq2b.c:68-69)"""
    code -= 1
    i, j = code >> 2, code & 3
    return (u & ~(3 << (2 * i)) & 0xFFFFFFFF) | (j << (2 * i))


def whitelist_table(barcodes):
    """read10xWhitelist (fq2b.c:71-94) as a dict over the touched slots: later lines overwrite earlier ones"""
    table = {}
    for u in barcodes:
        for i in range(16):
            ui = 1 + i * 4 + ((u >> (2 * i)) & 3)
            for j in range(4):
                table[switch_base(u, 1 + i * 4 + j)] = ui
    return table


def read_fastq(text, entry0=1, entry_step=1):
    """gzReadFastq (fq2b.c:180-208) over a whole text: [(seq, qual)]; the sequence length is the first entry's"""
    out, p, n, slen, entry = [], 0, len(text), 0, entry0
    while True:
        e = text.find(b"\n", p)
        if e < 0:                                   # gzeof inside the id line: silent end
            break
        if text[p:p + 1] != b"@":
            raise FastqError("fastq id line for entry %d does not start with @" % entry)
        p = e + 1
        if slen:
            s = text[p:p + slen + 1]
            if len(s) != slen + 1:
                raise FastqError("bad seq gzread entry %d" % entry)
        else:
            e = text.find(b"\n", p)
            slen = (e if e >= 0 else n) - p
            s = text[p:p + slen + 1]
        if s[slen:slen + 1] != b"\n":
            raise FastqError("fastq entry %d seq line does not end in \\n" % entry)
        p += slen + 1
        if text[p:p + 2] != b"+\n":
            raise FastqError("bad + fastq line entry %d" % entry)
        p += 2
        q = text[p:p + slen + 1]
        if len(q) != slen + 1:
            raise FastqError("bad qual gzread entry %d" % entry)
        if q[slen:slen + 1] != b"\n":
            raise FastqError("fastq entry %d qual line does not end in \\n" % entry)
        p += slen + 1
        out.append((s[:slen], q[:slen]))
        entry += entry_step
    return out, slen


def fq2b(fq1, fq2=None, whitelist=None):
    """main (fq2b.c:108-178): -> (records uint32 [n, recWords], stats dict)"""
    two = fq2 is not None
    r1, l1 = read_fastq(fq1, 1, 2 if two else 1)
    r2, l2 = read_fastq(fq2, 2, 2) if two else ([], 0)
    if two and len(r2) < len(r1):
        raise FastqError("second fastq file terminated early at %d" % len(r2))
    table = whitelist_table(whitelist) if whitelist is not None else None
    recs, n_bad, n_fixed, fix_base = [], 0, 0, [0] * 16
    for i, (s1, q1) in enumerate(r1):
        u1 = seq_pack(s1) + qual_pack(q1)
        if table is not None:
            c = table.get(u1[0], 0)
            if not c:
                n_bad += 1
                continue
            v = switch_base(u1[0], c)
            if v != u1[0]:
                n_fixed += 1
                fix_base[15 - (c - 1) // 4] += 1
                u1[0] = v
        if two:
            u1 = u1 + seq_pack(r2[i][0]) + qual_pack(r2[i][1])
        recs.append(u1)
    w = (l1 + 15) // 16 + (l1 + 31) // 32 + ((l2 + 15) // 16 + (l2 + 31) // 32 if two else 0)
    a = np.array(recs, dtype=np.uint32).reshape(len(recs), w) if recs else np.zeros((0, w), np.uint32)
    return a, dict(nRead=len(r1), nRecords=len(recs), nBad=n_bad, nFixed=n_fixed, nFixBase=fix_base, s1Len=l1, s2Len=l2,
                   recWords=w)


def bsort(recs):
    """`bsort -k 4 -r <record bytes>`: records ordered by their first four bytes as memcmp sees them, i.e. by the
    byte-swapped first word; stable here (bsort's order inside a key is not specified: hash10x only needs runs)"""
    if recs.shape[0] == 0:
        return recs
    key = recs[:, 0].astype(np.uint32).byteswap()
    return recs[np.argsort(key, kind="stable")]


def pack_barcode(s):
    assert len(s) == 16
    return seq_pack(s.encode() if isinstance(s, str) else s)[0]


def synth_fastq(seed, n, l1=151, l2=151, whitelist=None, p_err=0.3, p_bad=0.1, p_n=0.02, lower=False):
    """two FASTQ texts of n entries: read 1 starts with a whitelist barcode (some with one substitution, some random)"""
    rng = np.random.default_rng(seed)
    alpha = b"acgt" if lower else b"ACGT"
    f1, f2 = [], []
    for i in range(n):
        s1 = bytearray(alpha[x] for x in rng.integers(0, 4, l1))
        s2 = bytearray(alpha[x] for x in rng.integers(0, 4, l2))
        if whitelist and l1 >= 16:
            r = rng.random()
            if r >= p_bad:
                bc = bytearray(whitelist[int(rng.integers(0, len(whitelist)))].encode())
                if r < p_bad + p_err:
                    k = int(rng.integers(0, 16))
                    bc[k] = b"ACGT"[(b"ACGT".index(bc[k]) + int(rng.integers(1, 4))) % 4]
                s1[:16] = bc
        for s in (s1, s2):
            for k in np.nonzero(rng.random(len(s)) < p_n)[0]:
                s[k] = ord("N")
        q1 = bytes(int(x) for x in rng.integers(33, 75, l1))
        q2 = bytes(int(x) for x in rng.integers(33, 75, l2))
        name = b"@read%d/%s" % (i, b"x" * int(rng.integers(0, 9)))
        f1.append(name + b" 1\n" + bytes(s1) + b"\n+\n" + q1 + b"\n")
        f2.append(name + b" 2\n" + bytes(s2) + b"\n+\n" + q2 + b"\n")
    return b"".join(f1), b"".join(f2)


def synth_whitelist(seed, n):
    """n distinct random barcodes plus a few pairs one and two substitutions apart (their variants collide)"""
    rng = np.random.default_rng(seed)
    out, seen = [], set()
    while len(out) < n:
        s = "".join("ACGT"[x] for x in rng.integers(0, 4, 16))
        if s not in seen:
            seen.add(s)
            out.append(s)
    for k in range(min(4, n)):
        s = list(out[k])
        s[3 + k] = "ACGT"[("ACGT".index(s[3 + k]) + 1) % 4]
        out.append("".join(s))                       # distance 1: each is a variant of the other
        s[9] = "ACGT"[("ACGT".index(s[9]) + 2) % 4]
        out.append("".join(s))                       # distance 2 from out[k]: they share variants
    return out
