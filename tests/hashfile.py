"""Parser and canonicaliser for hash10x `.hash` files (SURVEY.md Appendix B;
reference writeHashFile hash10x.c:244-267, ArrayStruct array.h:41-50).  Test helper."""
import struct

import numpy as np


class HashFile:
    pass


def _read_array(buf, off, want_size):
    magic, _p0, _base, dim, size, mx, _p1 = struct.unpack_from("<iiQiiii", buf, off)
    assert size == want_size, (size, want_size)
    off += 32
    data = buf[off:off + size * dim]
    assert len(data) == size * dim, "truncated array"
    return magic, dim, mx, data, off + size * dim


def parse(path, keep_table=True):
    buf = memoryview(open(path, "rb").read())
    hf = HashFile()
    assert bytes(buf[0:4]) == b"10XH"
    hf.version, hf.chSize, hf.cbSize, hf.B = struct.unpack_from("<IHHi", buf, 4)
    assert (hf.version, hf.chSize, hf.cbSize) == (2, 8, 32)
    off = 16
    n = 1 << hf.B
    hf.hashIndex = np.frombuffer(buf, np.uint32, n, off).copy() if keep_table else None
    off += 4 * n
    (hf.hashNumber,) = struct.unpack_from("<I", buf, off)
    off += 4
    hf.hashValue = np.frombuffer(buf, np.uint64, hf.hashNumber, off).copy()
    off += 8 * hf.hashNumber
    hf.depthMagic, hf.depthDim, hf.depthMax, d, off = _read_array(buf, off, 4)
    hf.hashDepth = np.frombuffer(d, np.uint32)[:hf.depthMax].copy()
    hf.blkMagic, hf.blkDim, hf.nBlocksMax, d, off = _read_array(buf, off, 32)
    cb = np.frombuffer(d, np.uint32).reshape(-1, 8)[:hf.nBlocksMax]
    hf.blkNRead = cb[:, 0].copy()
    hf.blkNHash = cb[:, 1].copy()
    hf.blkNSub = cb[:, 2].copy()
    hf.blkParent = cb[:, 3].copy()
    hf.blkPointToMin = np.frombuffer(d, np.float64).reshape(-1, 4)[:hf.nBlocksMax, 3].copy()
    hf.nHashes = int(hf.blkNHash[1:].sum()) if hf.nBlocksMax > 1 else 0
    raw = np.frombuffer(buf, np.uint64, hf.nHashes, off).copy()
    off += 8 * hf.nHashes
    assert off == len(buf), "trailing bytes: %d != %d" % (off, len(buf))
    hf.clusRaw = raw
    hf.clusIdx = (raw & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hf.clusRead = ((raw >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.uint16)
    hf.clusSub = ((raw >> np.uint64(48)) & np.uint64(0xFF)).astype(np.uint8)
    hf.blkOff = np.zeros(hf.nBlocksMax + 1, np.uint64)
    if hf.nBlocksMax > 1:
        hf.blkOff[2:] = np.cumsum(hf.blkNHash[1:].astype(np.uint64))
        hf.blkOff[1] = 0
    hf.size = len(buf)
    return hf


def from_index(ix):
    """Wrap an oracle/GPU Index (numpy fields) in the same attribute names as parse()."""
    hf = HashFile()
    hf.B = ix.B
    hf.hashIndex = getattr(ix, "hashIndex", None)
    hf.hashNumber = int(ix.hashNumber)
    hf.hashValue = np.asarray(ix.hashValue, np.uint64)
    hf.hashDepth = np.asarray(ix.hashDepth, np.uint32)
    hf.depthMax = hf.hashNumber if hf.hashNumber > 1 else 0
    hf.nBlocksMax = int(ix.nBlocksMax)
    hf.blkNRead = np.asarray(ix.blkNRead, np.uint32)
    hf.blkNHash = np.asarray(ix.blkNHash, np.uint32)
    hf.nHashes = int(ix.nHashes)
    raw = np.asarray(ix.clus, np.uint64)
    hf.clusRaw = raw
    hf.clusIdx = (raw & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hf.clusRead = ((raw >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.uint16)
    hf.blkOff = np.asarray(ix.blkOff, np.uint64)
    return hf


def hash_to_code_lists(hf):
    """fillHashTable (hash10x.c:317-347) as CSR (codeOff, codes) derived from the block lists."""
    hn = hf.hashNumber
    blk = np.repeat(np.arange(hf.nBlocksMax, dtype=np.uint32), np.r_[0, hf.blkNHash[1:]].astype(np.int64)
                    if hf.nBlocksMax > 1 else np.zeros(hf.nBlocksMax, np.int64))
    order = np.argsort(hf.clusIdx, kind="stable")
    codes = blk[order]
    counts = np.bincount(hf.clusIdx, minlength=hn).astype(np.uint64)
    codeOff = np.zeros(hn + 1, np.uint64)
    codeOff[1:] = np.cumsum(counts)
    return codeOff, codes


def check_table(hf):
    """Every bin id is reachable by hashIndexFind (hash10x.c:139-152) and stored exactly once."""
    B = hf.B
    mask = np.uint64((1 << B) - 1)
    hv = hf.hashValue[1:]
    ids = np.arange(1, hf.hashNumber, dtype=np.uint32)
    off = hv & mask
    diff = ((hv >> np.uint64(B)) & mask) | np.uint64(1)
    todo = np.arange(hv.size)
    tab = hf.hashIndex
    for _ in range(1 << 12):
        if todo.size == 0:
            break
        got = tab[off[todo]]
        assert (got != 0).all(), "probe path crosses an empty slot"
        hit = got == ids[todo]
        todo = todo[~hit]
        off[todo] = (off[todo] + diff[todo]) & mask
    assert todo.size == 0
    assert int((tab != 0).sum()) == hf.hashNumber - 1


def canonical(hf):
    """Bin-id independent content: sorted (hash value, depth); per block nRead and
    the (hash value, read) list sorted by hash value; per hash value the ascending block list."""
    hv = hf.hashValue
    pairs = np.stack([hv[1:], hf.hashDepth[1:hf.hashNumber].astype(np.uint64)], 1)
    pairs = pairs[np.argsort(pairs[:, 0], kind="stable")]
    vals = hv[hf.clusIdx]
    blk = np.repeat(np.arange(hf.nBlocksMax, dtype=np.uint64),
                    np.r_[0, hf.blkNHash[1:]].astype(np.int64) if hf.nBlocksMax > 1
                    else np.zeros(hf.nBlocksMax, np.int64))
    o = np.lexsort((vals, blk))
    per_block = np.stack([blk[o], vals[o], hf.clusRead[o].astype(np.uint64)], 1)
    o2 = np.lexsort((blk, vals))
    per_hash = np.stack([vals[o2], blk[o2]], 1)
    return {"bins": pairs, "nRead": hf.blkNRead.copy(), "nHash": hf.blkNHash.copy(),
            "per_block": per_block, "per_hash": per_hash}


def assert_canonical_equal(a, b):
    ca, cb = canonical(a), canonical(b)
    for key in ("nRead", "nHash", "bins", "per_block", "per_hash"):
        assert ca[key].shape == cb[key].shape, (key, ca[key].shape, cb[key].shape)
        assert np.array_equal(ca[key], cb[key]), key


def assert_strict_equal(a, b, table=True):
    """Strict mode: identical bin ids, hashValue order, depths, per-block (idx, read) sequences
    and (when table=True) a byte-identical hashIndex."""
    assert a.B == b.B
    assert a.hashNumber == b.hashNumber, (a.hashNumber, b.hashNumber)
    assert a.nBlocksMax == b.nBlocksMax, (a.nBlocksMax, b.nBlocksMax)
    assert np.array_equal(a.blkNRead, b.blkNRead)
    assert np.array_equal(a.blkNHash, b.blkNHash)
    assert np.array_equal(a.hashValue, b.hashValue)
    assert np.array_equal(a.hashDepth[:a.hashNumber], b.hashDepth[:b.hashNumber])
    assert np.array_equal(a.clusIdx, b.clusIdx)
    assert np.array_equal(a.clusRead, b.clusRead)
    # bytes 6-7 (subCluster, flags) are uninitialised malloc in the reference (hash10x.c:175);
    # they are compared only when both sides are ours
    if table and a.hashIndex is not None and b.hashIndex is not None:
        assert np.array_equal(a.hashIndex, b.hashIndex)
