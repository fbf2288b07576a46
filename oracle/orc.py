"""ctypes loader for oracle/liborc.so - the CPU restatement and the synthetic FQB generator.

TEST INFRASTRUCTURE ONLY (see the header of h10x_oracle.c): imported by tests/, by
__graft_entry__.smoke() and by bench.py's CPU-baseline legs.  Never imported by hash10x_b200.
"""
import ctypes as C
import os
import subprocess

import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "liborc.so")
REF_DIR = os.path.join(_HERE, "_ref")

STATUS = {0: "ok", 1: "hashTableSize is too small", 2: "chunkSize too small", 3: "bad parameter",
          4: "out of memory", 5: "io"}

DEFAULT_FACTOR1 = 0x49308BB9003CB3AD  # seed 17, SURVEY.md Appendix E


def build_lib(force=False):
    """(Re)build liborc.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB


class OrcIndex(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("B", C.c_int32), ("status", C.c_int32),
                ("factor1", C.c_uint64), ("hashNumber", C.c_uint32), ("nBlocksMax", C.c_uint32),
                ("nReads", C.c_uint64), ("nHashes", C.c_uint64),
                ("hashIndex", C.POINTER(C.c_uint32)), ("hashValue", C.POINTER(C.c_uint64)),
                ("hashDepth", C.POINTER(C.c_uint32)), ("blkNRead", C.POINTER(C.c_uint32)),
                ("blkNHash", C.POINTER(C.c_uint32)), ("blkOff", C.POINTER(C.c_uint64)),
                ("clus", C.POINTER(C.c_uint64)), ("codeOff", C.POINTER(C.c_uint64)),
                ("codes", C.POINTER(C.c_uint32))]


from hash10x_b200.synth import SynthParams, make_params as _make_synth_params  # shared parameter struct


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(_LIB)
        L.orc_factor1.restype = C.c_uint64
        L.orc_factor1.argtypes = [C.c_int]
        L.orc_build.restype = C.POINTER(OrcIndex)
        L.orc_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                C.c_int64, C.c_int, C.c_int, C.c_int]
        L.orc_free.argtypes = [C.POINTER(OrcIndex)]
        L.orc_write_hash.argtypes = [C.POINTER(OrcIndex), C.c_char_p]
        L.orc_record_moshes.restype = C.c_int
        L.orc_record_moshes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int]
        L.orc_seq_moshes.restype = C.c_int
        L.orc_seq_moshes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
        L.orc_kmer_hashes.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_good_hashes.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cluster_stale_labels.restype = C.c_uint64
        L.orc_digest_u32.restype = C.c_uint64
        L.orc_digest_u32.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_digest_u64.restype = C.c_uint64
        L.orc_digest_u64.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.synth_layout.restype = C.c_uint64
        L.synth_layout.argtypes = [C.POINTER(SynthParams), C.c_void_p]
        L.synth_fill.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def factor1(seed=17):
    return int(lib().orc_factor1(seed))


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Index:
    """Plain-numpy view of a built index (same fields for the oracle and the GPU build)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def status_text(self):
        return STATUS.get(self.status, "status %d" % self.status)


def build(recs, k=21, w=31, factor1_=DEFAULT_FACTOR1, B=24, N=0, chunk=100000, minB=20, maxB=30,
          keep_table=True):
    """Run the oracle on an in-memory FQB (uint32 array of 30*n words)."""
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
    n = recs.size // 30
    L = lib()
    p = L.orc_build(recs.ctypes.data, n, k, w, factor1_, B, N, chunk, minB, maxB)
    try:
        ix = p.contents
        out = Index(k=ix.k, w=ix.w, B=ix.B, status=ix.status, factor1=ix.factor1,
                    hashNumber=ix.hashNumber, nBlocksMax=ix.nBlocksMax, nReads=ix.nReads,
                    nHashes=ix.nHashes)
        if ix.status in (0,):
            hn, nb = ix.hashNumber, ix.nBlocksMax
            out.hashIndex = _np(ix.hashIndex, 1 << ix.B, np.uint32) if keep_table else None
            out.hashValue = _np(ix.hashValue, hn, np.uint64)
            out.hashDepth = _np(ix.hashDepth, hn, np.uint32)
            out.blkNRead = _np(ix.blkNRead, nb, np.uint32)
            out.blkNHash = _np(ix.blkNHash, nb, np.uint32)
            out.blkOff = _np(ix.blkOff, nb + 1, np.uint64)
            out.clus = _np(ix.clus, ix.nHashes, np.uint64)
            out.codeOff = _np(ix.codeOff, hn + 1, np.uint64)
            out.codes = _np(ix.codes, ix.nHashes, np.uint32)
        return out
    finally:
        L.orc_free(p)


def build_and_write(recs, path, **kw):
    """Oracle build written as a .hash file by the oracle's own writer; returns status."""
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
    L = lib()
    p = L.orc_build(recs.ctypes.data, recs.size // 30, kw.get("k", 21), kw.get("w", 31),
                    kw.get("factor1_", DEFAULT_FACTOR1), kw.get("B", 24), kw.get("N", 0),
                    kw.get("chunk", 100000), kw.get("minB", 20), kw.get("maxB", 30))
    try:
        st = p.contents.status
        if st == 0:
            st = L.orc_write_hash(p, path.encode())
        return st
    finally:
        L.orc_free(p)


def time_build(recs, repeat=1, **kw):
    """Wall-clock seconds of the oracle build alone (arrays are not copied out)."""
    import time
    recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
    L = lib()
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        p = L.orc_build(recs.ctypes.data, recs.size // 30, kw.get("k", 21), kw.get("w", 31),
                        kw.get("factor1_", DEFAULT_FACTOR1), kw.get("B", 24), kw.get("N", 0),
                        kw.get("chunk", 100000), 20, 30)
        dt = time.perf_counter() - t0
        st = p.contents.status
        L.orc_free(p)
        if st:
            raise RuntimeError("oracle build failed: " + STATUS.get(st, str(st)))
        best = dt if best is None else min(best, dt)
    return best


def record_moshes(rec, k=21, w=31, factor1_=DEFAULT_FACTOR1):
    """(hash, pos, which) arrays for one 30-word record, in processBlock's generation order."""
    rec = np.ascontiguousarray(rec, dtype=np.uint32)
    cap = 320
    h = np.zeros(cap, np.uint64); pos = np.zeros(cap, np.int32); which = np.zeros(cap, np.uint8)
    n = lib().orc_record_moshes(rec.ctypes.data, k, w, factor1_, h.ctypes.data, pos.ctypes.data,
                                which.ctypes.data, cap)
    return h[:n], pos[:n], which[:n]


def seq_moshes(codes, k=21, w=31, factor1_=DEFAULT_FACTOR1):
    """(hash, pos, isForward) for a base string given as 2-bit codes (uint8)."""
    s = np.ascontiguousarray(codes, dtype=np.uint8)
    cap = max(1, s.size)
    h = np.zeros(cap, np.uint64); pos = np.zeros(cap, np.int32); fwd = np.zeros(cap, np.uint8)
    n = lib().orc_seq_moshes(s.ctypes.data, s.size, k, w, factor1_, h.ctypes.data, pos.ctypes.data,
                             fwd.ctypes.data, cap)
    return h[:n], pos[:n], fwd[:n]


def kmer_hashes(h, hrc, k=21, factor1_=DEFAULT_FACTOR1):
    a, b = C.c_uint64(), C.c_uint64()
    lib().orc_kmer_hashes(h, hrc, factor1_, k, C.byref(a), C.byref(b))
    return a.value, b.value


def good_hashes(ix, dmin, dmax, within=None):
    """hashWithinRangeBuild + goodHashesBuild (hash10x.c:528-539,738-766) on an Index:
    returns (within flags, goodOff[nBlocksMax+1], good u16 indices)."""
    hn, nb = int(ix.hashNumber), int(ix.nBlocksMax)
    if within is None:
        within = np.zeros(hn, np.uint8)
    depth = np.ascontiguousarray(ix.hashDepth, np.uint32)
    nh = np.ascontiguousarray(ix.blkNHash, np.uint32)
    off = np.ascontiguousarray(ix.blkOff, np.uint64)
    clus = np.ascontiguousarray(ix.clus, np.uint64)
    good_off = np.zeros(nb + 1, np.uint64)
    good = np.zeros(max(1, clus.size), np.uint16)
    st = lib().orc_good_hashes(hn, depth.ctypes.data, nb, nh.ctypes.data, off.ctypes.data, clus.ctypes.data,
                               dmin, dmax, within.ctypes.data, good_off.ctypes.data, good.ctypes.data)
    assert st == 0
    return within, good_off, good[:int(good_off[nb])]


def cluster(ix, good_off, good, code_min=0, code_max=0, threshold=5, clus=None, n_sub=None, point_to_min=None):
    """`--cluster codeMin codeMax` (codeClusterFind + codeClusterReadMerge, hash10x.c:770-868) on an Index and its
    goodHashes lists: returns (clus with the subCluster bytes set, nSubCluster[nBlocksMax], pointToMin[nBlocksMax]).
    clus / n_sub / point_to_min carry the state of an earlier --cluster command."""
    nb = int(ix.nBlocksMax)
    clus = np.array(ix.clus if clus is None else clus, np.uint64, copy=True)
    if clus.size == 0:
        clus = np.zeros(1, np.uint64)
    n_sub = np.zeros(nb, np.uint32) if n_sub is None else np.array(n_sub, np.uint32, copy=True)
    ptm = np.zeros(nb, np.float64) if point_to_min is None else np.array(point_to_min, np.float64, copy=True)
    depth = np.ascontiguousarray(ix.hashDepth, np.uint32)
    code_off = np.ascontiguousarray(ix.codeOff, np.uint64)
    codes = np.ascontiguousarray(ix.codes, np.uint32)
    nr = np.ascontiguousarray(ix.blkNRead, np.uint32)
    nh = np.ascontiguousarray(ix.blkNHash, np.uint32)
    off = np.ascontiguousarray(ix.blkOff, np.uint64)
    good_off = np.ascontiguousarray(good_off, np.uint64)
    good = np.ascontiguousarray(good, np.uint16)
    st = lib().orc_cluster(depth.ctypes.data, code_off.ctypes.data, codes.ctypes.data, nb, nr.ctypes.data,
                           nh.ctypes.data, off.ctypes.data, clus.ctypes.data, good_off.ctypes.data,
                           good.ctypes.data if good.size else None, code_min, code_max, threshold,
                           n_sub.ctypes.data, ptm.ctypes.data)
    assert st == 0, st
    return clus[:int(ix.nHashes)], n_sub, ptm


def cluster_stale_labels():
    """entries whose label from an earlier --cluster exceeded the block's new nSubCluster (undefined behaviour in
    the reference, treated as unclustered here), summed over all cluster() calls of this process"""
    return int(lib().orc_cluster_stale_labels())


# ------------------------------------------------------------------ synthetic FQB (CPU)

def synth_params(**kw):
    """parameters of the synthetic data set (see hash10x_b200/synth.py::make_params)"""
    return _make_synth_params(**kw)


def synth_layout(p):
    off = np.zeros(p.nBarcodes + 1, np.uint64)
    n = lib().synth_layout(C.byref(p), off.ctypes.data)
    return int(n), off


def synth_fqb(p, r0=0, r1=None):
    """Records r0..r1-1 of the synthetic data set as a (n,30) uint32 array."""
    n, off = synth_layout(p)
    if r1 is None:
        r1 = n
    out = np.zeros((r1 - r0, 30), np.uint32)
    lib().synth_fill(C.byref(p), off.ctypes.data, r0, r1, out.ctypes.data)
    return out


# ------------------------------------------------------------------ the compiled reference

def ref_binary(name="hash10x"):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run_reference(fqb_path, hash_path=None, B=24, extra=(), k=None, w=None, r=None, N=None,
                  chunk=None, binary="hash10x", timeout=600):
    """Run oracle/_ref/hash10x on an FQB file; returns CompletedProcess (stdout text)."""
    exe = ref_binary(binary)
    if exe is None:
        raise FileNotFoundError("oracle/_ref/%s not built" % binary)
    cmd = [exe]
    for flag, v in (("-k", k), ("-w", w), ("-r", r), ("-N", N), ("-c", chunk)):
        if v is not None:
            cmd += [flag, str(v)]
    cmd += ["-B", str(B), "--readFQB", fqb_path]
    if hash_path:
        cmd += ["--writeHash", hash_path]
    cmd += list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


def digest(a, base=0, mask=0xFFFFFFFFFFFFFFFF):
    """position-salted sum digest (hash10x_b200/csrc/h10x_digest.h) of a uint32 / uint64 array on the host"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return int(lib().orc_digest_u32(a.ctypes.data, a.size, base))
    assert a.dtype == np.uint64, a.dtype
    return int(lib().orc_digest_u64(a.ctypes.data, a.size, base, mask))


def index_digests(ix):
    """the digests h10x_gpu_index_digest reports, computed from an Index on the host"""
    d = {"hashValue": digest(ix.hashValue), "hashDepth": digest(ix.hashDepth), "blkNRead": digest(ix.blkNRead),
         "blkNHash": digest(ix.blkNHash), "clusHash": digest(ix.clus, mask=0x0000FFFFFFFFFFFF)}
    if getattr(ix, "hashIndex", None) is not None:
        d["hashIndex"] = digest(ix.hashIndex)
    if getattr(ix, "codes", None) is not None:
        d["codes"] = digest(ix.codes)
        d["codeOff"] = digest(ix.codeOff)
    return d


# ------------------------------------------------------------------ --clusterSplit (hash10x.c:956-1013)

class SplitIndex:
    """what clusterSplitCodes leaves in clusterBlocks: per-block arrays (nBlocksMax = old + sum of nSubCluster) and the
    ClusterHash stream (8-byte words: bin id | read << 32, subCluster and flags 0), block after block"""
    pass


def cluster_split(ix, clus, n_sub, point_to_min):
    """clusterSplitCodes (hash10x.c:956-1013), restated.  TEST INFRASTRUCTURE.  Every sub-cluster j of block i becomes a
    new block behind the original ones (:963-964: new2 starts at nCodes - 1 and is indexed from 1, so the clusters of
    block i sit at nCodes + sum of nSubCluster of the blocks before i, in label order) holding the block's entries with
    that label in list order, subCluster wiped (:981), reads renumbered in order of first appearance (:983-985: ONE
    readMap per original block, so the number a read gets comes from the counter of the cluster of its FIRST clustered
    entry); clusterParent = i + 1 (:976).  The parent keeps the unclustered entries and its old nRead (:991) in a fresh,
    otherwise zero ClusterBlock; blocks without sub-clusters are copied as they are (:998)."""
    nb = int(ix.nBlocksMax)
    clus = np.asarray(clus, np.uint64)
    n_sub = np.asarray(n_sub, np.uint32)
    off = np.asarray(ix.blkOff, np.uint64).astype(np.int64)
    nh = np.asarray(ix.blkNHash, np.uint32).astype(np.int64)
    total_sub = int(n_sub.astype(np.int64).sum())
    out = SplitIndex()
    out.nBlocksMax = nb + total_sub
    out.blkNRead = np.zeros(out.nBlocksMax, np.uint32)
    out.blkNHash = np.zeros(out.nBlocksMax, np.uint32)
    out.blkNSub = np.zeros(out.nBlocksMax, np.uint32)
    out.blkParent = np.zeros(out.nBlocksMax, np.uint32)
    out.blkPointToMin = np.zeros(out.nBlocksMax, np.float64)
    lists = [None] * out.nBlocksMax
    nxt = nb                                             # new2 + 1 in the reference's terms
    for i in range(nb):
        e = clus[off[i]:off[i] + nh[i]] if i else clus[0:0]        # block 0 is the dummy
        ns = int(n_sub[i])
        if not ns:
            lists[i] = e.copy()
            out.blkNRead[i], out.blkNHash[i] = ix.blkNRead[i], nh[i]
            out.blkPointToMin[i] = point_to_min[i]                  # *new1++ = *old carries every field (:998)
            continue
        lab = ((e >> np.uint64(48)) & np.uint64(0xFF)).astype(np.int64)
        read = ((e >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
        base = e & np.uint64(0xFF00FFFFFFFFFFFF)                    # c->subCluster = 0 (:981); the flags byte travels
        read_map = {}
        n_read = [0] * (ns + 1)
        parts = [[] for _ in range(ns + 1)]
        for j in range(e.size):
            c = int(lab[j])
            if c:
                r = int(read[j])
                if r not in read_map:
                    n_read[c] += 1
                    read_map[r] = n_read[c]
                w = (int(base[j]) & ~(0xFFFF << 32)) | ((read_map[r] - 1) << 32)
                parts[c].append(w)
            else:
                parts[0].append(int(base[j]))
        lists[i] = np.array(parts[0], np.uint64)
        out.blkNRead[i], out.blkNHash[i] = ix.blkNRead[i], len(parts[0])
        for c in range(1, ns + 1):
            k = nxt + c - 1
            lists[k] = np.array(parts[c], np.uint64)
            out.blkNRead[k], out.blkNHash[k], out.blkParent[k] = n_read[c], len(parts[c]), i + 1
        nxt += ns
    out.clus = np.concatenate([x for x in lists if x is not None and x.size] or [np.zeros(0, np.uint64)])
    out.blkOff = np.zeros(out.nBlocksMax + 1, np.uint64)
    out.blkOff[1:] = np.cumsum(out.blkNHash.astype(np.uint64))
    out.nHashes = int(out.clus.size)
    # fillHashTable (:317-347) over the new blocks: per bin the blocks holding it, ascending
    ids = (out.clus & np.uint64(0xFFFFFFFF)).astype(np.int64)
    blk = np.repeat(np.arange(out.nBlocksMax, dtype=np.int64), out.blkNHash.astype(np.int64))
    order = np.argsort(ids, kind="stable")
    out.codes = blk[order].astype(np.uint32)
    return out
