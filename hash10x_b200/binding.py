"""ctypes binding of include/h10x_gpu.h (the C ABI of libh10xgpu.so).

Mirrors the reference's seam for this path - `initialise(); readFQB(); fillHashTable();`
(hash10x.c:1200-1205) - with the same parameter names (-k -w -r -B -N -c) and the same error
texts ("hashTableSize is too small", "chunkSize too small", "hashTableBits %d out of range").
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_FACTOR1 = 0x49308BB9003CB3AD  # -r 17

H10X_NSTAGES = 12
FLAG_WIDE_B, FLAG_NO_TABLE, FLAG_NO_CODES, FLAG_GENERIC_ONLY, FLAG_LEGACY_TAIL, FLAG_LAZY_CODES = 1, 2, 4, 8, 16, 32


class H10xError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("h10x error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("factor1", C.c_uint64), ("B", C.c_int32),
                ("chunkSize", C.c_int32), ("N", C.c_int64), ("device", C.c_int32),
                ("flags", C.c_uint32)]


class CIndex(C.Structure):
    _fields_ = [("B", C.c_int32), ("hashNumber", C.c_uint32), ("nBlocksMax", C.c_uint32),
                ("reserved", C.c_uint32), ("nReads", C.c_uint64), ("nHashes", C.c_uint64),
                ("hashIndex", C.c_void_p), ("hashValue", C.c_void_p), ("hashDepth", C.c_void_p),
                ("blkNRead", C.c_void_p), ("blkNHash", C.c_void_p), ("blkOff", C.c_void_p),
                ("clusHash", C.c_void_p), ("codeOff", C.c_void_p), ("codes", C.c_void_p),
                ("onDevice", C.c_int32), ("pinned", C.c_int32),
                ("blkNSubCluster", C.c_void_p), ("blkPointToMin", C.c_void_p), ("blkClusterParent", C.c_void_p)]


class CStats(C.Structure):
    _fields_ = [("msTotal", C.c_double), ("msStage", C.c_double * H10X_NSTAGES),
                ("nRecords", C.c_uint64), ("nMoshes", C.c_uint64), ("nHashes", C.c_uint64),
                ("nBins", C.c_uint64), ("nBlocks", C.c_uint64), ("algorithmicBytes", C.c_uint64),
                ("kernelLaunches", C.c_uint64), ("fusedBlocks", C.c_uint64),
                ("genericBlocks", C.c_uint64), ("peakDeviceBytes", C.c_uint64), ("tailPath", C.c_uint64)]


class CDigest(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("hashIndex", "hashValue", "hashDepth", "blkNRead", "blkNHash", "clusHash",
                                          "codes", "codeOff", "codesMissing", "codesUnordered")] + \
               [("haveTable", C.c_int32), ("haveBins", C.c_int32), ("haveCodes", C.c_int32), ("reserved", C.c_int32)]


class CDistInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("blockBase", C.c_uint32),
                ("nBlocksGlobal", C.c_uint32), ("nReadsGlobal", C.c_uint64), ("nHashesGlobal", C.c_uint64),
                ("nLocalBins", C.c_uint32), ("reserved", C.c_uint32), ("localBinId", C.c_void_p),
                ("localCodeOff", C.c_void_p), ("localCodes", C.c_void_p)]


class CGood(C.Structure):
    _fields_ = [("nGood", C.c_uint64), ("hashNumber", C.c_uint32), ("nBlocksMax", C.c_uint32),
                ("within", C.c_void_p), ("goodOff", C.c_void_p), ("good", C.c_void_p)]


class CClusters(C.Structure):
    _fields_ = [("nBlocksMax", C.c_uint32), ("reserved", C.c_uint32), ("nHashes", C.c_uint64),
                ("nSubCluster", C.c_void_p), ("pointToMin", C.c_void_p), ("clusHash", C.c_void_p),
                ("msKernel", C.c_double)]


def lib_path():
    return os.path.join(_HERE, "libh10xgpu.so")


_lib = None


def load_library():
    """Load libh10xgpu.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise H10xError(7, "libh10xgpu.so is not built (run `make -C hash10x_b200/csrc` or "
                           "__graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    vp, u64, sz, cp = C.c_void_p, C.c_uint64, C.c_size_t, C.c_char_p
    L.h10x_abi_version.restype = C.c_int
    L.h10x_gpu_device_count.restype = C.c_int
    L.h10x_strerror.restype = cp
    L.h10x_strerror.argtypes = [C.c_int]
    L.h10x_stage_name.restype = cp
    L.h10x_stage_name.argtypes = [C.c_int]
    L.h10x_factor1_from_seed.restype = u64
    L.h10x_factor1_from_seed.argtypes = [C.c_int]
    L.h10x_gpu_create.restype = vp
    L.h10x_gpu_create.argtypes = [C.POINTER(Params), cp, sz]
    L.h10x_gpu_destroy.argtypes = [vp]
    L.h10x_gpu_build_device.argtypes = [vp, vp, u64, vp, cp, sz]
    L.h10x_gpu_index_device.argtypes = [vp, C.POINTER(CIndex)]
    L.h10x_gpu_download.argtypes = [vp, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_download_codes.argtypes = [vp, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_dist_global_codes.argtypes = [vp, cp, sz]
    L.h10x_gpu_fq2b.argtypes = [vp, vp, u64, vp, u64, vp, u64, C.c_uint32, C.POINTER(CFq2bOut), cp, sz]
    L.h10x_pack_barcode.argtypes = [cp, C.POINTER(C.c_uint32)]
    L.h10x_gpu_build_host.argtypes = [vp, vp, u64, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_build_file.argtypes = [vp, cp, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_stats.argtypes = [vp, C.POINTER(CStats)]
    L.h10x_gpu_index_digest.argtypes = [vp, u64, u64, C.c_int, C.POINTER(CDigest), cp, sz]
    L.h10x_index_free.argtypes = [C.POINTER(CIndex)]
    L.h10x_host_alloc.restype = vp
    L.h10x_host_alloc.argtypes = [sz]
    L.h10x_host_free.argtypes = [vp]
    L.h10x_gpu_record_moshes.argtypes = [vp, vp, u64, vp, vp, u64, cp, sz]
    L.h10x_dist_unique_id.argtypes = [vp, cp, sz]
    L.h10x_dist_init.argtypes = [vp, C.c_int, C.c_int, vp, cp, sz]
    L.h10x_gpu_build_device_dist.argtypes = [vp, vp, u64, vp, cp, sz]
    L.h10x_gpu_build_host_dist.argtypes = [vp, vp, u64, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_dist_info.argtypes = [vp, C.POINTER(CDistInfo)]
    L.h10x_gpu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.h10x_gpu_depth_range.argtypes = [vp, C.c_int, C.c_int, C.POINTER(CGood), cp, sz]
    L.h10x_gpu_load_index.argtypes = [vp, C.POINTER(CIndex), cp, sz]
    L.h10x_gpu_cluster.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(CClusters), cp, sz]
    L.h10x_gpu_cluster_split.argtypes = [vp, C.POINTER(CIndex), C.POINTER(C.c_uint32), cp, sz]
    L.h10x_gpu_depth_range_device.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint64), cp, sz]
    L.h10x_gpu_block_keys.argtypes = [vp, vp, u64, vp, vp, vp, u64, C.POINTER(C.c_uint32), C.POINTER(C.c_int), cp, sz]
    L.h10x_write_hash.argtypes = [C.POINTER(CIndex), cp]
    L.h10x_read_hash.argtypes = [cp, C.c_int32, C.POINTER(CIndex), cp, sz]
    _lib = L
    return L


def factor1_from_seed(seed=17):
    return int(load_library().h10x_factor1_from_seed(seed))


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class CFq2bOut(C.Structure):
    _fields_ = [("fqb", C.c_void_p), ("d_fqb", C.c_void_p), ("nRecords", C.c_uint64), ("nRead", C.c_uint64),
                ("recWords", C.c_uint32), ("s1Len", C.c_uint32), ("s2Len", C.c_uint32), ("reserved", C.c_uint32),
                ("nBad", C.c_uint64), ("nFixed", C.c_uint64), ("nFixBase", C.c_uint64 * 16)]


class Index:
    """Host copy of the state `--readFQB` leaves behind (see h10x_index in include/h10x_gpu.h)."""

    def __init__(self, ci, lib, keep_c=False):
        self.B, self.hashNumber, self.nBlocksMax = ci.B, ci.hashNumber, ci.nBlocksMax
        self.reserved = ci.reserved
        self.nReads, self.nHashes = ci.nReads, ci.nHashes
        hn, nb, H = ci.hashNumber, ci.nBlocksMax, ci.nHashes
        self.hashIndex = _arr(ci.hashIndex, 1 << ci.B, np.uint32) if ci.hashIndex else None
        self.hashValue = _arr(ci.hashValue, hn, np.uint64)
        self.hashDepth = _arr(ci.hashDepth, hn, np.uint32)
        self.blkNRead = _arr(ci.blkNRead, nb, np.uint32)
        self.blkNHash = _arr(ci.blkNHash, nb, np.uint32)
        self.blkOff = _arr(ci.blkOff, nb + 1, np.uint64)
        self.clus = _arr(ci.clusHash, H, np.uint64)
        self.codeOff = _arr(ci.codeOff, hn + 1, np.uint64) if ci.codeOff else None
        self.codes = _arr(ci.codes, H, np.uint32) if ci.codes else None
        self.blkNSub = _arr(ci.blkNSubCluster, nb, np.uint32) if ci.blkNSubCluster else None
        self.blkPointToMin = _arr(ci.blkPointToMin, nb, np.float64) if ci.blkPointToMin else None
        self.blkParent = _arr(ci.blkClusterParent, nb, np.uint32) if ci.blkClusterParent else None
        self.status = 0


class Hash10xGPU:
    """One build context on one GPU = the reference's initialise() (hash10x.c:1099-1118)."""

    def __init__(self, k=21, w=31, r=17, B=28, N=0, chunkSize=100000, device=0, flags=0,
                 factor1=None):
        self.lib = load_library()
        if factor1 is None:
            factor1 = DEFAULT_FACTOR1 if r == 17 else factor1_from_seed(r)
        self.params = Params(k, w, factor1, B, chunkSize, N, device, flags)
        err = C.create_string_buffer(512)
        self.ctx = self.lib.h10x_gpu_create(C.byref(self.params), err, len(err))
        if not self.ctx:
            msg = err.value.decode()
            code = 7 if "no CUDA device" in msg else (2 if "chunkSize" in msg else 3)
            raise H10xError(code, msg)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.h10x_gpu_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, err):
        if st != 0:
            raise H10xError(st, err.value.decode() or self.lib.h10x_strerror(st).decode())

    # --- the seam with host buffers (what the C host program calls) ---
    def build_host(self, recs, want_index=True):
        """recs: uint32 array of 30*n words (n FQB records) in host memory."""
        recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
        return self.build_host_ptr(recs.ctypes.data, recs.size // 30, want_index)

    def build_host_ptr(self, ptr, n, want_index=True):
        ci = CIndex()
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_build_host(self.ctx, ptr, n, C.byref(ci), err, len(err))
        self._check(st, err)
        try:
            return Index(ci, self.lib) if want_index else (ci.hashNumber, ci.nHashes, ci.nBlocksMax)
        finally:
            self.lib.h10x_index_free(C.byref(ci))

    def download_codes(self):
        """fillHashTable()'s lists of the resident index -> (codeOff u64, codes u32) host copies: what a context created
        with FLAG_LAZY_CODES leaves out of build_host / download"""
        ci = CIndex()
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_download_codes(self.ctx, C.byref(ci), err, len(err)), err)
        ix = CIndex()
        self.lib.h10x_gpu_index_device(self.ctx, C.byref(ix))
        code_off = _arr(ci.codeOff, ix.hashNumber + 1, np.uint64)
        return code_off, _arr(ci.codes, int(code_off[-1]), np.uint32)

    def fq2b(self, fq1, fq2=None, whitelist=None, sort=False, host=True):
        """fq2b (+ bsort when sort): FASTQ texts (bytes) -> (records uint32 [n, recWords] or None, stats, device pointer).
        whitelist: packed barcodes (uint32) in file order, or None"""
        co = CFq2bOut()
        err = C.create_string_buffer(512)
        wl = None if whitelist is None else np.ascontiguousarray(whitelist, dtype=np.uint32)
        b1 = C.create_string_buffer(fq1, len(fq1)) if len(fq1) else None
        b2 = None if fq2 is None else (C.create_string_buffer(fq2, len(fq2)) if len(fq2) else C.create_string_buffer(1))
        st = self.lib.h10x_gpu_fq2b(self.ctx, C.cast(b1, C.c_void_p), len(fq1), C.cast(b2, C.c_void_p), 0 if fq2 is None else len(fq2),
                                    None if wl is None else wl.ctypes.data, 0 if wl is None else wl.size,
                                    (1 if sort else 0) | (0 if host else 2), C.byref(co), err, len(err))
        self._check(st, err)
        stats = dict(nRead=co.nRead, nRecords=co.nRecords, nBad=co.nBad, nFixed=co.nFixed, nFixBase=list(co.nFixBase),
                     s1Len=co.s1Len, s2Len=co.s2Len, recWords=co.recWords)
        recs = None
        if host:
            recs = _arr(co.fqb, co.nRecords * co.recWords, np.uint32).reshape(co.nRecords, max(co.recWords, 1))
        return recs, stats, co.d_fqb

    def build_file(self, path):
        ci = CIndex()
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_build_file(self.ctx, path.encode(), C.byref(ci), err, len(err))
        self._check(st, err)
        try:
            return Index(ci, self.lib)
        finally:
            self.lib.h10x_index_free(C.byref(ci))

    def build_file_to_hash(self, fqb_path, hash_path):
        """--readFQB fqb --writeHash hash, entirely through the C ABI."""
        ci = CIndex()
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_build_file(self.ctx, fqb_path.encode(), C.byref(ci), err, len(err))
        self._check(st, err)
        try:
            st = self.lib.h10x_write_hash(C.byref(ci), hash_path.encode())
            if st:
                raise H10xError(st, "write fail")
            return ci.hashNumber, ci.nHashes, ci.nBlocksMax
        finally:
            self.lib.h10x_index_free(C.byref(ci))

    # --- device-resident build (inputs already in HBM; results stay in HBM) ---
    def build_device(self, dev_ptr, n_records, stream=0):
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_build_device(self.ctx, dev_ptr, n_records, stream, err, len(err))
        self._check(st, err)

    def download(self):
        ci = CIndex()
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_download(self.ctx, C.byref(ci), err, len(err))
        self._check(st, err)
        try:
            return Index(ci, self.lib)
        finally:
            self.lib.h10x_index_free(C.byref(ci))

    def device_index(self):
        ci = CIndex()
        st = self.lib.h10x_gpu_index_device(self.ctx, C.byref(ci))
        if st:
            raise H10xError(st, "no index resident")
        return ci

    # --- multi-GPU: one context per rank, NCCL inside the library ---
    @staticmethod
    def dist_unique_id():
        L = load_library()
        buf = C.create_string_buffer(128)
        err = C.create_string_buffer(512)
        st = L.h10x_dist_unique_id(buf, err, len(err))
        if st:
            raise H10xError(st, err.value.decode())
        return buf.raw

    def dist_init(self, rank, nranks, id_bytes):
        err = C.create_string_buffer(512)
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        self._check(self.lib.h10x_dist_init(self.ctx, rank, nranks, buf, err, len(err)), err)

    def build_device_dist(self, dev_ptr, n_records, stream=0):
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_build_device_dist(self.ctx, dev_ptr, n_records, stream, err, len(err)), err)

    def build_host_dist_ptr(self, ptr, n):
        """collective; this rank's records in host memory -> (hashNumber, nHashes, nBlocksMax)"""
        ci = CIndex()
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_build_host_dist(self.ctx, ptr, n, C.byref(ci), err, len(err)), err)
        return ci.hashNumber, ci.nHashes, ci.nBlocksMax

    def dist_global_codes(self):
        """collective: hashDepth and the whole hash->code CSR on every rank, so that depth_range / cluster work per rank"""
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_dist_global_codes(self.ctx, err, len(err)), err)

    def dist_info(self, download=True):
        di = CDistInfo()
        if self.lib.h10x_gpu_dist_info(self.ctx, C.byref(di)):
            raise H10xError(3, "no distributed build")
        out = {f: getattr(di, f) for f, _ in CDistInfo._fields_ if not f.startswith("local") and f != "reserved"}
        out["nLocalBins"] = di.nLocalBins
        if download:
            def pull(ptr, n):
                a = np.zeros(n, np.uint32)
                if n:
                    st = self.lib.h10x_gpu_memcpy_d2h(self.ctx, a.ctypes.data, ptr, 4 * n)
                    if st:
                        raise H10xError(st, "d2h")
                return a
            off = pull(di.localCodeOff, di.nLocalBins + 1)
            out["localBinId"] = pull(di.localBinId, di.nLocalBins)
            out["localCodeOff"] = off
            out["localCodes"] = pull(di.localCodes, int(off[-1]) if off.size else 0)
        return out

    def depth_range(self, dmin, dmax, copy=True):
        """--hashDepthRange on the resident index -> (within u8, goodOff u64, good u16) host copies;
        copy=False returns only the number of good hashes (the lists stay in the context's pinned arena)"""
        cg = CGood()
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_depth_range(self.ctx, dmin, dmax, C.byref(cg), err, len(err)), err)
        if not copy:
            return int(cg.nGood)
        return (_arr(cg.within, cg.hashNumber, np.uint8), _arr(cg.goodOff, cg.nBlocksMax + 1, np.uint64),
                _arr(cg.good, cg.nGood, np.uint16))

    def depth_range_device(self, dmin, dmax):
        """--hashDepthRange with the lists left on the device (what --cluster reads) -> number of good hashes"""
        n = C.c_uint64(0)
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_depth_range_device(self.ctx, dmin, dmax, C.byref(n), err, len(err)), err)
        return int(n.value)

    def load_index(self, ix):
        """make a host index (numpy arrays as in Index, with codeOff / codes) the resident one: --readHash's counterpart"""
        keep = [np.ascontiguousarray(a) for a in (ix.hashValue, ix.hashDepth, ix.blkNRead, ix.blkNHash, ix.blkOff, ix.clus,
                                                  ix.codeOff, ix.codes)]
        tab = np.ascontiguousarray(ix.hashIndex) if getattr(ix, "hashIndex", None) is not None else None
        ci = CIndex(ix.B, ix.hashNumber, ix.nBlocksMax, 0, ix.nReads, ix.nHashes,
                    tab.ctypes.data if tab is not None else None, keep[0].ctypes.data, keep[1].ctypes.data,
                    keep[2].ctypes.data, keep[3].ctypes.data, keep[4].ctypes.data, keep[5].ctypes.data,
                    keep[6].ctypes.data, keep[7].ctypes.data, 0, 0, None, None, None)
        nsub, ptm = getattr(ix, "blkNSub", None), getattr(ix, "blkPointToMin", None)
        if nsub is not None and ptm is not None:
            keep += [np.ascontiguousarray(nsub, np.uint32), np.ascontiguousarray(ptm, np.float64)]
            ci.blkNSubCluster, ci.blkPointToMin = keep[-2].ctypes.data, keep[-1].ctypes.data
        if getattr(ix, "blkParent", None) is not None:
            keep.append(np.ascontiguousarray(ix.blkParent, np.uint32))
            ci.blkClusterParent = keep[-1].ctypes.data
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_load_index(self.ctx, C.byref(ci), err, len(err)), err)

    def cluster_split(self):
        """--clusterSplit (clusterSplitCodes, hash10x.c:956-1013) on the resident index -> (Index with blkParent, blocks added)"""
        ci = CIndex()
        n_new = C.c_uint32(0)
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_cluster_split(self.ctx, C.byref(ci), C.byref(n_new), err, len(err)), err)
        return Index(ci, self.lib), int(n_new.value)

    def cluster(self, code_min=0, code_max=0, threshold=5, copy=True):
        """--cluster codeMin codeMax (-ct threshold) on the resident index and goodHashes ->
        (clus u64 with the subCluster bytes set, nSubCluster u32, pointToMin f64, kernel ms) host copies"""
        cc = CClusters()
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_cluster(self.ctx, code_min, code_max, threshold, C.byref(cc), err, len(err)), err)
        if not copy:
            return (None, _arr(cc.nSubCluster, cc.nBlocksMax, np.uint32), None, cc.msKernel)
        return (_arr(cc.clusHash, cc.nHashes, np.uint64), _arr(cc.nSubCluster, cc.nBlocksMax, np.uint32),
                _arr(cc.pointToMin, cc.nBlocksMax, np.float64), cc.msKernel)

    def stats(self):
        cs = CStats()
        self.lib.h10x_gpu_stats(self.ctx, C.byref(cs))
        d = {f: getattr(cs, f) for f, _ in CStats._fields_ if f != "msStage"}
        d["msStage"] = {self.lib.h10x_stage_name(i).decode(): cs.msStage[i] for i in range(H10X_NSTAGES)}
        return d

    def digest(self, block_base=0, entry_base=0, with_block_zero=True):
        """h10x_gpu_index_digest of the resident index -> dict of ints (see include/h10x_gpu.h)"""
        d = CDigest()
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_index_digest(self.ctx, block_base, entry_base, 1 if with_block_zero else 0,
                                                   C.byref(d), err, len(err)), err)
        return {f: int(getattr(d, f)) for f, _ in CDigest._fields_ if f != "reserved"}

    def block_keys(self, recs):
        """what the fused kernel stored per block -> (offsets[nBlocks+1], hashes u64, reads u32, lean)"""
        recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
        n = recs.size // 30
        cap = max(1, n * 240)
        off = np.zeros(n + 2, np.uint64)
        hs, rd = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        nb, lean = C.c_uint32(0), C.c_int(0)
        err = C.create_string_buffer(512)
        self._check(self.lib.h10x_gpu_block_keys(self.ctx, recs.ctypes.data, n, off.ctypes.data, hs.ctypes.data, rd.ctypes.data,
                                                 cap, C.byref(nb), C.byref(lean), err, len(err)), err)
        m = int(off[nb.value])
        return off[:nb.value + 1].copy(), hs[:m].copy(), rd[:m].copy(), bool(lean.value)

    def record_moshes(self, recs):
        """K1 alone: per-record mosh hashes in generation order -> (offsets[n+1], hashes)."""
        recs = np.ascontiguousarray(recs, dtype=np.uint32).reshape(-1)
        n = recs.size // 30
        off = np.zeros(n + 1, np.uint64)
        cap = max(1, n * 237)
        out = np.zeros(cap, np.uint64)
        err = C.create_string_buffer(512)
        st = self.lib.h10x_gpu_record_moshes(self.ctx, recs.ctypes.data, n, off.ctypes.data,
                                             out.ctypes.data, cap, err, len(err))
        self._check(st, err)
        return off, out[:int(off[n])]


def write_hash(index_arrays, path):
    """h10x_write_hash on numpy arrays (B, hashNumber, ... as in Index)."""
    L = load_library()
    ix = index_arrays
    keep = [np.ascontiguousarray(a) for a in (ix.hashIndex, ix.hashValue, ix.hashDepth, ix.blkNRead,
                                              ix.blkNHash, ix.blkOff, ix.clus)]
    ci = CIndex(ix.B, ix.hashNumber, ix.nBlocksMax, int(getattr(ix, "reserved", 0)), ix.nReads, ix.nHashes,
                keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data,
                keep[4].ctypes.data, keep[5].ctypes.data, keep[6].ctypes.data, None, None, 0, 0, None, None, None)
    nsub, ptm = getattr(ix, "blkNSub", None), getattr(ix, "blkPointToMin", None)
    if nsub is not None and ptm is not None:
        keep += [np.ascontiguousarray(nsub, np.uint32), np.ascontiguousarray(ptm, np.float64)]
        ci.blkNSubCluster, ci.blkPointToMin = keep[-2].ctypes.data, keep[-1].ctypes.data
    if getattr(ix, "blkParent", None) is not None:
        keep.append(np.ascontiguousarray(ix.blkParent, np.uint32))
        ci.blkClusterParent = keep[-1].ctypes.data
    st = L.h10x_write_hash(C.byref(ci), path.encode())
    if st:
        raise H10xError(st, "write fail")
