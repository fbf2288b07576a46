/* h10x_gpu.cu - device pipeline and C ABI of libh10xgpu.so (sm_100a).
 *
 * Replaces, on one B200, the reference's `--readFQB` branch (hash10x.c:1200-1205):
 *   readFQB (hash10x.c:188-236)  ->  stage "runs"     barcode-run detection over word 0
 *   processBlock (:154-186)      ->  stages "moshes"  k-mer hash + modulo selection (seqhash.c:58-80,154-195)
 *                                              "blocksort"/"dedup"  per-barcode sort + first-occurrence dedup
 *   hashIndexFind (:139-152)     ->  stages "hashsort"/"binids"  bin ids in (first block, hash) order,
 *                                              per-bin depth = number of blocks holding the hash
 *   clusHash sort (:176-183)     ->  stage  "clusters" per-block (bin id, read16) lists sorted by bin id
 *   fillHashTable (:317-347)     ->  stage  "codes"    hash->barcode CSR, ascending block numbers
 *   hashIndex[] layout (:142-148)->  stage  "table"    min-id displacement insertion = the sequential layout
 *
 * Data layout in HBM: see DESIGN.md.  No CPU fallback: every entry point fails without a device.
 */
#include "../../include/h10x_gpu.h"
#include "h10x_common.cuh"
#include "h10x_fused.cuh"
#include "h10x_cluster.cuh"
#include "h10x_bucket.cuh"
#include "h10x_subcluster.cuh"
#include "h10x_tail.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <iterator>
#include <map>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <deque>
#include <functional>

/* ------------------------------------------------------------------ errors / memory */

struct H10xError : std::runtime_error {
  int code ;
  H10xError (int c, const std::string &m) : std::runtime_error (m), code (c) {}
} ;

#define CK(call) do { cudaError_t e_ = (call) ; if (e_ != cudaSuccess) \
  throw H10xError (e_ == cudaErrorMemoryAllocation ? H10X_ERR_NOMEM : H10X_ERR_CUDA, \
		   std::string (#call) + ": " + cudaGetErrorString (e_) + " (" __FILE__ ":" + std::to_string (__LINE__) + ")") ; } while (0)

/* Device workspace: one cudaMalloc'ed slab per context with a first-fit free list kept on the host.
   Every buffer of a build is carved out of it, so a build makes no driver allocation calls
   (cudaMallocAsync pools were measured to cost 100s of ms per build while they grow and split).
   All work of a build is enqueued on one stream, so a block freed on the host can be handed out
   again at once: whatever used it was enqueued earlier on the same stream.  When the slab is too
   small the build is abandoned (SlabFull), the slab is re-allocated larger and the build is re-run. */
struct SlabFull { size_t need ; } ;

struct MemTrack {
  char *base = nullptr ; size_t cap = 0 ;
  size_t cur = 0, peak = 0 ;
  std::map<size_t, size_t> freeList ;	/* offset -> size, coalesced */
  void reset () { freeList.clear () ; if (cap) freeList[0] = cap ; cur = 0 ; }
  void *take (size_t bytes)
  { bytes = (bytes + 511) & ~(size_t) 511 ;
    for (auto it = freeList.begin () ; it != freeList.end () ; ++it)
      if (it->second >= bytes)
	{ size_t off = it->first, sz = it->second ;
	  freeList.erase (it) ;
	  if (sz > bytes) freeList[off + bytes] = sz - bytes ;
	  cur += bytes ; if (cur > peak) peak = cur ;
	  return base + off ;
	}
    throw SlabFull { cur + bytes } ;
  }
  size_t largestFree () const { size_t m = 0 ; for (auto &f : freeList) m = std::max (m, f.second) ; return m ; }
  void give (void *p, size_t bytes)
  { bytes = (bytes + 511) & ~(size_t) 511 ;
    size_t off = (size_t) ((char*) p - base) ;
    cur -= bytes ;
    auto nx = freeList.lower_bound (off) ;
    if (nx != freeList.end () && off + bytes == nx->first) { bytes += nx->second ; nx = freeList.erase (nx) ; }
    if (nx != freeList.begin ())
      { auto pv = std::prev (nx) ;
	if (pv->first + pv->second == off) { pv->second += bytes ; return ; }
      }
    freeList[off] = bytes ;
  }
} ;

template <class T> struct DBuf {
  T *p = nullptr ; size_t n = 0 ; MemTrack *mt = nullptr ;
  DBuf () {}
  DBuf (size_t n_, cudaStream_t, MemTrack *mt_) { alloc (n_, 0, mt_) ; }
  DBuf (const DBuf&) = delete ; DBuf &operator= (const DBuf&) = delete ;
  void alloc (size_t n_, cudaStream_t, MemTrack *mt_)
  { release () ; mt = mt_ ; n = n_ ;
    p = (T*) mt->take ((n ? n : 1) * sizeof (T)) ;
  }
  void release ()
  { if (p) { mt->give (p, (n ? n : 1) * sizeof (T)) ; p = nullptr ; n = 0 ; } }
  void swap (DBuf &o) { std::swap (p, o.p) ; std::swap (n, o.n) ; std::swap (mt, o.mt) ; }
  ~DBuf () { release () ; }
} ;

static const char *kStageNames[H10X_NSTAGES] = {
  "runs", "moshes", "blocksort", "dedup", "hashsort", "binids", "entryids", "codes",
  "clusters", "table", "fused", "other" } ;
enum { ST_RUNS = 0, ST_MOSHES, ST_BLOCKSORT, ST_DEDUP, ST_HASHSORT, ST_BINIDS, ST_ENTRYIDS, ST_CODES,
       ST_CLUSTERS, ST_TABLE, ST_FUSED, ST_OTHER } ;

struct h10x_ctx {
  h10x_params P ;
  HashParams hp ;
  cudaStream_t own = 0 ;
  MemTrack mt ;
  /* resident result of the last build */
  DBuf<uint32_t> hashIndex, hashDepth, blkNRead, blkNHash, codes ;
  DBuf<uint64_t> hashValue, blkOff, codeOff, clus ;
  uint32_t hashNumber = 1, nBlocksMax = 2 ;
  uint64_t nReads = 0, nHashes = 0 ;
  bool haveIndex = false ;
  h10x_stats stats ;
  /* stage timing */
  std::vector<cudaEvent_t> evPool ; size_t evUsed = 0 ;
  struct Span { int stage ; cudaEvent_t a, b ; } ;
  std::vector<Span> spans ;
  uint64_t launches = 0 ;
  /* multi-GPU (h10x_dist.cuh) */
  bool slabClamped = false ;	/* the slab already takes all free device memory */
  /* host-buffer builds start the D2H of an index array as soon as it is final, on a second stream */
    /* h10x_gpu_histogram / h10x_gpu_crib_build results (host) */
  std::vector<int> histHost, cribHist ; std::vector<uint8_t> cribType ; std::vector<int16_t> cribChr ; std::vector<uint16_t> cribPos ;
  /* h10x_gpu_fq2b: the records of the last call and the whitelist table (2^32 words, kept while the whitelist is the same) */
  uint32_t *fqRecs = nullptr ; void *fqHost = nullptr ; size_t fqHostCap = 0 ;
  uint32_t *wlTable = nullptr ; uint64_t wlCount = 0, wlSum = 0 ;
  cudaStream_t ulStream = 0 ;	/* ... and copy the file up in slabs on a third one while the fused kernel hashes the slabs that landed */
  bool earlyDl = false ; cudaStream_t dlStream = 0 ; bool slotDone[9] = { false, false, false, false, false, false, false, false, false } ;
  DBuf<uint8_t> within ;	/* --hashDepthRange flags per bin; only ever set (hash10x.c:535) until the next build */
  /* goodHashes of the last --hashDepthRange (hash10x.c:722-766), resident for --cluster; ClusterBlock.nSubCluster /
     .pointToMin (hash10x.c:62-70) once a --cluster command ran */
  DBuf<uint64_t> goodOffD ; DBuf<uint16_t> goodD ; bool haveGood = false ;
  DBuf<uint32_t> blkNSub, blkParent ; DBuf<double> blkPtm ;
  /* tests: the key lists the mosh stage left per block, kept when dbgKeys is set (h10x_gpu_block_keys) */
  bool dbgKeys = false, dbgLean = false ; uint32_t dbgNBlk = 0 ; uint64_t dbgScratchLen = 0 ;
  std::vector<uint64_t> dbgSrcOff, dbgScratch ; std::vector<uint32_t> dbgBlkCnt ;
  void *clusSlot[3] = { nullptr, nullptr, nullptr } ; size_t clusCap[3] = { 0, 0, 0 } ;	/* pinned host: nSubCluster, pointToMin, clusterParent */
  void *goodSlot[3] = { nullptr, nullptr, nullptr } ; size_t goodCap[3] = { 0, 0, 0 } ;	/* pinned host: within, goodOff, good */
  struct DistState *dist = nullptr ;
  DBuf<uint32_t> localBinId, localCodeOff, localCodes ;	/* this rank's part of the hash->code lists */
  /* pinned host arena reused by h10x_gpu_download (one slot per index array) */
  void *hostSlot[9] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr } ;
  size_t hostCap[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 } ;
} ;

static cudaEvent_t ctx_event (h10x_ctx *c)
{ if (c->evUsed == c->evPool.size ()) { cudaEvent_t e ; CK (cudaEventCreate (&e)) ; c->evPool.push_back (e) ; }
  return c->evPool[c->evUsed++] ;
}

enum { SLOT_INDEX = 0, SLOT_VALUE, SLOT_DEPTH, SLOT_NREAD, SLOT_NHASH, SLOT_BLKOFF, SLOT_CLUS, SLOT_CODEOFF, SLOT_CODES } ;

static void *host_slot (h10x_ctx *c, int i, size_t bytes)
{ if (c->hostCap[i] < bytes || !c->hostSlot[i])
    { if (c->hostSlot[i]) cudaFreeHost (c->hostSlot[i]) ;
      c->hostSlot[i] = nullptr ; c->hostCap[i] = 0 ;
      void *p = nullptr ;
      CK (cudaHostAlloc (&p, bytes ? bytes : 1, cudaHostAllocDefault)) ;
      c->hostSlot[i] = p ; c->hostCap[i] = bytes ? bytes : 1 ;
    }
  return c->hostSlot[i] ;
}

/* during a host-buffer build: copy a finished index array to its pinned host slot behind an event,
   while later stages keep the SMs busy */
static void early_pull (h10x_ctx *c, cudaStream_t s, int slot, const void *src, size_t bytes)
{ if (!c->earlyDl || !src) return ;
  void *dst = host_slot (c, slot, bytes) ;
  cudaEvent_t ready = ctx_event (c) ;
  CK (cudaEventRecord (ready, s)) ;
  CK (cudaStreamWaitEvent (c->dlStream, ready, 0)) ;
  if (bytes) CK (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToHost, c->dlStream)) ;
  c->slotDone[slot] = true ;
}

struct StageTimer {
  h10x_ctx *c ; cudaStream_t s ; int stage ; cudaEvent_t a ;
  StageTimer (h10x_ctx *c_, cudaStream_t s_, int st) : c (c_), s (s_), stage (st)
  { a = ctx_event (c) ; cudaEventRecord (a, s) ; }
  ~StageTimer () { cudaEvent_t b = ctx_event (c) ; cudaEventRecord (b, s) ; c->spans.push_back ({stage, a, b}) ; }
} ;

#define LAUNCH(c, kernel, grid, block, smem, stream, ...) do { \
  kernel<<<(grid), (block), (smem), (stream)>>> (__VA_ARGS__) ; ++(c)->launches ; CK (cudaGetLastError ()) ; } while (0)

static inline unsigned gridFor (uint64_t n, unsigned block) { return (unsigned) std::max<uint64_t> (1, (n + block - 1) / block) ; }

template <class F> static void cubCall (h10x_ctx *c, cudaStream_t s, F f)
{ size_t bytes = 0 ;
  CK (f (nullptr, bytes)) ;
  DBuf<char> tmp (bytes, s, &c->mt) ;
  CK (f ((void*) tmp.p, bytes)) ;
  c->launches += 1 ;	/* counted as one library call */
}

struct CastU64 { __host__ __device__ uint64_t operator() (uint32_t x) const { return (uint64_t) x ; } } ;

#include "h10x_dist.cuh"
#include "h10x_fq2b.cuh"
#include "h10x_crib.cuh"
#include "h10x_split.cuh"

/* ------------------------------------------------------------------ kernels: runs */

/* flag[i] = 1 when record i starts a new barcode run (hash10x.c:213-220); notes whether any
   record carries barcode word 0, the only case in which the reference's chunk loop can merge runs */
__global__ void k_run_flags (const uint32_t *__restrict__ fqb, uint32_t n, uint32_t *__restrict__ flag,
			     int *__restrict__ anyZero)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  uint32_t w0 = fqb[(size_t) H10X_REC_WORDS * i] ;
  uint32_t prev = i ? fqb[(size_t) H10X_REC_WORDS * (i - 1)] : ~w0 ;
  flag[i] = (w0 != prev) ? 1u : 0u ;
  if (w0 == 0) *anyZero = 1 ;
}

/* the same for records r0 .. r0 + n - 1 of a file whose earlier records are resident: flag[i - r0] */
__global__ void k_run_flags_at (const uint32_t *__restrict__ fqb, uint32_t r0, uint32_t n, uint32_t *__restrict__ flag,
				int *__restrict__ anyZero)
{ uint32_t j = blockIdx.x * blockDim.x + threadIdx.x ;
  if (j >= n) return ;
  const uint32_t i = r0 + j ;
  uint32_t w0 = fqb[(size_t) H10X_REC_WORDS * i] ;
  uint32_t prev = i ? fqb[(size_t) H10X_REC_WORDS * (i - 1)] : ~w0 ;
  flag[j] = (w0 != prev) ? 1u : 0u ;
  if (w0 == 0) *anyZero = 1 ;
}

/* blkStart[b] = first record of 0-based run b; blkStart[nRuns] = n */
__global__ void k_run_starts (const uint32_t *__restrict__ flag, const uint32_t *__restrict__ incl, uint32_t n,
			      uint32_t *__restrict__ blkStart)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  if (flag[i]) blkStart[incl[i] - 1] = i ;
  if (i == n - 1) blkStart[incl[i]] = n ;
}

__global__ void k_flags_from_starts (const uint32_t *__restrict__ blkStart, uint32_t nBlk, uint32_t *__restrict__ flag)
{ uint32_t b = blockIdx.x * blockDim.x + threadIdx.x ;
  if (b < nBlk) flag[blkStart[b]] = 1u ;
}

__global__ void k_gather_word0 (const uint32_t *__restrict__ fqb, const uint32_t *__restrict__ blkStart, uint32_t nBlk,
				uint32_t *__restrict__ out)
{ uint32_t b = blockIdx.x * blockDim.x + threadIdx.x ;
  if (b < nBlk) out[b] = fqb[(size_t) H10X_REC_WORDS * blkStart[b]] ;
}

/* ------------------------------------------------------------------ kernels: generic mosh path */

/* one thread per read (2 per record).  EMIT=false counts, EMIT=true writes (hash, record) pairs at
   the offsets the scan of the counts produced, so generation order (read index ascending, read 1
   before read 2, position ascending - hash10x.c:160-164) is the memory order. */
template <bool EMIT>
__global__ void k_moshes (const uint32_t *__restrict__ fqb, uint32_t r0, uint32_t nb, HashParams hp,
			  uint32_t *__restrict__ cnt2, const uint32_t *__restrict__ blkIncl, uint32_t p0,
			  const uint32_t *__restrict__ phOff, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{ uint32_t t = blockIdx.x * blockDim.x + threadIdx.x ;
  if (t >= 2 * nb) return ;
  uint32_t rec = r0 + (t >> 1) ; int rd = t & 1 ;
  const uint32_t *src = fqb + (size_t) H10X_REC_WORDS * rec + 15 * rd ;
  uint32_t u[10] ;
#pragma unroll
  for (int i = 0 ; i < 10 ; ++i) u[i] = src[i] ;
  uint32_t n = 0 ;
  uint32_t base = 0 ;
  if (EMIT) base = cnt2[t] + phOff[blkIncl[rec] - 1 - p0] ;
  h10x_scan_read (u, rd ? H10X_R2_START : H10X_R1_START, rd ? H10X_R2_LEN : H10X_R1_LEN, hp,
		  [&] (uint64_t x) { if (EMIT) { keys[base + n] = x ; vals[base + n] = rec ; } ++n ; }) ;
  if (!EMIT) cnt2[t] = n ;
}

/* per block of the batch: number of moshes; ph = 1 for a block with none (it gets the phantom
   {hash 0, read 0} entry of hash10x.c:167-168) */
__global__ void k_blk_moshes (const uint32_t *__restrict__ blkStart, uint32_t p0, uint32_t nblk, uint32_t r0,
			      const uint32_t *__restrict__ off2, uint32_t *__restrict__ ph)
{ uint32_t p = blockIdx.x * blockDim.x + threadIdx.x ;
  if (p > nblk) return ;
  if (p == nblk) { ph[p] = 0 ; return ; }
  uint32_t a = off2[2 * (blkStart[p0 + p] - r0)], b = off2[2 * (blkStart[p0 + p + 1] - r0)] ;
  ph[p] = (a == b) ? 1u : 0u ;
}

__global__ void k_seg_off (const uint32_t *__restrict__ blkStart, uint32_t p0, uint32_t nblk, uint32_t r0,
			   const uint32_t *__restrict__ off2, const uint32_t *__restrict__ phOff,
			   const uint32_t *__restrict__ ph, uint32_t *__restrict__ segOff,
			   uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{ uint32_t p = blockIdx.x * blockDim.x + threadIdx.x ;
  if (p > nblk) return ;
  uint32_t so = off2[2 * (blkStart[p0 + p] - r0)] + phOff[p] ;
  segOff[p] = so ;
  if (keys && p < nblk && ph[p]) { keys[so] = 0 ; vals[so] = blkStart[p0 + p] ; }
}

/* first occurrence of each hash within its block, after the stable sort by hash */
__global__ void k_uniq_flag (const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t m,
			     const uint32_t *__restrict__ blkIncl, uint32_t *__restrict__ flag)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= m) return ;
  bool f = (i == 0) || keys[i] != keys[i-1] || blkIncl[vals[i]] != blkIncl[vals[i-1]] ;
  flag[i] = f ? 1u : 0u ;
}

__global__ void k_compact (const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t m,
			   const uint32_t *__restrict__ flag, const uint32_t *__restrict__ incl, uint64_t eBase,
			   uint64_t *__restrict__ eHash, uint32_t *__restrict__ eRec)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= m || !flag[i]) return ;
  uint64_t e = eBase + incl[i] - 1 ;
  eHash[e] = keys[i] ; eRec[e] = vals[i] ;
}

__global__ void k_blk_off (const uint32_t *__restrict__ segOff, const uint32_t *__restrict__ incl, uint32_t nblk,
			   uint64_t eBase, uint64_t *__restrict__ blkOffProc)
{ uint32_t p = blockIdx.x * blockDim.x + threadIdx.x ;
  if (p < nblk) blkOffProc[p] = eBase + incl[segOff[p]] - 1 ;
}

/* ------------------------------------------------------------------ kernels: bins */

__global__ void k_iota (uint32_t *__restrict__ a, uint64_t n)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; if (i < n) a[i] = (uint32_t) i ; }

__global__ void k_head_flag (const uint64_t *__restrict__ sh, uint64_t n, uint32_t *__restrict__ head)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) head[i] = (i == 0 || sh[i] != sh[i-1]) ? 1u : 0u ;
}

/* entries are stored in (block, hash) order, so the reference's insertion order of bins
   (hash10x.c:147: first block that holds the hash, then hash value) is the order of each bin's
   smallest entry index.  isFirst marks those entries. */
__global__ void k_seg_start (const uint32_t *__restrict__ head, const uint32_t *__restrict__ segIncl, uint64_t n,
			     const uint32_t *__restrict__ se, uint32_t *__restrict__ segStart,
			     uint32_t *__restrict__ isFirst)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  if (head[i]) { segStart[segIncl[i] - 1] = (uint32_t) i ; if (isFirst) isFirst[se[i]] = 1u ; }
  if (i == n - 1) segStart[segIncl[i]] = (uint32_t) n ;
}


__global__ void k_entry_ids (uint64_t n, const uint32_t *__restrict__ segIncl, const uint32_t *__restrict__ idOfSeg,
			     const uint32_t *__restrict__ se, uint32_t *__restrict__ entryId)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) entryId[se[i]] = idOfSeg[segIncl[i] - 1] ;
}


/* single-GPU tail.  Bin ids in insertion order (hash10x.c:147) = bins ordered by their first entry
   (entries are stored in (block, hash) order): firstE[s] is sorted, the rank is the id. */
__global__ void k_first_entry (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint32_t *__restrict__ se,
			       uint32_t *__restrict__ firstE, uint32_t *__restrict__ segIdx)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ;
  if (s < nSeg) { firstE[s] = se[segStart[s]] ; segIdx[s] = s ; }
}

__global__ void k_bins_by_rank (uint32_t nSeg, const uint32_t *__restrict__ sortedSeg, const uint32_t *__restrict__ segStart,
				const uint64_t *__restrict__ sh, uint64_t wMul, uint32_t *__restrict__ idOfSeg,
				uint64_t *__restrict__ hashValue, uint32_t *__restrict__ hashDepth)
{ uint32_t r = blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= nSeg) return ;
  uint32_t s = sortedSeg[r], id = r + 1u, i = segStart[s] ;
  idOfSeg[s] = id ;
  hashValue[id] = sh[i] * wMul ;		/* the sort key is hash / w */
  hashDepth[id] = segStart[s+1] - i ;	/* one entry per (block, hash): hash10x.c:178 */
}

/* fillHashTable (hash10x.c:317-347) and, in the same pass, the transposed view: entry (bin id, read) laid
   out bin-major next to its block number.  A stable sort of that by block number then yields every
   block's ClusterHash list already ordered by bin id (hash10x.c:183) - no per-block sort, no scatter of
   ids back to entry order. */
__global__ void k_codes_tr (uint64_t n, const uint32_t *__restrict__ segIncl, const uint32_t *__restrict__ segStart,
			    const uint32_t *__restrict__ idOfSeg, const uint32_t *__restrict__ se,
			    const uint64_t *__restrict__ eBR, const uint64_t *__restrict__ codeOff,
			    uint32_t *__restrict__ codes, uint64_t *__restrict__ idRead)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  uint32_t s = segIncl[i] - 1 ;
  uint32_t id = idOfSeg[s] ;
  uint64_t pos = codeOff[id] + (i - segStart[s]) ;
  uint64_t br = eBR[se[i]] ;
  codes[pos] = (uint32_t) br ;
  idRead[pos] = (uint64_t) id | (br & 0xffff00000000ull) ;
}


__global__ void k_clus_pack (uint64_t n, const uint32_t *__restrict__ ids, const uint16_t *__restrict__ read16,
			     uint64_t *__restrict__ clus)
{ uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (e < n) clus[e] = (uint64_t) ids[e] | ((uint64_t) read16[e] << 32) ;
}

__global__ void k_shift_offsets (const uint64_t *__restrict__ in, uint64_t base, uint32_t n, uint32_t *__restrict__ out)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) out[i] = (uint32_t) (in[i] - base) ;
}

/* ---- --hashDepthRange on the resident index ("next" row f1): hashWithinRangeBuild hash10x.c:528-539 ---- */
__global__ void k_within (uint32_t hashNumber, const uint32_t *__restrict__ depth, int dmin, int dmax, uint8_t *__restrict__ within)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= hashNumber) return ;
  int n = (int) depth[i] ;
  if (n >= dmin && n < dmax) within[i] = 1 ;	/* flags are only ever set */
}

/* goodHashesBuild hash10x.c:738-766, part 1: which entries of each block are good (bin within range);
   blocks with more than 65535 hashes are skipped (:748).  One CTA per block. */
__global__ void k_good_mark (uint32_t nBlocksMax, const uint64_t *__restrict__ blkOff, const uint32_t *__restrict__ blkNHash,
			     const uint64_t *__restrict__ clus, const uint8_t *__restrict__ within, uint32_t *__restrict__ flag)
{ for (uint32_t b = blockIdx.x ; b < nBlocksMax ; b += gridDim.x)
    { const uint32_t n = b ? blkNHash[b] : 0 ;
      const uint64_t off = blkOff[b] ;
      const bool skip = n > 65535u ;
      for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	flag[off + i] = (!skip && within[(uint32_t) clus[off + i]]) ? 1u : 0u ;
    }
}

/* part 2: compact (depth of the bin, index inside the block) of the good entries, block after block */
__global__ void k_good_compact (uint32_t nBlocksMax, const uint64_t *__restrict__ blkOff, const uint32_t *__restrict__ blkNHash,
				const uint64_t *__restrict__ clus, const uint32_t *__restrict__ depth,
				const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
				uint32_t *__restrict__ keyDepth, uint16_t *__restrict__ valIdx, uint64_t *__restrict__ goodOff,
				uint64_t nEntries, uint32_t nGood)
{ for (uint32_t b = blockIdx.x ; b <= nBlocksMax ; b += gridDim.x)
    { if (b == nBlocksMax) { if (threadIdx.x == 0) goodOff[b] = nGood ; continue ; }
      const uint32_t n = b ? blkNHash[b] : 0 ;
      const uint64_t off = blkOff[b] ;
      if (threadIdx.x == 0) goodOff[b] = off < nEntries ? pos[off] : nGood ;
      for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	if (flag[off + i]) { uint32_t p = pos[off + i] ; keyDepth[p] = depth[(uint32_t) clus[off + i]] ; valIdx[p] = (uint16_t) i ; }
    }
}

/* hashIndex[] (hash10x.c:139-152).  Sequential insertion puts bin n at the first slot of its probe
   sequence not held by a smaller id.  That fixed point is unique, so it can be reached in parallel:
   a bin claims a slot that is empty or holds a larger id (compare-and-swap); the displaced larger
   id is re-inserted from the start of its own sequence. */
__global__ void k_table_insert (uint32_t hashNumber, const uint64_t *__restrict__ hashValue,
				uint32_t *table, int B)
{ uint32_t id = 1u + blockIdx.x * blockDim.x + threadIdx.x ;
  if (id >= hashNumber) return ;
  const uint64_t mask = ((uint64_t) 1 << B) - 1 ;
  uint32_t cur = id ;
  for (;;)
    { uint64_t hv = hashValue[cur] ;
      uint64_t off = hv & mask, diff = ((hv >> B) & mask) | 1 ;
      uint32_t displaced = 0 ;
      for (;;)
	{ uint32_t old = *(volatile uint32_t*) &table[off] ;
	  for (;;)
	    { if (old != 0 && old < cur) break ;		/* held by a smaller id: probe on */
	      uint32_t prev = atomicCAS (&table[off], old, cur) ;
	      if (prev == old) { displaced = old ; old = 0xffffffffu ; break ; }
	      old = prev ;
	    }
	  if (old == 0xffffffffu) break ;
	  off = (off + diff) & mask ;
	}
      if (!displaced) return ;
      cur = displaced ;
    }
}

/* every (block, bin) of the ClusterHash lists must stand in the bin's barcode list: one CTA per block, a binary
   search per entry (the lists are ascending); counts the misses */
__global__ void k_check_codes (uint32_t nBlocksMax, const uint64_t *__restrict__ blkOff, const uint32_t *__restrict__ blkNHash,
			       const uint64_t *__restrict__ clus, const uint64_t *__restrict__ codeOff, const uint32_t *__restrict__ codes,
			       unsigned long long *__restrict__ missing)
{ unsigned long long bad = 0 ;
  for (uint32_t b = 1 + blockIdx.x ; b < nBlocksMax ; b += gridDim.x)
    { const uint32_t n = blkNHash[b] ;
      const uint64_t off = blkOff[b] ;
      for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	{ const uint32_t id = (uint32_t) clus[off + i] ;
	  uint64_t lo = codeOff[id], hi = codeOff[id + 1] ;
	  while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1) ; if (codes[mid] < b) lo = mid + 1 ; else hi = mid ; }
	  if (lo >= codeOff[id + 1] || codes[lo] != b) ++bad ;
	}
    }
  if (bad) atomicAdd (missing, bad) ;
}

/* each bin's list strictly ascending and as long as its depth */
__global__ void k_check_ascending (const uint64_t *__restrict__ codeOff, const uint32_t *__restrict__ codes,
				   const uint32_t *__restrict__ depth, uint32_t hashNumber, unsigned long long *__restrict__ unordered)
{ unsigned long long bad = 0 ;
  for (uint64_t id = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; id < hashNumber ; id += (uint64_t) gridDim.x * blockDim.x)
    { const uint64_t a = codeOff[id], b = codeOff[id + 1] ;
      if (b - a != depth[id]) ++bad ;
      for (uint64_t i = a + 1 ; i < b ; ++i) if (codes[i] <= codes[i - 1]) ++bad ;
    }
  if (bad) atomicAdd (unordered, bad) ;
}

/* ------------------------------------------------------------------ host: parameters */

static uint64_t inv64 (uint64_t a)	/* inverse of odd a modulo 2^64 (Newton) */
{ uint64_t x = a ; for (int i = 0 ; i < 6 ; ++i) x *= 2 - a * x ; return x ; }

static void make_hash_params (const h10x_params &P, HashParams &hp)
{ hp.k = P.k ; hp.w = P.w ; hp.factor1 = P.factor1 ;
  hp.shift = 64 - 2 * P.k ;
  hp.kmask = (((uint64_t) 1) << (2 * P.k)) - 1 ;
  hp.rcShift = 2 * (P.k - 1) ;
  uint64_t w = (uint64_t) P.w ; int tz = 0 ;
  while (!(w & 1)) { w >>= 1 ; ++tz ; }
  hp.wTz = tz ; hp.wTzMask = (((uint64_t) 1) << tz) - 1 ;
  hp.wInv = inv64 (w) ; hp.wLim = ~(uint64_t) 0 / w ;
}

/* ------------------------------------------------------------------ host: the build */

struct BlockTable {	/* 0-based barcode blocks of the consumed records */
  std::vector<uint32_t> start ;	/* nBlk + 1 */
  uint32_t nBlk = 0 ;
} ;

/* The reference's chunk loop (hash10x.c:202-223) at run granularity.  Needed only when some run
   carries barcode word 0: `if (!barcode) barcode = u[0]` at a chunk start then re-seeds the barcode
   and glues the open all-A run onto the run that follows.  Also decides "chunkSize too small". */
static int simulate_chunks (const std::vector<uint32_t> &runStart, const std::vector<uint32_t> &runWord,
			    uint32_t nRuns, uint64_t nRec, int chunkSize, int64_t N, BlockTable &bt)
{
  bt.start.clear () ;
  uint64_t pos = 0, nReads = 0, curN = 0 ;
  uint32_t barcode = 0, r = 0 ;	/* r = run containing pos */
  bt.start.push_back (0) ;		/* block 1 exists from the start with nRead 0 */
  while (!N || nReads < (uint64_t) N)
    { int64_t thisChunk = (int64_t) chunkSize - (int64_t) curN ;
      if (thisChunk <= 0) return H10X_ERR_CHUNK_TOO_SMALL ;
      if (N && nReads + (uint64_t) thisChunk > (uint64_t) N) thisChunk = N - (int64_t) nReads ;
      uint64_t n = std::min<uint64_t> ((uint64_t) thisChunk, nRec - pos) ;
      if (!n) break ;
      if (!barcode) barcode = runWord[r] ;
      uint64_t end = pos + n ;
      while (pos < end)
	{ uint64_t segEnd = std::min<uint64_t> (end, runStart[r+1]) ;
	  uint64_t len = segEnd - pos ;
	  if (runWord[r] == barcode) curN += len ;
	  else { bt.start.push_back ((uint32_t) pos) ; curN = len ; barcode = runWord[r] ; }
	  pos = segEnd ;
	  if (pos == runStart[r+1] && r + 1 < nRuns) ++r ;
	}
      nReads += n ;
    }
  bt.nBlk = (uint32_t) bt.start.size () ;
  bt.start.push_back ((uint32_t) nRec) ;
  return H10X_OK ;
}

static void reset_result (h10x_ctx *c)
{ c->hashIndex.release () ; c->hashDepth.release () ; c->blkNRead.release () ; c->blkNHash.release () ;
  c->localBinId.release () ; c->localCodeOff.release () ; c->localCodes.release () ; c->within.release () ;
  c->goodOffD.release () ; c->goodD.release () ; c->haveGood = false ; c->blkNSub.release () ; c->blkPtm.release () ; c->blkParent.release () ;
  c->codes.release () ; c->hashValue.release () ; c->blkOff.release () ; c->codeOff.release () ; c->clus.release () ;
  c->hashNumber = 1 ; c->nBlocksMax = 2 ; c->nReads = 0 ; c->nHashes = 0 ; c->haveIndex = false ;
  c->spans.clear () ; c->evUsed = 0 ; c->launches = 0 ; c->mt.peak = c->mt.cur ;
  for (int i = 0 ; i < 9 ; ++i) c->slotDone[i] = false ;
  memset (&c->stats, 0, sizeof (c->stats)) ;
}

/* segmented sort of (key, val) pairs whose segments are the barcode blocks; batched so that each
   CUB call stays inside its `int num_items` interface */
template <class K, class V>
static void segmented_sort_blocks (h10x_ctx *c, cudaStream_t s, const K *kin, K *kout, const V *vin, V *vout,
				   const std::vector<uint64_t> &hOff /* nSeg+1, host */, const uint64_t *dOff,
				   bool stable, size_t segFirst = 0, size_t segLast = (size_t) -1)
{
  const uint64_t maxItems = (uint64_t) 1 << 30 ;
  size_t nSeg = std::min (hOff.size () - 1, segLast), a = segFirst ;
  while (a < nSeg)
    { size_t b = a + 1 ;
      while (b < nSeg && hOff[b+1] - hOff[a] <= maxItems) ++b ;
      uint64_t base = hOff[a], items = hOff[b] - base ;
      if (items > 0x7fffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "a single barcode block exceeds 2^31 entries") ;
      if (items)
	{ uint32_t ns = (uint32_t) (b - a) ;
	  DBuf<uint32_t> rel (ns + 1, s, &c->mt) ;
	  LAUNCH (c, k_shift_offsets, gridFor (ns + 1, 256), 256, 0, s, dOff + a, base, ns + 1, rel.p) ;
	  cubCall (c, s, [&] (void *t, size_t &bytes)
	    { return stable
		? cub::DeviceSegmentedSort::StableSortPairs (t, bytes, kin + base, kout + base, vin + base, vout + base,
							      (int) items, (int) ns, rel.p, rel.p + 1, s)
		: cub::DeviceSegmentedSort::SortPairs (t, bytes, kin + base, kout + base, vin + base, vout + base,
						       (int) items, (int) ns, rel.p, rel.p + 1, s) ; }) ;
	}
      a = b ;
    }
}

/* H10X_TRACE=1: host wall-clock between checkpoints of a build, on stderr */
struct HostTrace {
  bool on ; std::chrono::steady_clock::time_point t0 ;
  HostTrace () : on (getenv ("H10X_TRACE") != nullptr), t0 (std::chrono::steady_clock::now ()) {}
  void mark (const char *what)
  { if (!on) return ;
    auto t1 = std::chrono::steady_clock::now () ;
    fprintf (stderr, "h10x-trace %-14s %9.3f ms\n", what, std::chrono::duration<double, std::milli> (t1 - t0).count ()) ;
    t0 = t1 ;
  }
} ;

/* ------------------------------------------------------------------ host: the hand-written tail (h10x_tail.cuh) */

static int bits_for (uint64_t maxValue) { int b = 1 ; while (b < 64 && (maxValue >> b)) ++b ; return b ; }

/* sub-range / range geometry for H entries whose q = hash / w is at most `top`; false = the packed 64-bit entry
   does not fit or the keys are too coarse to be cut into shared-memory pieces (the library path runs instead) */
static bool tail_geometry (uint64_t H, uint64_t top, uint32_t maxBlock, TailGeom &g)
{ g.sortBits = bits_for (top) ;
  g.blkBits = bits_for (maxBlock) ;
  g.eShift = g.blkBits + 16 ;
  int rem = g.sortBits ;
  double target = H10X_SR_TARGET ;
  if (const char *e = getenv ("H10X_SR_TARGET")) { double v = atof (e) ; if (v >= 64.0 && v <= 0.35 * H10X_SR_CAP) target = v ; }
  while (rem > 0 && (double) H / (double) ((top >> rem) + 1) > target) --rem ;
  if ((double) H / (double) ((top >> rem) + 1) > 0.5 * H10X_SR_CAP) return false ;	/* a handful of hash values: nothing to cut */
  int p2 = 0 ;
  for (;;)		/* ranges x digits must cover the sub-ranges; coarser sub-ranges when they cannot */
    { p2 = 0 ;
      while (rem + p2 < g.sortBits && (top >> (rem + p2)) + 1 > H10X_P1_MAX_RANGES) ++p2 ;
      if (((uint64_t) 1 << p2) <= H10X_PART_MAX_BINS) break ;
      ++rem ;
    }
  /* a 1024-digit partition pass costs 20.6 ms over the 1 Gb workload's entries, a 512-digit one 17.7 (profiles/README.md):
     one more bit goes to the shared-memory sort instead when the sub-ranges still fit */
  if (p2 == 10 && rem + 1 <= g.sortBits && (double) H / (double) ((top >> (rem + 1)) + 1) <= 0.35 * H10X_SR_CAP && !getenv ("H10X_SR_TARGET"))
    { ++rem ; --p2 ; }
  if ((double) H / (double) ((top >> rem) + 1) > 0.35 * H10X_SR_CAP) return false ;	/* twice that at the dense end */
  g.remBits = rem ; g.p2 = p2 ; g.lowBits = rem + p2 ;
  if (g.lowBits + g.eShift > 64) return false ;
  g.nRanges = (uint32_t) ((top >> g.lowBits) + 1) ;
  g.nSub = g.nRanges << p2 ;
  return true ;
}

__global__ void k_job_starts (uint64_t n, uint32_t nJobs, uint64_t chunk, uint64_t *__restrict__ jobStart)
{ uint32_t j = blockIdx.x * blockDim.x + threadIdx.x ;
  if (j <= nJobs) jobStart[j] = (j == nJobs) ? n : min (n, (uint64_t) j * chunk) ;
}

__global__ void k_u64_to_u32 (const uint64_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ; if (i < n) out[i] = (uint32_t) in[i] ; }

static int device_sms (h10x_ctx *c)
{ int nSM = 148 ; CK (cudaDeviceGetAttribute (&nSM, cudaDevAttrMultiProcessorCount, c->P.device)) ; return nSM ; }

/* the scatter kernel for a digit width: 16 warps per CTA above 512 digits (its shared memory then allows two CTAs
   per SM anyway), 8 warps and up to four CTAs below */
template <class L>
static void part_scatter_launch (h10x_ctx *c, cudaStream_t s, L ld, const uint64_t *jobStart, uint32_t nJobs, uint32_t nBins,
				 uint64_t strideBin, uint64_t strideJob, const uint32_t *off, uint64_t *out, uint32_t grid,
				 unsigned int *ticket = nullptr)
{ if (nBins > 512)
    { auto fn = k_part_scatter<L, 16, 8, 2> ;
      const size_t smem = h10x_part_smem (nBins, 16) ;
      CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
      LAUNCH (c, fn, grid, 16 * 32, smem, s, ld, jobStart, nJobs, nBins, strideBin, strideJob, off, out, ticket) ;
    }
  else
    { auto fn = k_part_scatter<L, 8, 8, 4> ;
      const size_t smem = h10x_part_smem (nBins, 8) ;
      CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
      LAUNCH (c, fn, grid, 8 * 32, smem, s, ld, jobStart, nJobs, nBins, strideBin, strideJob, off, out, ticket) ;
    }
}

template <class L>
static uint32_t part_resident_ctas (h10x_ctx *c, uint32_t nBins)
{ int occ = 1 ;
  if (nBins > 512)
    { auto fn = k_part_scatter<L, 16, 8, 2> ;
      CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h10x_part_smem (nBins, 16))) ;
      CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, 512, h10x_part_smem (nBins, 16))) ;
    }
  else
    { auto fn = k_part_scatter<L, 8, 8, 4> ;
      CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h10x_part_smem (nBins, 8))) ;
      CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, 256, h10x_part_smem (nBins, 8))) ;
    }
  return (uint32_t) (device_sms (c) * std::max (occ, 1)) ;
}

/* one stable partition pass over n elements seen through `ld`: out[] = the elements grouped by digit, each group
   in input order.  Jobs are equal chunks of the input, one per resident CTA. */
template <class L>
static void part_pass (h10x_ctx *c, cudaStream_t s, L ld, uint64_t n, uint32_t nBins, uint64_t *out)
{ if (!n) return ;
  MemTrack *mt = &c->mt ;
  const uint64_t tile = (uint64_t) (nBins > 512 ? 16 : 8) * 32 * 8 ;
  /* one job per resident CTA: the scatter runs as a single wave */
  uint32_t nJobs = (uint32_t) std::min<uint64_t> (part_resident_ctas<L> (c, nBins), (n + tile - 1) / tile) ;
  uint64_t chunk = (n + nJobs - 1) / nJobs ;
  chunk = (chunk + tile - 1) / tile * tile ;
  nJobs = (uint32_t) ((n + chunk - 1) / chunk) ;
  DBuf<uint64_t> jobStart ((size_t) nJobs + 1, s, mt) ;
  LAUNCH (c, k_job_starts, gridFor (nJobs + 1, 256), 256, 0, s, n, nJobs, chunk, jobStart.p) ;
  const size_t nh = (size_t) nBins * nJobs ;
  DBuf<uint32_t> hist (nh + 1, s, mt) ;
  CK (cudaMemsetAsync (hist.p + nh, 0, 4, s)) ;
  LAUNCH (c, k_part_hist<L>, nJobs, 512, (size_t) nBins * 4, s, ld, jobStart.p, nJobs, nBins, (uint64_t) nJobs, (uint64_t) 1, hist.p,
	  (unsigned int*) nullptr) ;
  cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, hist.p, hist.p, nh + 1, s) ; }) ;
  part_scatter_launch (c, s, ld, jobStart.p, nJobs, nBins, (uint64_t) nJobs, (uint64_t) 1, hist.p, out, nJobs) ;
}

struct TailSrc {	/* what the mosh stages left behind */
  uint32_t nProcBlk, blkBase ;
  const uint64_t *srcOff ; const uint32_t *blkCnt ; const uint64_t *scratch, *gHash ; const uint32_t *gRec, *blkStart ;
  uint64_t wInvFull, wDiv ;
} ;

/* P1: the entries, range-major, as packed words (stage "dedup") */
static void tail_p1 (h10x_ctx *c, cudaStream_t s, const TailGeom &g, const TailSrc &in, uint64_t H,
		     DBuf<uint64_t> &B, DBuf<uint64_t> &rangeStart)
{ MemTrack *mt = &c->mt ;
  const int nSM = device_sms (c) ;
  uint32_t G = (uint32_t) (((uint64_t) in.nProcBlk * g.nRanges + ((1u << 24) - 1)) >> 24) ;
  if (G < 1) G = 1 ;
  const uint32_t nTiles = (in.nProcBlk + G - 1) / G ;
  const size_t nc = (size_t) g.nRanges * nTiles ;
  DBuf<uint32_t> cnt (nc + 1, s, mt) ;
  CK (cudaMemsetAsync (cnt.p + nc, 0, 4, s)) ;
  LAUNCH (c, k_p1_count, std::min<uint32_t> (nTiles, (uint32_t) nSM * 8), 256, (size_t) g.nRanges * 4, s, in.nProcBlk, G, nTiles,
	  g.nRanges, g.lowBits, in.srcOff, in.blkCnt, in.scratch, in.gHash, in.wInvFull, cnt.p) ;
  cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, cnt.p, cnt.p, nc + 1, s) ; }) ;
  rangeStart.alloc ((size_t) g.nRanges + 1, s, mt) ;
  LAUNCH (c, k_p1_range_start, gridFor (g.nRanges + 1, 256), 256, 0, s, g.nRanges, nTiles, cnt.p, H, rangeStart.p) ;
  B.alloc (H, s, mt) ;
  const size_t smem = (size_t) g.nRanges * (8 * H10X_RING_STRIDE + 12) + 16 ;
  CK (cudaFuncSetAttribute (k_p1_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
  const uint32_t grid = std::min<uint32_t> (nTiles, (uint32_t) nSM) ;
  uint32_t flushEvery = 2 ;	/* measured at the 1 Gb workload: 21.0 ms every block, 20.1 every second, 23.9 every fourth (ring overflows) */
  if (const char *e = getenv ("H10X_P1_FLUSH")) { long v = atol (e) ; if (v >= 1 && v <= 64) flushEvery = (uint32_t) v ; }
  LAUNCH (c, k_p1_place, grid, H10X_P1_THREADS, smem, s, flushEvery, in.nProcBlk, G, nTiles, g.nRanges, g.lowBits, g.eShift, in.srcOff,
	  in.blkCnt, cnt.p, in.scratch, in.gHash, in.gRec, in.blkStart, in.blkBase, in.wInvFull, B.p) ;
}

/* P2, sub-range sort, bin ids, codes, ClusterHash lists: everything after P1.  Leaves hashValue / hashDepth /
   hashNumber / codeOff / codes / clus in the context; returns the number of bins.  H comes in as the number of
   entries P1 placed and leaves as the number of (hash, block) pairs: the sub-range sort drops the duplicates a lean
   fused kernel left in, blkDupHost[b] = how many of (global, 1-based) block b's. */
static void dist_bins (h10x_ctx *c, cudaStream_t s, uint64_t H, const uint64_t *sh, const uint32_t *se,
		       const uint32_t *segIncl, uint32_t Dl, const uint32_t *segStart, const uint32_t *entryBlk,
		       uint32_t nBlkGlobal, uint32_t *entryId, uint32_t &Dglobal, uint64_t wMul,
		       const uint64_t *preHash = nullptr, const uint32_t *preDepth = nullptr, const uint32_t *preFirst = nullptr) ;
struct TailDist { uint32_t nBlkGlobal ; } ;	/* multi-GPU build: the bins get their ids from the owners (dist_bins) */

static uint32_t tail_rest (h10x_ctx *c, cudaStream_t s, const TailGeom &g, uint64_t &H, uint64_t wDiv,
			   DBuf<uint64_t> &B, DBuf<uint64_t> &rangeStart, uint32_t nBlkNumbers, std::vector<uint32_t> &blkDupHost,
			   const TailDist *td = nullptr)
{ MemTrack *mt = &c->mt ;
  const h10x_params &P = c->P ;
  const int nSM = device_sms (c) ;
  const uint32_t blkMask = (uint32_t) ((((uint64_t) 1) << g.blkBits) - 1) ;
  DBuf<uint64_t> A ;
    DBuf<uint32_t> srStart ((size_t) g.nSub + 1, s, mt), stage, nHeads ((size_t) g.nSub + 1, s, mt), srCount ((size_t) g.nSub, s, mt) ;
  DBuf<uint32_t> blkDup ((size_t) nBlkNumbers + 1, s, mt) ; DBuf<unsigned long long> nDup (1, s, mt) ;
  uint32_t D = 0 ;
  { StageTimer tm (c, s, ST_HASHSORT) ;
    if (g.p2 > 0)
      { const uint32_t nBins = 1u << g.p2 ;
	LoadWord ld = { B.p, g.eShift + g.remBits, nBins - 1u, ~(uint64_t) 0 } ;
	CK (cudaMemsetAsync (srStart.p + g.nSub, 0, 4, s)) ;
	DBuf<unsigned int> tickets (2, s, mt) ;		/* ranges are handed out in order = largest first (h10x_next_job) */
	CK (cudaMemsetAsync (tickets.p, 0, 8, s)) ;
	unsigned int *tk = getenv ("H10X_P2_STATIC") ? nullptr : tickets.p ;
	LAUNCH (c, k_part_hist<LoadWord>, std::min<uint32_t> (g.nRanges, (uint32_t) nSM * 4), 512, (size_t) nBins * 4, s, ld, rangeStart.p,
		g.nRanges, nBins, (uint64_t) 1, (uint64_t) nBins, srStart.p, tk) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, srStart.p, srStart.p, (size_t) g.nSub + 1, s) ; }) ;
	A.alloc (H, s, mt) ;
	part_scatter_launch (c, s, ld, rangeStart.p, g.nRanges, nBins, (uint64_t) 1, (uint64_t) nBins, srStart.p, A.p,
			     std::min<uint32_t> (g.nRanges, part_resident_ctas<LoadWord> (c, nBins)), tk ? tk + 1 : nullptr) ;
	B.release () ;
      }
    else
      { LAUNCH (c, k_u64_to_u32, gridFor (g.nSub + 1, 256), 256, 0, s, rangeStart.p, g.nSub + 1, srStart.p) ;
	A.swap (B) ;
      }
    rangeStart.release () ;
    /* sub-range sort + first entries */
    stage.alloc (H, s, mt) ;
    DBuf<uint32_t> overList ((size_t) g.nSub, s, mt), jobList ((size_t) g.nSub, s, mt) ; DBuf<unsigned int> counters (3, s, mt) ;
    CK (cudaMemsetAsync (counters.p, 0, 12, s)) ;
        CK (cudaMemsetAsync (nHeads.p, 0, 4 * ((size_t) g.nSub + 1), s)) ;
    CK (cudaMemsetAsync (srCount.p, 0, 4 * (size_t) g.nSub, s)) ;
    CK (cudaMemsetAsync (blkDup.p, 0, 4 * ((size_t) nBlkNumbers + 1), s)) ;
    CK (cudaMemsetAsync (nDup.p, 0, 8, s)) ;
    SrArgs sa ;
    sa.A = A.p ; sa.srStart = srStart.p ; sa.stage = stage.p ; sa.nHeads = nHeads.p ; sa.overList = overList.p ; sa.jobList = jobList.p ;
    sa.nOver = counters.p ; sa.ticket = counters.p + 1 ; sa.nJobs = counters.p + 2 ; sa.nSub = g.nSub ;
    sa.cap = H10X_SR_CAP ;
    if (const char *e = getenv ("H10X_SR_CAP")) { long v = atol (e) ; if (v >= 1 && v < (long) H10X_SR_CAP) sa.cap = (uint32_t) v ; }	/* tests: force the big-sub-range path */
    sa.groupCap = sa.cap ;		/* as large as fits: 3584 cost 2.1 ms more at the 1 Gb workload, no grouping 9.3 ms more */ sa.maxLog = (uint32_t) std::min (2, g.p2) ;
        sa.eShift = g.eShift ; sa.remBits = g.remBits ;
    sa.srCount = srCount.p ; sa.blkDup = blkDup.p ; sa.nDup = nDup.p ; sa.blkMask = blkMask ;
    if (const char *e = getenv ("H10X_SR_GROUP")) { long v = atol (e) ; if (v >= 1) sa.groupCap = std::min<uint32_t> ((uint32_t) v, sa.cap) ; }
    LAUNCH (c, k_sr_jobs, gridFor (((uint64_t) g.nSub + 3) / 4, 256), 256, 0, s, sa) ;
    const int srThreads = H10X_SR_THREADS_DEFAULT ;
    const size_t smem = (size_t) 2 * sa.cap * 8 + ((size_t) 4 << H10X_SR_DIGIT) + ((size_t) (srThreads / 32) * 2 << H10X_SR_DIGIT) + 16 ;
    auto srFn = k_sr_sort<H10X_SR_THREADS_DEFAULT> ;	/* 256-thread CTAs (four per SM) were tried: the per-thread step arrays no longer fit the register budget */
    CK (cudaFuncSetAttribute (srFn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
    int occ = 1 ;
    CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, srFn, srThreads, smem)) ;
    if (occ < 1) occ = 1 ;
    LAUNCH (c, srFn, std::min<uint32_t> (g.nSub, (uint32_t) (nSM * occ)), srThreads, smem, s, sa) ;
    unsigned int nOver = 0 ;
    CK (cudaMemcpyAsync (&nOver, counters.p, 4, cudaMemcpyDeviceToHost, s)) ;
    CK (cudaStreamSynchronize (s)) ;
    if (nOver)
      { /* one segmented library sort for all of them (one sort + three host round trips per sub-range made a data set
	   with ~1400 oversized sub-ranges 440 ms slower) */
	std::vector<uint32_t> over (nOver), st2 ((size_t) g.nSub + 1) ;
	CK (cudaMemcpyAsync (over.data (), overList.p, 4 * (size_t) nOver, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaMemcpyAsync (st2.data (), srStart.p, 4 * ((size_t) g.nSub + 1), cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	std::sort (over.begin (), over.end ()) ;
	std::vector<uint64_t> segOff ((size_t) nOver + 1, 0) ;
	for (uint32_t x = 0 ; x < nOver ; ++x) segOff[x + 1] = segOff[x] + (st2[over[x] + 1] - st2[over[x]]) ;
	const uint64_t T = segOff[nOver] ;
	DBuf<uint64_t> cIn (T, s, mt), cOut (T, s, mt), dSegOff ((size_t) nOver + 1, s, mt) ;
	CK (cudaMemcpyAsync (overList.p, over.data (), 4 * (size_t) nOver, cudaMemcpyHostToDevice, s)) ;
	CK (cudaMemcpyAsync (dSegOff.p, segOff.data (), 8 * ((size_t) nOver + 1), cudaMemcpyHostToDevice, s)) ;
	LAUNCH (c, k_over_copy, nOver, 512, 0, s, A.p, srStart.p, overList.p, dSegOff.p, cIn.p, 1) ;
	/* the whole word orders a sub-range: q bits, then block, then read (the lowest read of a (hash, block) first) */
	cubCall (c, s, [&] (void *t, size_t &b)
	  { return cub::DeviceSegmentedSort::SortKeys (t, b, cIn.p, cOut.p, (::cuda::std::int64_t) T, (::cuda::std::int64_t) nOver,
						       dSegOff.p, dSegOff.p + 1, s) ; }) ;
	LAUNCH (c, k_over_copy, nOver, 512, 0, s, A.p, srStart.p, overList.p, dSegOff.p, cOut.p, 0) ;
	LAUNCH (c, k_sr_heads_big, nOver, 512, 0, s, sa, overList.p) ;
	CK (cudaStreamSynchronize (s)) ;	/* `over` / `segOff` are read by the async copies */
      }
    unsigned long long dups = 0 ;
    blkDupHost.assign ((size_t) nBlkNumbers + 1, 0) ;
    CK (cudaMemcpyAsync (&dups, nDup.p, 8, cudaMemcpyDeviceToHost, s)) ;
    CK (cudaMemcpyAsync (blkDupHost.data (), blkDup.p, 4 * ((size_t) nBlkNumbers + 1), cudaMemcpyDeviceToHost, s)) ;
    CK (cudaStreamSynchronize (s)) ;
    H -= dups ;
  }
  DBuf<uint32_t> segStart, segLen, idOfSeg, gidOfSeg, localDepth ;
  uint32_t Dglobal = 0 ;
  { std::unique_ptr<StageTimer> tm (new StageTimer (c, s, ST_BINIDS)) ;
    DBuf<uint32_t> binBase ((size_t) g.nSub + 1, s, mt) ;
    cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, nHeads.p, binBase.p, (size_t) g.nSub + 1, s) ; }) ;
    CK (cudaMemcpyAsync (&D, binBase.p + g.nSub, 4, cudaMemcpyDeviceToHost, s)) ;
    CK (cudaStreamSynchronize (s)) ;
    if (!td && (uint64_t) D + 1 > (((uint64_t) 1 << P.B) >> 2) - 2)	/* hash10x.c:149 (multi-GPU: dist_bins, on the global count) */
      throw H10xError (H10X_ERR_TABLE_TOO_SMALL, "hashTableSize is too small") ;
    segStart.alloc (D, s, mt) ; segLen.alloc (D, s, mt) ; idOfSeg.alloc (D, s, mt) ;
    DBuf<uint64_t> fkey (D, s, mt), fkey2 (D, s, mt), hvHash (D, s, mt) ;
    DBuf<uint32_t> firstBlk ;
    if (td) firstBlk.alloc (D, s, mt) ;
    LAUNCH (c, k_heads_compact, (uint32_t) std::min<uint64_t> (((uint64_t) g.nSub * 32 + 255) / 256, (uint64_t) nSM * 16), 256, 0, s, g.nSub,
	    srStart.p, nHeads.p, srCount.p, binBase.p, stage.p, A.p, g.eShift, g.lowBits, g.p2, blkMask, wDiv, segStart.p, segLen.p,
	    fkey.p, hvHash.p, firstBlk.p) ;
    stage.release () ; nHeads.release () ; binBase.release () ; srStart.release () ; srCount.release () ;
    if (td)
      { /* multi-GPU.  k_heads_compact has just left this rank's distinct hashes in ascending order with their local depth
	   and first block: exactly what the hash-range owners want.  The global ids come back per local bin; the local
	   lists are then laid out in id order (stable passes on the id bits), so that everything below - the local part of
	   fillHashTable, the transposition to ClusterHash lists - is the single-GPU code. */
	tm.reset () ;		/* dist_bins times itself under the same stage */
	dist_bins (c, s, H, nullptr, nullptr, nullptr, D, nullptr, nullptr, td->nBlkGlobal, nullptr, Dglobal, wDiv,
		   hvHash.p, segLen.p, firstBlk.p) ;
	tm.reset (new StageTimer (c, s, ST_BINIDS)) ;
	hvHash.release () ; firstBlk.release () ;
	gidOfSeg.swap (c->localBinId) ;		/* ids by local bin, hash order */
	LAUNCH (c, k_id_words, gridFor (D, 256), 256, 0, s, D, gidOfSeg.p, fkey.p) ;
	const int idBits = bits_for ((uint64_t) Dglobal + 1) ;
	const int nP = (idBits + 9) / 10 ;
	uint64_t *src = fkey.p, *dst = fkey2.p ;
	for (int i = 0, left = idBits, shift = 32 ; i < nP ; ++i)
	  { const int bits = (left + (nP - i) - 1) / (nP - i) ;
	    LoadWord lw = { src, shift, (1u << bits) - 1u, ~(uint64_t) 0 } ;
	    part_pass (c, s, lw, D, 1u << bits, dst) ;
	    shift += bits ; left -= bits ; std::swap (src, dst) ;
	  }
	c->localBinId.alloc (D, s, mt) ; localDepth.alloc ((size_t) D + 2, s, mt) ;
	CK (cudaMemsetAsync (localDepth.p, 0, 4, s)) ;
	CK (cudaMemsetAsync (localDepth.p + D + 1, 0, 4, s)) ;
	LAUNCH (c, k_local_ranks, gridFor (D, 256), 256, 0, s, D, src, segLen.p, idOfSeg.p, c->localBinId.p, localDepth.p) ;
      }
    else
      {
    /* bins stand in hash order; the id order is (first block, hash): stable passes on the first block */
    const int b1 = g.blkBits <= 9 ? g.blkBits : (g.blkBits + 1) / 2, b2 = g.blkBits - b1 ;
    if (b1 > 11) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^22 barcode blocks") ;
    const uint64_t *sorted = nullptr ;
    { LoadWord l1 = { fkey.p, 32, (1u << b1) - 1u, ~(uint64_t) 0 } ;
      part_pass (c, s, l1, D, 1u << b1, fkey2.p) ;
      sorted = fkey2.p ;
      if (b2 > 0)
	{ LoadWord l2 = { fkey2.p, 32 + b1, (1u << b2) - 1u, ~(uint64_t) 0 } ;
	  part_pass (c, s, l2, D, 1u << b2, fkey.p) ;
	  sorted = fkey.p ;
	}
    }
    c->hashNumber = D + 1 ;
    c->hashValue.alloc ((size_t) D + 1, s, mt) ;
    c->hashDepth.alloc ((size_t) D + 2, s, mt) ;	/* one spare 0 so the scan yields codeOff[hashNumber] */
    CK (cudaMemsetAsync (c->hashValue.p, 0, 8, s)) ;
    CK (cudaMemsetAsync (c->hashDepth.p, 0, 4, s)) ;
    CK (cudaMemsetAsync (c->hashDepth.p + D + 1, 0, 4, s)) ;
    LAUNCH (c, k_bins_by_rank_e, gridFor (D, 256), 256, 0, s, D, sorted, segLen.p, hvHash.p, idOfSeg.p, c->hashValue.p, c->hashDepth.p) ;
      }
  }
  early_pull (c, s, SLOT_VALUE, c->hashValue.p, 8 * (size_t) c->hashNumber) ;
  early_pull (c, s, SLOT_DEPTH, c->hashDepth.p, 4 * (size_t) c->hashNumber) ;
  /* hashIndex[] only needs the values in id order: built here, its copy to the host runs beside the stages below */
  if (!(P.flags & H10X_FLAG_NO_TABLE) && c->hashValue.p)	/* multi-GPU: rank 0 holds the values */
    { StageTimer tm (c, s, ST_TABLE) ;
      const size_t tableSize = (size_t) 1 << P.B ;
      c->hashIndex.alloc (tableSize, s, mt) ;
      CK (cudaMemsetAsync (c->hashIndex.p, 0, 4 * tableSize, s)) ;
      if (c->hashNumber > 1) LAUNCH (c, k_table_insert, gridFor (c->hashNumber - 1, 256), 256, 0, s, c->hashNumber, c->hashValue.p, c->hashIndex.p, P.B) ;
      early_pull (c, s, SLOT_INDEX, c->hashIndex.p, 4 * tableSize) ;
    }
  const size_t hn = td ? (size_t) D + 1 : (size_t) c->hashNumber ;
  DBuf<uint64_t> idRead (H, s, mt) ;
  const uint32_t *codesPtr = nullptr ;
  { StageTimer tm (c, s, ST_CODES) ;
    if (td)
      { /* this rank's part of every bin's barcode list (h10x_dist.cuh): local bin j (id order) is bin localBinId[j] and
	   holds the (global) blocks localCodes[localCodeOff[j] .. localCodeOff[j+1]), ascending */
	DBuf<uint64_t> codeOffLocal (hn + 1, s, mt) ;
	c->localCodes.alloc (H, s, mt) ;
	cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> depth64 (localDepth.p, CastU64 ()) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, depth64, codeOffLocal.p, hn + 1, s) ; }) ;
	LAUNCH (c, k_codes_seg_e, gridFor (((uint64_t) D + 31) / 32 * 32, 256), 256, 0, s, D, segStart.p, segLen.p, idOfSeg.p, gidOfSeg.p,
		A.p, blkMask, codeOffLocal.p, c->localCodes.p, idRead.p) ;
	c->localCodeOff.alloc ((size_t) D + 1, s, mt) ;
	LAUNCH (c, k_u64_to_u32_from, gridFor ((uint64_t) D + 1, 256), 256, 0, s, codeOffLocal.p + 1, D + 1, c->localCodeOff.p) ;
	codesPtr = c->localCodes.p ;
      }
    else
      { c->codeOff.alloc (hn + 1, s, mt) ;
	c->codes.alloc (H, s, mt) ;
	cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> depth64 (c->hashDepth.p, CastU64 ()) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, depth64, c->codeOff.p, hn + 1, s) ; }) ;
	LAUNCH (c, k_codes_seg_e, gridFor (((uint64_t) D + 31) / 32 * 32, 256), 256, 0, s, D, segStart.p, segLen.p, idOfSeg.p,
		(const uint32_t*) nullptr, A.p, blkMask, c->codeOff.p, c->codes.p, idRead.p) ;
	codesPtr = c->codes.p ;
      }
  }
    A.release () ; segStart.release () ; segLen.release () ; idOfSeg.release () ; gidOfSeg.release () ; localDepth.release () ;
    if (!(P.flags & (H10X_FLAG_NO_CODES | H10X_FLAG_LAZY_CODES)))
    { early_pull (c, s, SLOT_CODEOFF, c->codeOff.p, 8 * (hn + 1)) ; early_pull (c, s, SLOT_CODES, c->codes.p, 4 * H) ; }
  c->clus.alloc (H, s, mt) ;
  { StageTimer tm (c, s, ST_CLUSTERS) ;
    /* stable partition of the bin-major (id, read) words by block number, lowest digit first.  The block bits
       not yet used ride in the free top 16 bits of the word. */
    const int nP = (g.blkBits + 9) / 10 ;	/* 512-1024 digits per pass: the rings absorb everything; with few digits most
						   entries overflow them and take the slow direct path (3 passes of 6-7 bits: 50 ms) */
    int bits[4] = { 0, 0, 0, 0 } ;
    for (int i = 0, left = g.blkBits ; i < nP ; ++i) { bits[i] = (left + (nP - i) - 1) / (nP - i) ; left -= bits[i] ; }
    if (const char *e = getenv ("H10X_T_FIRST")) { int v = atoi (e) ; if (nP == 2 && v >= 1 && v <= 10 && g.blkBits - v >= 1 && g.blkBits - v <= 10) { bits[0] = v ; bits[1] = g.blkBits - v ; } }
    LoadCodes l1 = { codesPtr, idRead.p, (1u << bits[0]) - 1u, bits[0] } ;
    if (nP == 1) part_pass (c, s, l1, H, 1u << bits[0], c->clus.p) ;
    else
      { DBuf<uint64_t> X (H, s, mt), Y ;
	part_pass (c, s, l1, H, 1u << bits[0], X.p) ;
	idRead.release () ;
	uint64_t *src = X.p ;
	int shift = 48 ;
	for (int i = 1 ; i < nP ; ++i)
	  { const bool last = (i == nP - 1) ;
	    uint64_t *dst = last ? c->clus.p : (src == X.p ? (Y.alloc (H, s, mt), Y.p) : X.p) ;
	    LoadWord lw = { src, shift, (1u << bits[i]) - 1u, last ? 0x0000ffffffffffffull : ~(uint64_t) 0 } ;
	    part_pass (c, s, lw, H, 1u << bits[i], dst) ;
	    shift += bits[i] ; src = dst ;
	  }
      }
  }
  if (P.flags & H10X_FLAG_NO_CODES) { c->codes.release () ; c->codeOff.release () ; }
  early_pull (c, s, SLOT_CLUS, c->clus.p, 8 * H) ;
  return td ? Dglobal : D ;
}


static void dist_agree (h10x_ctx *c, cudaStream_t s, int localErr, const uint32_t mine[4], std::vector<uint32_t> &all) ;

/* ------------------------------------------------------------------ host: the fused mosh stage (h10x_fused.cuh) */

struct FusedClass { uint32_t cap, nbuck, lb, threads, rowCap ; } ;
/* shared memory per CTA = 8 * cap + 4 * nbuck (+ 1 KB the driver reserves): 5, 5, 4, 2 and 1 CTAs per SM */
static const FusedClass kClasses[5] = { { 1024, 512, 9, 128, 32 }, { 4096, 2048, 11, 256, 48 }, { 5632, 2048, 11, 384, 48 },
					{ 12288, 4096, 12, 512, 64 }, { 24576, 8192, 13, 1024, 64 } } ;
static const int kNClasses = 5 ;

/* Everything the fused kernel's launches of one build share: kernel variants, grids, the staging area, the scratch
   slab the block lists go to and its cursor.  A launch takes a list of blocks per size class; the blocks of a build
   may come in several launches (h10x_gpu_build_host hashes each slab of the file while the next one is copied). */
struct FusedEngine {
  bool ready = false, lean = false ;
  int n1 = 0, n2 = 0, c1 = 0, c2 = 0 ;
  double perPair = 0 ;
  const void *fn[5] = { nullptr, nullptr, nullptr, nullptr, nullptr } ; size_t smemB[5] = { 0, 0, 0, 0, 0 } ; uint32_t maxGrid[5] = { 0, 0, 0, 0, 0 } ;
  DBuf<uint64_t> scratch, stage ; DBuf<unsigned long long> cursor ; DBuf<unsigned int> work ;
  uint64_t scratchCap = 0 ; uint32_t workUsed = 0, workCap = 0 ;
  std::deque<DBuf<uint32_t>> dLists ;

  static bool usable (const h10x_params &P)
  { const int n1 = H10X_R1_LEN - P.k + 1, n2 = H10X_R2_LEN - P.k + 1 ;
    return !(P.flags & H10X_FLAG_GENERIC_ONLY) && P.k >= 13 && P.k <= 23 && (n1 + 7) / 8 + (n2 + 7) / 8 <= 32 && n1 >= 8 ;
  }
  /* nProc: records of the blocks that will be hashed; nLaunchGroups: how many times launch () may be called */
  void init (h10x_ctx *c, cudaStream_t s, bool lean_, uint64_t nProc, uint32_t nLaunchGroups)
  { const h10x_params &P = c->P ; MemTrack *mt = &c->mt ;
    lean = lean_ ;
    n1 = H10X_R1_LEN - P.k + 1 ; n2 = H10X_R2_LEN - P.k + 1 ; c1 = (n1 + 7) / 8 ; c2 = (n2 + 7) / 8 ;
    perPair = (double) (n1 + n2) / (double) P.w ;	/* expected moshes per read pair */
    scratchCap = (uint64_t) (perPair * 1.15 * nProc) + (1u << 20) ;
    workCap = nLaunchGroups * kNClasses ; workUsed = 0 ;
    scratch.alloc (scratchCap, s, mt) ; cursor.alloc (2, s, mt) ; work.alloc (workCap, s, mt) ;	/* cursor[1]: mosh count */
    CK (cudaMemsetAsync (cursor.p, 0, 16, s)) ;
    CK (cudaMemsetAsync (work.p, 0, 4 * (size_t) workCap, s)) ;
    int nSM = 148 ;
    CK (cudaDeviceGetAttribute (&nSM, cudaDevAttrMultiProcessorCount, P.device)) ;
    const bool wodd = (c->hp.wTz == 0 && P.w >= 3) ;
    const bool k21 = (P.k == 21 && wodd) ;
    size_t stageKeys = 0 ;
    for (int ci = 0 ; ci < kNClasses ; ++ci)
      { const FusedClass &fc = kClasses[ci] ;
	smemB[ci] = (size_t) fc.cap * 8 + 4 * ((size_t) fc.nbuck + 1) + 16 ;
#define FUSED_FN1(T, L) (k21 ? (const void*) k_fused_block<T, true, 21, L> : wodd ? (const void*) k_fused_block<T, true, 0, L> \
		     : (const void*) k_fused_block<T, false, 0, L>)
#define FUSED_FN(T) (lean ? FUSED_FN1 (T, true) : FUSED_FN1 (T, false))
	fn[ci] = fc.threads == 128 ? FUSED_FN (128) : fc.threads == 256 ? FUSED_FN (256) : fc.threads == 384 ? FUSED_FN (384)
	  : fc.threads == 512 ? FUSED_FN (512) : FUSED_FN (1024) ;
#undef FUSED_FN
#undef FUSED_FN1
	CK (cudaFuncSetAttribute (fn[ci], cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemB[ci])) ;
	int occ = 1 ;
	CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn[ci], (int) fc.threads, smemB[ci])) ;
	if (occ < 1) occ = 1 ;
	maxGrid[ci] = (uint32_t) nSM * (uint32_t) occ ;
	stageKeys = std::max<size_t> (stageKeys, (size_t) maxGrid[ci] * fc.threads * fc.rowCap) ;
      }
    stage.alloc (stageKeys, s, mt) ;	/* launches run one after another and share it */
    ready = true ;
  }
  /* size class of a block of nRead read pairs, -1: not for this path */
  int classOf (const h10x_params &P, uint32_t nRead) const
  { const double e = perPair * nRead, need = e * 1.12 + 6.0 * sqrt (e) + 16.0 ;
    if (((uint64_t) nRead >> (64 - 2 * P.k)) != 0) return -1 ;	/* read index must fit under the hash bits */
    for (int ci = 0 ; ci < kNClasses ; ++ci)
      if (need <= kClasses[ci].cap && (int) kClasses[ci].lb <= 2 * P.k) return ci ;
    return -1 ;
  }
  void launch (h10x_ctx *c, cudaStream_t s, const std::vector<uint32_t> lists[5], const uint32_t *fqb, const uint32_t *blkStart,
	       uint64_t *srcOff, uint32_t *blkCnt)
  { if (workUsed + kNClasses > workCap) throw H10xError (H10X_ERR_CUDA, "internal: fused launch groups exhausted") ;
    for (int ci = 0 ; ci < kNClasses ; ++ci)
      { if (lists[ci].empty ()) continue ;
	const FusedClass &fc = kClasses[ci] ;
	dLists.emplace_back () ;
	DBuf<uint32_t> &dl = dLists.back () ;
	dl.alloc (lists[ci].size (), s, &c->mt) ;
	CK (cudaMemcpyAsync (dl.p, lists[ci].data (), 4 * lists[ci].size (), cudaMemcpyHostToDevice, s)) ;
	FusedArgs fa ;
	fa.fqb = fqb ; fa.list = dl.p ; fa.blkStart = blkStart ; fa.stage = stage.p ; fa.scratch = scratch.p ;
	fa.cursor = cursor.p ; fa.work = work.p + workUsed + ci ; fa.scratchCap = scratchCap ; fa.srcOff = srcOff ; fa.blkCnt = blkCnt ;
	fa.nList = (uint32_t) lists[ci].size () ; fa.cap = fc.cap ; fa.nbuck = fc.nbuck ; fa.lb = fc.lb ; fa.rowCap = fc.rowCap ;
	fa.c1 = c1 ; fa.c2 = c2 ; fa.n1 = n1 ; fa.n2 = n2 ; fa.moshCount = cursor.p + 1 ;
	HashParams hpv = c->hp ;
	void *args[2] = { (void*) &fa, (void*) &hpv } ;
	const uint32_t grid = (uint32_t) std::min<size_t> (lists[ci].size (), maxGrid[ci]) ;
	CK (cudaLaunchKernel (fn[ci], dim3 (grid), dim3 (fc.threads), args, smemB[ci], s)) ;
	++c->launches ;
      }
    workUsed += kNClasses ;
  }
  void release ()
  { scratch.release () ; stage.release () ; work.release () ; cursor.release () ; dLists.clear () ; ready = false ; }
  void swap (FusedEngine &o)
  { std::swap (ready, o.ready) ; std::swap (lean, o.lean) ; std::swap (n1, o.n1) ; std::swap (n2, o.n2) ; std::swap (c1, o.c1) ;
    std::swap (c2, o.c2) ; std::swap (perPair, o.perPair) ;
    for (int i = 0 ; i < 5 ; ++i) { std::swap (fn[i], o.fn[i]) ; std::swap (smemB[i], o.smemB[i]) ; std::swap (maxGrid[i], o.maxGrid[i]) ; }
    scratch.swap (o.scratch) ; stage.swap (o.stage) ; cursor.swap (o.cursor) ; work.swap (o.work) ;
    std::swap (scratchCap, o.scratchCap) ; std::swap (workUsed, o.workUsed) ; std::swap (workCap, o.workCap) ; dLists.swap (o.dLists) ;
  }
} ;

/* What h10x_gpu_build_host's streamed front half hands to build_device_impl: the barcode runs of the whole file and
   the fused kernel's lists of every run but the last, produced slab by slab while the file was still being copied.
   Runs are blocks unless a barcode word is 0 (simulate_chunks); then, or when the tail's geometry turns the lean
   lists down, the lists are dropped and the classic stages run. */
struct Prefuse {
  bool valid = false, anyZero = false ;
  uint32_t nRuns = 0 ;
  std::vector<uint32_t> runStart ;	/* nRuns + 1 */
  DBuf<uint32_t> dBlkStart ;		/* the same on the device */
  DBuf<uint64_t> srcOff ; DBuf<uint32_t> blkCnt ;	/* per run */
  FusedEngine eng ;
} ;

static bool tail_geometry (uint64_t H, uint64_t top, uint32_t maxBlock, TailGeom &g) ;
static bool lean_wanted (h10x_ctx *c, bool dist, uint64_t nProc, uint32_t nBlocks)
{ const h10x_params &P = c->P ;
  if ((P.flags & H10X_FLAG_LEGACY_TAIL) || getenv ("H10X_LEGACY_TAIL") || getenv ("H10X_NO_LEAN")) return false ;
  const uint64_t wDiv = c->hp.wTz ? 1 : (uint64_t) P.w ;
  const uint64_t topQ = (((uint64_t) 1 << (2 * P.k)) - 1) / wDiv ;
  const double perPair = (double) (H10X_R1_LEN + H10X_R2_LEN - 2 * P.k + 2) / (double) P.w ;
  TailGeom te ;
  return tail_geometry ((uint64_t) (perPair * 1.1 * nProc) + nBlocks + 1, topQ, nBlocks, te) ;
}

static void build_device_impl (h10x_ctx *c, const uint32_t *fqb, uint64_t nFile, cudaStream_t s, bool reset = true,
			       bool dist = false, Prefuse *pre = nullptr)
{
  const h10x_params &P = c->P ;
  HostTrace tr ;
  if (reset) reset_result (c) ;
  cudaEvent_t evA = ctx_event (c), evB = ctx_event (c) ;
  CK (cudaEventRecord (evA, s)) ;

  uint64_t nRec64 = (P.N > 0 && (uint64_t) P.N < nFile) ? (uint64_t) P.N : nFile ;
  if (nRec64 >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 records on one device") ;
  uint32_t nRec = (uint32_t) nRec64 ;
  c->nReads = nRec ;
  MemTrack *mt = &c->mt ;

  /* ---------------- runs -> barcode blocks ---------------- */
  BlockTable bt ;
  DBuf<uint32_t> blkIncl (nRec, s, mt) ;	/* 1-based block number of every record */
  DBuf<uint32_t> dBlkStart ;
  int sawZeroBarcode = 0 ;
  auto doRuns = [&] ()
  {
  if (nRec == 0)
    { bt.nBlk = 1 ; bt.start = {0, 0} ; }	/* block 1 exists with nRead 0 (hash10x.c:200-201) */
  else
    { StageTimer tm (c, s, ST_RUNS) ;
      DBuf<uint32_t> flag (nRec, s, mt) ;
      DBuf<int> anyZero (1, s, mt) ;
      CK (cudaMemsetAsync (anyZero.p, 0, sizeof (int), s)) ;
      LAUNCH (c, k_run_flags, gridFor (nRec, 256), 256, 0, s, fqb, nRec, flag.p, anyZero.p) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, flag.p, blkIncl.p, nRec, s) ; }) ;
      uint32_t nRuns = 0 ; int hAnyZero = 0 ;
      CK (cudaMemcpyAsync (&nRuns, blkIncl.p + (nRec - 1), 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaMemcpyAsync (&hAnyZero, anyZero.p, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      sawZeroBarcode = hAnyZero ;
      dBlkStart.alloc ((size_t) nRuns + 1, s, mt) ;
      LAUNCH (c, k_run_starts, gridFor (nRec, 256), 256, 0, s, flag.p, blkIncl.p, nRec, dBlkStart.p) ;
      std::vector<uint32_t> runStart ((size_t) nRuns + 1) ;
      CK (cudaMemcpyAsync (runStart.data (), dBlkStart.p, 4 * ((size_t) nRuns + 1), cudaMemcpyDeviceToHost, s)) ;
      if (!hAnyZero)
	{ CK (cudaStreamSynchronize (s)) ;
	  /* blocks are the runs; "chunkSize too small" (hash10x.c:206) fires when one run fills the
	     whole chunk buffer and the loop comes round again (chunk semantics) */
	  for (uint32_t r = 0 ; r < nRuns ; ++r)
	    { uint64_t st = runStart[r], len = runStart[r+1] - st ;
	      if (len >= (uint64_t) P.chunkSize && (P.N == 0 || st + (uint64_t) P.chunkSize < (uint64_t) P.N))
		throw H10xError (H10X_ERR_CHUNK_TOO_SMALL, "chunkSize too small") ;
	    }
	  bt.nBlk = nRuns ; bt.start.swap (runStart) ;
	}
      else
	{ DBuf<uint32_t> dWord (nRuns, s, mt) ;
	  LAUNCH (c, k_gather_word0, gridFor (nRuns, 256), 256, 0, s, fqb, dBlkStart.p, nRuns, dWord.p) ;
	  std::vector<uint32_t> runWord (nRuns) ;
	  CK (cudaMemcpyAsync (runWord.data (), dWord.p, 4 * (size_t) nRuns, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaStreamSynchronize (s)) ;
	  int st = simulate_chunks (runStart, runWord, nRuns, nRec, P.chunkSize, P.N, bt) ;
	  if (st) throw H10xError (st, h10x_strerror (st)) ;
	  if (bt.nBlk != nRuns)	/* some runs were glued: rebuild the per-record block numbers */
	    { dBlkStart.alloc ((size_t) bt.nBlk + 1, s, mt) ;
	      CK (cudaMemcpyAsync (dBlkStart.p, bt.start.data (), 4 * ((size_t) bt.nBlk + 1), cudaMemcpyHostToDevice, s)) ;
	      CK (cudaMemsetAsync (flag.p, 0, 4 * (size_t) nRec, s)) ;
	      LAUNCH (c, k_flags_from_starts, gridFor (bt.nBlk, 256), 256, 0, s, dBlkStart.p, bt.nBlk, flag.p) ;
	      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, flag.p, blkIncl.p, nRec, s) ; }) ;
	      CK (cudaStreamSynchronize (s)) ;	/* bt.start is read by the async copy above */
	    }
	}
    }
  } ;

  /* multi-GPU: ranks own consecutive barcode-run ranges; agree on errors and on global block numbers */
    bool lastRank = true ; uint32_t blkBase = 0, nBlkGlobal = 0 ;
  bool blkInclReady = true ;		/* the per-record block numbers, which only the generic mosh path reads */
  bool preRuns = false ;
  if (pre && pre->valid && !pre->anyZero && nRec && !dist)
    { /* the streamed front half found the runs already; blocks are the runs (no barcode word is 0) */
      for (uint32_t r = 0 ; r < pre->nRuns ; ++r)
	{ uint64_t st = pre->runStart[r], len = pre->runStart[r+1] - st ;
	  if (len >= (uint64_t) P.chunkSize && (P.N == 0 || st + (uint64_t) P.chunkSize < (uint64_t) P.N))
	    throw H10xError (H10X_ERR_CHUNK_TOO_SMALL, "chunkSize too small") ;
	}
      bt.nBlk = pre->nRuns ; bt.start = pre->runStart ;
      dBlkStart.swap (pre->dBlkStart) ;
      blkInclReady = false ; preRuns = true ;
    }
    else if (pre)
    { pre->valid = false ; pre->eng.release () ; pre->srcOff.release () ; pre->blkCnt.release () ; pre->dBlkStart.release () ; }
  auto ensureBlkIncl = [&] ()
    { if (blkInclReady) return ;
      DBuf<uint32_t> flag (nRec, s, mt) ;
      CK (cudaMemsetAsync (flag.p, 0, 4 * (size_t) nRec, s)) ;
      LAUNCH (c, k_flags_from_starts, gridFor (bt.nBlk, 256), 256, 0, s, dBlkStart.p, bt.nBlk, flag.p) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, flag.p, blkIncl.p, nRec, s) ; }) ;
      CK (cudaStreamSynchronize (s)) ;
      blkInclReady = true ;
    } ;
  if (preRuns) {}
  else if (!dist) doRuns () ;
  else
    { int localErr = 0 ;
      try { doRuns () ; }
      catch (const H10xError &e) { localErr = e.code ; bt.nBlk = 1 ; bt.start = {0, nRec} ; }
      uint32_t words[2] = { 0, 0 } ;
      if (nRec)
	{ CK (cudaMemcpyAsync (&words[0], fqb, 4, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaMemcpyAsync (&words[1], fqb + (size_t) H10X_REC_WORDS * (nRec - 1), 4, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaStreamSynchronize (s)) ;
	}
      if (!localErr && (nRec == 0 || P.N != 0 || sawZeroBarcode)) localErr = H10X_ERR_UNSUPPORTED ;
      const uint32_t mine[4] = { bt.nBlk, words[0], words[1], nRec } ;
      std::vector<uint32_t> all ;
      c->dist->owesAgreement = false ; c->dist->globalCodes = false ;
      dist_agree (c, s, localErr, mine, all) ;		/* throws the same error on every rank */
      c->dist->owesAgreement = true ;			/* the next one is at the entry of dist_bins */
      const int R = c->dist->rank, NR = c->dist->nranks ;
      uint64_t readsGlobal = 0 ;
      for (int r = 0 ; r < NR ; ++r)
	{ if (r < R) blkBase += all[4*r] ;
	  nBlkGlobal += all[4*r] ; readsGlobal += all[4*r + 3] ;
	  if (r + 1 < NR && all[4*r + 2] == all[4*(r+1) + 1])
	    throw H10xError (H10X_ERR_UNSUPPORTED, "multi-GPU shards must be cut at barcode-run boundaries") ;
	}
      lastRank = (R == NR - 1) ;
      if (const char *e = getenv ("H10X_TEST_FAIL_RANK"))	/* tests: a rank-local failure between two agreements */
	if (atoi (e) == R) throw H10xError (H10X_ERR_UNSUPPORTED, "forced failure (H10X_TEST_FAIL_RANK)") ;
      c->dist->blockBase = blkBase ; c->dist->nBlocksGlobal = nBlkGlobal ; c->dist->nReadsGlobal = readsGlobal ;
    }

  tr.mark ("runs") ;
  const uint32_t nBlk = bt.nBlk ;
  const uint32_t nProcBlk = nBlk - (lastRank ? 1 : 0) ;	/* the (globally) final run is never hashed (hash10x.c:209,216) */
  const uint32_t nProc = bt.start[nProcBlk] ;	/* records of the processed blocks */
  c->nBlocksMax = nBlk + 1 ;

  /* ---------------- moshes -> per-block unique (hash, read) lists ---------------- */
  DBuf<uint64_t> blkOffProc ((size_t) nProcBlk + 1, s, mt) ;
  DBuf<uint64_t> srcOff ((size_t) nProcBlk + 1, s, mt) ;
  DBuf<uint32_t> blkCnt ((size_t) nProcBlk + 1, s, mt) ;
  std::vector<uint64_t> hBlkOff ((size_t) nProcBlk + 1, 0) ;
  std::vector<uint32_t> hCnt ((size_t) nProcBlk + 1, H10X_BLK_FALLBACK) ;
  uint64_t H = 0, totalMoshes = 0 ;

    /* -- fused path: one CTA per block, everything in shared memory (h10x_fused.cuh) -- */
  const bool fusedOK = FusedEngine::usable (P) && nProcBlk > 0 ;
  FusedEngine eng ;
  DBuf<uint64_t> &scratch = eng.scratch ;
  uint64_t nFused = 0 ;
  DBuf<uint64_t> gHash ; DBuf<uint32_t> gRec ;
  uint64_t G = 0 ; size_t gCap = 0 ;
  /* the global sort runs on hash / w: exact division = multiplication by w^-1 mod 2^64 (w odd part) and a
     shift (power-of-two part); quotients of multiples keep the order and need fewer radix passes */
  const uint64_t wInvFull = c->hp.wTz ? 1 : c->hp.wInv ;
  const uint64_t wDiv = c->hp.wTz ? 1 : (uint64_t) P.w ;
  const uint64_t topQ = (((uint64_t) 1 << (2 * P.k)) - 1) / wDiv ;
  /* single-GPU: the hand-written tail (h10x_tail.cuh) whenever its packed 64-bit entry fits; the library-sort tail
     of round 1 stays as the general path (multi-GPU, very wide hashes) and as a cross-check (H10X_FLAG_LEGACY_TAIL).
     The hand-written tail sorts every key itself, so the fused kernel then runs LEAN (no in-CTA sort / dedup); should
     the real key count turn the tail's geometry down after all, the mosh stage is repeated the classic way. */
  TailGeom tg ; memset (&tg, 0, sizeof (tg)) ;
  const bool tailWanted = !(P.flags & H10X_FLAG_LEGACY_TAIL) && !getenv ("H10X_LEGACY_TAIL") ;
  bool tail2 = false, lean = false ;
  for (int attempt = 0 ; ; ++attempt)
  {
  const bool usePre = attempt == 0 && preRuns && pre->eng.ready && fusedOK ;
  lean = usePre ? pre->eng.lean : (attempt == 0 && lean_wanted (c, dist, nProc, blkBase + nProcBlk)) ;
  std::fill (hCnt.begin (), hCnt.end (), H10X_BLK_FALLBACK) ;
  H = 0 ; totalMoshes = 0 ; nFused = 0 ; G = 0 ;
  if (usePre)
    { /* the lists of every run but the last are there already (and the last run is never hashed) */
      eng.swap (pre->eng) ; srcOff.swap (pre->srcOff) ; blkCnt.swap (pre->blkCnt) ;
    }
  else if (fusedOK)
    { StageTimer tm (c, s, ST_FUSED) ;
      eng.release () ;
      eng.init (c, s, lean, nProc, 1) ;
      std::vector<uint32_t> lists[5] ;
      for (uint32_t p = 0 ; p < nProcBlk ; ++p)
	{ const int ci = eng.classOf (P, bt.start[p+1] - bt.start[p]) ;
	  if (ci >= 0) lists[ci].push_back (p) ;
	}
      CK (cudaMemsetAsync (blkCnt.p, 0xff, 4 * ((size_t) nProcBlk + 1), s)) ;
      eng.launch (c, s, lists, fqb, dBlkStart.p, srcOff.p, blkCnt.p) ;
      CK (cudaStreamSynchronize (s)) ;	/* the lists are read by the async copies of launch () */
    }
  if (fusedOK)
    { CK (cudaMemcpyAsync (hCnt.data (), blkCnt.p, 4 * (size_t) nProcBlk, cudaMemcpyDeviceToHost, s)) ;
      unsigned long long fusedMoshes = 0 ;
      CK (cudaMemcpyAsync (&fusedMoshes, eng.cursor.p + 1, 8, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      totalMoshes += fusedMoshes ;
      for (uint32_t p = 0 ; p < nProcBlk ; ++p) if (hCnt[p] != H10X_BLK_FALLBACK) ++nFused ;
    }
  c->stats.fusedBlocks = nFused ; c->stats.genericBlocks = nProcBlk - nFused ;
  tr.mark ("fused") ;

  /* -- generic path for whatever the fused path did not take: global-memory segmented sort -- */
    auto ensureG = [&] (uint64_t need)
    { if (need <= gCap) return ;
      size_t cap = std::max<size_t> ((size_t) need, gCap + gCap / 2) ;
      DBuf<uint64_t> nh (cap, s, mt) ; DBuf<uint32_t> nr (cap, s, mt) ;
      if (G) { CK (cudaMemcpyAsync (nh.p, gHash.p, 8 * G, cudaMemcpyDeviceToDevice, s)) ;
	       CK (cudaMemcpyAsync (nr.p, gRec.p, 4 * G, cudaMemcpyDeviceToDevice, s)) ; }
      gHash.swap (nh) ; gRec.swap (nr) ; gCap = cap ;
    } ;
  const uint32_t batchRecs = 8u << 20 ;		/* 8M records: at most 2^31 moshes even if every k-mer is one */
  uint32_t q0 = 0 ;
  while (q0 < nProcBlk)
    { if (hCnt[q0] != H10X_BLK_FALLBACK) { ++q0 ; continue ; }
      /* a maximal run of consecutive generic blocks, cut into batches of at most batchRecs records */
      uint32_t p0 = q0, p1 = p0 + 1 ;
      while (p1 < nProcBlk && hCnt[p1] == H10X_BLK_FALLBACK && bt.start[p1+1] - bt.start[p0] <= batchRecs) ++p1 ;
      q0 = p1 ;
            uint32_t r0 = bt.start[p0], nb = bt.start[p1] - r0, nblk = p1 - p0 ;
      ensureBlkIncl () ;
      if ((uint64_t) nb * 237 > 0x7fffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "barcode block too large for the generic path") ;
      DBuf<uint32_t> cnt2 ((size_t) 2 * nb + 1, s, mt), off2 ((size_t) 2 * nb + 1, s, mt) ;
      DBuf<uint32_t> ph ((size_t) nblk + 1, s, mt), phOff ((size_t) nblk + 1, s, mt), segOff ((size_t) nblk + 1, s, mt) ;
      DBuf<uint64_t> gOff ((size_t) nblk + 1, s, mt) ;
      uint32_t Mb = 0 ;
      { StageTimer tm (c, s, ST_MOSHES) ;
	CK (cudaMemsetAsync (cnt2.p + 2 * (size_t) nb, 0, 4, s)) ;
	LAUNCH (c, k_moshes<false>, gridFor (2 * (uint64_t) nb, 128), 128, 0, s, fqb, r0, nb, c->hp, cnt2.p,
		(const uint32_t*) nullptr, p0, (const uint32_t*) nullptr, (uint64_t*) nullptr, (uint32_t*) nullptr) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, cnt2.p, off2.p, 2 * (size_t) nb + 1, s) ; }) ;
	LAUNCH (c, k_blk_moshes, gridFor (nblk + 1, 256), 256, 0, s, dBlkStart.p, p0, nblk, r0, off2.p, ph.p) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, ph.p, phOff.p, (size_t) nblk + 1, s) ; }) ;
	LAUNCH (c, k_seg_off, gridFor (nblk + 1, 256), 256, 0, s, dBlkStart.p, p0, nblk, r0, off2.p, phOff.p, ph.p, segOff.p,
		(uint64_t*) nullptr, (uint32_t*) nullptr) ;
	uint32_t rawMoshes = 0 ;	/* without the phantom entries of empty blocks */
	CK (cudaMemcpyAsync (&Mb, segOff.p + nblk, 4, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaMemcpyAsync (&rawMoshes, off2.p + 2 * (size_t) nb, 4, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	totalMoshes += rawMoshes ;
      }
      DBuf<uint64_t> keys (Mb, s, mt), keysS (Mb, s, mt) ;
      DBuf<uint32_t> vals (Mb, s, mt), valsS (Mb, s, mt) ;
      { StageTimer tm (c, s, ST_MOSHES) ;
	LAUNCH (c, k_moshes<true>, gridFor (2 * (uint64_t) nb, 128), 128, 0, s, fqb, r0, nb, c->hp, off2.p,
		blkIncl.p, p0, phOff.p, keys.p, vals.p) ;
	LAUNCH (c, k_seg_off, gridFor (nblk + 1, 256), 256, 0, s, dBlkStart.p, p0, nblk, r0, off2.p, phOff.p, ph.p, segOff.p,
		keys.p, vals.p) ;
      }
      { StageTimer tm (c, s, ST_BLOCKSORT) ;
	/* stable: among equal hashes the first generated (lowest read index) stays first, which is
	   the entry glibc's stable qsort leaves first for the dedup of hash10x.c:166-172 */
	cubCall (c, s, [&] (void *t, size_t &b)
	  { return cub::DeviceSegmentedSort::StableSortPairs (t, b, keys.p, keysS.p, vals.p, valsS.p, (int) Mb, (int) nblk,
							       segOff.p, segOff.p + 1, s) ; }) ;
      }
      uint32_t Hb = 0 ;
      { StageTimer tm (c, s, ST_DEDUP) ;
	keys.release () ; vals.release () ;
	DBuf<uint32_t> uflag (Mb, s, mt), uincl (Mb, s, mt) ;
	LAUNCH (c, k_uniq_flag, gridFor (Mb, 256), 256, 0, s, keysS.p, valsS.p, Mb, blkIncl.p, uflag.p) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, uflag.p, uincl.p, Mb, s) ; }) ;
	CK (cudaMemcpyAsync (&Hb, uincl.p + (Mb - 1), 4, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	ensureG (G + Hb) ;
	LAUNCH (c, k_compact, gridFor (Mb, 256), 256, 0, s, keysS.p, valsS.p, Mb, uflag.p, uincl.p, G, gHash.p, gRec.p) ;
	LAUNCH (c, k_blk_off, gridFor (nblk, 256), 256, 0, s, segOff.p, uincl.p, nblk, G, gOff.p) ;
	std::vector<uint64_t> hOff ((size_t) nblk + 1) ;
	CK (cudaMemcpyAsync (hOff.data (), gOff.p, 8 * (size_t) nblk, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	hOff[nblk] = G + Hb ;
	std::vector<uint64_t> src (nblk) ;
	for (uint32_t i = 0 ; i < nblk ; ++i)
	  { hCnt[p0 + i] = (uint32_t) (hOff[i+1] - hOff[i]) ; src[i] = hOff[i] | ((uint64_t) 1 << 63) ; }
	CK (cudaMemcpyAsync (srcOff.p + p0, src.data (), 8 * (size_t) nblk, cudaMemcpyHostToDevice, s)) ;
	CK (cudaMemcpyAsync (blkCnt.p + p0, hCnt.data () + p0, 4 * (size_t) nblk, cudaMemcpyHostToDevice, s)) ;
	CK (cudaStreamSynchronize (s)) ;
      }
      G += Hb ;
    }

    tr.mark ("generic") ;
  for (uint32_t p = 0 ; p < nProcBlk ; ++p) { hBlkOff[p] = H ; H += hCnt[p] ; }
  hBlkOff[nProcBlk] = H ;
  tail2 = tailWanted && H > 0 && tail_geometry (H, topQ, blkBase + nProcBlk, tg) ;
    if (lean && !tail2) { eng.release () ; continue ; }
  break ;
  }
  int sortBits = 2 * P.k ;
  { uint64_t top = (((uint64_t) 1 << (2 * P.k)) - 1) / wDiv ; sortBits = 1 ; while (sortBits < 64 && (top >> sortBits)) ++sortBits ; }
  /* -- final placement in block order: eHash / eRead / entryBlk -- */
    c->nHashes = H ;
  if (H >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 block-unique hashes on one device") ;
  /* single-GPU: bucket-major (key, value) pairs (h10x_bucket.cuh); multi-GPU: block-major hash / read / block arrays */
  int blkBits = 1 ; while (((uint64_t) 1 << blkBits) < (uint64_t) nBlk + 2) ++blkBits ;
  const int nbBits = sortBits > 32 ? sortBits - 32 : 0 ;
    c->stats.tailPath = tail2 ? 2 : 1 ;
  const bool bucketed = !tail2 && !dist && nbBits <= 8 && sortBits + blkBits <= 64 && H < 0x7fffffffull && H > 0 ;
  uint32_t nBuck = 1 ;
  std::vector<uint64_t> hBucketBase ;
  DBuf<uint64_t> eHash, eBR, bucketBase ; DBuf<uint16_t> eRead ; DBuf<uint32_t> entryBlk, key32 ;
  DBuf<uint64_t> tailB, tailRangeStart ;
  if (dist && !tail2) { eHash.alloc (H, s, mt) ; eRead.alloc (H, s, mt) ; entryBlk.alloc (H, s, mt) ; }
  else if (!bucketed && !tail2) { eHash.alloc (H, s, mt) ; eBR.alloc (H, s, mt) ; }
  { StageTimer tm (c, s, ST_DEDUP) ;
    CK (cudaMemcpyAsync (blkOffProc.p, hBlkOff.data (), 8 * ((size_t) nProcBlk + 1), cudaMemcpyHostToDevice, s)) ;
    if (c->dbgKeys)
      { unsigned long long used = 0 ;
	CK (cudaMemcpyAsync (&used, eng.cursor.p, 8, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	c->dbgNBlk = nProcBlk ; c->dbgLean = lean ; c->dbgScratchLen = used ;
	c->dbgSrcOff.assign (nProcBlk, 0) ; c->dbgBlkCnt.assign (nProcBlk, 0) ; c->dbgScratch.assign ((size_t) used, 0) ;
	CK (cudaMemcpyAsync (c->dbgSrcOff.data (), srcOff.p, 8 * (size_t) nProcBlk, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaMemcpyAsync (c->dbgBlkCnt.data (), blkCnt.p, 4 * (size_t) nProcBlk, cudaMemcpyDeviceToHost, s)) ;
	if (used) CK (cudaMemcpyAsync (c->dbgScratch.data (), scratch.p, 8 * (size_t) used, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
      }
    if (tail2)
      { TailSrc in = { nProcBlk, blkBase, srcOff.p, blkCnt.p, scratch.p, gHash.p, gRec.p, dBlkStart.p, wInvFull, wDiv } ;
	tail_p1 (c, s, tg, in, H, tailB, tailRangeStart) ;
      }
    else if (bucketed)
      { uint64_t top = (((uint64_t) 1 << (2 * P.k)) - 1) / wDiv ;
	nBuck = (uint32_t) (top >> 32) + 1 ;
	const size_t nCnt = (size_t) nBuck * nProcBlk ;
	DBuf<uint32_t> cnt (nCnt + 1, s, mt) ; DBuf<uint64_t> off (nCnt + 1, s, mt) ;
	CK (cudaMemsetAsync (cnt.p + nCnt, 0, 4, s)) ;
	LAUNCH (c, k_bucket_count, gridFor ((uint64_t) nProcBlk * 32, 256), 256, 0, s, nProcBlk, nBuck, srcOff.p, blkCnt.p,
		scratch.p, gHash.p, wDiv, cnt.p) ;
	cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> cnt64 (cnt.p, CastU64 ()) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, cnt64, off.p, nCnt + 1, s) ; }) ;
	bucketBase.alloc ((size_t) nBuck + 1, s, mt) ;
	LAUNCH (c, k_bucket_base, gridFor (nBuck + 1, 256), 256, 0, s, nBuck, nProcBlk, off.p, H, bucketBase.p) ;
	hBucketBase.resize ((size_t) nBuck + 1) ;
	CK (cudaMemcpyAsync (hBucketBase.data (), bucketBase.p, 8 * ((size_t) nBuck + 1), cudaMemcpyDeviceToHost, s)) ;
	key32.alloc (H, s, mt) ; eBR.alloc (H, s, mt) ;
	LAUNCH (c, k_place_bucketed_kv, std::min<uint32_t> (nProcBlk, 148 * 16), 256, 0, s, nProcBlk, nBuck, srcOff.p, blkCnt.p,
		cnt.p, off.p, scratch.p, gHash.p, gRec.p, dBlkStart.p, blkBase, wInvFull, key32.p, eBR.p) ;
	CK (cudaStreamSynchronize (s)) ;
      }
    else
      { if (nProcBlk)
	  LAUNCH (c, k_place, std::min<uint32_t> (nProcBlk, 148 * 16), 256, 0, s, nProcBlk, srcOff.p, blkCnt.p, blkOffProc.p,
		  scratch.p, gHash.p, gRec.p, dBlkStart.p, blkBase, wInvFull, eHash.p, eRead.p, entryBlk.p, eBR.p) ;
	CK (cudaStreamSynchronize (s)) ;	/* hBlkOff is read by the async copy */
      }
  }
    eng.release () ; gHash.release () ; gRec.release () ; srcOff.release () ; blkCnt.release () ;

  tr.mark ("place") ;
  /* ---------------- bins: ids, values, depths ---------------- */
  uint32_t D = 0 ;
  DBuf<uint32_t> entryId ;
  if (dist && !tail2) entryId.alloc (H, s, mt) ;
  DBuf<uint32_t> se, segIncl, segStart, idOfSeg, sk ;
  DBuf<uint64_t> sv ;
  if (!bucketed && !tail2) segIncl.alloc (H, s, mt) ;
    if (tail2)
    { std::vector<uint32_t> blkDup ;
      TailDist td = { nBlkGlobal } ;
      D = tail_rest (c, s, tg, H, wDiv, tailB, tailRangeStart, blkBase + nProcBlk, blkDup, dist ? &td : nullptr) ;
      /* per-block unique counts, now that the duplicates are gone */
      uint64_t run = 0 ;
      for (uint32_t p = 0 ; p < nProcBlk ; ++p) { hBlkOff[p] = run ; hCnt[p] -= blkDup[blkBase + p + 1] ; run += hCnt[p] ; }
      hBlkOff[nProcBlk] = run ;
      if (run != H) throw H10xError (H10X_ERR_CUDA, "internal: duplicate count mismatch in the grouping tail") ;
      c->nHashes = H ;
    }
  else if (bucketed)
    { sk.alloc (H, s, mt) ; sv.alloc (H, s, mt) ;
      { StageTimer tm (c, s, ST_HASHSORT) ;
	const int endBit = std::min (32, sortBits) ;
	for (uint32_t v = 0 ; v < nBuck ; ++v)
	  { uint64_t b0 = hBucketBase[v], n = hBucketBase[v+1] - b0 ;
	    if (!n) continue ;
	    /* stable: inside a bin the placement order = ascending block is kept */
	    cubCall (c, s, [&] (void *t, size_t &b)
	      { return cub::DeviceRadixSort::SortPairs (t, b, key32.p + b0, sk.p + b0, eBR.p + b0, sv.p + b0, n, 0, endBit, s) ; }) ;
	  }
      }
      key32.release () ; eBR.release () ;
      { StageTimer tm (c, s, ST_BINIDS) ;
	/* first position of every bin, in one selection pass over the sorted keys (the bound on D is only known
	   afterwards, so the output is sized for the worst case and released with the stage) */
	segStart.alloc ((size_t) H + 1, s, mt) ;
	DBuf<uint32_t> dD (1, s, mt) ;
	HeadPredKV pred = { sk.p, bucketBase.p, nBuck, (uint64_t) ((((unsigned __int128) nBuck) << 64) / H) } ;
	cub::CountingInputIterator<uint32_t> iota (0u) ;
	cubCall (c, s, [&] (void *t, size_t &b)
	  { return cub::DeviceSelect::If (t, b, iota, segStart.p, dD.p, (::cuda::std::int64_t) H, pred, s) ; }) ;
	CK (cudaMemcpyAsync (&D, dD.p, 4, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	if ((uint64_t) D + 1 > (((uint64_t) 1 << P.B) >> 2) - 2)	/* hash10x.c:149 */
	  throw H10xError (H10X_ERR_TABLE_TOO_SMALL, "hashTableSize is too small") ;
	const uint32_t H32 = (uint32_t) H ;
	CK (cudaMemcpyAsync (segStart.p + D, &H32, 4, cudaMemcpyHostToDevice, s)) ;	/* pageable source: copied before the call returns */
	idOfSeg.alloc (D, s, mt) ;
	DBuf<uint64_t> fkey (D, s, mt), fkeyS (D, s, mt) ; DBuf<uint32_t> segIdx (D, s, mt), sortedSeg (D, s, mt) ;
	LAUNCH (c, k_first_key_kv, gridFor (D, 256), 256, 0, s, D, segStart.p, sk.p, sv.p, bucketBase.p, nBuck, sortBits, fkey.p, segIdx.p) ;
	cubCall (c, s, [&] (void *t, size_t &b)
	  { return cub::DeviceRadixSort::SortPairs (t, b, fkey.p, fkeyS.p, segIdx.p, sortedSeg.p, D, 0, sortBits + blkBits, s) ; }) ;
	c->hashNumber = D + 1 ;
	c->hashValue.alloc ((size_t) D + 1, s, mt) ;
	c->hashDepth.alloc ((size_t) D + 2, s, mt) ;
	CK (cudaMemsetAsync (c->hashValue.p, 0, 8, s)) ;
	CK (cudaMemsetAsync (c->hashDepth.p, 0, 4, s)) ;
	CK (cudaMemsetAsync (c->hashDepth.p + D + 1, 0, 4, s)) ;
	LAUNCH (c, k_bins_by_rank_kv, gridFor (D, 256), 256, 0, s, D, sortedSeg.p, fkeyS.p, segStart.p, sortBits, wDiv,
		idOfSeg.p, c->hashValue.p, c->hashDepth.p) ;
      }
    }
  else if (H)
    { se.alloc (H, s, mt) ;
      DBuf<uint64_t> sh (H, s, mt) ;
      { StageTimer tm (c, s, ST_HASHSORT) ;
	DBuf<uint32_t> iota (H, s, mt) ;
	LAUNCH (c, k_iota, gridFor (H, 256), 256, 0, s, iota.p, H) ;
	/* stable LSD radix sort over the 2k hash bits: inside a bin, entries keep ascending index */
	cubCall (c, s, [&] (void *t, size_t &b)
	  { return cub::DeviceRadixSort::SortPairs (t, b, eHash.p, sh.p, iota.p, se.p, H, 0, sortBits, s) ; }) ;
      }
      { StageTimer tm (c, s, ST_BINIDS) ;
	DBuf<uint32_t> head (H, s, mt) ;
	LAUNCH (c, k_head_flag, gridFor (H, 256), 256, 0, s, sh.p, H, head.p) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, head.p, segIncl.p, H, s) ; }) ;
	CK (cudaMemcpyAsync (&D, segIncl.p + (H - 1), 4, cudaMemcpyDeviceToHost, s)) ;
	CK (cudaStreamSynchronize (s)) ;
	segStart.alloc ((size_t) D + 1, s, mt) ;
	if (!dist)
	  { /* hash10x.c:149: die once hashNumber exceeds 2^(B-2) - 2 */
	    if ((uint64_t) D + 1 > (((uint64_t) 1 << P.B) >> 2) - 2)
	      throw H10xError (H10X_ERR_TABLE_TOO_SMALL, "hashTableSize is too small") ;
	    idOfSeg.alloc (D, s, mt) ;
	    LAUNCH (c, k_seg_start, gridFor (H, 256), 256, 0, s, head.p, segIncl.p, H, se.p, segStart.p, (uint32_t*) nullptr) ;
	    head.release () ;
	    DBuf<uint32_t> firstE (D, s, mt), segIdx (D, s, mt), firstS (D, s, mt), sortedSeg (D, s, mt) ;
	    LAUNCH (c, k_first_entry, gridFor (D, 256), 256, 0, s, D, segStart.p, se.p, firstE.p, segIdx.p) ;
	    int eBits = 1 ; while (((uint64_t) 1 << eBits) < H) ++eBits ;
	    cubCall (c, s, [&] (void *t, size_t &b)
	      { return cub::DeviceRadixSort::SortPairs (t, b, firstE.p, firstS.p, segIdx.p, sortedSeg.p, D, 0, eBits, s) ; }) ;
	    c->hashNumber = D + 1 ;
	    c->hashValue.alloc ((size_t) D + 1, s, mt) ;
	    c->hashDepth.alloc ((size_t) D + 2, s, mt) ;	/* one spare 0 so the scan yields codeOff[hashNumber] */
	    CK (cudaMemsetAsync (c->hashValue.p, 0, 8, s)) ;
	    CK (cudaMemsetAsync (c->hashDepth.p, 0, 4, s)) ;
	    CK (cudaMemsetAsync (c->hashDepth.p + D + 1, 0, 4, s)) ;
	    LAUNCH (c, k_bins_by_rank, gridFor (D, 256), 256, 0, s, D, sortedSeg.p, segStart.p, sh.p, wDiv, idOfSeg.p,
		    c->hashValue.p, c->hashDepth.p) ;
	  }
	else
	  LAUNCH (c, k_seg_start, gridFor (H, 256), 256, 0, s, head.p, segIncl.p, H, se.p, segStart.p, (uint32_t*) nullptr) ;
      }
      if (dist)
	{ uint32_t Dl = D ;
	  dist_bins (c, s, H, sh.p, se.p, segIncl.p, Dl, segStart.p, entryBlk.p, nBlkGlobal, entryId.p, D, wDiv) ;
	}
    }
  else if (!dist)
    { c->hashNumber = 1 ;
      c->hashValue.alloc (1, s, mt) ; c->hashDepth.alloc (2, s, mt) ;
      CK (cudaMemsetAsync (c->hashValue.p, 0, 8, s)) ; CK (cudaMemsetAsync (c->hashDepth.p, 0, 8, s)) ;
    }
  else
    { segStart.alloc (1, s, mt) ; CK (cudaMemsetAsync (segStart.p, 0, 4, s)) ;
      dist_bins (c, s, 0, nullptr, nullptr, nullptr, 0, segStart.p, nullptr, nBlkGlobal, nullptr, D, wDiv) ;
    }

  if (!dist && !tail2) { early_pull (c, s, SLOT_VALUE, c->hashValue.p, 8 * (size_t) c->hashNumber) ; early_pull (c, s, SLOT_DEPTH, c->hashDepth.p, 4 * (size_t) c->hashNumber) ; }
  tr.mark ("bins-enq") ;
  /* ---------------- hash -> code CSR ---------------- */
  if (dist && tail2) ;	/* tail_rest has left localBinId / localCodeOff / localCodes and the ClusterHash lists */
  else if (dist)
    { /* this rank's part of every bin's barcode list: bin localBinId[j] holds the (global) blocks
	 localCodes[localCodeOff[j] .. localCodeOff[j+1]), ascending; the full list of a bin is the
	 concatenation over ranks in rank order, because ranks own ascending block ranges */
      StageTimer tm (c, s, ST_CODES) ;
      c->localCodes.alloc (H, s, mt) ;
      uint32_t Dl = c->dist->nLocalBins ;
      c->localCodeOff.alloc ((size_t) Dl + 1, s, mt) ;
      CK (cudaMemcpyAsync (c->localCodeOff.p, segStart.p, 4 * ((size_t) Dl + 1), cudaMemcpyDeviceToDevice, s)) ;
      if (H) LAUNCH (c, k_gather_u32, gridFor (H, 256), 256, 0, s, H, se.p, entryBlk.p, c->localCodes.p) ;
    }
  else if (!tail2)
    { /* codes (bin-major block lists) are needed as the sort key even under H10X_FLAG_NO_CODES */
      size_t hn = c->hashNumber ;
      DBuf<uint64_t> idRead (H, s, mt) ;
      { StageTimer tm (c, s, ST_CODES) ;
	c->codeOff.alloc (hn + 1, s, mt) ;
	c->codes.alloc (H, s, mt) ;
	cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> depth64 (c->hashDepth.p, CastU64 ()) ;
	cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, depth64, c->codeOff.p, hn + 1, s) ; }) ;
	if (H && bucketed)
	  LAUNCH (c, k_codes_seg_kv, gridFor (((uint64_t) D + 31) / 32 * 32, 256), 256, 0, s, D, segStart.p, idOfSeg.p, sv.p,
		  c->codeOff.p, c->codes.p, idRead.p) ;
	else if (H)
	  LAUNCH (c, k_codes_tr, gridFor (H, 256), 256, 0, s, H, segIncl.p, segStart.p, idOfSeg.p, se.p, eBR.p,
		  c->codeOff.p, c->codes.p, idRead.p) ;
      }
      se.release () ; sk.release () ; sv.release () ; segIncl.release () ; segStart.release () ; idOfSeg.release () ;
      eHash.release () ; eBR.release () ;
      /* the block sort below only reads codes[]: its copy to the host runs beside the sort */
            if (!(P.flags & (H10X_FLAG_NO_CODES | H10X_FLAG_LAZY_CODES)))
	{ early_pull (c, s, SLOT_CODEOFF, c->codeOff.p, 8 * (hn + 1)) ; early_pull (c, s, SLOT_CODES, c->codes.p, 4 * H) ; }
      c->clus.alloc (H, s, mt) ;
      if (H)
	{ StageTimer tm (c, s, ST_CLUSTERS) ;
	  DBuf<uint32_t> keysOut (H, s, mt) ;
	  int bBits = 1 ; while (((uint64_t) 1 << bBits) < (uint64_t) nBlk + 2) ++bBits ;
	  /* stable: inside a block the bin-major order, i.e. ascending bin id, is kept */
	  cubCall (c, s, [&] (void *t, size_t &b)
	    { return cub::DeviceRadixSort::SortPairs (t, b, c->codes.p, keysOut.p, idRead.p, c->clus.p, H, 0, bBits, s) ; }) ;
	}
      if (P.flags & H10X_FLAG_NO_CODES) { c->codes.release () ; c->codeOff.release () ; }
      early_pull (c, s, SLOT_CLUS, c->clus.p, 8 * H) ;
    }
  se.release () ; segIncl.release () ; segStart.release () ; idOfSeg.release () ; eHash.release () ; entryBlk.release () ;

  /* ---------------- code -> hash lists (multi-GPU: ids came back per entry; sort inside each block) ---------------- */
  if (dist && !tail2) c->clus.alloc (H, s, mt) ;
  if (dist && !tail2 && H)
    { StageTimer tm (c, s, ST_CLUSTERS) ;
      struct ClusClass { uint32_t cap, threads ; } ;
      static const ClusClass kCC[3] = { { 1024, 128 }, { 4096, 256 }, { 12288, 512 } } ;
      int idBits = 1 ; while (((uint64_t) 1 << idBits) < (uint64_t) c->hashNumber) ++idBits ;
      uint32_t digitBits = 8, passes = (idBits + 7) / 8 ;
      for (uint32_t db = 9 ; db <= 10 ; ++db) if ((idBits + db - 1) / db < passes) { digitBits = db ; passes = (idBits + db - 1) / db ; }
      std::vector<uint32_t> lists[3] ;
      std::vector<std::pair<uint32_t, uint32_t>> bigRuns ;	/* blocks beyond the largest class */
      for (uint32_t p = 0 ; p < nProcBlk ; ++p)
	{ uint32_t n = hCnt[p] ; int ci = n <= kCC[0].cap ? 0 : n <= kCC[1].cap ? 1 : n <= kCC[2].cap ? 2 : -1 ;
	  if (ci >= 0) lists[ci].push_back (p) ;
	  else if (!bigRuns.empty () && bigRuns.back ().second == p) bigRuns.back ().second = p + 1 ;
	  else bigRuns.push_back ({ p, p + 1 }) ;
	}
      int nSM = 148 ;
      CK (cudaDeviceGetAttribute (&nSM, cudaDevAttrMultiProcessorCount, P.device)) ;
      DBuf<unsigned int> cwork (3, s, mt) ;
      CK (cudaMemsetAsync (cwork.p, 0, 12, s)) ;
      std::vector<DBuf<uint32_t>> dl (3) ;
      for (int ci = 0 ; ci < 3 ; ++ci)
	{ if (lists[ci].empty ()) continue ;
	  const ClusClass &cc = kCC[ci] ;
	  dl[ci].alloc (lists[ci].size (), s, mt) ;
	  CK (cudaMemcpyAsync (dl[ci].p, lists[ci].data (), 4 * lists[ci].size (), cudaMemcpyHostToDevice, s)) ;
	  const uint32_t nd = 1u << digitBits ;
	  size_t smem = (size_t) cc.cap * 12 + 4 + (size_t) nd * 4 + (size_t) (cc.threads / 32) * nd * 2 + 16 ;
	  const void *fn = cc.threads == 128 ? (const void*) k_cluster_sort<128> : cc.threads == 256 ? (const void*) k_cluster_sort<256>
	    : (const void*) k_cluster_sort<512> ;
	  CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
	  int occ = 1 ;
	  CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, (int) cc.threads, smem)) ;
	  if (occ < 1) occ = 1 ;
	  ClusterArgs ca ;
	  ca.list = dl[ci].p ; ca.blkOff = blkOffProc.p ; ca.entryId = entryId.p ; ca.eRead = eRead.p ; ca.clus = c->clus.p ; ca.valsOut = nullptr ;
	  ca.work = cwork.p + ci ; ca.nList = (uint32_t) lists[ci].size () ; ca.cap = cc.cap ; ca.digitBits = digitBits ; ca.passes = passes ;
	  void *args[1] = { (void*) &ca } ;
	  uint32_t grid = (uint32_t) std::min<size_t> (lists[ci].size (), (size_t) nSM * occ) ;
	  CK (cudaLaunchKernel (fn, dim3 (grid), dim3 (cc.threads), args, smem, s)) ;
	  ++c->launches ;
	}
      if (!bigRuns.empty ())
	{ DBuf<uint16_t> rdS (H, s, mt) ;
	  DBuf<uint32_t> idS (H, s, mt) ;
	  for (auto &r : bigRuns)
	    { segmented_sort_blocks<uint32_t, uint16_t> (c, s, entryId.p, idS.p, eRead.p, rdS.p, hBlkOff, blkOffProc.p, false, r.first, r.second) ;
	      uint64_t a0 = hBlkOff[r.first], n = hBlkOff[r.second] - a0 ;
	      if (n) LAUNCH (c, k_clus_pack, gridFor (n, 256), 256, 0, s, n, idS.p + a0, rdS.p + a0, c->clus.p + a0) ;
	    }
	}
      CK (cudaStreamSynchronize (s)) ;	/* block lists are read by the async copies */
    }
  entryId.release () ; eRead.release () ; entryBlk.release () ;

  /* ---------------- hashIndex[] ---------------- */
  if (!(P.flags & H10X_FLAG_NO_TABLE) && c->hashValue.p && !c->hashIndex.p)
    { StageTimer tm (c, s, ST_TABLE) ;
      size_t tableSize = (size_t) 1 << P.B ;
      c->hashIndex.alloc (tableSize, s, mt) ;
      CK (cudaMemsetAsync (c->hashIndex.p, 0, 4 * tableSize, s)) ;
      if (D) LAUNCH (c, k_table_insert, gridFor (D, 256), 256, 0, s, c->hashNumber, c->hashValue.p, c->hashIndex.p, P.B) ;
    }

  tr.mark ("rest-enq") ;
  /* ---------------- block table in the reference's numbering ---------------- */
  { StageTimer tm (c, s, ST_OTHER) ;
    std::vector<uint32_t> nRead ((size_t) nBlk + 1, 0), nHash ((size_t) nBlk + 1, 0) ;
    std::vector<uint64_t> off ((size_t) nBlk + 2, 0) ;
    for (uint32_t b = 1 ; b <= nBlk ; ++b)
      { nRead[b] = bt.start[b] - bt.start[b-1] ;
	if (b <= nProcBlk) { nHash[b] = (uint32_t) (hBlkOff[b] - hBlkOff[b-1]) ; off[b] = hBlkOff[b-1] ; }
	else off[b] = H ;
      }
    off[nBlk + 1] = H ;
    c->blkNRead.alloc ((size_t) nBlk + 1, s, mt) ; c->blkNHash.alloc ((size_t) nBlk + 1, s, mt) ;
    c->blkOff.alloc ((size_t) nBlk + 2, s, mt) ;
    CK (cudaMemcpyAsync (c->blkNRead.p, nRead.data (), 4 * ((size_t) nBlk + 1), cudaMemcpyHostToDevice, s)) ;
    CK (cudaMemcpyAsync (c->blkNHash.p, nHash.data (), 4 * ((size_t) nBlk + 1), cudaMemcpyHostToDevice, s)) ;
    CK (cudaMemcpyAsync (c->blkOff.p, off.data (), 8 * ((size_t) nBlk + 2), cudaMemcpyHostToDevice, s)) ;
    CK (cudaStreamSynchronize (s)) ;
  }

  CK (cudaEventRecord (evB, s)) ;
  CK (cudaEventSynchronize (evB)) ;
  tr.mark ("final-sync") ;
  float ms = 0 ; CK (cudaEventElapsedTime (&ms, evA, evB)) ;
  h10x_stats &st = c->stats ;
  st.msTotal = ms ;
  for (auto &sp : c->spans) { float t = 0 ; cudaEventElapsedTime (&t, sp.a, sp.b) ; st.msStage[sp.stage] += t ; }
  st.nRecords = nRec ; st.nMoshes = totalMoshes ; st.nHashes = H ; st.nBins = D ; st.nBlocks = nBlk ;
  st.algorithmicBytes = 120ull * nRec + 12ull * H + 12ull * D + ((P.flags & H10X_FLAG_NO_TABLE) ? 0 : (4ull << P.B))
    + 32ull * ((uint64_t) nBlk + 1) ;
  st.kernelLaunches = c->launches ;
  st.peakDeviceBytes = c->mt.peak ;
  c->haveIndex = true ;
}

/* ------------------------------------------------------------------ multi-GPU tail */

/* all-gather (nBlk, first word 0, last word 0, nRec, error) of every rank; the lowest rank's error, if
   any, is thrown on every rank so that nobody is left waiting in a collective */
static void dist_agree (h10x_ctx *c, cudaStream_t s, int localErr, const uint32_t mine[4], std::vector<uint32_t> &all)
{ DistState *d = c->dist ; const int NR = d->nranks ;
  DBuf<uint32_t> sb (5, s, &c->mt), rb ((size_t) 5 * NR, s, &c->mt) ;
  uint32_t h[5] = { mine[0], mine[1], mine[2], mine[3], (uint32_t) localErr } ;
  CK (cudaMemcpyAsync (sb.p, h, 20, cudaMemcpyHostToDevice, s)) ;
  NCK (gNccl.AllGather (sb.p, rb.p, 5, ncclUint32, d->comm, s)) ;
  std::vector<uint32_t> r ((size_t) 5 * NR) ;
  CK (cudaMemcpyAsync (r.data (), rb.p, 20 * (size_t) NR, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  all.resize ((size_t) 4 * NR) ;
  for (int i = 0 ; i < NR ; ++i)
    { for (int j = 0 ; j < 4 ; ++j) all[4*i + j] = r[5*i + j] ;
      if (r[5*i + 4])
	throw H10xError ((int) r[5*i + 4], std::string (h10x_strerror ((int) r[5*i + 4])) + " (rank " + std::to_string (i) + ")") ;
    }
}

static void dist_bins (h10x_ctx *c, cudaStream_t s, uint64_t H, const uint64_t *sh, const uint32_t *se,
		       const uint32_t *segIncl, uint32_t Dl, const uint32_t *segStart, const uint32_t *entryBlk,
		       uint32_t nBlkGlobal, uint32_t *entryId, uint32_t &Dglobal, uint64_t wMul,
		       const uint64_t *preHash, const uint32_t *preDepth, const uint32_t *preFirst)
{
  DistState *d = c->dist ; const int R = d->rank, NR = d->nranks ;
  MemTrack *mt = &c->mt ; const h10x_params &P = c->P ;
  StageTimer tm (c, s, ST_BINIDS) ;
  d->nLocalBins = Dl ;
  /* Error protocol.  A rank-local failure (in practice: the workspace slab is full) must not leave the peers waiting in a
     collective, so every group of data collectives below is preceded by an AGREEMENT (dist_agree: an all-gather of one
     error word) and everything a group allocates is allocated before its agreement.  A rank that throws between two
     agreements owes the peers its word at the next one (owesAgreement); the caller's handler delivers it
     (dist_fail_agree) and every rank leaves with that error.  Entry agreement: every rank got through its own mosh
     stage and grouping. */
  const uint32_t none[4] = { 0, 0, 0, 0 } ; std::vector<uint32_t> agreed ;
  int agreements = 0, failAt = -1 ;
  if (const char *e = getenv ("H10X_DIST_FAIL"))	/* tests: "rank:n" = this rank runs out of workspace just before its n-th agreement */
    { int fr = -1, fa = -1 ; if (sscanf (e, "%d:%d", &fr, &fa) == 2 && fr == R) failAt = fa ; }
  auto agree = [&] (bool last = false)
    { if (++agreements == failAt) throw SlabFull { 0 } ;
      d->owesAgreement = false ; dist_agree (c, s, 0, none, agreed) ; d->owesAgreement = !last ;
    } ;
  DBuf<uint64_t> dThr ((size_t) NR + 1, s, mt), dSendOff ((size_t) NR + 1, s, mt), dCnt (NR, s, mt), dMat ((size_t) NR * NR, s, mt) ;
  agree () ;
  HostTrace tr ;
  auto mark = [&] (const char *w) { if (tr.on) { cudaStreamSynchronize (s) ; tr.mark (w) ; } } ;

  /* 2. owner = hash range, monotone in the hash so the reference id order composes.  A mosh is min (hash, hashRC) of
	two roughly uniform values, so its density over [0, 2^2k) is 2 (1 - x), not flat: equal-width ranges would give
	owner 0 three quarters of the bins at 2 ranks (23 % instead of 12.5 % at 8).  The thresholds are the quantiles
	of that density, x_o = 1 - sqrt (1 - o / NR); any monotone choice is correct, this one balances the owners.
	H10X_FLAT_OWNERS=1 restores the equal-width cut (tests run both). */
  if (NR > H10X_MAX_RANKS) throw H10xError (H10X_ERR_UNSUPPORTED, "more ranks than H10X_MAX_RANKS") ;
  std::vector<uint64_t> thr ((size_t) NR + 1) ;
  h10x_dist_owner_thresholds (P.k, NR, getenv ("H10X_FLAT_OWNERS") ? 1 : 0, thr.data ()) ;
  CK (cudaMemcpyAsync (dThr.p, thr.data (), 8 * ((size_t) NR + 1), cudaMemcpyHostToDevice, s)) ;
  /* the rank-distinct (hash, local depth, local first block) triples either come ready from the hand-written tail
     (preHash / preDepth / preFirst, hash-ascending) or are read off the library-sorted entries (sh, se, segStart) */
  if (preHash) LAUNCH (c, k_lower_bounds, 1, 64, 0, s, preHash, Dl, dThr.p, (uint32_t) NR + 1, dSendOff.p) ;
  else LAUNCH (c, k_lower_bounds_seg, 1, 64, 0, s, segStart, sh, wMul, Dl, dThr.p, (uint32_t) NR + 1, dSendOff.p) ;
  std::vector<uint64_t> sendOff ((size_t) NR + 1) ;
  CK (cudaMemcpyAsync (sendOff.data (), dSendOff.p, 8 * ((size_t) NR + 1), cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  sendOff[0] = 0 ; sendOff[NR] = Dl ;

  mark ("d1-bounds") ;
  /* 3. counts */
  std::vector<uint64_t> sendCnt (NR), cntMat ((size_t) NR * NR) ;
  for (int o = 0 ; o < NR ; ++o) sendCnt[o] = sendOff[o+1] - sendOff[o] ;
  CK (cudaMemcpyAsync (dCnt.p, sendCnt.data (), 8 * (size_t) NR, cudaMemcpyHostToDevice, s)) ;
  NCK (gNccl.AllGather (dCnt.p, dMat.p, NR, ncclUint64, d->comm, s)) ;
  CK (cudaMemcpyAsync (cntMat.data (), dMat.p, 8 * (size_t) NR * NR, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  std::vector<uint64_t> recvCnt (NR), recvOff ((size_t) NR + 1, 0) ;
  for (int r = 0 ; r < NR ; ++r) { recvCnt[r] = cntMat[(size_t) r * NR + R] ; recvOff[r+1] = recvOff[r] + recvCnt[r] ; }
  const uint64_t Ro64 = recvOff[NR] ;
  for (int o = 0 ; o < NR ; ++o)	/* the count matrix is the same everywhere: every rank throws, or none */
    { uint64_t tot = 0 ; for (int r = 0 ; r < NR ; ++r) tot += cntMat[(size_t) r * NR + o] ;
      if (tot >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 rank-distinct hashes for one owner") ;
    }
  const uint32_t Ro = (uint32_t) Ro64 ;

  mark ("d3-counts") ;
  /* 4. all-to-all-v of (hash, depth, first block) to the hash-range owners.  Preferred: one kernel that
	computes the rank-distinct values and stores them into the owners' receive arrays over NVLink (peer
	memory through CUDA IPC); otherwise k_local_distinct + ncclSend/ncclRecv. */
  DBuf<uint64_t> rHash (Ro, s, mt) ; DBuf<uint32_t> rDepth (Ro, s, mt), rFirst (Ro, s, mt) ;
  c->localBinId.alloc (Dl, s, mt) ;
  DBuf<unsigned char> dInfo (sizeof (PeerInfo), s, mt), dInfos (sizeof (PeerInfo) * (size_t) NR, s, mt) ;
  DBuf<uint32_t> dOk (1, s, mt), dOks (NR, s, mt) ;
  DBuf<uint64_t> dHash ; DBuf<uint32_t> dDepth, dFirst ;	/* library-sort tail: the triples are materialised here first */
  if (!preHash) { dHash.alloc (Dl, s, mt) ; dDepth.alloc (Dl, s, mt) ; dFirst.alloc (Dl, s, mt) ; }
  agree () ;		/* 1: the receive arrays exist everywhere */
  bool pushed = false ;
  std::vector<PeerInfo> infos (NR) ;
  /* peer copies run on side streams (copy engines) and are joined back into s */
  auto peerCopy = [&] (int peer, void *dst, const void *src, size_t bytes)
    { if (!bytes) return ;
      if (d->copyStreams.empty ())
	for (int i = 0 ; i < NR ; ++i) { cudaStream_t cs ; CK (cudaStreamCreateWithFlags (&cs, cudaStreamNonBlocking)) ; d->copyStreams.push_back (cs) ; }
      cudaStream_t cs = d->copyStreams[peer] ;
      cudaEvent_t ready = ctx_event (c), done = ctx_event (c) ;
      CK (cudaEventRecord (ready, s)) ; CK (cudaStreamWaitEvent (cs, ready, 0)) ;
      CK (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToDevice, cs)) ;
      CK (cudaEventRecord (done, cs)) ; CK (cudaStreamWaitEvent (s, done, 0)) ;
    } ;
  if (d->pushState >= 0 && !getenv ("H10X_NO_PEER_PUSH"))
    { PeerInfo mine ; memset (&mine, 0, sizeof (mine)) ;
      mine.ok = (cudaIpcGetMemHandle (&mine.handle, c->mt.base) == cudaSuccess) ? 1 : 0 ;
      if (!mine.ok) cudaGetLastError () ;
      mine.pid = (uint64_t) getpid () ; mine.base = (uint64_t) c->mt.base ; mine.device = P.device ;
      mine.offHash = (uint64_t) ((char*) rHash.p - c->mt.base) ; mine.offDepth = (uint64_t) ((char*) rDepth.p - c->mt.base) ;
      mine.offFirst = (uint64_t) ((char*) rFirst.p - c->mt.base) ;
      mine.offBinId = (uint64_t) ((char*) c->localBinId.p - c->mt.base) ;
      CK (cudaMemcpyAsync (dInfo.p, &mine, sizeof (PeerInfo), cudaMemcpyHostToDevice, s)) ;
      NCK (gNccl.AllGather (dInfo.p, dInfos.p, sizeof (PeerInfo), ncclUint8, d->comm, s)) ;
      CK (cudaMemcpyAsync (infos.data (), dInfos.p, sizeof (PeerInfo) * (size_t) NR, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      /* map every peer's slab (cached while its handle stays the same) */
      uint32_t ok = 1 ;
      for (int r = 0 ; r < NR && ok ; ++r)
	{ PeerMap &pm = d->peers[r] ;
	  if (!infos[r].ok) { ok = 0 ; break ; }
	  if (r == R) { pm.mapped = c->mt.base ; pm.viaIpc = false ; continue ; }
	  bool same = pm.mapped && pm.pid == infos[r].pid && pm.base == infos[r].base
	    && !memcmp (&pm.handle, &infos[r].handle, sizeof (cudaIpcMemHandle_t)) ;
	  if (same) continue ;
	  if (pm.mapped && pm.viaIpc) cudaIpcCloseMemHandle (pm.mapped) ;
	  pm.mapped = nullptr ; pm.viaIpc = false ;
	  if (infos[r].pid == mine.pid)		/* ranks are threads of one process: plain peer access */
	    { int can = 0 ;
	      cudaDeviceCanAccessPeer (&can, P.device, infos[r].device) ;
	      cudaError_t e = can ? cudaDeviceEnablePeerAccess (infos[r].device, 0) : cudaErrorInvalidDevice ;
	      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError () ; e = cudaSuccess ; }
	      if (e != cudaSuccess) { cudaGetLastError () ; ok = 0 ; break ; }
	      pm.mapped = (char*) infos[r].base ;
	    }
	  else
	    { void *ptr = nullptr ;
	      cudaError_t e = cudaIpcOpenMemHandle (&ptr, infos[r].handle, cudaIpcMemLazyEnablePeerAccess) ;
	      if (e != cudaSuccess)
		{ if (tr.on) fprintf (stderr, "h10x-trace rank %d: cudaIpcOpenMemHandle(rank %d) failed: %s\n", R, r, cudaGetErrorString (e)) ;
		  cudaGetLastError () ; ok = 0 ; break ;
		}
	      pm.mapped = (char*) ptr ; pm.viaIpc = true ;
	    }
	  pm.handle = infos[r].handle ; pm.pid = infos[r].pid ; pm.base = infos[r].base ;
	}
      /* every rank must take the same path */
      std::vector<uint32_t> oks (NR) ;
      CK (cudaMemcpyAsync (dOk.p, &ok, 4, cudaMemcpyHostToDevice, s)) ;
      NCK (gNccl.AllGather (dOk.p, dOks.p, 1, ncclUint32, d->comm, s)) ;
      CK (cudaMemcpyAsync (oks.data (), dOks.p, 4 * (size_t) NR, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      bool all = true ; for (int r = 0 ; r < NR ; ++r) all = all && oks[r] ;
      d->pushState = all ? 1 : -1 ;
      if (tr.on && R == 0) fprintf (stderr, "h10x-trace peer push %s\n", all ? "active" : "unavailable, using ncclSend/ncclRecv") ;
      if (all)
	{ PushArgs pa ; memset (&pa, 0, sizeof (pa)) ;
	  pa.nranks = NR ;
	  for (int o = 0 ; o < NR ; ++o)
	    { char *pb = d->peers[o].mapped ;
	      pa.hash[o] = (uint64_t*) (pb + infos[o].offHash) ; pa.depth[o] = (uint32_t*) (pb + infos[o].offDepth) ;
	      pa.first[o] = (uint32_t*) (pb + infos[o].offFirst) ;
	      pa.sendOff[o] = sendOff[o] ;
	      uint64_t before = 0 ; for (int src = 0 ; src < R ; ++src) before += cntMat[(size_t) src * NR + o] ;
	      pa.dstOff[o] = before ;
	    }
	  pa.sendOff[NR] = Dl ;
	  /* two ways to move the data: SM stores from the producing kernel (k_push_distinct), or local
	     materialisation + copy engines.  Measured on 8 B200s (1.3 GB leaving each rank): SM stores 14.4 ms
	     (~110 GB/s per GPU, no better than ncclSend/ncclRecv), copy engines 6.3 ms; at 2 ranks the kernel
	     wins (3.3 ms).  H10X_PEER_KERNEL / H10X_PEER_COPY force one. */
	  const bool useCopy = getenv ("H10X_PEER_COPY") || (NR > 2 && !getenv ("H10X_PEER_KERNEL")) ;
	  if (useCopy)
	    { const uint64_t *xh = preHash ; const uint32_t *xd = preDepth, *xf = preFirst ;
	      if (!preHash)
		{ if (Dl) LAUNCH (c, k_local_distinct, gridFor (Dl, 256), 256, 0, s, Dl, segStart, sh, se, entryBlk, wMul, dHash.p, dDepth.p, dFirst.p) ;
		  xh = dHash.p ; xd = dDepth.p ; xf = dFirst.p ;
		}
	      for (int k = 1 ; k <= NR ; ++k)
		{ int o = (R + k) % NR ;		/* start with the next rank: spread the load over the links */
		  uint64_t n = sendCnt[o] ;
		  peerCopy (o, pa.hash[o] + pa.dstOff[o], xh + sendOff[o], 8 * n) ;
		  peerCopy (o, pa.depth[o] + pa.dstOff[o], xd + sendOff[o], 4 * n) ;
		  peerCopy (o, pa.first[o] + pa.dstOff[o], xf + sendOff[o], 4 * n) ;
		}
	      CK (cudaStreamSynchronize (s)) ;	/* the staging arrays go out of scope */
	    }
	  else if (Dl && preHash) LAUNCH (c, k_push_triples, gridFor (Dl, 256), 256, 0, s, Dl, preHash, preDepth, preFirst, pa) ;
	  else if (Dl) LAUNCH (c, k_push_distinct, gridFor (Dl, 256), 256, 0, s, Dl, segStart, sh, se, entryBlk, wMul, pa) ;
	  /* nobody reads its receive arrays before every rank's stores have landed: the collective is
	     enqueued behind the kernel on each rank's stream */
	  NCK (gNccl.AllGather (dOk.p, dOks.p, 1, ncclUint32, d->comm, s)) ;
	  pushed = true ;
	}
    }
  if (!pushed)
    { const uint64_t *xh = preHash ; const uint32_t *xd = preDepth, *xf = preFirst ;
      if (!preHash)
	{ if (Dl) LAUNCH (c, k_local_distinct, gridFor (Dl, 256), 256, 0, s, Dl, segStart, sh, se, entryBlk, wMul, dHash.p, dDepth.p, dFirst.p) ;
	  xh = dHash.p ; xd = dDepth.p ; xf = dFirst.p ;
	}
      NCK (gNccl.GroupStart ()) ;
      for (int peer = 0 ; peer < NR ; ++peer)
	{ if (sendCnt[peer])
	    { NCK (gNccl.Send (xh + sendOff[peer], sendCnt[peer], ncclUint64, peer, d->comm, s)) ;
	      NCK (gNccl.Send (xd + sendOff[peer], sendCnt[peer], ncclUint32, peer, d->comm, s)) ;
	      NCK (gNccl.Send (xf + sendOff[peer], sendCnt[peer], ncclUint32, peer, d->comm, s)) ;
	    }
	  if (recvCnt[peer])
	    { NCK (gNccl.Recv (rHash.p + recvOff[peer], recvCnt[peer], ncclUint64, peer, d->comm, s)) ;
	      NCK (gNccl.Recv (rDepth.p + recvOff[peer], recvCnt[peer], ncclUint32, peer, d->comm, s)) ;
	      NCK (gNccl.Recv (rFirst.p + recvOff[peer], recvCnt[peer], ncclUint32, peer, d->comm, s)) ;
	    }
	}
      NCK (gNccl.GroupEnd ()) ;
      CK (cudaStreamSynchronize (s)) ;	/* the send buffers are freed on leaving this scope */
    }
  dHash.release () ; dDepth.release () ; dFirst.release () ;
  mark ("d4-alltoall") ;

  /* 5. owner merge: depth = sum, first block = min over the (at most NR) copies of a hash */
  const uint32_t nB2 = nBlkGlobal + 2 ;
  uint32_t Do = 0 ;
  DBuf<uint64_t> gHash ; DBuf<uint32_t> gDepth, gFirst, oi, oSegIncl, segOf ;
  DBuf<uint32_t> newCnt (nB2, s, mt) ;
  CK (cudaMemsetAsync (newCnt.p, 0, 4 * (size_t) nB2, s)) ;
  bool merged = false ;
  if (Ro && !getenv ("H10X_OWNER_SORT"))
    { /* the received runs are sorted: merge them tile by tile (h10x_dist.cuh, "owner merge without a sort") */
      uint32_t S = std::max<uint32_t> (16u, H10X_MERGE_CAP / (2u * (uint32_t) NR)) ;
      if (const char *e = getenv ("H10X_MERGE_S")) { long v = atol (e) ; if (v >= 16 && v <= (long) H10X_MERGE_CAP) S = (uint32_t) v ; }
      std::vector<uint32_t> candOff ((size_t) NR + 1, 0) ;
      for (int r = 0 ; r < NR ; ++r) candOff[r + 1] = candOff[r] + (uint32_t) (recvCnt[r] ? (recvCnt[r] - 1) / S : 0) ;
      const uint32_t nCand = candOff[NR], nTiles = nCand / (uint32_t) NR + 1 ;
      DBuf<uint32_t> dCandOff ((size_t) NR + 1, s, mt), bnd (((size_t) nTiles + 1) * NR, s, mt) ;
      DBuf<unsigned long long> tileState (nTiles, s, mt) ;
      DBuf<uint64_t> cand (nCand, s, mt), candS (nCand, s, mt) ; DBuf<unsigned int> ovf (2, s, mt) ;
      /* the number of bins is only known afterwards: sized for the worst case (every received copy its own bin) */
      gHash.alloc (Ro, s, mt) ; gDepth.alloc (Ro, s, mt) ; gFirst.alloc (Ro, s, mt) ; segOf.alloc (Ro, s, mt) ;
      MergeArgs ma ; memset (&ma, 0, sizeof (ma)) ;
      ma.rHash = rHash.p ; ma.rDepth = rDepth.p ; ma.rFirst = rFirst.p ;
      for (int r = 0 ; r <= NR ; ++r) ma.recvOff[r] = recvOff[r] ;
      ma.bnd = bnd.p ; ma.tileState = tileState.p ; ma.newCnt = newCnt.p ; ma.overflow = ovf.p ; ma.ticket = ovf.p + 1 ;
      ma.nTiles = nTiles ; ma.nranks = NR ;
      ma.gHash = gHash.p ; ma.gDepth = gDepth.p ; ma.gFirst = gFirst.p ; ma.segOf = segOf.p ;
      CK (cudaMemcpyAsync (dCandOff.p, candOff.data (), 4 * ((size_t) NR + 1), cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemsetAsync (ovf.p, 0, 8, s)) ;
      CK (cudaMemsetAsync (tileState.p, 0, 8 * (size_t) nTiles, s)) ;
      if (nCand)
	{ const uint32_t gx = (uint32_t) std::min<uint64_t> (gridFor (nCand / NR + 1, 256), 1024) ;
	  k_merge_candidates<<<dim3 (gx, (unsigned) NR), 256, 0, s>>> (ma, S, dCandOff.p, cand.p) ; ++c->launches ;
	  cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceRadixSort::SortKeys (t, b, cand.p, candS.p, nCand, 0, 2 * P.k, s) ; }) ;
	}
      LAUNCH (c, k_merge_bounds, gridFor (((uint64_t) nTiles + 1) * NR, 256), 256, 0, s, ma, candS.p, nCand, bnd.p) ;
      const int nSM = device_sms (c) ;
      CK (cudaFuncSetAttribute (k_owner_merge_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h10x_merge_smem ())) ;
      int occ = 1 ;
      CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, k_owner_merge_tiles, H10X_MERGE_THREADS, h10x_merge_smem ())) ;
      /* every CTA of the grid must be resident: a waiting tile's predecessor is then always running or done */
      LAUNCH (c, k_owner_merge_tiles, std::min<uint32_t> (nTiles, (uint32_t) (nSM * std::max (occ, 1))), H10X_MERGE_THREADS, h10x_merge_smem (), s, ma) ;
      unsigned int over = 0 ; unsigned long long last = 0 ;
      CK (cudaMemcpyAsync (&last, tileState.p + (nTiles - 1), 8, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaMemcpyAsync (&over, ovf.p, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      if (!over) { Do = (uint32_t) (last & 0x3fffffffffffffffull) ; merged = true ; }
      else
	{ Do = 0 ; gHash.release () ; gDepth.release () ; gFirst.release () ; segOf.release () ;
	  CK (cudaMemsetAsync (newCnt.p, 0, 4 * (size_t) nB2, s)) ;	/* the tiles before the overflow have counted already */
	}
    }
  if (Ro && !merged)
    { oi.alloc (Ro, s, mt) ; oSegIncl.alloc (Ro, s, mt) ;
      DBuf<uint64_t> oh (Ro, s, mt) ;
      DBuf<uint32_t> iota (Ro, s, mt), head (Ro, s, mt), oSegStart ;
      LAUNCH (c, k_iota, gridFor (Ro, 256), 256, 0, s, iota.p, (uint64_t) Ro) ;
      cubCall (c, s, [&] (void *t, size_t &b)
	{ return cub::DeviceRadixSort::SortPairs (t, b, rHash.p, oh.p, iota.p, oi.p, Ro, 0, 2 * P.k, s) ; }) ;
      LAUNCH (c, k_head_flag, gridFor (Ro, 256), 256, 0, s, oh.p, (uint64_t) Ro, head.p) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveSum (t, b, head.p, oSegIncl.p, Ro, s) ; }) ;
      CK (cudaMemcpyAsync (&Do, oSegIncl.p + (Ro - 1), 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      oSegStart.alloc ((size_t) Do + 1, s, mt) ;
      LAUNCH (c, k_seg_start, gridFor (Ro, 256), 256, 0, s, head.p, oSegIncl.p, (uint64_t) Ro, oi.p, oSegStart.p, (uint32_t*) nullptr) ;
      gHash.alloc (Do, s, mt) ; gDepth.alloc (Do, s, mt) ; gFirst.alloc (Do, s, mt) ;
      LAUNCH (c, k_owner_merge, gridFor (Do, 256), 256, 0, s, Do, oSegStart.p, oh.p, oi.p, rDepth.p, rFirst.p,
	      gHash.p, gDepth.p, gFirst.p, newCnt.p) ;
    }
  rHash.release () ; rDepth.release () ; rFirst.release () ;

  mark ("d5-merge") ;
  /* 6. global id bases from every owner's per-block new-hash counts */
  DBuf<uint32_t> newMat ((size_t) NR * nB2, s, mt), colSum (nB2, s, mt), below (nB2, s, mt), prefixAll (nB2, s, mt) ;
  agree () ;		/* 2: every owner has merged what it received */
  NCK (gNccl.AllGather (newCnt.p, newMat.p, nB2, ncclUint32, d->comm, s)) ;
  LAUNCH (c, k_id_base, gridFor (nB2, 256), 256, 0, s, nB2, (uint32_t) NR, (uint32_t) R, newMat.p, colSum.p, below.p) ;
  cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, colSum.p, prefixAll.p, nB2, s) ; }) ;
  uint32_t tail[2] = { 0, 0 } ;
  CK (cudaMemcpyAsync (&tail[0], prefixAll.p + (nB2 - 1), 4, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaMemcpyAsync (&tail[1], colSum.p + (nB2 - 1), 4, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  Dglobal = tail[0] + tail[1] ;
  if ((uint64_t) Dglobal + 1 > (((uint64_t) 1 << P.B) >> 2) - 2)	/* hash10x.c:149, the same on every rank */
    throw H10xError (H10X_ERR_TABLE_TOO_SMALL, "hashTableSize is too small") ;
  c->hashNumber = Dglobal + 1 ;

  mark ("d6-idbase") ;
  /* 7. ids of this owner's hashes: order by (first block, hash) = stable sort by first block of the
	hash-sorted list; then the id of every received copy */
  DBuf<uint32_t> gId (Do, s, mt), ans (Ro, s, mt), sId (Do, s, mt), sDepth (Do, s, mt) ;
  DBuf<uint64_t> sHash (Do, s, mt) ;
  if (Do)
    { DBuf<uint32_t> iota (Do, s, mt), sf (Do, s, mt), sg (Do, s, mt), headPos (Do, s, mt), groupStart (Do, s, mt) ;
      int bits = 1 ; while (((uint64_t) 1 << bits) < nB2) ++bits ;
      LAUNCH (c, k_iota, gridFor (Do, 256), 256, 0, s, iota.p, (uint64_t) Do) ;
      cubCall (c, s, [&] (void *t, size_t &b)
	{ return cub::DeviceRadixSort::SortPairs (t, b, gFirst.p, sf.p, iota.p, sg.p, Do, 0, bits, s) ; }) ;
      LAUNCH (c, k_group_head, gridFor (Do, 256), 256, 0, s, sf.p, Do, headPos.p) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::InclusiveScan (t, b, headPos.p, groupStart.p, MaxOp (), Do, s) ; }) ;
      LAUNCH (c, k_owner_ids, gridFor (Do, 256), 256, 0, s, Do, sf.p, sg.p, groupStart.p, prefixAll.p, below.p, gId.p,
	      gHash.p, gDepth.p, sId.p, sHash.p, sDepth.p) ;
      if (merged) LAUNCH (c, k_answer_ids_seg, gridFor (Ro, 256), 256, 0, s, Ro, segOf.p, gId.p, ans.p) ;
      else LAUNCH (c, k_answer_ids, gridFor (Ro, 256), 256, 0, s, Ro, oSegIncl.p, oi.p, gId.p, ans.p) ;
    }

  DBuf<uint32_t> tId, tDepth ; DBuf<uint64_t> tHash ;
  if (R == 0)
    { tId.alloc (Dglobal, s, mt) ; tDepth.alloc (Dglobal, s, mt) ; tHash.alloc (Dglobal, s, mt) ;
      c->hashValue.alloc ((size_t) Dglobal + 1, s, mt) ; c->hashDepth.alloc ((size_t) Dglobal + 2, s, mt) ;
    }
  DBuf<uint64_t> d2 (5, s, mt), dAll2 ((size_t) 5 * NR, s, mt) ;
  DBuf<uint32_t> b1 (1, s, mt), bN (NR, s, mt) ;
  agree (true) ;	/* 3, the last: every owner has its ids; nothing below allocates */
  mark ("d7-ids") ;
  /* 8. reverse all-to-all-v: the bin id of every rank-distinct hash, in the order it was sent */
  if (pushed)
    { for (int k = 1 ; k <= NR ; ++k)
	{ int src = (R + k) % NR ;
	  uint64_t before = 0 ; for (int o = 0 ; o < R ; ++o) before += cntMat[(size_t) src * NR + o] ;	/* src's sendOff[R] */
	  uint32_t *dst = (uint32_t*) (d->peers[src].mapped + infos[src].offBinId) + before ;
	  peerCopy (src, dst, ans.p + recvOff[src], 4 * recvCnt[src]) ;
	}
    }
  else
    { NCK (gNccl.GroupStart ()) ;
      for (int peer = 0 ; peer < NR ; ++peer)
	{ if (recvCnt[peer]) NCK (gNccl.Send (ans.p + recvOff[peer], recvCnt[peer], ncclUint32, peer, d->comm, s)) ;
	  if (sendCnt[peer]) NCK (gNccl.Recv (c->localBinId.p + sendOff[peer], sendCnt[peer], ncclUint32, peer, d->comm, s)) ;
	}
      NCK (gNccl.GroupEnd ()) ;
    }

  mark ("d8-reverse") ;
  /* 9. (id, hash, depth) of every bin to rank 0, which owns hashValue / hashDepth / hashIndex */
  if (R == 0)
    { CK (cudaMemsetAsync (c->hashValue.p, 0, 8, s)) ;
      CK (cudaMemsetAsync (c->hashDepth.p, 0, 4, s)) ;
      CK (cudaMemsetAsync (c->hashDepth.p + Dglobal + 1, 0, 4, s)) ;
    }
  std::vector<uint64_t> mine2 = { Do, H, 0, 0, 0 }, all2 ((size_t) 5 * NR) ;
  if (R == 0)
    { mine2[2] = (uint64_t) ((char*) tId.p - c->mt.base) ; mine2[3] = (uint64_t) ((char*) tHash.p - c->mt.base) ;
      mine2[4] = (uint64_t) ((char*) tDepth.p - c->mt.base) ;
    }
  CK (cudaMemcpyAsync (d2.p, mine2.data (), 40, cudaMemcpyHostToDevice, s)) ;
  /* this collective is also the barrier behind the reverse copies: after it every rank's localBinId is complete */
  NCK (gNccl.AllGather (d2.p, dAll2.p, 5, ncclUint64, d->comm, s)) ;
  CK (cudaMemcpyAsync (all2.data (), dAll2.p, 40 * (size_t) NR, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  mark ("d9a-sync") ;
  if (H && entryId) LAUNCH (c, k_entry_ids, gridFor (H, 256), 256, 0, s, H, segIncl, c->localBinId.p, se, entryId) ;
  mark ("d9b-entryids") ;
  d->nHashesGlobal = 0 ;
  for (int r = 0 ; r < NR ; ++r) d->nHashesGlobal += all2[5*r + 1] ;
  if (pushed)
    { uint64_t before = 0 ; for (int r = 0 ; r < R ; ++r) before += all2[5*r] ;
      char *z = d->peers[0].mapped ;
      peerCopy (0, (uint32_t*) (z + all2[2]) + before, sId.p, 4 * (size_t) Do) ;
      peerCopy (0, (uint64_t*) (z + all2[3]) + before, sHash.p, 8 * (size_t) Do) ;
      peerCopy (0, (uint32_t*) (z + all2[4]) + before, sDepth.p, 4 * (size_t) Do) ;
      CK (cudaMemsetAsync (b1.p, 0, 4, s)) ;
      NCK (gNccl.AllGather (b1.p, bN.p, 1, ncclUint32, d->comm, s)) ;	/* barrier: rank 0 scatters only complete data */
      CK (cudaStreamSynchronize (s)) ;
    }
  else
    { NCK (gNccl.GroupStart ()) ;
      if (Do)
	{ NCK (gNccl.Send (sId.p, Do, ncclUint32, 0, d->comm, s)) ;
	  NCK (gNccl.Send (sHash.p, Do, ncclUint64, 0, d->comm, s)) ;
	  NCK (gNccl.Send (sDepth.p, Do, ncclUint32, 0, d->comm, s)) ;
	}
      if (R == 0)
	{ uint64_t off = 0 ;
	  for (int r = 0 ; r < NR ; ++r)
	    { uint64_t n = all2[5*r] ;
	      if (n)
		{ NCK (gNccl.Recv (tId.p + off, n, ncclUint32, r, d->comm, s)) ;
		  NCK (gNccl.Recv (tHash.p + off, n, ncclUint64, r, d->comm, s)) ;
		  NCK (gNccl.Recv (tDepth.p + off, n, ncclUint32, r, d->comm, s)) ;
		}
	      off += n ;
	    }
	}
      NCK (gNccl.GroupEnd ()) ;
    }
  mark ("d9c-gather") ;
  if (R == 0 && Dglobal)
    LAUNCH (c, k_scatter_bins, gridFor (Dglobal, 256), 256, 0, s, Dglobal, tId.p, tHash.p, tDepth.p, c->hashValue.p, c->hashDepth.p) ;
  CK (cudaStreamSynchronize (s)) ;	/* sends read gId/gHash/gDepth, freed on return */
  mark ("d9-gather0") ;
}

/* ------------------------------------------------------------------ slab sizing / retry */

static void slab_free (h10x_ctx *c)
{ if (c->dlStream) cudaStreamSynchronize (c->dlStream) ;	/* early downloads read the slab */
  reset_result (c) ;
  if (c->mt.base) { cudaFree (c->mt.base) ; c->mt.base = nullptr ; }
  c->mt.cap = 0 ; c->mt.reset () ;
}

static void slab_resize (h10x_ctx *c, size_t want)
{ slab_free (c) ;
  size_t freeB = 0, totalB = 0 ;
  CK (cudaMemGetInfo (&freeB, &totalB)) ;
  size_t limit = freeB - std::min<size_t> (freeB / 50, (size_t) 512 << 20) ;	/* leave a little headroom */
  want = (want + 511) & ~(size_t) 511 ;
  if (want > limit) want = limit & ~(size_t) 511 ;
  c->slabClamped = (want >= (limit & ~(size_t) 511)) ;
  CK (cudaMalloc ((void**) &c->mt.base, want)) ;
  c->mt.cap = want ; c->mt.reset () ;
}

/* run `body` (a complete build) inside the slab, growing the slab and re-running when it is too small */
template <class F> static void with_slab (h10x_ctx *c, cudaStream_t s, size_t estimate, F body)
{ if (c->mt.cap < estimate && !c->slabClamped) { reset_result (c) ; slab_resize (c, estimate) ; }
  for (int attempt = 0 ; ; ++attempt)
    { try { body () ; return ; }
      catch (const SlabFull &f)
	{ cudaStreamSynchronize (s) ;
	  size_t old = c->mt.cap ;
	  size_t want = std::max<size_t> (f.need + f.need / 4, old + old / 2) ;
	  slab_resize (c, want) ;
	  if (c->mt.cap <= old || attempt > 12)
	    throw H10xError (H10X_ERR_NOMEM, "device workspace: need " + std::to_string (f.need) + " bytes, have " + std::to_string (c->mt.cap)) ;
	}
    }
}

static size_t slab_estimate (const h10x_params &P, uint64_t nRec, bool withFqb)
{ /* never more than the device can give: slab_resize clamps to free memory, and a clamped slab must
     not look "too small" again on the next call */
  return (size_t) 420 * nRec + ((P.flags & H10X_FLAG_NO_TABLE) ? 0 : ((size_t) 4 << P.B)) + (withFqb ? 120 * nRec : 0)
    + ((size_t) 64 << 20) ;
}

/* ------------------------------------------------------------------ C ABI */

extern "C" { static void *pinned_alloc (size_t bytes) ; }
static void set_err (char *err, size_t errlen, const char *msg)
{ if (err && errlen) { strncpy (err, msg, errlen - 1) ; err[errlen - 1] = 0 ; } }

template <class F> static int guarded (char *err, size_t errlen, F f)
{ try { f () ; set_err (err, errlen, "") ; return H10X_OK ; }
  catch (const H10xError &e) { set_err (err, errlen, e.what ()) ; return e.code ; }
  catch (const SlabFull &f) { set_err (err, errlen, ("device workspace too small: need " + std::to_string (f.need) + " bytes").c_str ()) ; return H10X_ERR_NOMEM ; }
  catch (const std::bad_alloc &) { set_err (err, errlen, "host out of memory") ; return H10X_ERR_NOMEM ; }
  catch (const std::exception &e) { set_err (err, errlen, e.what ()) ; return H10X_ERR_CUDA ; }
}

extern "C" {

int h10x_abi_version (void) { return H10X_ABI_VERSION ; }

int h10x_gpu_device_count (void)
{ int n = 0 ; if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError () ; return 0 ; } return n ; }

const char *h10x_strerror (int code)
{ switch (code)
    { case H10X_OK: return "ok" ;
    case H10X_ERR_TABLE_TOO_SMALL: return "hashTableSize is too small" ;	/* hash10x.c:149 */
    case H10X_ERR_CHUNK_TOO_SMALL: return "chunkSize too small" ;		/* hash10x.c:206 */
    case H10X_ERR_BAD_PARAM: return "bad parameter" ;
    case H10X_ERR_NOMEM: return "out of memory" ;
    case H10X_ERR_IO: return "file read problem" ;				/* hash10x.c:209 */
    case H10X_ERR_CUDA: return "CUDA error" ;
    case H10X_ERR_NO_DEVICE: return "no CUDA device: libh10xgpu has no CPU fallback" ;
    case H10X_ERR_UNSUPPORTED: return "size not supported on one device" ;
    }
  return "unknown error" ;
}

const char *h10x_stage_name (int stage) { return (stage >= 0 && stage < H10X_NSTAGES) ? kStageNames[stage] : "" ; }

uint64_t h10x_factor1_from_seed (int seed)
{ srandom ((unsigned) seed) ;
  uint64_t hi = (uint64_t) random () ; uint64_t lo = (uint64_t) random () ;
  return (hi << 32) | lo | 1 ;
}

h10x_ctx *h10x_gpu_create (const h10x_params *p, char *err, size_t errlen)
{
  h10x_ctx *c = nullptr ;
  int st = guarded (err, errlen, [&] ()
    { if (!p) throw H10xError (H10X_ERR_BAD_PARAM, "null params") ;
      if (p->k < 1 || p->k >= 32) throw H10xError (H10X_ERR_BAD_PARAM, "seqhash k " + std::to_string (p->k) + " must be between 1 and 32") ;
      if (p->w < 1) throw H10xError (H10X_ERR_BAD_PARAM, "seqhash w " + std::to_string (p->w) + " must be positive") ;
      int maxB = (p->flags & H10X_FLAG_WIDE_B) ? 34 : 30 ;
      if (p->B < 20 || p->B > maxB)
	throw H10xError (H10X_ERR_BAD_PARAM, "hashTableBits " + std::to_string (p->B) + " out of range 20-" + std::to_string (maxB)) ;
      if (p->chunkSize < 1) throw H10xError (H10X_ERR_CHUNK_TOO_SMALL, "chunkSize too small") ;
      if (p->N < 0) throw H10xError (H10X_ERR_BAD_PARAM, "negative N") ;
      int n = h10x_gpu_device_count () ;
      if (n <= 0 || p->device < 0 || p->device >= n) throw H10xError (H10X_ERR_NO_DEVICE, h10x_strerror (H10X_ERR_NO_DEVICE)) ;
      CK (cudaSetDevice (p->device)) ;
      c = new h10x_ctx () ;
      c->P = *p ;
      make_hash_params (*p, c->hp) ;
      memset (&c->stats, 0, sizeof (c->stats)) ;
      CK (cudaStreamCreateWithFlags (&c->own, cudaStreamNonBlocking)) ;
    }) ;
  if (st != H10X_OK) { delete c ; return nullptr ; }
  return c ;
}

void h10x_gpu_destroy (h10x_ctx *c)
{ if (!c) return ;
  cudaSetDevice (c->P.device) ;
  if (c->own) cudaStreamSynchronize (c->own) ;
    if (c->dlStream) { cudaStreamSynchronize (c->dlStream) ; cudaStreamDestroy (c->dlStream) ; c->dlStream = 0 ; }
  if (c->ulStream) { cudaStreamSynchronize (c->ulStream) ; cudaStreamDestroy (c->ulStream) ; c->ulStream = 0 ; }
  if (c->fqRecs) cudaFree (c->fqRecs) ;
  if (c->wlTable) cudaFree (c->wlTable) ;
  if (c->fqHost) cudaFreeHost (c->fqHost) ;
  if (c->dist)
    { for (int r = 0 ; r < H10X_MAX_RANKS ; ++r)
	if (c->dist->peers[r].mapped && c->dist->peers[r].viaIpc) cudaIpcCloseMemHandle (c->dist->peers[r].mapped) ;
      for (auto cs : c->dist->copyStreams) cudaStreamDestroy (cs) ;
      if (c->dist->comm) gNccl.CommDestroy (c->dist->comm) ;
      delete c->dist ; c->dist = nullptr ;
    }
  slab_free (c) ;
  for (auto e : c->evPool) cudaEventDestroy (e) ;
  for (int i = 0 ; i < 9 ; ++i) if (c->hostSlot[i]) cudaFreeHost (c->hostSlot[i]) ;
  for (int i = 0 ; i < 3 ; ++i) if (c->goodSlot[i]) cudaFreeHost (c->goodSlot[i]) ;
  for (int i = 0 ; i < 3 ; ++i) if (c->clusSlot[i]) cudaFreeHost (c->clusSlot[i]) ;
  if (c->own) { cudaStreamSynchronize (c->own) ; cudaStreamDestroy (c->own) ; }
  delete c ;
}

int h10x_gpu_build_device (h10x_ctx *c, const void *d_fqb, uint64_t nRecords, void *stream, char *err, size_t errlen)
{ if (!c) { set_err (err, errlen, "null context") ; return H10X_ERR_BAD_PARAM ; }
  HostTrace apiTrace ;
  struct Done { HostTrace &t ; ~Done () { t.mark ("api-total") ; } } done { apiTrace } ;
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = stream ? (cudaStream_t) stream : c->own ;
      with_slab (c, s, slab_estimate (c->P, nRecords, false),
		 [&] () { build_device_impl (c, (const uint32_t*) d_fqb, nRecords, s) ; }) ;
    }) ;
  if (st != H10X_OK) { cudaStreamSynchronize (stream ? (cudaStream_t) stream : c->own) ; cudaGetLastError () ; c->haveIndex = false ; }
  return st ;
}

int h10x_gpu_index_device (h10x_ctx *c, h10x_index *out)
{ if (!c || !out || !c->haveIndex) return H10X_ERR_BAD_PARAM ;
  memset (out, 0, sizeof (*out)) ;
  out->B = c->P.B ; out->hashNumber = c->hashNumber ; out->nBlocksMax = c->nBlocksMax ;
  out->nReads = c->nReads ; out->nHashes = c->nHashes ;
  out->hashIndex = c->hashIndex.p ; out->hashValue = c->hashValue.p ; out->hashDepth = c->hashDepth.p ;
  out->blkNRead = c->blkNRead.p ; out->blkNHash = c->blkNHash.p ; out->blkOff = c->blkOff.p ;
  out->clusHash = (h10x_cluster_hash*) c->clus.p ; out->codeOff = c->codeOff.p ; out->codes = c->codes.p ;
  out->onDevice = 1 ;
  return H10X_OK ;
}

static void *pinned_alloc (size_t bytes)
{ void *p = nullptr ; CK (cudaHostAlloc (&p, bytes ? bytes : 1, cudaHostAllocDefault)) ; return p ; }

void h10x_index_free (h10x_index *ix)
{ if (!ix || ix->onDevice || ix->pinned == 2) return ;	/* pinned == 2: arrays live in the context's arena */
  void *ps[] = { ix->hashIndex, ix->hashValue, ix->hashDepth, ix->blkNRead, ix->blkNHash, ix->blkOff,
		 ix->clusHash, ix->codeOff, ix->codes } ;
  for (void *p : ps) if (p) { if (ix->pinned) cudaFreeHost (p) ; else free (p) ; }
  if (!ix->pinned) { free (ix->blkNSubCluster) ; free (ix->blkPointToMin) ; free (ix->blkClusterParent) ; }	/* h10x_read_hash's; otherwise borrowed */
  ix->blkNSubCluster = nullptr ; ix->blkPointToMin = nullptr ; ix->blkClusterParent = nullptr ;
  ix->hashIndex = nullptr ; ix->hashValue = nullptr ; ix->hashDepth = nullptr ; ix->blkNRead = nullptr ;
  ix->blkNHash = nullptr ; ix->blkOff = nullptr ; ix->clusHash = nullptr ; ix->codeOff = nullptr ; ix->codes = nullptr ;
}

int h10x_gpu_download (h10x_ctx *c, h10x_index *out, char *err, size_t errlen)
{ if (!c || !out || !c->haveIndex) { set_err (err, errlen, "no index resident") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      out->B = c->P.B ; out->hashNumber = c->hashNumber ; out->nBlocksMax = c->nBlocksMax ;
      out->nReads = c->nReads ; out->nHashes = c->nHashes ; out->pinned = 2 ;
      size_t hn = c->hashNumber, nb = c->nBlocksMax, H = c->nHashes ;
      int slot = 0 ;
      auto pull = [&] (void **dst, const void *src, size_t bytes)
	{ int i = slot++ ;
	  if (!src) { *dst = nullptr ; return ; }
	  if (c->slotDone[i]) { *dst = c->hostSlot[i] ; return ; }	/* already on its way (early_pull) */
	  *dst = host_slot (c, i, bytes) ;
	  if (bytes) CK (cudaMemcpyAsync (*dst, src, bytes, cudaMemcpyDeviceToHost, s)) ;
	} ;
      pull ((void**) &out->hashIndex, c->hashIndex.p, c->hashIndex.p ? ((size_t) 4 << c->P.B) : 0) ;
      pull ((void**) &out->hashValue, c->hashValue.p, 8 * hn) ;
      pull ((void**) &out->hashDepth, c->hashDepth.p, 4 * hn) ;
      pull ((void**) &out->blkNRead, c->blkNRead.p, 4 * nb) ;
      pull ((void**) &out->blkNHash, c->blkNHash.p, 4 * nb) ;
      pull ((void**) &out->blkOff, c->blkOff.p, 8 * (nb + 1)) ;
      pull ((void**) &out->clusHash, c->clus.p, 8 * H) ;
            const bool lazy = (c->P.flags & H10X_FLAG_LAZY_CODES) != 0 ;
      pull ((void**) &out->codeOff, lazy ? nullptr : c->codeOff.p, c->codeOff.p ? 8 * (hn + 1) : 0) ;
      pull ((void**) &out->codes, lazy ? nullptr : c->codes.p, c->codes.p ? 4 * H : 0) ;
      CK (cudaStreamSynchronize (s)) ;
      if (c->dlStream) CK (cudaStreamSynchronize (c->dlStream)) ;
      for (int i = 0 ; i < 9 ; ++i) c->slotDone[i] = false ;
    }) ;
  if (st != H10X_OK) memset (out, 0, sizeof (*out)) ;
  return st ;
}

/* The front half of a host-buffer build, overlapped with the copy of the file (hash10x.c:197-223 reads and hashes the
   file chunk by chunk; here the chunks are slabs of the pinned host buffer going up on their own stream): as soon as a
   slab has landed its barcode runs are found and the fused kernel hashes the runs that are complete, while the next
   slabs are on their way.  What build_device_impl gets is the run table of the whole file and the per-run key lists. */
static void prefuse_streamed (h10x_ctx *c, cudaStream_t s, const void *hostFqb, uint32_t *dFqb, uint64_t n64, Prefuse &pf)
{ const h10x_params &P = c->P ; MemTrack *mt = &c->mt ;
  const uint32_t n = (uint32_t) n64 ;
  uint32_t slabRecs = 4u << 20 ;		/* 4M records = 503 MB: 9 ms of PCIe time, ~1.3 ms of hashing */
  if (const char *e = getenv ("H10X_STREAM_SLAB")) { long v = atol (e) ; if (v >= 1) slabRecs = (uint32_t) v ; }
  const uint32_t nSlabs = (n + slabRecs - 1) / slabRecs ;
  std::vector<cudaEvent_t> landed (nSlabs) ;
  for (uint32_t k = 0 ; k < nSlabs ; ++k)
    { const uint64_t r0 = (uint64_t) k * slabRecs, nr = std::min<uint64_t> (slabRecs, n - r0) ;
      CK (cudaMemcpyAsync (dFqb + r0 * H10X_REC_WORDS, (const char*) hostFqb + r0 * 120, nr * 120, cudaMemcpyHostToDevice, c->ulStream)) ;
      landed[k] = ctx_event (c) ;
      CK (cudaEventRecord (landed[k], c->ulStream)) ;
    }
    /* beyond that (runs of fewer than 8 read pairs on average) the classic stages take over */
  const uint32_t runCap = std::max<uint32_t> (n / 8 + 65536, std::min<uint32_t> (slabRecs, n)) ;
  pf.dBlkStart.alloc ((size_t) runCap + 2, s, mt) ; pf.srcOff.alloc ((size_t) runCap + 2, s, mt) ; pf.blkCnt.alloc ((size_t) runCap + 2, s, mt) ;
  CK (cudaMemsetAsync (pf.blkCnt.p, 0xff, 4 * ((size_t) runCap + 2), s)) ;
  DBuf<uint32_t> flag (std::min<uint32_t> (slabRecs, n), s, mt), dCount (1, s, mt) ;
  DBuf<int> anyZero (1, s, mt) ;
  CK (cudaMemsetAsync (anyZero.p, 0, sizeof (int), s)) ;
  pf.runStart.clear () ; pf.nRuns = 0 ;
  uint32_t done = 0 ;		/* runs handed to the fused kernel */
  bool ok = true ;
  for (uint32_t k = 0 ; k < nSlabs ; ++k)
    { const uint32_t r0 = k * slabRecs, nr = std::min<uint32_t> (slabRecs, n - r0) ;
            CK (cudaStreamWaitEvent (s, landed[k], 0)) ;
      if (ok && (uint64_t) pf.nRuns + nr > (uint64_t) runCap) ok = false ;	/* the selection below may write a start per record */
      if (!ok) continue ;
      LAUNCH (c, k_run_flags_at, gridFor (nr, 256), 256, 0, s, dFqb, r0, nr, flag.p, anyZero.p) ;
      cub::CountingInputIterator<uint32_t> recNo (r0) ;
      cubCall (c, s, [&] (void *t, size_t &b)
	{ return cub::DeviceSelect::Flagged (t, b, recNo, flag.p, pf.dBlkStart.p + pf.nRuns, dCount.p, (int) nr, s) ; }) ;
      
      uint32_t cnt = 0 ;
      CK (cudaMemcpyAsync (&cnt, dCount.p, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      
      pf.runStart.resize ((size_t) pf.nRuns + cnt) ;
      if (cnt)
	{ CK (cudaMemcpyAsync (pf.runStart.data () + pf.nRuns, pf.dBlkStart.p + pf.nRuns, 4 * (size_t) cnt, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaStreamSynchronize (s)) ;
	}
      pf.nRuns += cnt ;
      if (!pf.eng.ready)
	{ if (!FusedEngine::usable (P)) { ok = false ; continue ; }
	  const uint32_t runsGuess = (uint32_t) std::min<uint64_t> ((uint64_t) pf.nRuns * nSlabs * 3 / 2 + 16, 0x7fffffffull) ;
	  pf.eng.init (c, s, lean_wanted (c, false, n, runsGuess), n, nSlabs + 1) ;
	}
      if (pf.nRuns > done + 1)		/* run i is complete once run i + 1 has started */
	{ std::vector<uint32_t> lists[5] ;
	  for (uint32_t r = done ; r + 1 < pf.nRuns ; ++r)
	    { const int ci = pf.eng.classOf (P, pf.runStart[r + 1] - pf.runStart[r]) ;
	      if (ci >= 0) lists[ci].push_back (r) ;
	    }
	  pf.eng.launch (c, s, lists, dFqb, pf.dBlkStart.p, pf.srcOff.p, pf.blkCnt.p) ;
	  CK (cudaStreamSynchronize (s)) ;	/* the lists are read by the async copies of launch (); the slab after this one is on its way */
	  done = pf.nRuns - 1 ;
	}
    }
  int hAnyZero = 0 ;
  CK (cudaMemcpyAsync (&hAnyZero, anyZero.p, 4, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaMemcpyAsync (pf.dBlkStart.p + pf.nRuns, &n, 4, cudaMemcpyHostToDevice, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  pf.runStart.push_back (n) ;
  pf.anyZero = hAnyZero != 0 ;
  pf.valid = ok && pf.nRuns > 0 ;
  if (!pf.valid) { pf.eng.release () ; pf.srcOff.release () ; pf.blkCnt.release () ; pf.dBlkStart.release () ; }
}

/* ------------------------------------------------------------------ --hashStats / --codeStats / --cribBuild (h10x_crib.cuh) */

int h10x_gpu_histogram (h10x_ctx *c, int which, const int **hist, int *n, char *err, size_t errlen)
{ if (!c || !hist || !n || !c->haveIndex || which < 0 || which > 2) { set_err (err, errlen, "no index resident") ; return H10X_ERR_BAD_PARAM ; }
  const uint32_t *v = which == 0 ? c->hashDepth.p : which == 1 ? c->blkNHash.p : c->blkNSub.p ;
  const uint64_t cnt = which == 0 ? c->hashNumber : c->nBlocksMax ;
  if (!v) { set_err (err, errlen, which == 2 ? "no sub-clusters yet" : "array not resident on this rank") ; return H10X_ERR_BAD_PARAM ; }
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      DBuf<uint32_t> dTop (1, s, mt) ;
      CK (cudaMemsetAsync (dTop.p, 0, 4, s)) ;
      const unsigned grid = (unsigned) std::min<uint64_t> (gridFor (cnt, 256), (uint64_t) device_sms (c) * 8) ;
      LAUNCH (c, k_max_u32, grid, 256, 0, s, v, cnt, dTop.p) ;
      uint32_t top = 0 ;
      CK (cudaMemcpyAsync (&top, dTop.p, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      DBuf<int> dHist ((size_t) top + 1, s, mt) ;
      CK (cudaMemsetAsync (dHist.p, 0, 4 * ((size_t) top + 1), s)) ;
      LAUNCH (c, k_hist_u32, grid, 256, 0, s, v, cnt, top + 1, dHist.p) ;
      c->histHost.resize ((size_t) top + 1) ;
      CK (cudaMemcpyAsync (c->histHost.data (), dHist.p, 4 * ((size_t) top + 1), cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      *hist = c->histHost.data () ; *n = (int) top + 1 ;
    }) ;
}

int h10x_gpu_crib_build (h10x_ctx *c, const uint8_t *g1, const uint64_t *off1, uint32_t nSeq1,
			 const uint8_t *g2, const uint64_t *off2, uint32_t nSeq2, h10x_crib *out, char *err, size_t errlen)
{ if (!c || !out || !c->haveIndex || c->dist) { set_err (err, errlen, "no single-GPU index resident") ; return H10X_ERR_BAD_PARAM ; }
  if (!c->hashIndex.p || !c->hashValue.p || !c->hashDepth.p) { set_err (err, errlen, "the bin table is not resident (H10X_FLAG_NO_TABLE)") ; return H10X_ERR_BAD_PARAM ; }
  if ((nSeq1 && (!g1 || !off1)) || (nSeq2 && (!g2 || !off2))) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      const uint32_t hn = c->hashNumber ;
      DBuf<uint32_t> cnt[2] ; DBuf<unsigned long long> first[2] ; DBuf<CribCounts> cc (2, s, mt) ;
      CK (cudaMemsetAsync (cc.p, 0, 2 * sizeof (CribCounts), s)) ;
      CribCounts hcc[2] ;
      for (int g = 0 ; g < 2 ; ++g)
	{ const uint8_t *codes = g ? g2 : g1 ; const uint64_t *off = g ? off2 : off1 ; const uint32_t nSeq = g ? nSeq2 : nSeq1 ;
	  cnt[g].alloc (hn, s, mt) ; first[g].alloc (hn, s, mt) ;
	  CK (cudaMemsetAsync (cnt[g].p, 0, 4 * (size_t) hn, s)) ;
	  CK (cudaMemsetAsync (first[g].p, 0xff, 8 * (size_t) hn, s)) ;
	  const uint64_t total = nSeq ? off[nSeq] : 0 ;
	  if (total)
	    { DBuf<uint8_t> dCodes (total, s, mt) ; DBuf<uint64_t> dOff ((size_t) nSeq + 1, s, mt) ;
	      CK (cudaMemcpyAsync (dCodes.p, codes, total, cudaMemcpyHostToDevice, s)) ;
	      CK (cudaMemcpyAsync (dOff.p, off, 8 * ((size_t) nSeq + 1), cudaMemcpyHostToDevice, s)) ;
	      const unsigned grid = (unsigned) std::min<uint64_t> (gridFor (total, 256), (uint64_t) device_sms (c) * 16) ;
	      LAUNCH (c, k_crib_scan, grid, 256, 0, s, dCodes.p, total, dOff.p, nSeq, c->hp, c->hashIndex.p, c->P.B, c->hashValue.p,
		      cnt[g].p, first[g].p, cc.p + g) ;
	      CK (cudaStreamSynchronize (s)) ;		/* the staging buffers go out of scope */
	    }
	  out->nSeq[g] = (int32_t) nSeq ;
	}
      CK (cudaMemcpyAsync (hcc, cc.p, sizeof (hcc), cudaMemcpyDeviceToHost, s)) ;
      /* the groups' histograms by bin depth */
      DBuf<uint32_t> dTop (1, s, mt) ;
      CK (cudaMemsetAsync (dTop.p, 0, 4, s)) ;
      LAUNCH (c, k_max_u32, (unsigned) std::min<uint64_t> (gridFor (hn, 256), (uint64_t) device_sms (c) * 8), 256, 0, s, c->hashDepth.p, (uint64_t) hn, dTop.p) ;
      uint32_t top = 0 ;
      CK (cudaMemcpyAsync (&top, dTop.p, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      const uint32_t histLen = top + 1 ;
      DBuf<int> dHist ((size_t) 4 * histLen, s, mt) ; DBuf<uint8_t> dType (hn, s, mt) ; DBuf<int16_t> dChr (hn, s, mt) ; DBuf<uint16_t> dPos (hn, s, mt) ;
      CK (cudaMemsetAsync (dHist.p, 0, 16 * (size_t) histLen, s)) ;
      LAUNCH (c, k_crib_classify, gridFor (hn, 256), 256, 0, s, hn, cnt[0].p, first[0].p, cnt[1].p, first[1].p, c->hashDepth.p, histLen,
	      dType.p, dChr.p, dPos.p, dHist.p) ;
      c->cribHist.resize ((size_t) 4 * histLen) ; c->cribType.resize (hn) ; c->cribChr.resize (hn) ; c->cribPos.resize (hn) ;
      CK (cudaMemcpyAsync (c->cribHist.data (), dHist.p, 16 * (size_t) histLen, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaMemcpyAsync (c->cribType.data (), dType.p, hn, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaMemcpyAsync (c->cribChr.data (), dChr.p, 2 * (size_t) hn, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaMemcpyAsync (c->cribPos.data (), dPos.p, 2 * (size_t) hn, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      out->hashNumber = hn ; out->histLen = (int32_t) histLen ;
      out->type = c->cribType.data () ; out->chr = c->cribChr.data () ; out->pos = c->cribPos.data () ;
      for (int g = 0 ; g < 2 ; ++g) { out->nPresent[g] = (int32_t) hcc[g].nPresent ; out->nAbsent[g] = (int32_t) hcc[g].nAbsent ; }
      for (int t = 0 ; t < 4 ; ++t)
	{ const int *h = c->cribHist.data () + (size_t) t * histLen ;
	  out->hist[t] = h ;
	  int mx = 0 ; for (uint32_t d = 0 ; d < histLen ; ++d) if (h[d]) mx = (int) d + 1 ;
	  out->histMax[t] = mx ;
	}
    }) ;
}

/* ------------------------------------------------------------------ fq2b + bsort (h10x_fq2b.cuh) */

int h10x_pack_barcode (const char *s, uint32_t *out)
{ if (!s || !out || strlen (s) != 16) return -1 ;
  uint32_t u = 0 ;
  for (int i = 0 ; i < 16 ; ++i)
    { const char ch = s[i] ;
      u = (u << 2) | ((ch == 'c' || ch == 'C') ? 1u : (ch == 'g' || ch == 'G') ? 2u : (ch == 't' || ch == 'T') ? 3u : 0u) ;
    }
  *out = u ;
  return 0 ;
}

struct FqFile { DBuf<char> text ; DBuf<unsigned long long> nl ; uint64_t nLines = 0, nRec = 0 ; uint32_t L = 0 ; } ;

/* text -> device, newline positions, line length of the first sequence line, gzReadFastq's checks */
static void fq_index (h10x_ctx *c, cudaStream_t s, const char *host, uint64_t n, FqFile &f, int entryMul, int entryAdd)
{ MemTrack *mt = &c->mt ;
  f.text.alloc (n + 1, s, mt) ;
  if (n) CK (cudaMemcpyAsync (f.text.p, host, n, cudaMemcpyHostToDevice, s)) ;
  f.nl.alloc (n / 2 + 4, s, mt) ;		/* at most every other byte is a newline of a well-formed entry (checked below) */
  DBuf<unsigned long long> dCount (1, s, mt) ;
  unsigned long long cnt = 0 ;
  if (n)
    { /* a text with more newlines than n/2 is malformed; count first so that the selection cannot overrun */
      FqIsNewline pred = { f.text.p } ;
      cub::CountingInputIterator<unsigned long long> pos (0ull) ;
      cub::TransformInputIterator<unsigned long long, FqNlCount, cub::CountingInputIterator<unsigned long long>> ones (pos, FqNlCount { f.text.p }) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceReduce::Sum (t, b, ones, dCount.p, (::cuda::std::int64_t) n, s) ; }) ;
      CK (cudaMemcpyAsync (&cnt, dCount.p, 8, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      if (cnt > n / 2 + 4) throw H10xError (H10X_ERR_IO, "fastq id line for entry " + std::to_string (entryAdd) + " does not start with @") ;
      cubCall (c, s, [&] (void *t, size_t &b)
	{ return cub::DeviceSelect::If (t, b, pos, f.nl.p, dCount.p, (::cuda::std::int64_t) n, pred, s) ; }) ;
    }
  f.nLines = cnt ; f.nRec = cnt / 4 ;
  char last = '\n' ;
  unsigned long long first2[2] = { 0, 0 } ;
  if (n) CK (cudaMemcpyAsync (&last, f.text.p + (n - 1), 1, cudaMemcpyDeviceToHost, s)) ;
  if (cnt >= 2) CK (cudaMemcpyAsync (first2, f.nl.p, 16, cudaMemcpyDeviceToHost, s)) ;
  CK (cudaStreamSynchronize (s)) ;
  if ((cnt & 3) || last != '\n')	/* the reference dies inside the unfinished entry; which message depends on where it stops */
    throw H10xError (H10X_ERR_IO, "truncated fastq entry " + std::to_string ((long long) f.nRec * entryMul + entryAdd)) ;
  f.L = f.nRec ? (uint32_t) (first2[1] - first2[0] - 1) : 0 ;
  if (f.nRec && (f.L == 0 || f.L > 1023)) throw H10xError (H10X_ERR_IO, "fastq sequence lines of 1..1023 bases expected (fq2b.c:142)") ;
  if (f.nRec)
    { DBuf<unsigned long long> dErr (1, s, mt) ;
      CK (cudaMemsetAsync (dErr.p, 0xff, 8, s)) ;
      LAUNCH (c, k_fq_check, gridFor (f.nRec, 256), 256, 0, s, f.text.p, f.nl.p, f.nRec, f.L, dErr.p) ;
      unsigned long long e = 0 ;
      CK (cudaMemcpyAsync (&e, dErr.p, 8, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      if (e != ~0ull)
	{ const long long entry = (long long) (e >> 3) * entryMul + entryAdd ; const int code = (int) (e & 7) ;
	  const std::string en = std::to_string (entry) ;
	  throw H10xError (H10X_ERR_IO, code == FQ_ERR_ID ? "fastq id line for entry " + en + " does not start with @"
			   : code == FQ_ERR_SEQ ? "fastq entry " + en + " seq line does not end in \\n"
			   : code == FQ_ERR_PLUS ? "bad + fastq line entry " + en
			   : "fastq entry " + en + " qual line does not end in \\n") ;
	}
    }
}

int h10x_gpu_fq2b (h10x_ctx *c, const char *fq1, uint64_t n1, const char *fq2, uint64_t n2,
		   const uint32_t *whitelist, uint64_t nWhitelist, uint32_t flags, h10x_fq2b_out *out, char *err, size_t errlen)
{ if (!c || !out || (!fq1 && n1) || (!fq2 && n2) || (!whitelist && nWhitelist)) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  if (nWhitelist >= (1ull << 24)) { set_err (err, errlen, "more than 2^24-1 whitelist barcodes") ; return H10X_ERR_UNSUPPORTED ; }
  memset (out, 0, sizeof (*out)) ;
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      /* the whitelist table outlives the call: 16 GiB are not worth re-filling for every chunk of a run */
      if (whitelist)
	{ uint64_t sum = 1469598103934665603ull ;
	  for (uint64_t i = 0 ; i < nWhitelist ; ++i) sum = (sum ^ whitelist[i]) * 1099511628211ull ;
	  if (!c->wlTable || c->wlCount != nWhitelist || c->wlSum != sum)
	    { if (!c->wlTable) CK (cudaMalloc (&c->wlTable, (size_t) 4 << 32)) ;
	      CK (cudaMemsetAsync (c->wlTable, 0, (size_t) 4 << 32, s)) ;
	      uint32_t *dWl = nullptr ;
	      CK (cudaMalloc (&dWl, 4 * (size_t) (nWhitelist ? nWhitelist : 1))) ;
	      CK (cudaMemcpyAsync (dWl, whitelist, 4 * (size_t) nWhitelist, cudaMemcpyHostToDevice, s)) ;
	      if (nWhitelist) LAUNCH (c, k_wl_build, gridFor (nWhitelist * 64, 256), 256, 0, s, dWl, nWhitelist, c->wlTable) ;
	      CK (cudaStreamSynchronize (s)) ;
	      cudaFree (dWl) ;
	      c->wlCount = nWhitelist ; c->wlSum = sum ;
	    }
	}
      if (c->fqRecs) { cudaFree (c->fqRecs) ; c->fqRecs = nullptr ; }
      with_slab (c, s, 3 * (n1 + n2) + ((size_t) 64 << 20), [&] ()
	{ MemTrack *mt = &c->mt ;
	  if (c->fqRecs) { cudaFree (c->fqRecs) ; c->fqRecs = nullptr ; }
	  FqFile f1, f2 ;
	  fq_index (c, s, fq1, n1, f1, fq2 ? 2 : 1, 1) ;
	  if (fq2)
	    { fq_index (c, s, fq2, n2, f2, 2, 2) ;
	      if (f2.nRec < f1.nRec) throw H10xError (H10X_ERR_IO, "second fastq file terminated early at " + std::to_string (f2.nRec)) ;
	    }
	  const uint64_t nRec = f1.nRec ;
	  const uint32_t w1 = (f1.L + 15) / 16 + (f1.L + 31) / 32, w2 = fq2 ? (f2.L + 15) / 16 + (f2.L + 31) / 32 : 0 ;
	  const uint32_t recWords = w1 + w2 ;
	  out->nRead = nRec ; out->recWords = recWords ; out->s1Len = f1.L ; out->s2Len = fq2 ? f2.L : 0 ;
	  if (nRec >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 fastq entries in one call") ;
	  DBuf<uint32_t> recs ((size_t) nRec * recWords, s, mt) ;
	  if (nRec)
	    { LAUNCH (c, k_fq_pack, gridFor (nRec * w1, 256), 256, 0, s, f1.text.p, f1.nl.p, nRec, f1.L, recWords, 0u, recs.p) ;
	      if (fq2) LAUNCH (c, k_fq_pack, gridFor (nRec * w2, 256), 256, 0, s, f2.text.p, f2.nl.p, nRec, f2.L, recWords, w1, recs.p) ;
	    }
	  f1.text.release () ; f1.nl.release () ; f2.text.release () ; f2.nl.release () ;
	  /* barcode correction; the kept records, in input order */
	  uint64_t nKeep = nRec ;
	  DBuf<unsigned long long> idx ;
	  if (whitelist && nRec)
	    { DBuf<uint32_t> keep (nRec, s, mt) ; DBuf<FqStats> st (1, s, mt) ; DBuf<unsigned long long> dN (1, s, mt) ;
	      CK (cudaMemsetAsync (st.p, 0, sizeof (FqStats), s)) ;
	      LAUNCH (c, k_wl_apply, gridFor (nRec, 256), 256, 0, s, recs.p, nRec, recWords, c->wlTable, keep.p, st.p) ;
	      idx.alloc (nRec, s, mt) ;
	      cub::CountingInputIterator<unsigned long long> pos (0ull) ;
	      cubCall (c, s, [&] (void *t, size_t &b)
		{ return cub::DeviceSelect::Flagged (t, b, pos, keep.p, idx.p, dN.p, (::cuda::std::int64_t) nRec, s) ; }) ;
	      FqStats hs ; unsigned long long hn = 0 ;
	      CK (cudaMemcpyAsync (&hs, st.p, sizeof (FqStats), cudaMemcpyDeviceToHost, s)) ;
	      CK (cudaMemcpyAsync (&hn, dN.p, 8, cudaMemcpyDeviceToHost, s)) ;
	      CK (cudaStreamSynchronize (s)) ;
	      nKeep = hn ; out->nBad = hs.nBad ; out->nFixed = hs.nFixed ;
	      for (int i = 0 ; i < 16 ; ++i) out->nFixBase[i] = hs.nFixBase[i] ;
	    }
	  out->nRecords = nKeep ;
	  CK (cudaMalloc (&c->fqRecs, 4 * (size_t) (nKeep ? nKeep : 1) * (recWords ? recWords : 1))) ;
	  if (nKeep)
	    { DBuf<uint64_t> wa, wb ;
	      const uint64_t *order = nullptr ;
	      if (flags & H10X_FQ2B_SORT)
		{ wa.alloc (nKeep, s, mt) ; wb.alloc (nKeep, s, mt) ;
		  LAUNCH (c, k_fq_sort_words, gridFor (nKeep, 256), 256, 0, s, recs.p, recWords, idx.p, nKeep, wa.p) ;
		  /* least significant digit first, four stable passes of 8 bits over the byte-swapped first word */
		  uint64_t *src = wa.p, *dst = wb.p ;
		  for (int p = 0 ; p < 4 ; ++p)
		    { LoadWord ld = { src, 32 + 8 * p, 255u, ~(uint64_t) 0 } ;
		      part_pass (c, s, ld, nKeep, 256u, dst) ;
		      std::swap (src, dst) ;
		    }
		  order = src ;
		}
	      LAUNCH (c, k_fq_gather, gridFor (nKeep * 32, 256), 256, 0, s, recs.p, recWords, idx.p, order, nKeep, c->fqRecs) ;
	      CK (cudaStreamSynchronize (s)) ;
	    }
	}) ;
      out->d_fqb = c->fqRecs ;
      if (!(flags & H10X_FQ2B_NO_HOST))
	{ const size_t bytes = 4 * (size_t) out->nRecords * out->recWords ;
	  if (c->fqHostCap < bytes || !c->fqHost)
	    { if (c->fqHost) cudaFreeHost (c->fqHost) ;
	      c->fqHost = nullptr ; c->fqHostCap = 0 ;
	      CK (cudaHostAlloc (&c->fqHost, bytes ? bytes : 1, cudaHostAllocDefault)) ;
	      c->fqHostCap = bytes ? bytes : 1 ;
	    }
	  if (bytes) CK (cudaMemcpyAsync (c->fqHost, c->fqRecs, bytes, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaStreamSynchronize (s)) ;
	  out->fqb = c->fqHost ;
	}
    }) ;
}

int h10x_gpu_download_codes (h10x_ctx *c, h10x_index *out, char *err, size_t errlen)
{ if (!c || !out || !c->haveIndex) { set_err (err, errlen, "no index resident") ; return H10X_ERR_BAD_PARAM ; }
  if (!c->codes.p || !c->codeOff.p) { set_err (err, errlen, "the hash->code lists were not built (H10X_FLAG_NO_CODES)") ; return H10X_ERR_BAD_PARAM ; }
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      const size_t hn = c->hashNumber, H = c->codes.n ;	/* after h10x_gpu_dist_global_codes: all ranks' pairs */
      out->codeOff = (uint64_t*) host_slot (c, SLOT_CODEOFF, 8 * (hn + 1)) ;
      out->codes = (uint32_t*) host_slot (c, SLOT_CODES, 4 * H) ;
      CK (cudaMemcpyAsync (out->codeOff, c->codeOff.p, 8 * (hn + 1), cudaMemcpyDeviceToHost, s)) ;
      if (H) CK (cudaMemcpyAsync (out->codes, c->codes.p, 4 * H, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
    }) ;
}

int h10x_gpu_build_host (h10x_ctx *c, const void *fqb, uint64_t nRecords, h10x_index *out, char *err, size_t errlen)
{ if (!c || !out || (!fqb && nRecords)) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      /* only the first N records are ever looked at (hash10x.c:202,207) */
      uint64_t n = (c->P.N > 0 && (uint64_t) c->P.N < nRecords) ? (uint64_t) c->P.N : nRecords ;
      with_slab (c, s, slab_estimate (c->P, n, true), [&] ()
		{ reset_result (c) ;
	  DBuf<uint32_t> d ((size_t) n * H10X_REC_WORDS, s, &c->mt) ;
	  if (!c->dlStream) CK (cudaStreamCreateWithFlags (&c->dlStream, cudaStreamNonBlocking)) ;
	  Prefuse pf ;
	  const bool streamed = !c->dist && n > 0 && n < 0xffffffffull && FusedEngine::usable (c->P) && !getenv ("H10X_NO_STREAM") ;
	  if (streamed)
	    { if (!c->ulStream) CK (cudaStreamCreateWithFlags (&c->ulStream, cudaStreamNonBlocking)) ;
	      try { prefuse_streamed (c, s, fqb, d.p, n, pf) ; }
	      catch (...) { cudaStreamSynchronize (c->ulStream) ; throw ; }
	    }
	  else if (n) CK (cudaMemcpyAsync (d.p, fqb, (size_t) n * 120, cudaMemcpyHostToDevice, s)) ;
	  c->earlyDl = true ;
	  try { build_device_impl (c, d.p, n, s, false, false, streamed ? &pf : nullptr) ; }
	  catch (...) { c->earlyDl = false ; cudaStreamSynchronize (c->dlStream) ; throw ; }
	  c->earlyDl = false ;
	}) ;
    }) ;
  if (st != H10X_OK) { cudaStreamSynchronize (c->own) ; cudaGetLastError () ; c->haveIndex = false ; return st ; }
  return h10x_gpu_download (c, out, err, errlen) ;
}

/* readHashFile()'s counterpart for the device (hash10x.c:269-315 + fillHashTable :317-347 done by the caller): a host
   index - from h10x_read_hash, with the hash->code lists rebuilt - becomes the resident index, so that
   --hashDepthRange and --cluster run on the GPU in a session that starts with --readHash as the README pipelines do */
int h10x_gpu_load_index (h10x_ctx *c, const h10x_index *h, char *err, size_t errlen)
{ if (!c || !h || h->onDevice) { set_err (err, errlen, "null or device index") ; return H10X_ERR_BAD_PARAM ; }
  if (c->dist) { set_err (err, errlen, "not on a distributed context") ; return H10X_ERR_BAD_PARAM ; }
  if (h->B != c->P.B) { set_err (err, errlen, "incompatible hash table size") ; return H10X_ERR_BAD_PARAM ; }
  if (!h->hashValue || !h->hashDepth || !h->blkNRead || !h->blkNHash || !h->blkOff || (h->nHashes && !h->clusHash)
      || h->hashNumber < 1 || h->nBlocksMax < 1)
    { set_err (err, errlen, "incomplete index") ; return H10X_ERR_BAD_PARAM ; }
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      const size_t hn = h->hashNumber, nb = h->nBlocksMax, H = h->nHashes ;
      /* the arrays themselves plus what --hashDepthRange needs beside them */
      const size_t estimate = 64 * H + 32 * hn + 64 * nb + (h->hashIndex ? ((size_t) 4 << h->B) : 0) + ((size_t) 64 << 20) ;
      with_slab (c, s, estimate, [&] ()
	{ reset_result (c) ;
	  MemTrack *mt = &c->mt ;
	  auto up = [&] (auto &buf, const void *src, size_t n, size_t elem)
	    { buf.alloc (n, s, mt) ; if (n) CK (cudaMemcpyAsync (buf.p, src, n * elem, cudaMemcpyHostToDevice, s)) ; } ;
	  if (h->hashIndex) up (c->hashIndex, h->hashIndex, (size_t) 1 << h->B, 4) ;
	  up (c->hashValue, h->hashValue, hn, 8) ;
	  c->hashDepth.alloc (hn + 1, s, mt) ;		/* one spare 0 as after a build */
	  CK (cudaMemcpyAsync (c->hashDepth.p, h->hashDepth, 4 * hn, cudaMemcpyHostToDevice, s)) ;
	  CK (cudaMemsetAsync (c->hashDepth.p + hn, 0, 4, s)) ;
	  up (c->blkNRead, h->blkNRead, nb, 4) ; up (c->blkNHash, h->blkNHash, nb, 4) ;
	  up (c->blkOff, h->blkOff, nb + 1, 8) ;
	  up (c->clus, h->clusHash, H, 8) ;
	  if (h->codeOff && h->codes) { up (c->codeOff, h->codeOff, hn + 1, 8) ; up (c->codes, h->codes, H, 4) ; }
	  if (h->blkNSubCluster && h->blkPointToMin)	/* what an earlier --cluster left in the file */
	    { up (c->blkNSub, h->blkNSubCluster, nb, 4) ; up (c->blkPtm, h->blkPointToMin, nb, 8) ; }
	  if (h->blkClusterParent) up (c->blkParent, h->blkClusterParent, nb, 4) ;
	  c->hashNumber = (uint32_t) hn ; c->nBlocksMax = (uint32_t) nb ; c->nReads = h->nReads ; c->nHashes = H ;
	  CK (cudaStreamSynchronize (s)) ;
	  c->haveIndex = true ;
	}) ;
    }) ;
  if (st != H10X_OK) { cudaStreamSynchronize (c->own) ; cudaGetLastError () ; c->haveIndex = false ; }
  return st ;
}

int h10x_gpu_build_file (h10x_ctx *c, const char *path, h10x_index *out, char *err, size_t errlen)
{ if (!c || !path || !out) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  FILE *f = fopen (path, "rb") ;
  if (!f) { set_err (err, errlen, "failed to open fqb file") ; return H10X_ERR_IO ; }
  uint32_t *d_fqb = nullptr ; uint64_t n = 0 ;
  void *stage[2] = { nullptr, nullptr } ;
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      if (fseeko (f, 0, SEEK_END)) throw H10xError (H10X_ERR_IO, "file read problem") ;
      uint64_t bytes = (uint64_t) ftello (f) ; rewind (f) ;
      n = bytes / 120 ;			/* fread(u,120,..) drops a trailing partial record */
      if (c->P.N > 0 && (uint64_t) c->P.N < n) n = (uint64_t) c->P.N ;
      CK (cudaMalloc ((void**) &d_fqb, std::max<uint64_t> (n * 120, 16))) ;
      const size_t chunkRecs = 1u << 19 ;	/* 60 MB pinned staging, double buffered */
      stage[0] = pinned_alloc (chunkRecs * 120) ; stage[1] = pinned_alloc (chunkRecs * 120) ;
      cudaEvent_t done[2] ; CK (cudaEventCreate (&done[0])) ; CK (cudaEventCreate (&done[1])) ;
      uint64_t pos = 0 ; int cur = 0 ; bool used[2] = { false, false } ;
      while (pos < n)
	{ size_t want = (size_t) std::min<uint64_t> (chunkRecs, n - pos) ;
	  if (used[cur]) CK (cudaEventSynchronize (done[cur])) ;
	  size_t got = fread (stage[cur], 120, want, f) ;
	  if (got != want) throw H10xError (H10X_ERR_IO, "file read problem") ;
	  CK (cudaMemcpyAsync ((char*) d_fqb + pos * 120, stage[cur], got * 120, cudaMemcpyHostToDevice, s)) ;
	  CK (cudaEventRecord (done[cur], s)) ; used[cur] = true ;
	  pos += got ; cur ^= 1 ;
	}
      CK (cudaStreamSynchronize (s)) ;
      cudaEventDestroy (done[0]) ; cudaEventDestroy (done[1]) ;
      with_slab (c, s, slab_estimate (c->P, n, false), [&] () { build_device_impl (c, d_fqb, n, s) ; }) ;
    }) ;
  fclose (f) ;
  if (stage[0]) cudaFreeHost (stage[0]) ;
  if (stage[1]) cudaFreeHost (stage[1]) ;
  if (st != H10X_OK) { cudaStreamSynchronize (c->own) ; cudaGetLastError () ; c->haveIndex = false ; }
  if (d_fqb) cudaFree (d_fqb) ;
  if (st != H10X_OK) return st ;
  return h10x_gpu_download (c, out, err, errlen) ;
}

static int depth_range_impl (h10x_ctx *c, int dmin, int dmax, h10x_good_hashes *out, bool download, char *err, size_t errlen)
{ if (!c || !out || !c->haveIndex || (c->dist && !c->dist->globalCodes))
    { set_err (err, errlen, "no index resident (after a distributed build: h10x_gpu_dist_global_codes first)") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      const uint32_t hn = c->hashNumber, nb = c->nBlocksMax ;
      const uint64_t H = c->nHashes ;
      if (!c->within.p) { c->within.alloc (hn, s, mt) ; CK (cudaMemsetAsync (c->within.p, 0, hn, s)) ; }
      LAUNCH (c, k_within, gridFor (hn, 256), 256, 0, s, hn, c->hashDepth.p, dmin, dmax, c->within.p) ;
      uint32_t nGood = 0 ;
      DBuf<uint32_t> flag (H + 1, s, mt), pos (H + 1, s, mt) ;
      c->haveGood = false ;
      c->goodD.release () ; c->goodOffD.alloc ((size_t) nb + 1, s, mt) ;
      DBuf<uint64_t> &goodOff = c->goodOffD ;
      CK (cudaMemsetAsync (flag.p + H, 0, 4, s)) ;
      LAUNCH (c, k_good_mark, std::min<uint32_t> (nb, 148 * 16), 256, 0, s, nb, c->blkOff.p, c->blkNHash.p, c->clus.p, c->within.p, flag.p) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, flag.p, pos.p, H + 1, s) ; }) ;
      CK (cudaMemcpyAsync (&nGood, pos.p + H, 4, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      DBuf<uint32_t> keyDepth (nGood, s, mt), keyS (nGood, s, mt) ;
      DBuf<uint16_t> valIdx (nGood, s, mt) ;
      c->goodD.alloc (nGood, s, mt) ;
      DBuf<uint16_t> &valS = c->goodD ;
      LAUNCH (c, k_good_compact, std::min<uint32_t> (nb + 1, 148 * 16), 256, 0, s, nb, c->blkOff.p, c->blkNHash.p, c->clus.p,
	      c->hashDepth.p, flag.p, pos.p, keyDepth.p, valIdx.p, goodOff.p, H, nGood) ;
      flag.release () ; pos.release () ;
      std::vector<uint64_t> hOff ((size_t) nb + 1) ;
      CK (cudaMemcpyAsync (hOff.data (), goodOff.p, 8 * ((size_t) nb + 1), cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      /* sort by increasing depth inside every block; stable = ties keep list order, as glibc's qsort does.  One CTA per
	 block, LSD radix sort in shared memory on the depth bits (k_cluster_sort: one pass for a range like 30..100);
	 the library's segmented sort, which took 1.1 s of this command's 1.2 s at the 1 Gb workload, only gets the blocks
	 with more good hashes than the largest shared-memory class (H10X_GOOD_LIBSORT=1: all of them, for A/B) */
      if (nGood && getenv ("H10X_GOOD_LIBSORT"))
	segmented_sort_blocks<uint32_t, uint16_t> (c, s, keyDepth.p, keyS.p, valIdx.p, valS.p, hOff, goodOff.p, true) ;
      else if (nGood)
	{ struct ClusClass { uint32_t cap, threads ; } ;
	  static const ClusClass kCC[3] = { { 1024, 128 }, { 4096, 256 }, { 12288, 512 } } ;
	  const int keyBits = bits_for (dmax > 1 ? (uint64_t) dmax - 1 : 1) ;
	  uint32_t digitBits = 8, passes = (keyBits + 7) / 8 ;
	  for (uint32_t db = 9 ; db <= 10 ; ++db) if ((keyBits + db - 1) / db < passes) { digitBits = db ; passes = (keyBits + db - 1) / db ; }
	  std::vector<uint32_t> lists[3] ;
	  std::vector<std::pair<uint32_t, uint32_t>> bigRuns ;
	  for (uint32_t p = 0 ; p < nb ; ++p)
	    { const uint64_t n = hOff[p + 1] - hOff[p] ;
	      if (!n) continue ;
	      const int ci = n <= kCC[0].cap ? 0 : n <= kCC[1].cap ? 1 : n <= kCC[2].cap ? 2 : -1 ;
	      if (ci >= 0) lists[ci].push_back (p) ;
	      else if (!bigRuns.empty () && bigRuns.back ().second == p) bigRuns.back ().second = p + 1 ;
	      else bigRuns.push_back ({ p, p + 1 }) ;
	    }
	  const int nSM = device_sms (c) ;
	  DBuf<unsigned int> cwork (3, s, mt) ;
	  CK (cudaMemsetAsync (cwork.p, 0, 12, s)) ;
	  std::vector<DBuf<uint32_t>> dl (3) ;
	  for (int ci = 0 ; ci < 3 ; ++ci)
	    { if (lists[ci].empty ()) continue ;
	      const ClusClass &cc = kCC[ci] ;
	      dl[ci].alloc (lists[ci].size (), s, mt) ;
	      CK (cudaMemcpyAsync (dl[ci].p, lists[ci].data (), 4 * lists[ci].size (), cudaMemcpyHostToDevice, s)) ;
	      const uint32_t nd = 1u << digitBits ;
	      const size_t smem = (size_t) cc.cap * 12 + 4 + (size_t) nd * 4 + (size_t) (cc.threads / 32) * nd * 2 + 16 ;
	      const void *fn = cc.threads == 128 ? (const void*) k_cluster_sort<128> : cc.threads == 256 ? (const void*) k_cluster_sort<256>
		: (const void*) k_cluster_sort<512> ;
	      CK (cudaFuncSetAttribute (fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
	      int occ = 1 ;
	      CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, fn, (int) cc.threads, smem)) ;
	      if (occ < 1) occ = 1 ;
	      ClusterArgs ca ;
	      ca.list = dl[ci].p ; ca.blkOff = goodOff.p ; ca.entryId = keyDepth.p ; ca.eRead = valIdx.p ; ca.clus = nullptr ; ca.valsOut = valS.p ;
	      ca.work = cwork.p + ci ; ca.nList = (uint32_t) lists[ci].size () ; ca.cap = cc.cap ; ca.digitBits = digitBits ; ca.passes = passes ;
	      void *args[1] = { (void*) &ca } ;
	      const uint32_t grid = (uint32_t) std::min<size_t> (lists[ci].size (), (size_t) nSM * occ) ;
	      CK (cudaLaunchKernel (fn, dim3 (grid), dim3 (cc.threads), args, smem, s)) ;
	      ++c->launches ;
	    }
	  for (auto &r : bigRuns)
	    segmented_sort_blocks<uint32_t, uint16_t> (c, s, keyDepth.p, keyS.p, valIdx.p, valS.p, hOff, goodOff.p, true, r.first, r.second) ;
	  CK (cudaStreamSynchronize (s)) ;	/* the block lists are read by the async copies */
	}
      auto pull = [&] (int slot, const void *src, size_t bytes) -> void*
	{ if (c->goodCap[slot] < bytes || !c->goodSlot[slot])
	    { if (c->goodSlot[slot]) cudaFreeHost (c->goodSlot[slot]) ;
	      c->goodSlot[slot] = nullptr ; c->goodCap[slot] = 0 ;
	      c->goodSlot[slot] = pinned_alloc (bytes) ; c->goodCap[slot] = bytes ? bytes : 1 ;
	    }
	  if (bytes) CK (cudaMemcpyAsync (c->goodSlot[slot], src, bytes, cudaMemcpyDeviceToHost, s)) ;
	  return c->goodSlot[slot] ;
	} ;
      if (download)
	{ out->within = (uint8_t*) pull (0, c->within.p, hn) ;
	  out->goodOff = (uint64_t*) pull (1, goodOff.p, 8 * ((size_t) nb + 1)) ;
	  out->good = (uint16_t*) pull (2, valS.p, 2 * (size_t) nGood) ;
	}
      out->nGood = nGood ; out->hashNumber = hn ; out->nBlocksMax = nb ;
      CK (cudaStreamSynchronize (s)) ;
      c->haveGood = true ;
    }) ;
}

int h10x_gpu_depth_range (h10x_ctx *c, int dmin, int dmax, h10x_good_hashes *out, char *err, size_t errlen)
{ return depth_range_impl (c, dmin, dmax, out, true, err, errlen) ; }

/* the same, the lists staying where h10x_gpu_cluster reads them: no pinned host copies (2.7 GB of them at the 1 Gb
   workload, whose first cudaHostAlloc alone takes a second) */
int h10x_gpu_depth_range_device (h10x_ctx *c, int dmin, int dmax, uint64_t *nGood, char *err, size_t errlen)
{ h10x_good_hashes g ;
  int st = depth_range_impl (c, dmin, dmax, &g, false, err, errlen) ;
  if (st == H10X_OK && nGood) *nGood = g.nGood ;
  return st ;
}

/* --cluster codeMin codeMax (hash10x.c:1241-1256) on the resident index and goodHashes: h10x_subcluster.cuh */
int h10x_gpu_cluster (h10x_ctx *c, int codeMin, int codeMax, int clusterThreshold, h10x_clusters *out, char *err, size_t errlen)
{ if (!c || !out) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  if (!c->haveIndex || (c->dist && !c->dist->globalCodes))
    { set_err (err, errlen, "no index resident (after a distributed build: h10x_gpu_dist_global_codes first)") ; return H10X_ERR_BAD_PARAM ; }
  if (!c->haveGood) { set_err (err, errlen, "you must set hashDepthRange before cluster") ; return H10X_ERR_BAD_PARAM ; }	/* hash10x.c:1258 */
  if (!c->codes.p || !c->codeOff.p) { set_err (err, errlen, "the hash->code lists were not built (H10X_FLAG_NO_CODES)") ; return H10X_ERR_BAD_PARAM ; }
  if (clusterThreshold < 1) { set_err (err, errlen, "clusterThreshold must be at least 1") ; return H10X_ERR_BAD_PARAM ; }
  if (!codeMin) codeMin = 1 ;
  if (!codeMax) codeMax = (int) c->nBlocksMax ;
  if (codeMin < 1 || codeMax > (int) c->nBlocksMax)
    { set_err (err, errlen, "code range outside the barcode blocks") ; return H10X_ERR_BAD_PARAM ; }
  void *scratch = nullptr, *slabWork = nullptr ; size_t slabWorkBytes = 0 ;
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      const uint32_t nb = c->nBlocksMax ;
      const uint64_t H = c->nHashes ;
      if (!c->blkNSub.p)
	{ c->blkNSub.alloc (nb, s, mt) ; c->blkPtm.alloc (nb, s, mt) ;
	  CK (cudaMemsetAsync (c->blkNSub.p, 0, 4 * (size_t) nb, s)) ;
	  CK (cudaMemsetAsync (c->blkPtm.p, 0, 8 * (size_t) nb, s)) ;
	}
      cudaEvent_t evA = ctx_event (c), evB = ctx_event (c) ;
      if (codeMax > codeMin)
	{ int nSM = 148 ;
	  CK (cudaDeviceGetAttribute (&nSM, cudaDevAttrMultiProcessorCount, c->P.device)) ;
	  int occ = 1 ;
	  CK (cudaFuncSetAttribute (k_subcluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) H10X_SC_DYN_SMEM)) ;
	  CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, k_subcluster, H10X_SC_THREADS, H10X_SC_DYN_SMEM)) ;
	  if (occ < 1) throw H10xError (H10X_ERR_CUDA, "k_subcluster does not fit on this device") ;
	  uint32_t cap = 1024 ; int lg = 10 ;
	  const uint64_t nbAll = c->dist ? (uint64_t) c->dist->nBlocksGlobal + 2 : nb ;	/* barcodes a block can share hashes with */
	  while (cap < 2 * nbAll && lg < 31) { cap <<= 1 ; ++lg ; }
	  /* per CTA: the global table, 32 warps of deep-bin counters, bin depth / offset, the per-step results, and the
	     global-memory versions of the per-step arrays and read labels for blocks that do not fit in shared memory */
	  const size_t perCta = (size_t) cap * 8 + (size_t) H10X_SC_WARPS * 65536 * 4 + (size_t) 2 * 65536 * 4 + (size_t) 6 * 65536 * 4
	    + (size_t) 65536 * (2 + 4 + 2 + 1) + (size_t) 65536 * 4 ;
	  /* the workspace comes out of the context's slab when the build's temporaries have left room there (no driver
	     call, and the slab may hold nearly all of the device's memory); a driver allocation otherwise */
	  size_t freeB = 0, totalB = 0 ;
	  CK (cudaMemGetInfo (&freeB, &totalB)) ;
	  const size_t devBudget = freeB - std::min<size_t> (freeB / 8, (size_t) 1 << 30) ;
	  const size_t slabBudget = c->mt.largestFree () > 1024 ? c->mt.largestFree () - 1024 : 0 ;
	  const uint32_t want = (uint32_t) std::min<size_t> ((size_t) nSM * occ, (size_t) (codeMax - codeMin)) ;
	  const bool fromSlab = slabBudget >= perCta * want + 256 || slabBudget >= devBudget ;
	  uint32_t grid = (uint32_t) std::min<size_t> (want, (fromSlab ? slabBudget - 256 : devBudget) / perCta) ;
	  if (grid < 1) throw H10xError (H10X_ERR_NOMEM, "not enough device memory for the cluster workspace") ;
	  const size_t bytes = perCta * grid + 256 ;
	  if (fromSlab) { slabWork = c->mt.take (bytes) ; slabWorkBytes = bytes ; }
	  else CK (cudaMalloc (&scratch, bytes)) ;
	  char *p = (char*) (fromSlab ? slabWork : scratch) ;
	  CK (cudaMemsetAsync (p, 0, bytes, s)) ;
	  SubClusterArgs a ;
	  a.work = (unsigned int*) p ; p += 256 ;
	  a.table = (unsigned long long*) p ; p += (size_t) cap * 8 * grid ;
	  a.cnt = (uint32_t*) p ; p += (size_t) H10X_SC_WARPS * 65536 * 4 * grid ;
	  a.pre = (uint32_t*) p ; p += (size_t) 2 * 65536 * 4 * grid ;
	  a.res = (uint32_t*) p ; p += (size_t) 6 * 65536 * 4 * grid ;
	  a.firstG = (uint32_t*) p ; p += (size_t) 65536 * 4 * grid ;
	  a.readLabG = (int*) p ; p += (size_t) 65536 * 4 * grid ;
	  a.parG = (uint16_t*) p ; p += (size_t) 65536 * 2 * grid ;
	  a.fscanG = (uint16_t*) p ; p += (size_t) 65536 * 2 * grid ;
	  a.gsubG = (uint8_t*) p ;
	  a.tableCap = cap ; a.tableShift = (uint32_t) (32 - lg) ;
	  a.clus = (unsigned long long*) c->clus.p ; a.blkOff = c->blkOff.p ; a.blkNHash = c->blkNHash.p ; a.blkNRead = c->blkNRead.p ;
	  a.hashDepth = c->hashDepth.p ; a.codeOff = c->codeOff.p ; a.codes = c->codes.p ;
	  a.goodOff = c->goodOffD.p ; a.good = c->goodD.p ;
	  a.nSub = c->blkNSub.p ; a.pointToMin = c->blkPtm.p ;
	  a.codeMin = (uint32_t) codeMin ; a.codeMax = (uint32_t) codeMax ; a.threshold = clusterThreshold ;
	  a.codeBase = c->dist ? c->dist->blockBase : 0u ;
	  CK (cudaEventRecord (evA, s)) ;
	  LAUNCH (c, k_subcluster, grid, H10X_SC_THREADS, H10X_SC_DYN_SMEM, s, a) ;
	  CK (cudaEventRecord (evB, s)) ;
	}
      auto pull = [&] (int slot, const void *src, size_t bytes) -> void*
	{ if (c->clusCap[slot] < bytes || !c->clusSlot[slot])
	    { if (c->clusSlot[slot]) cudaFreeHost (c->clusSlot[slot]) ;
	      c->clusSlot[slot] = nullptr ; c->clusCap[slot] = 0 ;
	      c->clusSlot[slot] = pinned_alloc (bytes) ; c->clusCap[slot] = bytes ? bytes : 1 ;
	    }
	  if (bytes) CK (cudaMemcpyAsync (c->clusSlot[slot], src, bytes, cudaMemcpyDeviceToHost, s)) ;
	  return c->clusSlot[slot] ;
	} ;
      out->nSubCluster = (uint32_t*) pull (0, c->blkNSub.p, 4 * (size_t) nb) ;
      out->pointToMin = (double*) pull (1, c->blkPtm.p, 8 * (size_t) nb) ;
      out->clusHash = (h10x_cluster_hash*) host_slot (c, SLOT_CLUS, 8 * H) ;
      if (H) CK (cudaMemcpyAsync (out->clusHash, c->clus.p, 8 * H, cudaMemcpyDeviceToHost, s)) ;
      out->nBlocksMax = nb ; out->nHashes = H ;
      CK (cudaStreamSynchronize (s)) ;
      if (codeMax > codeMin) { float ms = 0 ; CK (cudaEventElapsedTime (&ms, evA, evB)) ; out->msKernel = ms ; }
    }) ;
  if (scratch) { cudaStreamSynchronize (c->own) ; cudaFree (scratch) ; }
  if (slabWork) c->mt.give (slabWork, slabWorkBytes) ;		/* same stream as whatever takes it next */
  if (st != H10X_OK) { cudaGetLastError () ; memset (out, 0, sizeof (*out)) ; }
  return st ;
}

/* --clusterSplit: see include/h10x_gpu.h and h10x_split.cuh */
int h10x_gpu_cluster_split (h10x_ctx *c, h10x_index *out, uint32_t *nNew, char *err, size_t errlen)
{ if (!c || !out || !nNew) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  if (!c->haveIndex || c->dist) { set_err (err, errlen, "no single-GPU index resident") ; return H10X_ERR_BAD_PARAM ; }
  if (!c->codes.p || !c->codeOff.p) { set_err (err, errlen, "the context was created without the hash->code lists") ; return H10X_ERR_BAD_PARAM ; }
  *nNew = 0 ;
  void *scratch = nullptr ; MemTrack scratchMt ;
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      const uint32_t nbOld = c->nBlocksMax ;
      const uint64_t H = c->nHashes ;
      /* new block numbers of every block's first cluster (hash10x.c:963-964,993) */
      std::vector<uint32_t> nSub (nbOld, 0), nReadOld (nbOld, 0), base (nbOld, 0) ;
      std::vector<double> ptmOld (nbOld, 0.0) ;
      if (c->blkNSub.p)
	{ CK (cudaMemcpyAsync (nSub.data (), c->blkNSub.p, 4 * (size_t) nbOld, cudaMemcpyDeviceToHost, s)) ;
	  CK (cudaMemcpyAsync (ptmOld.data (), c->blkPtm.p, 8 * (size_t) nbOld, cudaMemcpyDeviceToHost, s)) ;
	}
      CK (cudaMemcpyAsync (nReadOld.data (), c->blkNRead.p, 4 * (size_t) nbOld, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      uint64_t add = 0 ;
      for (uint32_t i = 0 ; i < nbOld ; ++i) { base[i] = (uint32_t) (nbOld + add) ; add += nSub[i] ; }
      if ((uint64_t) nbOld + add >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 barcode blocks after the split") ;
      const uint32_t nbNew = (uint32_t) (nbOld + add) ;
      std::vector<uint32_t> parent (nbNew, 0), nReadNew (nbNew, 0) ;
      std::vector<double> ptmNew (nbNew, 0.0) ;
      for (uint32_t i = 0 ; i < nbOld ; ++i)
	{ nReadNew[i] = nReadOld[i] ;				/* :991 for a split block, :998 otherwise */
	  if (!nSub[i]) ptmNew[i] = ptmOld[i] ;			/* a split block starts from a zeroed ClusterBlock */
	  for (uint32_t j = 0 ; j < nSub[i] ; ++j) parent[base[i] + j] = i + 1 ;	/* :976 */
	}
      /* everything that stays is allocated first and swapped in at the end: a full slab leaves the old index intact */
      DBuf<uint32_t> dNSub (nbOld, s, mt), dBase (nbOld, s, mt), newNHash ((size_t) nbNew + 1, s, mt), newNRead (nbNew, s, mt) ;
      DBuf<uint32_t> newParent (nbNew, s, mt), newNSub (nbNew, s, mt) ; DBuf<double> newPtm (nbNew, s, mt) ;
      DBuf<uint64_t> newOff ((size_t) nbNew + 1, s, mt) ;
      DBuf<uint64_t> newClus (H, s, mt) ;
      DBuf<unsigned int> ticket (1, s, mt) ;
      uint32_t maxRead = 1 ;
      for (uint32_t i = 0 ; i < nbOld ; ++i) maxRead = std::max (maxRead, nReadOld[i]) ;
      const uint32_t tblSize = std::min<uint32_t> (65536u, maxRead) ;
      const int nSM = device_sms (c) ;
      int occ = 1 ;
      CK (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&occ, k_split<true>, H10X_SPLIT_WARPS * 32, 0)) ;
      const uint32_t grid = (uint32_t) std::min<uint64_t> ((uint64_t) nSM * std::max (occ, 1), ((uint64_t) nbOld + H10X_SPLIT_WARPS - 1) / H10X_SPLIT_WARPS) ;
      /* the temporaries (per-warp read tables, the (bin, block) words of the transposition) come from the slab when it has
	 room, from one driver allocation otherwise (a context sized for a small build) */
      const size_t nTbl = (size_t) grid * H10X_SPLIT_WARPS * tblSize ;
      MemTrack *tm = mt ;
      if (mt->largestFree () < 8 * nTbl + 16 * (size_t) H + (1u << 20) + (mt->cap - mt->cur) / 4)
	{ const size_t bytes = 8 * nTbl + 16 * (size_t) H + (4u << 20) ;
	  CK (cudaMalloc (&scratch, bytes)) ;
	  scratchMt.base = (char*) scratch ; scratchMt.cap = bytes ; scratchMt.reset () ;
	  tm = &scratchMt ;
	}
      DBuf<uint32_t> first (nTbl, s, tm), number (nTbl, s, tm) ;
      DBuf<uint64_t> idBlock (H, s, tm), idBlock2 (H, s, tm) ;
      CK (cudaMemcpyAsync (dNSub.p, nSub.data (), 4 * (size_t) nbOld, cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemcpyAsync (dBase.p, base.data (), 4 * (size_t) nbOld, cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemcpyAsync (newNRead.p, nReadNew.data (), 4 * (size_t) nbNew, cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemcpyAsync (newParent.p, parent.data (), 4 * (size_t) nbNew, cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemcpyAsync (newPtm.p, ptmNew.data (), 8 * (size_t) nbNew, cudaMemcpyHostToDevice, s)) ;
      CK (cudaMemsetAsync (newNHash.p, 0, 4 * ((size_t) nbNew + 1), s)) ;
      CK (cudaMemsetAsync (newNSub.p, 0, 4 * (size_t) nbNew, s)) ;
      SplitArgs a ; memset (&a, 0, sizeof (a)) ;
      a.clus = (const unsigned long long*) c->clus.p ; a.blkOff = c->blkOff.p ; a.blkNHash = c->blkNHash.p ; a.nSub = dNSub.p ;
      a.clusterBase = dBase.p ; a.nBlocksOld = nbOld ; a.first = first.p ; a.number = number.p ; a.tblSize = tblSize ;
      a.newNHash = newNHash.p ; a.newNRead = newNRead.p ; a.newOff = newOff.p ; a.newClus = (unsigned long long*) newClus.p ;
      a.idBlock = idBlock.p ; a.ticket = ticket.p ;
      CK (cudaMemsetAsync (ticket.p, 0, 4, s)) ;
      LAUNCH (c, k_split<false>, grid, H10X_SPLIT_WARPS * 32, 0, s, a) ;
      cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> nh64 (newNHash.p, CastU64 ()) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, nh64, newOff.p, (size_t) nbNew + 1, s) ; }) ;
      CK (cudaMemsetAsync (ticket.p, 0, 4, s)) ;
      LAUNCH (c, k_split<true>, grid, H10X_SPLIT_WARPS * 32, 0, s, a) ;
      /* fillHashTable again (:1012): stable passes on the bin id turn the block-major (bin, new block) words bin-major,
	 new blocks ascending inside a bin; hashDepth and therefore codeOff do not change */
      if (H)
	{ const int idBits = bits_for (c->hashNumber) ;
	  const int nP = (idBits + 9) / 10 ;
	  uint64_t *src = idBlock.p, *dst = idBlock2.p ;
	  for (int i = 0, left = idBits, shift = 32 ; i < nP ; ++i)
	    { const int bits = (left + (nP - i) - 1) / (nP - i) ;
	      LoadWord lw = { src, shift, (1u << bits) - 1u, ~(uint64_t) 0 } ;
	      part_pass (c, s, lw, H, 1u << bits, dst) ;
	      shift += bits ; left -= bits ; std::swap (src, dst) ;
	    }
	  LAUNCH (c, k_split_codes, gridFor (H, 256), 256, 0, s, H, src, c->codes.p) ;
	}
      CK (cudaStreamSynchronize (s)) ;
      /* the new index replaces the old one (the old arrays leave with the locals) */
      c->clus.swap (newClus) ; c->blkOff.swap (newOff) ; c->blkNHash.swap (newNHash) ; c->blkNRead.swap (newNRead) ;
      c->blkNSub.swap (newNSub) ; c->blkPtm.swap (newPtm) ; c->blkParent.swap (newParent) ;
      c->goodOffD.release () ; c->goodD.release () ; c->haveGood = false ;
      c->nBlocksMax = nbNew ;
      *nNew = (uint32_t) add ;
    }) ;
  if (scratch) { cudaStreamSynchronize (c->own) ; cudaFree (scratch) ; }
  if (st != H10X_OK) return st ;
  st = h10x_gpu_download (c, out, err, errlen) ;
  if (st != H10X_OK) return st ;
  return guarded (err, errlen, [&] ()
    { /* the ClusterBlock fields beside the block table, in the context's pinned cluster slots */
      const size_t nb = c->nBlocksMax ;
      auto pull = [&] (int slot, const void *src, size_t bytes) -> void*
	{ if (c->clusCap[slot] < bytes || !c->clusSlot[slot])
	    { if (c->clusSlot[slot]) cudaFreeHost (c->clusSlot[slot]) ;
	      c->clusSlot[slot] = nullptr ; c->clusCap[slot] = 0 ;
	      c->clusSlot[slot] = pinned_alloc (bytes) ; c->clusCap[slot] = bytes ? bytes : 1 ;
	    }
	  if (bytes) CK (cudaMemcpyAsync (c->clusSlot[slot], src, bytes, cudaMemcpyDeviceToHost, c->own)) ;
	  return c->clusSlot[slot] ;
	} ;
      out->blkNSubCluster = (uint32_t*) pull (0, c->blkNSub.p, 4 * nb) ;
      out->blkPointToMin = (double*) pull (1, c->blkPtm.p, 8 * nb) ;
      out->blkClusterParent = (uint32_t*) pull (2, c->blkParent.p, 4 * nb) ;
      out->reserved |= H10X_INDEX_EXACT_BLOCKS ;
      CK (cudaStreamSynchronize (c->own)) ;
    }) ;
}

int h10x_gpu_stats (h10x_ctx *c, h10x_stats *out)
{ if (!c || !out) return H10X_ERR_BAD_PARAM ; *out = c->stats ; return H10X_OK ; }

/* position-salted sum digests (h10x_digest.h) of the resident index, computed where it lies; and the check that
   codes[] is exactly the transposition of the ClusterHash lists (fillHashTable, hash10x.c:317-347) */
int h10x_gpu_index_digest (h10x_ctx *c, uint64_t blockBase, uint64_t entryBase, int withBlockZero, h10x_digest *out,
			   char *err, size_t errlen)
{ if (!c || !out || !c->haveIndex) { set_err (err, errlen, "no index resident") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      CK (cudaStreamSynchronize (s)) ;
      const size_t hn = c->hashNumber, nb = c->nBlocksMax, H = c->nHashes ;
      unsigned long long *acc = nullptr ;
      CK (cudaMalloc ((void**) &acc, 16 * sizeof (unsigned long long))) ;
      struct Free { void *p ; ~Free () { cudaFree (p) ; } } freer { acc } ;
      CK (cudaMemsetAsync (acc, 0, 16 * sizeof (unsigned long long), s)) ;
      const int grid = 148 * 8 ;
      const uint64_t all = ~(uint64_t) 0 ;
      if (c->hashIndex.p) k_digest<uint32_t><<<grid, 256, 0, s>>> (c->hashIndex.p, (uint64_t) 1 << c->P.B, 0, all, acc + 0) ;
      if (c->hashValue.p) k_digest<uint64_t><<<grid, 256, 0, s>>> (c->hashValue.p, hn, 0, all, acc + 1) ;
      if (c->hashDepth.p) k_digest<uint32_t><<<grid, 256, 0, s>>> (c->hashDepth.p, hn, 0, all, acc + 2) ;
      const size_t i0 = withBlockZero ? 0 : 1 ;
      if (nb > i0)
	{ k_digest<uint32_t><<<grid, 256, 0, s>>> (c->blkNRead.p + i0, nb - i0, blockBase + i0, all, acc + 3) ;
	  k_digest<uint32_t><<<grid, 256, 0, s>>> (c->blkNHash.p + i0, nb - i0, blockBase + i0, all, acc + 4) ;
	}
      if (H) k_digest<uint64_t><<<grid, 256, 0, s>>> (c->clus.p, H, entryBase, H10X_DG_CLUS_MASK, acc + 5) ;
      if (c->codes.p && c->codeOff.p && !c->dist)
	{ if (H) k_digest<uint32_t><<<grid, 256, 0, s>>> (c->codes.p, H, 0, all, acc + 6) ;
	  k_digest<uint64_t><<<grid, 256, 0, s>>> (c->codeOff.p, hn + 1, 0, all, acc + 7) ;
	  k_check_codes<<<(unsigned) std::min<size_t> (nb, 148 * 16), 256, 0, s>>> ((uint32_t) nb, c->blkOff.p, c->blkNHash.p, c->clus.p, c->codeOff.p, c->codes.p, acc + 8) ;
	  if (H > 1) k_check_ascending<<<grid, 256, 0, s>>> (c->codeOff.p, c->codes.p, c->hashDepth.p, (uint32_t) hn, acc + 9) ;
	  out->haveCodes = 1 ;
	}
      CK (cudaGetLastError ()) ;
      unsigned long long h[16] ;
      CK (cudaMemcpyAsync (h, acc, sizeof (h), cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      out->hashIndex = h[0] ; out->hashValue = h[1] ; out->hashDepth = h[2] ; out->blkNRead = h[3] ; out->blkNHash = h[4] ;
      out->clusHash = h[5] ; out->codes = h[6] ; out->codeOff = h[7] ; out->codesMissing = h[8] ; out->codesUnordered = h[9] ;
      out->haveTable = c->hashIndex.p ? 1 : 0 ; out->haveBins = c->hashValue.p ? 1 : 0 ;
    }) ;
}

void *h10x_host_alloc (size_t bytes)
{ void *p = nullptr ; if (cudaHostAlloc (&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError () ; return nullptr ; } return p ; }

void h10x_host_free (void *p) { if (p) cudaFreeHost (p) ; }

/* thr[o] = first hash of owner o's range, thr[nranks] = 2^(2k): see dist_bins ("owner = hash range") */
int h10x_dist_owner_thresholds (int k, int nranks, int flat, uint64_t *thr)
{ if (!thr || nranks < 1 || k < 1 || k > 31) return H10X_ERR_BAD_PARAM ;
  for (int o = 0 ; o <= nranks ; ++o)
    { if (flat || o == 0 || o == nranks)
	{ unsigned __int128 t = ((unsigned __int128) o << (2 * k)) + (unsigned) (nranks - 1) ; thr[o] = (uint64_t) (t / (unsigned) nranks) ; }
      else
	{ const long double x = 1.0L - sqrtl (1.0L - (long double) o / (long double) nranks) ;
	  thr[o] = (uint64_t) (x * (long double) ((uint64_t) 1 << (2 * k))) ;	/* same bits on every rank: same code, same CPU */
	  if (thr[o] < thr[o-1]) thr[o] = thr[o-1] ;
	}
    }
  return H10X_OK ;
}

int h10x_dist_unique_id (void *id128, char *err, size_t errlen)
{ if (!id128) return H10X_ERR_BAD_PARAM ;
  return guarded (err, errlen, [&] ()
    { std::string why ;
      if (!gNccl.load (why)) throw H10xError (H10X_ERR_UNSUPPORTED, why) ;
      ncclUniqueId id ;
      NCK (gNccl.GetUniqueId (&id)) ;
      memcpy (id128, &id, sizeof (id)) ;
    }) ;
}

int h10x_dist_init (h10x_ctx *c, int rank, int nranks, const void *id128, char *err, size_t errlen)
{ if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_err (err, errlen, "bad rank / nranks") ; return H10X_ERR_BAD_PARAM ; }
  return guarded (err, errlen, [&] ()
    { std::string why ;
      if (!gNccl.load (why)) throw H10xError (H10X_ERR_UNSUPPORTED, why) ;
      CK (cudaSetDevice (c->P.device)) ;
      if (!c->dist) c->dist = new DistState () ;
      if (c->dist->comm) { gNccl.CommDestroy (c->dist->comm) ; c->dist->comm = nullptr ; }
      ncclUniqueId id ; memcpy (&id, id128, sizeof (id)) ;
      NCK (gNccl.CommInitRank (&c->dist->comm, nranks, id, rank)) ;
      c->dist->rank = rank ; c->dist->nranks = nranks ;
    }) ;
}

int h10x_gpu_dist_global_codes (h10x_ctx *c, char *err, size_t errlen)
{ if (!c || !c->dist || !c->dist->comm || !c->haveIndex) { set_err (err, errlen, "no distributed index resident") ; return H10X_ERR_BAD_PARAM ; }
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      MemTrack *mt = &c->mt ;
      DistState *d = c->dist ;
      const int R = d->rank, NR = d->nranks ;
      const size_t hn = c->hashNumber ;
      /* local allocations first, then one agreement: nobody enters a broadcast that a peer cannot follow */
      int localErr = 0 ;
      DBuf<uint32_t> fill ;
      try
	{ if (R != 0 || !c->hashDepth.p) c->hashDepth.alloc (hn + 1, s, mt) ;
	  c->codeOff.alloc (hn + 1, s, mt) ; fill.alloc (hn, s, mt) ;
	}
      catch (const SlabFull&) { localErr = H10X_ERR_NOMEM ; }
      const uint32_t mine[4] = { d->nLocalBins, (uint32_t) c->nHashes, 0, 0 } ;
      std::vector<uint32_t> all ;
      dist_agree (c, s, localErr, mine, all) ;
      uint64_t Hg = 0 ; uint32_t maxBins = 0, maxH = 0 ;
      for (int r = 0 ; r < NR ; ++r) { Hg += all[4*r + 1] ; maxBins = std::max (maxBins, all[4*r]) ; maxH = std::max (maxH, all[4*r + 1]) ; }
      if (Hg >= 0xffffffffull) throw H10xError (H10X_ERR_UNSUPPORTED, "more than 2^32-2 (block, hash) pairs over all ranks") ;
      DBuf<uint32_t> tBin, tOff, tCodes ;
      try { c->codes.alloc (Hg, s, mt) ; tBin.alloc (maxBins, s, mt) ; tOff.alloc ((size_t) maxBins + 1, s, mt) ; tCodes.alloc (maxH, s, mt) ; }
      catch (const SlabFull&) { localErr = H10X_ERR_NOMEM ; }
      dist_agree (c, s, localErr, mine, all) ;
      /* depths from rank 0, offsets by the same scan as a single-GPU build */
      NCK (gNccl.Broadcast (c->hashDepth.p, c->hashDepth.p, hn, ncclUint32, 0, d->comm, s)) ;
      CK (cudaMemsetAsync (c->hashDepth.p + hn, 0, 4, s)) ;
      cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> depth64 (c->hashDepth.p, CastU64 ()) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, depth64, c->codeOff.p, hn + 1, s) ; }) ;
      CK (cudaMemsetAsync (fill.p, 0, 4 * hn, s)) ;
      for (int r = 0 ; r < NR ; ++r)
	{ const uint32_t nb = all[4*r], nh = all[4*r + 1] ;
	  const uint32_t *bin = (r == R) ? c->localBinId.p : tBin.p, *off = (r == R) ? c->localCodeOff.p : tOff.p,
	    *piece = (r == R) ? c->localCodes.p : tCodes.p ;
	  NCK (gNccl.GroupStart ()) ;
	  if (nb) NCK (gNccl.Broadcast (bin, (void*) bin, nb, ncclUint32, r, d->comm, s)) ;
	  NCK (gNccl.Broadcast (off, (void*) off, (size_t) nb + 1, ncclUint32, r, d->comm, s)) ;
	  if (nh) NCK (gNccl.Broadcast (piece, (void*) piece, nh, ncclUint32, r, d->comm, s)) ;
	  NCK (gNccl.GroupEnd ()) ;
	  if (nb) LAUNCH (c, k_place_piece, gridFor ((uint64_t) nb * 32, 256), 256, 0, s, nb, bin, off, piece, c->codeOff.p, fill.p, c->codes.p) ;
	}
      CK (cudaStreamSynchronize (s)) ;
      d->globalCodes = true ;
    }) ;
}

/* a rank-local failure between two agreements of a distributed build: hand the error to the peers, who are (or will
   be) waiting at the next one.  Not after a CUDA error - the context cannot run a collective any more. */
static void dist_fail_agree (h10x_ctx *c, cudaStream_t s, int code)
{ if (!c->dist || !c->dist->owesAgreement || code == H10X_ERR_CUDA) return ;
  c->dist->owesAgreement = false ;
  try { const uint32_t none[4] = { 0, 0, 0, 0 } ; std::vector<uint32_t> all ; dist_agree (c, s, code, none, all) ; }
  catch (...) {}
}

int h10x_gpu_build_device_dist (h10x_ctx *c, const void *d_fqb, uint64_t nRecords, void *stream, char *err, size_t errlen)
{ if (!c || !c->dist || !c->dist->comm) { set_err (err, errlen, "h10x_dist_init has not been called") ; return H10X_ERR_BAD_PARAM ; }
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = stream ? (cudaStream_t) stream : c->own ;
      /* the slab is sized generously up front: a retry after SlabFull would have to be collective */
      size_t est = slab_estimate (c->P, nRecords, false) ;
      est += est / 2 + ((size_t) 8 * c->dist->nranks << 22) ;
      if (c->mt.cap < est && !c->slabClamped) { reset_result (c) ; slab_resize (c, est) ; }
      try { build_device_impl (c, (const uint32_t*) d_fqb, nRecords, s, true, true) ; }
      catch (const SlabFull &f)
	{ dist_fail_agree (c, s, H10X_ERR_NOMEM) ;
	  throw H10xError (H10X_ERR_NOMEM, "device workspace too small in a distributed build (need " + std::to_string (f.need) + " bytes)") ;
	}
      catch (const H10xError &e) { dist_fail_agree (c, s, e.code) ; throw ; }
    }) ;
  if (st != H10X_OK) { cudaStreamSynchronize (stream ? (cudaStream_t) stream : c->own) ; cudaGetLastError () ; c->haveIndex = false ; }
  return st ;
}

int h10x_gpu_build_host_dist (h10x_ctx *c, const void *fqb, uint64_t nRecords, h10x_index *out, char *err, size_t errlen)
{ if (!c || !out || (!fqb && nRecords) || !c->dist || !c->dist->comm) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  int st = guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      size_t est = slab_estimate (c->P, nRecords, true) ;
      est += est / 2 + ((size_t) 8 * c->dist->nranks << 22) ;
      if (c->mt.cap < est && !c->slabClamped) { reset_result (c) ; slab_resize (c, est) ; }
      reset_result (c) ;
      try
	{ DBuf<uint32_t> d ((size_t) nRecords * H10X_REC_WORDS, s, &c->mt) ;
	  if (nRecords) CK (cudaMemcpyAsync (d.p, fqb, (size_t) nRecords * 120, cudaMemcpyHostToDevice, s)) ;
	  build_device_impl (c, d.p, nRecords, s, false, true) ;
	}
      catch (const SlabFull &f)
	{ dist_fail_agree (c, s, H10X_ERR_NOMEM) ;
	  throw H10xError (H10X_ERR_NOMEM, "device workspace too small in a distributed build (need " + std::to_string (f.need) + " bytes)") ;
	}
      catch (const H10xError &e) { dist_fail_agree (c, s, e.code) ; throw ; }
    }) ;
  if (st != H10X_OK) { cudaStreamSynchronize (c->own) ; cudaGetLastError () ; c->haveIndex = false ; return st ; }
  return h10x_gpu_download (c, out, err, errlen) ;
}

/* ---- the whole seam on several GPUs of one node, from one process: one thread + one context per GPU ----
   The file is cut into nGpus record ranges at barcode-run boundaries; every thread reads its range, joins
   the NCCL communicator and runs the distributed build; the per-rank pieces are then stitched into one
   host index (block table and ClusterHash slabs concatenate in rank order; rank 0 holds the bin table).
   The hash->code lists stay distributed (codes / codeOff are NULL): --writeHash, --hashStats, --codeStats
   and --hashDepthRange do not read them, and readHashFile rebuilds them anyway (hash10x.c:269-315). */

/* ------------------------------------------------------------------ a multi-GPU session: build, then the commands that follow */

struct h10x_multi {
  std::vector<h10x_ctx*> ctxs ;		/* one per GPU, rank order = block order */
  std::vector<uint32_t> nbLocal ;	/* nBlocksMax of every rank (its blocks are 1 .. nbLocal - 1) */
  std::vector<uint64_t> hLocal ;
  uint32_t nb = 0, hn = 0 ; uint64_t H = 0 ;
  /* stitched results, owned by the session */
  std::vector<uint8_t> within ; std::vector<uint64_t> goodOff ; std::vector<uint16_t> good ;
  std::vector<uint64_t> clus ; std::vector<uint32_t> nSub ; std::vector<double> ptm ;
} ;

static int multi_build_file (const h10x_params *p, int nGpus, const char *path, h10x_index *out, h10x_multi **session, char *err, size_t errlen)
{ if (!p || !path || !out || nGpus < 1) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  int nDev = h10x_gpu_device_count () ;
  if (nDev <= 0) { set_err (err, errlen, h10x_strerror (H10X_ERR_NO_DEVICE)) ; return H10X_ERR_NO_DEVICE ; }
  if (nGpus > nDev) { set_err (err, errlen, "more GPUs requested than visible") ; return H10X_ERR_BAD_PARAM ; }
  if (p->N != 0) { set_err (err, errlen, "-N is not supported with several GPUs") ; return H10X_ERR_UNSUPPORTED ; }
  FILE *f = fopen (path, "rb") ;
  if (!f) { set_err (err, errlen, "failed to open fqb file") ; return H10X_ERR_IO ; }
  if (fseeko (f, 0, SEEK_END)) { fclose (f) ; set_err (err, errlen, "file read problem") ; return H10X_ERR_IO ; }
  const uint64_t nRec = (uint64_t) ftello (f) / 120 ;
  /* cuts: the first record at or after r*nRec/nGpus whose barcode differs from its predecessor's */
  std::vector<uint64_t> cut ((size_t) nGpus + 1, 0) ;
  cut[nGpus] = nRec ;
  bool ioOk = true ;
  for (int r = 1 ; r < nGpus && ioOk ; ++r)
    { uint64_t pos = std::max<uint64_t> (nRec * (uint64_t) r / (uint64_t) nGpus, cut[r-1] + 1) ;
      uint32_t prev = 0, cur = 0 ;
      if (pos >= nRec) { cut[r] = nRec ; continue ; }
      if (fseeko (f, (off_t) ((pos - 1) * 120), SEEK_SET) || fread (&prev, 4, 1, f) != 1) { ioOk = false ; break ; }
      for ( ; pos < nRec ; ++pos)
	{ if (fseeko (f, (off_t) (pos * 120), SEEK_SET) || fread (&cur, 4, 1, f) != 1) { ioOk = false ; break ; }
	  if (cur != prev) break ;
	}
      cut[r] = pos ;
    }
  fclose (f) ;
  if (!ioOk) { set_err (err, errlen, "file read problem") ; return H10X_ERR_IO ; }
  int used = nGpus ;
  while (used > 1 && cut[used-1] >= nRec) --used ;		/* fewer runs than GPUs: drop the empty tail ranks */
  cut[used] = nRec ;

  unsigned char id[H10X_DIST_ID_BYTES] ;
  char e0[512] = "" ;
  int st0 = (used > 1 || true) ? h10x_dist_unique_id (id, e0, sizeof (e0)) : H10X_OK ;
  if (st0) { set_err (err, errlen, e0) ; return st0 ; }

  std::vector<h10x_ctx*> ctxs (used, nullptr) ;
  std::vector<int> status (used, H10X_OK) ;
  std::vector<std::string> msgs (used) ;
  std::vector<h10x_index> parts (used) ;
  std::vector<uint32_t*> dFqb (used, nullptr) ;
  std::atomic<int> arrived (0), failedEarly (0) ;
  auto worker = [&] (int r)
    { char e[512] = "" ;
      h10x_params pr = *p ; pr.device = p->device + r ;
      uint64_t n = cut[r+1] - cut[r] ;
      void *stage[2] = { nullptr, nullptr } ;
      int st = H10X_OK ;
      ctxs[r] = h10x_gpu_create (&pr, e, sizeof (e)) ;
      if (!ctxs[r]) st = H10X_ERR_BAD_PARAM ;
      if (!st) st = h10x_dist_init (ctxs[r], r, used, id, e, sizeof (e)) ;
      if (!st)
	st = guarded (e, sizeof (e), [&] ()
	  { h10x_ctx *c = ctxs[r] ;
	    CK (cudaSetDevice (c->P.device)) ;
	    FILE *fr = fopen (path, "rb") ;
	    if (!fr) throw H10xError (H10X_ERR_IO, "failed to open fqb file") ;
	    struct Closer { FILE *f ; ~Closer () { fclose (f) ; } } closer { fr } ;
	    if (fseeko (fr, (off_t) (cut[r] * 120), SEEK_SET)) throw H10xError (H10X_ERR_IO, "file read problem") ;
	    CK (cudaMalloc ((void**) &dFqb[r], std::max<uint64_t> (n * 120, 16))) ;
	    const size_t chunkRecs = 1u << 19 ;
	    stage[0] = pinned_alloc (chunkRecs * 120) ; stage[1] = pinned_alloc (chunkRecs * 120) ;
	    cudaEvent_t done[2] ; CK (cudaEventCreate (&done[0])) ; CK (cudaEventCreate (&done[1])) ;
	    uint64_t pos = 0 ; int cur = 0 ; bool usedBuf[2] = { false, false } ;
	    while (pos < n)
	      { size_t want = (size_t) std::min<uint64_t> (chunkRecs, n - pos) ;
		if (usedBuf[cur]) CK (cudaEventSynchronize (done[cur])) ;
		if (fread (stage[cur], 120, want, fr) != want) throw H10xError (H10X_ERR_IO, "file read problem") ;
		CK (cudaMemcpyAsync ((char*) dFqb[r] + pos * 120, stage[cur], want * 120, cudaMemcpyHostToDevice, c->own)) ;
		CK (cudaEventRecord (done[cur], c->own)) ; usedBuf[cur] = true ;
		pos += want ; cur ^= 1 ;
	      }
	    CK (cudaStreamSynchronize (c->own)) ;
	    cudaEventDestroy (done[0]) ; cudaEventDestroy (done[1]) ;
	  }) ;
      if (stage[0]) cudaFreeHost (stage[0]) ;
      if (stage[1]) cudaFreeHost (stage[1]) ;
      /* nobody enters the collective build unless every rank got this far */
      if (st) failedEarly.fetch_add (1) ;
      arrived.fetch_add (1) ;
      while (arrived.load () < used) std::this_thread::yield () ;
      if (!st && failedEarly.load () == 0)
	{ st = h10x_gpu_build_device_dist (ctxs[r], dFqb[r], n, nullptr, e, sizeof (e)) ;
	  if (!st) st = h10x_gpu_download (ctxs[r], &parts[r], e, sizeof (e)) ;
	  /* a session goes on to --hashDepthRange / --cluster: every rank needs the depths and the whole hash->code CSR */
	  if (!st && session) st = h10x_gpu_dist_global_codes (ctxs[r], e, sizeof (e)) ;
	}
      else if (!st) { st = H10X_ERR_IO ; snprintf (e, sizeof (e), "another rank failed before the build") ; }
      status[r] = st ; msgs[r] = e ;
    } ;
  std::vector<std::thread> th ;
  for (int r = 0 ; r < used ; ++r) th.emplace_back (worker, r) ;
  for (auto &t : th) t.join () ;

  int st = H10X_OK ;
  for (int r = 0 ; r < used && !st ; ++r) if (status[r]) { st = status[r] ; set_err (err, errlen, msgs[r].c_str ()) ; }
  if (!st)
    st = guarded (err, errlen, [&] ()
      { /* stitch: bin table from rank 0, block table + ClusterHash slabs in rank order */
	uint64_t H = 0, reads = 0 ; uint32_t nb = 1 ;
	for (int r = 0 ; r < used ; ++r) { H += parts[r].nHashes ; reads += parts[r].nReads ; nb += parts[r].nBlocksMax - 1 ; }
	const h10x_index &z = parts[0] ;
	out->B = z.B ; out->hashNumber = z.hashNumber ; out->nBlocksMax = nb ; out->nReads = reads ; out->nHashes = H ;
	auto dup = [] (const void *src, size_t bytes) -> void*
	  { void *d = malloc (bytes ? bytes : 1) ; if (!d) throw std::bad_alloc () ; if (bytes) memcpy (d, src, bytes) ; return d ; } ;
	if (z.hashIndex) out->hashIndex = (uint32_t*) dup (z.hashIndex, (size_t) 4 << z.B) ;
	out->hashValue = (uint64_t*) dup (z.hashValue, 8 * (size_t) z.hashNumber) ;
	out->hashDepth = (uint32_t*) dup (z.hashDepth, 4 * (size_t) z.hashNumber) ;
	out->blkNRead = (uint32_t*) calloc (nb, 4) ; out->blkNHash = (uint32_t*) calloc (nb, 4) ;
	out->blkOff = (uint64_t*) calloc ((size_t) nb + 1, 8) ;
	out->clusHash = (h10x_cluster_hash*) malloc ((H ? H : 1) * 8) ;
	if (!out->blkNRead || !out->blkNHash || !out->blkOff || !out->clusHash) throw std::bad_alloc () ;
	uint32_t b = 1 ; uint64_t e = 0 ;
	for (int r = 0 ; r < used ; ++r)
	  { const h10x_index &q = parts[r] ;
	    for (uint32_t i = 1 ; i < q.nBlocksMax ; ++i, ++b)
	      { out->blkNRead[b] = q.blkNRead[i] ; out->blkNHash[b] = q.blkNHash[i] ; out->blkOff[b] = e + q.blkOff[i] ; }
	    if (q.nHashes) memcpy (out->clusHash + e, q.clusHash, 8 * q.nHashes) ;
	    e += q.nHashes ;
	  }
	out->blkOff[nb] = H ;
	out->pinned = 0 ; out->onDevice = 0 ;
      }) ;
  for (int r = 0 ; r < used ; ++r) if (dFqb[r]) { cudaSetDevice (p->device + r) ; cudaFree (dFqb[r]) ; }
  if (!st && session)
    { h10x_multi *m = new h10x_multi ;
      m->ctxs = ctxs ; m->nb = out->nBlocksMax ; m->H = out->nHashes ; m->hn = out->hashNumber ;
      for (int r = 0 ; r < used ; ++r) { m->nbLocal.push_back (parts[r].nBlocksMax) ; m->hLocal.push_back (parts[r].nHashes) ; }
      *session = m ;
    }
  else
    for (int r = 0 ; r < used ; ++r) if (ctxs[r]) h10x_gpu_destroy (ctxs[r]) ;
  if (st) { h10x_index_free (out) ; memset (out, 0, sizeof (*out)) ; }
  return st ;
}


int h10x_gpu_build_file_multi (const h10x_params *p, int nGpus, const char *path, h10x_index *out, char *err, size_t errlen)
{ return multi_build_file (p, nGpus, path, out, nullptr, err, errlen) ; }

int h10x_multi_build_file (const h10x_params *p, int nGpus, const char *path, h10x_multi **session, h10x_index *out,
			   char *err, size_t errlen)
{ if (!session) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  *session = nullptr ;
  return multi_build_file (p, nGpus, path, out, session, err, errlen) ;
}

void h10x_multi_destroy (h10x_multi *m)
{ if (!m) return ;
  for (h10x_ctx *c : m->ctxs) if (c) h10x_gpu_destroy (c) ;
  delete m ;
}

/* one thread per rank; the first failure is reported */
static int multi_each (h10x_multi *m, char *err, size_t errlen, const std::function<int (int, char*, size_t)> &f)
{ const int n = (int) m->ctxs.size () ;
  std::vector<int> status (n, H10X_OK) ; std::vector<std::string> msgs (n) ;
  std::vector<std::thread> th ;
  for (int r = 0 ; r < n ; ++r)
    th.emplace_back ([&, r] () { char e[512] = "" ; status[r] = f (r, e, sizeof (e)) ; msgs[r] = e ; }) ;
  for (auto &t : th) t.join () ;
  for (int r = 0 ; r < n ; ++r) if (status[r]) { set_err (err, errlen, msgs[r].c_str ()) ; return status[r] ; }
  return H10X_OK ;
}

int h10x_multi_depth_range (h10x_multi *m, int dmin, int dmax, h10x_good_hashes *out, char *err, size_t errlen)
{ if (!m || !out) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  const int n = (int) m->ctxs.size () ;
  std::vector<h10x_good_hashes> part (n) ;
  int st = multi_each (m, err, errlen, [&] (int r, char *e, size_t el) { return h10x_gpu_depth_range (m->ctxs[r], dmin, dmax, &part[r], e, el) ; }) ;
  if (st) return st ;
  return guarded (err, errlen, [&] ()
    { m->within.assign (part[0].within, part[0].within + m->hn) ;	/* the same on every rank: depths are global */
      m->goodOff.assign ((size_t) m->nb + 1, 0) ; m->good.clear () ;
      uint32_t b = 1 ;
      for (int r = 0 ; r < n ; ++r)
	{ const h10x_good_hashes &g = part[r] ;
	  for (uint32_t i = 1 ; i < m->nbLocal[r] ; ++i, ++b) m->goodOff[b] = m->good.size () + (g.goodOff[i] - g.goodOff[1]) ;
	  m->good.insert (m->good.end (), g.good + g.goodOff[1], g.good + g.goodOff[m->nbLocal[r]]) ;
	}
      m->goodOff[0] = 0 ; m->goodOff[m->nb] = m->good.size () ;
      out->within = m->within.data () ; out->goodOff = m->goodOff.data () ; out->good = m->good.data () ;
      out->nGood = m->good.size () ; out->hashNumber = m->hn ; out->nBlocksMax = m->nb ;
    }) ;
}

int h10x_multi_cluster (h10x_multi *m, int codeMin, int codeMax, int clusterThreshold, h10x_clusters *out, char *err, size_t errlen)
{ if (!m || !out) { set_err (err, errlen, "bad argument") ; return H10X_ERR_BAD_PARAM ; }
  memset (out, 0, sizeof (*out)) ;
  if (!codeMin) codeMin = 1 ;
  if (!codeMax) codeMax = (int) m->nb ;
  if (codeMin < 1 || codeMax > (int) m->nb) { set_err (err, errlen, "code range outside the barcode blocks") ; return H10X_ERR_BAD_PARAM ; }
  const int n = (int) m->ctxs.size () ;
  std::vector<h10x_clusters> part (n) ;
  std::vector<uint32_t> base (n + 1, 0) ;
  for (int r = 0 ; r < n ; ++r) base[r + 1] = base[r] + m->nbLocal[r] - 1 ;
  int st = multi_each (m, err, errlen, [&] (int r, char *e, size_t el)
    { /* global blocks base[r] + 1 .. base[r + 1] are this rank's blocks 1 .. nbLocal - 1; an empty local range still
	 returns the rank's arrays (codeMax == codeMin launches nothing) */
      long lo = std::max<long> ((long) codeMin - (long) base[r], 1), hi = std::min<long> ((long) codeMax - (long) base[r], (long) m->nbLocal[r]) ;
      if (hi < lo) hi = lo ;
      return h10x_gpu_cluster (m->ctxs[r], (int) lo, (int) hi, clusterThreshold, &part[r], e, el) ;
    }) ;
  if (st) return st ;
  return guarded (err, errlen, [&] ()
    { m->clus.resize (m->H ? m->H : 1) ; m->nSub.assign (m->nb, 0) ; m->ptm.assign (m->nb, 0.0) ;
      uint64_t e = 0 ; uint32_t b = 1 ; float ms = 0 ;
      for (int r = 0 ; r < n ; ++r)
	{ const h10x_clusters &q = part[r] ;
	  if (m->hLocal[r]) memcpy (m->clus.data () + e, q.clusHash, 8 * m->hLocal[r]) ;
	  e += m->hLocal[r] ;
	  for (uint32_t i = 1 ; i < m->nbLocal[r] ; ++i, ++b) { m->nSub[b] = q.nSubCluster[i] ; m->ptm[b] = q.pointToMin[i] ; }
	  ms = std::max (ms, (float) q.msKernel) ;
	}
      out->clusHash = (h10x_cluster_hash*) m->clus.data () ; out->nSubCluster = m->nSub.data () ; out->pointToMin = m->ptm.data () ;
      out->nBlocksMax = m->nb ; out->nHashes = m->H ; out->msKernel = ms ;
    }) ;
}

int h10x_gpu_dist_info (h10x_ctx *c, h10x_dist_info *out)
{ if (!c || !out || !c->dist) return H10X_ERR_BAD_PARAM ;
  memset (out, 0, sizeof (*out)) ;
  out->rank = c->dist->rank ; out->nranks = c->dist->nranks ;
  out->blockBase = c->dist->blockBase ; out->nBlocksGlobal = c->dist->nBlocksGlobal ;
  out->nReadsGlobal = c->dist->nReadsGlobal ; out->nHashesGlobal = c->dist->nHashesGlobal ;
  out->nLocalBins = c->dist->nLocalBins ;
  out->localBinId = c->localBinId.p ; out->localCodeOff = c->localCodeOff.p ; out->localCodes = c->localCodes.p ;
  return H10X_OK ;
}

int h10x_gpu_memcpy_d2h (h10x_ctx *c, void *dst, const void *src, size_t bytes)
{ if (!c || (!dst && bytes) || (!src && bytes)) return H10X_ERR_BAD_PARAM ;
  if (cudaSetDevice (c->P.device) != cudaSuccess) return H10X_ERR_CUDA ;
  if (bytes && cudaMemcpy (dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError () ; return H10X_ERR_CUDA ; }
  return H10X_OK ;
}

/* tests: what the FUSED kernel itself selected.  Builds the index of the host records and hands back, per processed
   block, the keys the fused kernel stored for it (hash, read index within the block) before any grouping: in lean mode
   (*lean = 1: the hand-written tail follows) that is every mosh of the block in no particular order, duplicates included
   - directly comparable with seqAddHashes / moshRCnext (hash10x.c:123-132, seqhash.c:154-195) record by record; in
   classic mode the sorted unique list.  A block the generic path took has outOff[b+1] == outOff[b]. */
int h10x_gpu_block_keys (h10x_ctx *c, const void *fqb, uint64_t nRecords, uint64_t *outOff, uint64_t *outHash, uint32_t *outRead,
			 uint64_t cap, uint32_t *nBlocks, int *lean, char *err, size_t errlen)
{ if (!c || !outOff || !nBlocks || (!fqb && nRecords)) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  h10x_index ix ;
  c->dbgKeys = true ;
  int st = h10x_gpu_build_host (c, fqb, nRecords, &ix, err, errlen) ;
  c->dbgKeys = false ;
  if (st != H10X_OK) return st ;
  uint64_t n = 0 ;
  outOff[0] = 0 ;
  for (uint32_t b = 0 ; b < c->dbgNBlk ; ++b)
    { const uint64_t so = c->dbgSrcOff[b] ;
      const uint32_t cnt = c->dbgBlkCnt[b] ;
      if (!(so >> 63) && cnt != 0xffffffffu)
	{ const int sh = (int) (so >> 56) ;
	  const uint64_t o = so & 0x00ffffffffffffffull, rmask = ((uint64_t) 1 << sh) - 1 ;
	  if (n + cnt > cap || o + cnt > c->dbgScratchLen) { set_err (err, errlen, "block_keys: output capacity too small") ; return H10X_ERR_BAD_PARAM ; }
	  for (uint32_t i = 0 ; i < cnt ; ++i)
	    { const uint64_t key = c->dbgScratch[o + i] ;
	      if (outHash) outHash[n] = key >> sh ;
	      if (outRead) outRead[n] = (uint32_t) (key & rmask) ;
	      ++n ;
	    }
	}
      outOff[b + 1] = n ;
    }
  *nBlocks = c->dbgNBlk ;
  if (lean) *lean = c->dbgLean ? 1 : 0 ;
  c->dbgSrcOff.clear () ; c->dbgScratch.clear () ; c->dbgBlkCnt.clear () ;
  return H10X_OK ;
}

int h10x_gpu_record_moshes (h10x_ctx *c, const void *fqb, uint64_t nRecords, uint64_t *outOff, uint64_t *outHash,
			    uint64_t cap, char *err, size_t errlen)
{ if (!c || !outOff || (!fqb && nRecords)) { set_err (err, errlen, "null argument") ; return H10X_ERR_BAD_PARAM ; }
  return guarded (err, errlen, [&] ()
    { CK (cudaSetDevice (c->P.device)) ;
      cudaStream_t s = c->own ;
      if (nRecords > (8u << 20)) throw H10xError (H10X_ERR_UNSUPPORTED, "record_moshes: at most 8M records per call") ;
      uint32_t nb = (uint32_t) nRecords ;
      outOff[0] = 0 ;
      if (!nb) return ;
      reset_result (c) ;
      size_t est = (size_t) nb * (120 + 24 + 237 * 12) + ((size_t) 16 << 20) ;
      if (c->mt.cap < est) slab_resize (c, est) ;
      DBuf<uint32_t> d ((size_t) nb * H10X_REC_WORDS, s, &c->mt) ;
      CK (cudaMemcpyAsync (d.p, fqb, (size_t) nb * 120, cudaMemcpyHostToDevice, s)) ;
      DBuf<uint32_t> cnt2 ((size_t) 2 * nb + 1, s, &c->mt), off2 ((size_t) 2 * nb + 1, s, &c->mt) ;
      DBuf<uint32_t> zero (1, s, &c->mt) ;
      CK (cudaMemsetAsync (cnt2.p + 2 * (size_t) nb, 0, 4, s)) ;
      CK (cudaMemsetAsync (zero.p, 0, 4, s)) ;
      LAUNCH (c, k_moshes<false>, gridFor (2 * (uint64_t) nb, 128), 128, 0, s, d.p, 0u, nb, c->hp, cnt2.p,
	      (const uint32_t*) nullptr, 0u, (const uint32_t*) nullptr, (uint64_t*) nullptr, (uint32_t*) nullptr) ;
      cubCall (c, s, [&] (void *t, size_t &b) { return cub::DeviceScan::ExclusiveSum (t, b, cnt2.p, off2.p, 2 * (size_t) nb + 1, s) ; }) ;
      std::vector<uint32_t> h ((size_t) 2 * nb + 1) ;
      CK (cudaMemcpyAsync (h.data (), off2.p, 4 * h.size (), cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
      uint32_t M = h[2 * (size_t) nb] ;
      for (uint32_t i = 0 ; i <= nb ; ++i) outOff[i] = h[2 * (size_t) i] ;
      if (M > cap) throw H10xError (H10X_ERR_BAD_PARAM, "record_moshes: output capacity too small") ;
      if (!M) return ;
      /* blkIncl = all ones and p0 = 0 make every record "block 0" whose phantom offset is 0 */
      DBuf<uint32_t> onesv (nb, s, &c->mt) ;
      std::vector<uint32_t> hv (nb, 1u) ;
      CK (cudaMemcpyAsync (onesv.p, hv.data (), 4 * (size_t) nb, cudaMemcpyHostToDevice, s)) ;
      DBuf<uint64_t> keys (M, s, &c->mt) ; DBuf<uint32_t> vals (M, s, &c->mt) ;
      LAUNCH (c, k_moshes<true>, gridFor (2 * (uint64_t) nb, 128), 128, 0, s, d.p, 0u, nb, c->hp, off2.p,
	      onesv.p, 0u, zero.p, keys.p, vals.p) ;
      CK (cudaMemcpyAsync (outHash, keys.p, 8 * (size_t) M, cudaMemcpyDeviceToHost, s)) ;
      CK (cudaStreamSynchronize (s)) ;
    }) ;
}

} /* extern "C" */
