/* h10x_dist.cuh - the multi-GPU part of the build (one process or thread per GPU, NCCL over NVLink).
 *
 * The reference has no distributed mode; this is new work.  Barcode blocks are
 * embarrassingly parallel up to the bin table, so:
 *   - rank r owns a contiguous range of barcode runs (the input is grouped by barcode); global block
 *     number = blockBase_r + local number; only the globally last run stays unhashed (hash10x.c:209,216);
 *   - each rank runs the fused kernel on its blocks and reduces its block-unique entries to
 *     RANK-DISTINCT hashes (hash, localDepth, localFirstBlock) - NVLink is ~9x below aggregate HBM, so
 *     only these (16 B per rank-distinct hash, not 12 B per entry) cross it;
 *   - all-to-all-v (ncclSend/ncclRecv in one group) to the hash-RANGE owner, owner(hash) monotone in
 *     hash; the owner merges: depth = sum, firstBlock = min;
 *   - bin ids in the reference's order (first block, then hash; hash10x.c:147): per-owner counts of
 *     "new hashes of block b" are all-gathered; id = 1 + sum_{b'<b} new[b'] + sum_{o'<o} new_o'[b] + rank
 *     inside (b, o) by hash;
 *   - reverse all-to-all-v returns the bin id of every rank-distinct hash; the read index never leaves
 *     the rank that owns the block;
 *   - (id, hash, depth) triples go to rank 0, which materialises hashValue / hashDepth / hashIndex.
 *
 * NCCL is loaded with dlopen so that the single-GPU library has no NCCL dependency.
 */
#pragma once
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

struct NcclApi {
  void *handle = nullptr ;
  decltype (&ncclGetUniqueId) GetUniqueId = nullptr ;
  decltype (&ncclCommInitRank) CommInitRank = nullptr ;
  decltype (&ncclCommDestroy) CommDestroy = nullptr ;
  decltype (&ncclGroupStart) GroupStart = nullptr ;
  decltype (&ncclGroupEnd) GroupEnd = nullptr ;
  decltype (&ncclSend) Send = nullptr ;
  decltype (&ncclRecv) Recv = nullptr ;
  decltype (&ncclAllGather) AllGather = nullptr ;
  decltype (&ncclAllReduce) AllReduce = nullptr ;
  decltype (&ncclBroadcast) Broadcast = nullptr ;
  decltype (&ncclGetErrorString) GetErrorString = nullptr ;
  bool load (std::string &why)
  { if (handle) return true ;
    const char *names[] = { "libnccl.so.2", "libnccl.so" } ;
    for (const char *n : names) if ((handle = dlopen (n, RTLD_NOW | RTLD_GLOBAL))) break ;
    if (!handle) { why = std::string ("cannot load NCCL: ") + dlerror () ; return false ; }
#define H10X_SYM(field, name) field = (decltype (field)) dlsym (handle, name) ; if (!field) { why = "NCCL symbol missing: " name ; return false ; }
    H10X_SYM (GetUniqueId, "ncclGetUniqueId") H10X_SYM (CommInitRank, "ncclCommInitRank") H10X_SYM (CommDestroy, "ncclCommDestroy")
    H10X_SYM (GroupStart, "ncclGroupStart") H10X_SYM (GroupEnd, "ncclGroupEnd") H10X_SYM (Send, "ncclSend") H10X_SYM (Recv, "ncclRecv")
    H10X_SYM (AllGather, "ncclAllGather") H10X_SYM (AllReduce, "ncclAllReduce") H10X_SYM (Broadcast, "ncclBroadcast") H10X_SYM (GetErrorString, "ncclGetErrorString")
#undef H10X_SYM
    return true ;
  }
} ;

static NcclApi gNccl ;

#define NCK(call) do { ncclResult_t r_ = (call) ; if (r_ != ncclSuccess) \
  throw H10xError (H10X_ERR_CUDA, std::string (#call) + ": " + gNccl.GetErrorString (r_)) ; } while (0)

#define H10X_MAX_RANKS 16

/* what a rank tells the others so that they can write straight into its receive buffers over NVLink */
struct PeerInfo {
  cudaIpcMemHandle_t handle ;	/* of the rank's workspace slab (one cudaMalloc) */
  uint64_t pid ;		/* ranks living in the same process use the pointer itself */
  uint64_t base ;		/* slab base address in the owning process */
  uint64_t offHash, offDepth, offFirst ;	/* byte offsets of this build's receive arrays inside the slab */
  uint64_t offBinId ;		/* ... and of the array that takes the bin ids coming back */
  int32_t device ;
  int32_t ok ;
} ;

struct PeerMap {		/* a peer's slab as mapped into this process */
  cudaIpcMemHandle_t handle ;
  uint64_t pid = 0, base = 0 ;
  char *mapped = nullptr ;
  bool viaIpc = false ;
} ;

struct DistState {
  int rank = 0, nranks = 1 ;
  ncclComm_t comm = nullptr ;
  int pushState = 0 ;		/* 0 untried, 1 peer stores work, -1 fall back to ncclSend/ncclRecv */
  bool globalCodes = false ;	/* h10x_gpu_dist_global_codes ran: hashDepth and the whole hash->code CSR are on this rank too */
  bool owesAgreement = false ;	/* the peers will wait for this rank's word at the next dist_agree (dist_bins' entry): a rank
				   that fails before it must still deliver it, or the others block in the collective for ever */
  PeerMap peers[H10X_MAX_RANKS] ;
  std::vector<cudaStream_t> copyStreams ;
  /* results of the last distributed build */
  uint32_t blockBase = 0, nBlocksGlobal = 0, nLocalBins = 0 ;
  uint64_t nReadsGlobal = 0, nHashesGlobal = 0 ;
} ;

/* ---- kernels ---- */

/* fillHashTable (hash10x.c:317-347) over the ranks' pieces: rank r's part of bin binId[j] goes behind the parts of the
   ranks before it (fill[] counts them; ranks own ascending block ranges, so every list ends up ascending).  A warp per
   local bin; the pieces of the ranks are placed one kernel after the other. */
__global__ void k_place_piece (uint32_t nBins, const uint32_t *__restrict__ binId, const uint32_t *__restrict__ off,
			       const uint32_t *__restrict__ piece, const uint64_t *__restrict__ codeOff, uint32_t *__restrict__ fill,
			       uint32_t *__restrict__ codes)
{ const uint32_t lane = threadIdx.x & 31 ;
  const uint64_t j = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
  if (j >= nBins) return ;
  const uint32_t bin = binId[j], o = off[j], n = off[j + 1] - o ;
  const uint64_t dst = codeOff[bin] + fill[bin] ;
  for (uint32_t x = lane ; x < n ; x += 32) codes[dst + x] = piece[o + x] ;
  __syncwarp () ;
  if (lane == 0) fill[bin] += n ;
}

/* rank-distinct hashes from the sorted entries: value, number of local blocks, first (global) block */
__global__ void k_local_distinct (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint64_t *__restrict__ sh,
				  const uint32_t *__restrict__ se, const uint32_t *__restrict__ entryBlk, uint64_t wMul,
				  uint64_t *__restrict__ dHash, uint32_t *__restrict__ dDepth, uint32_t *__restrict__ dFirst)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ;
  if (s >= nSeg) return ;
  uint32_t i = segStart[s] ;
  dHash[s] = sh[i] * wMul ; dDepth[s] = segStart[s+1] - i ; dFirst[s] = entryBlk[se[i]] ;
}

/* The forward exchange as ONE kernel: every rank-distinct hash is computed from the sorted entries and
   stored straight into the receive arrays of its hash-range owner - a peer GPU's memory mapped through
   CUDA IPC (or plain peer access inside one process) - so the NVLink transfer overlaps the gathers that
   produce the values and no send buffers or ncclSend/ncclRecv pairs exist.  Consecutive segments go to
   consecutive remote slots, so the stores coalesce into full NVLink packets. */
struct PushArgs {
  uint64_t *hash[H10X_MAX_RANKS] ; uint32_t *depth[H10X_MAX_RANKS] ; uint32_t *first[H10X_MAX_RANKS] ;
  uint64_t sendOff[H10X_MAX_RANKS + 1] ;	/* my segments [sendOff[o], sendOff[o+1]) belong to owner o */
  uint64_t dstOff[H10X_MAX_RANKS] ;		/* where my share starts inside owner o's receive arrays */
  int nranks ;
} ;

__global__ void k_push_distinct (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint64_t *__restrict__ sh,
				 const uint32_t *__restrict__ se, const uint32_t *__restrict__ entryBlk, uint64_t wMul,
				 PushArgs a)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ;
  if (s >= nSeg) return ;
  uint32_t i = segStart[s] ;
  int o = 0 ;
  while (o + 1 < a.nranks && (uint64_t) s >= a.sendOff[o + 1]) ++o ;
  uint64_t dst = a.dstOff[o] + ((uint64_t) s - a.sendOff[o]) ;
  a.hash[o][dst] = sh[i] * wMul ;
  a.depth[o][dst] = segStart[s + 1] - i ;
  a.first[o][dst] = entryBlk[se[i]] ;
}

/* off[o] = first segment whose hash reaches thr[o]; the segment hashes are sh[segStart[s]] * wMul, ascending */
__global__ void k_lower_bounds_seg (const uint32_t *__restrict__ segStart, const uint64_t *__restrict__ sh, uint64_t wMul,
				    uint32_t n, const uint64_t *__restrict__ thr, uint32_t nThr, uint64_t *__restrict__ off)
{ uint32_t o = blockIdx.x * blockDim.x + threadIdx.x ;
  if (o >= nThr) return ;
  uint64_t t = thr[o] ;
  uint32_t lo = 0, hi = n ;
  while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2 ; if (sh[segStart[mid]] * wMul < t) lo = mid + 1 ; else hi = mid ; }
  off[o] = lo ;
}

/* off[o] = first index whose hash >= thr[o] (hashes ascending); one thread per threshold */
__global__ void k_lower_bounds (const uint64_t *__restrict__ dHash, uint32_t n, const uint64_t *__restrict__ thr,
				uint32_t nThr, uint64_t *__restrict__ off)
{ uint32_t o = blockIdx.x * blockDim.x + threadIdx.x ;
  if (o >= nThr) return ;
  uint64_t t = thr[o] ;
  uint32_t lo = 0, hi = n ;
  while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2 ; if (dHash[mid] < t) lo = mid + 1 ; else hi = mid ; }
  off[o] = lo ;
}

/* owner side: one thread per distinct hash over its <= nranks received copies */
__global__ void k_owner_merge (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint64_t *__restrict__ oh,
			       const uint32_t *__restrict__ oi, const uint32_t *__restrict__ rDepth,
			       const uint32_t *__restrict__ rFirst, uint64_t *__restrict__ gHash,
			       uint32_t *__restrict__ gDepth, uint32_t *__restrict__ gFirst, uint32_t *__restrict__ newCnt)
{ uint32_t g = blockIdx.x * blockDim.x + threadIdx.x ;
  if (g >= nSeg) return ;
  uint32_t a = segStart[g], b = segStart[g+1] ;
  uint32_t depth = 0, first = 0xffffffffu ;
  for (uint32_t j = a ; j < b ; ++j) { uint32_t i = oi[j] ; depth += rDepth[i] ; first = min (first, rFirst[i]) ; }
  gHash[g] = oh[a] ; gDepth[g] = depth ; gFirst[g] = first ;
  atomicAdd (&newCnt[first], 1u) ;
}

/* per block b: all owners' new-hash counts, and those of the owners before this one */
__global__ void k_id_base (uint32_t nB, uint32_t nranks, uint32_t rank, const uint32_t *__restrict__ newMat,
			   uint32_t *__restrict__ colSum, uint32_t *__restrict__ below)
{ uint32_t b = blockIdx.x * blockDim.x + threadIdx.x ;
  if (b >= nB) return ;
  uint32_t all = 0, bel = 0 ;
  for (uint32_t o = 0 ; o < nranks ; ++o) { uint32_t v = newMat[(size_t) o * nB + b] ; all += v ; if (o < rank) bel += v ; }
  colSum[b] = all ; below[b] = bel ;
}

__global__ void k_group_head (const uint32_t *__restrict__ sf, uint32_t n, uint32_t *__restrict__ headPos)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) headPos[i] = (i == 0 || sf[i] != sf[i-1]) ? i : 0u ;
}

struct MaxOp { __host__ __device__ uint32_t operator() (uint32_t a, uint32_t b) const { return a > b ? a : b ; } } ;

/* sorted by (first block, hash): id = 1 + new hashes of earlier blocks + those of earlier owners in this
   block + rank inside the group */
__global__ void k_owner_ids (uint32_t n, const uint32_t *__restrict__ sf, const uint32_t *__restrict__ sg,
			     const uint32_t *__restrict__ groupStart, const uint32_t *__restrict__ prefixAll,
			     const uint32_t *__restrict__ below, uint32_t *__restrict__ gId)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  uint32_t b = sf[i] ;
  gId[sg[i]] = 1u + prefixAll[b] + below[b] + (i - groupStart[i]) ;
}

/* answer for every received copy, in the order it was received */
__global__ void k_answer_ids (uint32_t n, const uint32_t *__restrict__ segIncl, const uint32_t *__restrict__ oi,
			      const uint32_t *__restrict__ gId, uint32_t *__restrict__ ans)
{ uint32_t j = blockIdx.x * blockDim.x + threadIdx.x ;
  if (j < n) ans[oi[j]] = gId[segIncl[j] - 1] ;
}

__global__ void k_scatter_bins (uint32_t n, const uint32_t *__restrict__ id, const uint64_t *__restrict__ hash,
				const uint32_t *__restrict__ depth, uint64_t *__restrict__ hashValue, uint32_t *__restrict__ hashDepth)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) { hashValue[id[i]] = hash[i] ; hashDepth[id[i]] = depth[i] ; }
}

__global__ void k_gather_u32 (uint64_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ src, uint32_t *__restrict__ dst)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) dst[i] = src[idx[i]] ;
}
