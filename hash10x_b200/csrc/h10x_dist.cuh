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

/* the same push when the rank-distinct triples already exist (hand-written tail: k_heads_compact leaves them) */
__global__ void k_push_triples (uint32_t nSeg, const uint64_t *__restrict__ dHash, const uint32_t *__restrict__ dDepth,
				const uint32_t *__restrict__ dFirst, PushArgs a)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ;
  if (s >= nSeg) return ;
  int o = 0 ;
  while (o + 1 < a.nranks && (uint64_t) s >= a.sendOff[o + 1]) ++o ;
  uint64_t dst = a.dstOff[o] + ((uint64_t) s - a.sendOff[o]) ;
  a.hash[o][dst] = dHash[s] ; a.depth[o][dst] = dDepth[s] ; a.first[o][dst] = dFirst[s] ;
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

/* ---- owner merge without a sort --------------------------------------------------------------------------------------
   What an owner receives is NR runs (one per source rank), each already ascending in hash.  Instead of radix-sorting
   the concatenation over all 2k bits (6 library passes over 12 B) the runs are MERGED tile by tile:
     - every S-th element of every run is a splitter candidate; the (few) candidates are sorted and every NR-th one is a
       tile boundary, so a tile [B_t, B_t+1) holds on average NR*S elements and provably fewer than 3 NR S (the host aims at half the
       shared-memory capacity on average, S = CAP / (2 NR): data whose runs are so differently distributed that a tile
       overflows sends the whole merge to the library sort instead): a run has
       fewer than S elements between two of its own candidates, a tile holds at most NR candidates plus up to NR - 1 more
       whose value equals its lower boundary, and k candidates of a run inside a tile mean at most k + 1 such stretches
       (rank-distinct hashes: no run holds a value twice);
     - k_merge_bounds finds, per tile and run, where the tile starts in the run (binary search);
     - k_owner_merge_tiles loads the <= NR pieces of a tile into shared memory; an element's place in the merged order is
       its index in its own piece plus, for every other piece, the number of smaller elements there (ties: the lower source
       rank first) - binary searches in shared memory; equal hashes then stand side by side: depth = sum, first block =
       min (hash10x.c:178 / :147 seen from the owner), and every received copy learns the number of its bin (segOf).
   One launch: the tiles are handed out in order from a ticket and the number of bins before a tile comes from a decoupled
   look-back over the tile words (tileState).  Everything read or written in global memory is a contiguous piece. */
#define H10X_MERGE_CAP 4096u
#define H10X_MERGE_THREADS 256

struct MergeArgs {
  const uint64_t *rHash ; const uint32_t *rDepth, *rFirst ;	/* the receive arrays, run r at [recvOff[r], recvOff[r+1]) */
  uint64_t recvOff[H10X_MAX_RANKS + 1] ;
  const uint32_t *bnd ;		/* (nTiles + 1) * nranks: start of tile t inside run r (relative to the run) */
  unsigned long long *tileState ;	/* decoupled look-back: flag << 62 | bins of tile t (flag 1) or up to and including it (flag 2) */
  unsigned int *ticket ;
  uint64_t *gHash ; uint32_t *gDepth, *gFirst, *newCnt, *segOf ;
  unsigned int *overflow ;
  uint32_t nTiles ; int nranks ;
} ;

__global__ void k_merge_candidates (MergeArgs a, uint32_t S, const uint32_t *__restrict__ candOff, uint64_t *__restrict__ cand)
{ const int r = blockIdx.y ;
  const uint64_t n = a.recvOff[r + 1] - a.recvOff[r] ;
  const uint64_t nc = n ? (n - 1) / S : 0 ;
  for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; j < nc ; j += (uint64_t) gridDim.x * blockDim.x)
    cand[candOff[r] + j] = a.rHash[a.recvOff[r] + (j + 1) * S] ;
}

/* bnd[t * NR + r] = first element of run r that is >= B_t;  B_0 = 0, B_t = candSorted[t * NR - 1], B_nTiles = infinity */
__global__ void k_merge_bounds (MergeArgs a, const uint64_t *__restrict__ candSorted, uint32_t nCand, uint32_t *__restrict__ bnd)
{ const uint64_t x = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  const uint64_t tot = ((uint64_t) a.nTiles + 1) * a.nranks ;
  if (x >= tot) return ;
  const uint32_t t = (uint32_t) (x / a.nranks) ; const int r = (int) (x % a.nranks) ;
  const uint64_t *h = a.rHash + a.recvOff[r] ;
  const uint32_t n = (uint32_t) (a.recvOff[r + 1] - a.recvOff[r]) ;
  uint32_t lo = 0, hi = n ;
  if (t == 0) hi = 0 ;
  else if (t == a.nTiles) lo = n ;
  else
    { const uint64_t B = candSorted[min ((uint64_t) t * a.nranks - 1, (uint64_t) nCand - 1)] ;
      while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2 ; if (h[mid] < B) lo = mid + 1 ; else hi = mid ; }
    }
  bnd[x] = lo ;
}

static inline size_t h10x_merge_smem () { return (size_t) H10X_MERGE_CAP * 18 ; }

__global__ void __launch_bounds__ (H10X_MERGE_THREADS)
k_owner_merge_tiles (MergeArgs a)
{ extern __shared__ __align__ (16) unsigned char mergeRaw[] ;
  uint64_t *hs = (uint64_t*) mergeRaw ;			/* CAP: the pieces, run after run */
  uint16_t *order = (uint16_t*) (hs + H10X_MERGE_CAP) ;	/* CAP: merged position -> slot in hs[] */
  uint32_t *dep = (uint32_t*) (order + H10X_MERGE_CAP), *fst = dep + H10X_MERGE_CAP ;	/* CAP each */
  __shared__ uint32_t pStart[H10X_MAX_RANKS + 1], pSrc[H10X_MAX_RANKS], pLen[H10X_MAX_RANKS], warpTmp[33], sTile, sBase ;
  const uint32_t t = threadIdx.x ;
  const int NR = a.nranks ;
  constexpr uint32_t PER = H10X_MERGE_CAP / H10X_MERGE_THREADS ;
  for (;;)
    { __syncthreads () ;
      if (t == 0) sTile = atomicAdd (a.ticket, 1u) ;	/* tiles are taken in order: the chained scan below cannot deadlock */
      __syncthreads () ;
      const uint32_t tile = sTile ;
      if (tile >= a.nTiles) break ;
      if (t < (uint32_t) NR)		/* the piece lengths in parallel (two dependent global loads each), then a tiny prefix */
	{ const uint32_t lo = a.bnd[(size_t) tile * NR + t], hi = a.bnd[(size_t) (tile + 1) * NR + t] ;
	  pSrc[t] = lo ; pLen[t] = hi - lo ;
	}
      __syncthreads () ;
      if (t == 0)
	{ uint32_t run = 0 ;
	  for (int r = 0 ; r < NR ; ++r) { pStart[r] = run ; run += pLen[r] ; }
	  pStart[NR] = run ;
	}
      __syncthreads () ;
      uint32_t n = pStart[NR] ;
      const bool over = n > H10X_MERGE_CAP ;	/* the host re-runs with the library sort; the chain below still gets this tile's word */
      if (over) { if (t == 0) atomicExch (a.overflow, 1u) ; n = 0 ; }
      for (int r = 0 ; r < NR && !over ; ++r)
	{ const uint32_t ps = pStart[r], len = pStart[r + 1] - ps ;
	  const uint64_t g0 = a.recvOff[r] + pSrc[r] ;
	  for (uint32_t x = t ; x < len ; x += H10X_MERGE_THREADS)
	    { hs[ps + x] = a.rHash[g0 + x] ; dep[ps + x] = a.rDepth[g0 + x] ; fst[ps + x] = a.rFirst[g0 + x] ; }
	}
      __syncthreads () ;
      for (uint32_t slot = t ; slot < n ; slot += H10X_MERGE_THREADS)
	{ int r = 0 ; while (slot >= pStart[r + 1]) ++r ;
	  const uint64_t h = hs[slot] ;
	  uint32_t pos = slot - pStart[r] ;
	  for (int q = 0 ; q < NR ; ++q)
	    { if (q == r) continue ;
	      uint32_t lo = pStart[q], hi = pStart[q + 1] ;
	      if (q < r) { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1 ; if (hs[mid] <= h) lo = mid + 1 ; else hi = mid ; } }
	      else       { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1 ; if (hs[mid] <  h) lo = mid + 1 ; else hi = mid ; } }
	      pos += lo - pStart[q] ;
	    }
	  order[pos] = (uint16_t) slot ;
	}
      __syncthreads () ;
      /* heads of the runs of equal hash; a thread owns PER consecutive merged positions */
      const uint32_t p0 = min (t * PER, n), p1 = min (p0 + PER, n) ;
      uint32_t heads = 0 ;
      for (uint32_t p = p0 ; p < p1 ; ++p) if (p == 0 || hs[order[p]] != hs[order[p - 1]]) ++heads ;
      uint32_t total ;
      uint32_t k = sr_cta_exclusive_scan (heads, warpTmp, total) ;	/* bins before p0 in this tile */
      /* bins before this tile: decoupled look-back over the tile words (2 flag bits | count).  A tile publishes its own
	 count at once (flag 1 = aggregate), warp 0 then walks back 32 tiles at a time adding aggregates until it meets a
	 word that already holds an inclusive prefix (flag 2), and publishes its own inclusive prefix. */
      if (t < 32)
	{ volatile unsigned long long *st = a.tileState ;
	  if (t == 0) st[tile] = (1ull << 62) | (unsigned long long) total ;
	  unsigned long long sum = 0 ;
	  long long idx = (long long) tile - 1 ;
	  bool done = (tile == 0) ;
	  while (!done)
	    { const long long my = idx - (long long) t ;
	      const unsigned long long w = (my >= 0) ? st[my] : (2ull << 62) ;	/* before tile 0: prefix 0 */
	      const uint32_t flag = (uint32_t) (w >> 62) ;
	      const uint32_t ready = __ballot_sync (0xffffffffu, flag != 0), pm = __ballot_sync (0xffffffffu, flag == 2) ;
	      uint32_t take = 0 ;
	      if (pm) { const int fp = __ffs (pm) - 1 ; const uint32_t need = (fp == 31) ? 0xffffffffu : ((1u << (fp + 1)) - 1u) ;
			if ((ready & need) == need) { take = need ; done = true ; } }
	      else if (ready == 0xffffffffu) { take = 0xffffffffu ; idx -= 32 ; }
	      if (take)
		{ unsigned long long v = ((take >> t) & 1u) ? (w & 0x3fffffffffffffffull) : 0ull ;
#pragma unroll
		  for (int o = 16 ; o ; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o) ;
		  sum += v ;
		}
	    }
	  if (t == 0) { sBase = (uint32_t) sum ; st[tile] = (2ull << 62) | (sum + (unsigned long long) total) ; }
	}
      __syncthreads () ;
      const uint32_t base = sBase ;
      for (uint32_t p = p0 ; p < p1 ; ++p)
	{ const uint32_t slot = order[p] ;
	  const bool head = (p == 0 || hs[slot] != hs[order[p - 1]]) ;
	  if (head)
	    { uint32_t d = dep[slot], f = fst[slot] ;
	      for (uint32_t e = p + 1 ; e < n && hs[order[e]] == hs[slot] ; ++e) { d += dep[order[e]] ; f = min (f, fst[order[e]]) ; }
	      const uint32_t g = base + k ;
	      a.gHash[g] = hs[slot] ; a.gDepth[g] = d ; a.gFirst[g] = f ;
	      atomicAdd (&a.newCnt[f], 1u) ;
	      ++k ;
	    }
	  int r = 0 ; while (slot >= pStart[r + 1]) ++r ;
	  a.segOf[a.recvOff[r] + pSrc[r] + (slot - pStart[r])] = base + k - 1 ;	/* k counts the heads up to and including p's */
	}
    }
}

/* answer for every received copy, in the order it was received (merge path: segOf = the copy's bin) */
__global__ void k_answer_ids_seg (uint32_t n, const uint32_t *__restrict__ segOf, const uint32_t *__restrict__ gId, uint32_t *__restrict__ ans)
{ uint32_t j = blockIdx.x * blockDim.x + threadIdx.x ;
  if (j < n) ans[j] = gId[segOf[j]] ;
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
			     const uint32_t *__restrict__ below, uint32_t *__restrict__ gId,
			     const uint64_t *__restrict__ gHash, const uint32_t *__restrict__ gDepth,
			     uint32_t *__restrict__ sId, uint64_t *__restrict__ sHash, uint32_t *__restrict__ sDepth)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i >= n) return ;
  const uint32_t b = sf[i], g = sg[i] ;
  const uint32_t id = 1u + prefixAll[b] + below[b] + (i - groupStart[i]) ;
  gId[g] = id ;
  /* the same bins in id order (ascending inside an owner): what goes to rank 0, so that its scatter into hashValue[] /
     hashDepth[] writes whole runs of consecutive ids (one run per (block, owner)) instead of single random slots */
  sId[i] = id ; sHash[i] = gHash[g] ; sDepth[i] = gDepth[g] ;
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
