/* h10x_fused.cuh - the hot kernel: one CTA per barcode block does the whole of processBlock's
 * first half (hash10x.c:154-172) on chip:
 *
 *   ingest     one warp per read pair; lane l loads word l of the 30-word FQB record (coalesced);
 *   windows    the pair's 237 k-mers are split into chunks of 8 consecutive start positions, one
 *              chunk per lane (14 chunks of read 1, 17 of read 2 for k=21); the three 32-bit words
 *              that hold a chunk's 8+k-1 bases come from the owning lanes by warp shuffle and are
 *              funnel-shifted into one 64-bit window W; its 2-bit-reversed complement WR gives the
 *              reverse-complement k-mers, so no per-base rolling state exists at all;
 *   hash       h_j = (W >> (64-2k-2j)) & mask, hRC_j = (WR >> 2j) & mask, both times factor1
 *              (seqhash.c:58-59), canonical = smaller of the two top-2k-bit products (seqhash.c:67),
 *              "mosh" iff divisible by w (seqhash.c:171,189) - tested without division on the
 *              unshifted product;
 *   collect    selected (hash << rb | readIndex) keys go to a shared-memory list and a bucket
 *              histogram on the top hash bits;
 *   sort       bucket scatter + per-bucket insertion sort (expected O(n): hashes are uniform);
 *   dedup      first of each run of equal hashes = lowest read index, which is what the reference's
 *              stable qsort + dedup loop keeps (hash10x.c:166-172); an empty block yields the
 *              phantom {hash 0, read 0} entry (hash10x.c:167-168);
 *   output     the block's sorted unique list goes to a global scratch slab at an atomically
 *              reserved offset; k_place later moves it to its final place in block order.
 *
 * A block whose keys do not fit (more moshes than `cap`, an over-full bucket from low-complexity
 * reads, scratch exhausted) is flagged H10X_BLK_FALLBACK and re-done by the generic global-memory
 * path, so results never depend on which path ran.
 */
#pragma once
#include "h10x_common.cuh"

#define H10X_BLK_FALLBACK 0xffffffffu
#define H10X_BUCKET_LIMIT 48u		/* longest bucket the insertion sort will take */

__device__ __forceinline__ uint32_t h10x_swap_pairs (uint32_t x)
{ return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1) ; }

/* exclusive scan in place of a[0..n) (n a multiple of nothing in particular) by the whole CTA;
   returns the total to every thread.  warpTmp: 33 words of shared memory. */
template <int THREADS>
__device__ __forceinline__ uint32_t cta_exclusive_scan (uint32_t *a, uint32_t n, uint32_t *warpTmp)
{
  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5 ;
  const uint32_t per = (n + THREADS - 1) / THREADS ;
  const uint32_t lo = min (t * per, n), hi = min (lo + per, n) ;
  uint32_t sum = 0 ;
  for (uint32_t i = lo ; i < hi ; ++i) sum += a[i] ;
  uint32_t inc = sum ;
#pragma unroll
  for (int d = 1 ; d < 32 ; d <<= 1) { uint32_t v = __shfl_up_sync (0xffffffffu, inc, d) ; if (lane >= d) inc += v ; }
  if (lane == 31) warpTmp[wid] = inc ;
  __syncthreads () ;
  if (wid == 0)
    { uint32_t v = (lane < THREADS / 32) ? warpTmp[lane] : 0 ;
      uint32_t iv = v ;
#pragma unroll
      for (int d = 1 ; d < 32 ; d <<= 1) { uint32_t u = __shfl_up_sync (0xffffffffu, iv, d) ; if (lane >= d) iv += u ; }
      if (lane < THREADS / 32) warpTmp[lane] = iv - v ;
      if (lane == 31) warpTmp[32] = iv ;
    }
  __syncthreads () ;
  uint32_t run = warpTmp[wid] + inc - sum ;
  for (uint32_t i = lo ; i < hi ; ++i) { uint32_t v = a[i] ; a[i] = run ; run += v ; }
  uint32_t total = warpTmp[32] ;
  __syncthreads () ;
  return total ;
}

struct FusedArgs {
  const uint32_t *fqb ;		/* records */
  const uint32_t *list ;	/* 0-based block numbers handled by this launch */
  const uint32_t *blkStart ;	/* first record of every block */
  uint64_t *scratch ;		/* unique-key slab */
  unsigned long long *cursor ;	/* next free scratch slot */
  uint64_t scratchCap ;
  uint64_t *srcOff ;		/* per block: where its list went */
  uint32_t *blkCnt ;		/* per block: unique count, or H10X_BLK_FALLBACK */
  uint32_t nList ;
  uint32_t cap ;		/* key capacity of the shared-memory list */
  uint32_t nbuck ;		/* buckets (power of two) */
  uint32_t lb ;			/* log2 (nbuck) */
  uint32_t c1, c2 ;		/* 8-k-mer chunks of read 1 / read 2 */
  uint32_t n1, n2 ;		/* k-mers of read 1 / read 2 */
} ;

template <int THREADS, bool WODD>
__global__ void __launch_bounds__ (THREADS)
k_fused_block (FusedArgs a, HashParams hp)
{
  extern __shared__ __align__ (16) unsigned char smemRaw[] ;
  uint64_t *A = (uint64_t*) smemRaw ;
  uint64_t *S = A + a.cap ;
  uint32_t *start = (uint32_t*) (S + a.cap) ;	/* nbuck + 1 */
  uint32_t *cur = start + a.nbuck + 1 ;		/* nbuck */
  __shared__ uint32_t sCount, sBig, sUnique, warpTmp[33] ;
  __shared__ unsigned long long sBase ;

  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5 ;
  const uint32_t blk = a.list[blockIdx.x] ;
  const uint32_t r0 = a.blkStart[blk], nRead = a.blkStart[blk + 1] - r0 ;
  const uint32_t rb = (nRead > 1) ? 32 - __clz (nRead - 1) : 0 ;	/* bits of the read index */
  const int k2 = 2 * hp.k ;
  const int buckShift = k2 - (int) a.lb ;

  for (uint32_t i = t ; i <= a.nbuck ; i += THREADS) start[i] = 0 ;
  if (t == 0) { sCount = 0 ; sBig = 0 ; }
  __syncthreads () ;

  /* ---- lane's chunk: which read, first k-mer start position, source lanes of its 3 words ---- */
  const bool isR2 = lane >= a.c1 ;
  const uint32_t chunk = isR2 ? lane - a.c1 : lane ;
  const bool laneActive = lane < a.c1 + a.c2 ;
  const uint32_t p0 = (isR2 ? H10X_R2_START : H10X_R1_START) + 8 * chunk ;	/* unpacked position */
  const uint32_t nk = isR2 ? a.n2 : a.n1 ;		/* k-mers of this read */
  const uint32_t first = 8 * chunk ;			/* k-mer number of j = 0 */
  const uint32_t wi = p0 >> 4, sh = 2 * (p0 & 15) ;
  const uint32_t base = isR2 ? 15u : 0u ;
  const uint32_t src0 = base + wi, src1 = base + min (wi + 1, 9u), src2 = base + min (wi + 2, 9u) ;
  const bool has1 = wi + 1 <= 9, has2 = wi + 2 <= 9 ;
  const uint64_t f = hp.factor1 ;
  const uint64_t topMask = ~(((uint64_t) 1 << hp.shift) - 1) ;	/* the 2k hash bits of a product */
  const uint64_t tzMaskSh = hp.wTzMask << hp.shift ;

  for (uint32_t pr = wid ; pr < nRead ; pr += THREADS / 32)
    { const uint32_t *rec = a.fqb + (size_t) H10X_REC_WORDS * (r0 + pr) ;
      uint32_t word = (lane < H10X_REC_WORDS) ? __ldg (rec + lane) : 0u ;
      uint32_t w0 = __shfl_sync (0xffffffffu, word, src0) ;
      uint32_t w1 = __shfl_sync (0xffffffffu, word, src1) ;
      uint32_t w2 = __shfl_sync (0xffffffffu, word, src2) ;
      if (!has1) w1 = 0 ;
      if (!has2) w2 = 0 ;
      if (!laneActive) continue ;
      uint32_t Whi = __funnelshift_l (w1, w0, sh), Wlo = __funnelshift_l (w2, w1, sh) ;
      uint64_t W = ((uint64_t) Whi << 32) | Wlo ;		/* bases p0 .. p0+31, first base on top */
      uint64_t WR = ((uint64_t) h10x_swap_pairs (__brev (~Wlo)) << 32) | h10x_swap_pairs (__brev (~Whi)) ;
#pragma unroll
      for (int j = 0 ; j < 8 ; ++j)
	{ uint64_t h = (W >> (hp.shift - 2 * j)) & hp.kmask ;
	  uint64_t hrc = (WR >> (2 * j)) & hp.kmask ;
	  uint64_t pf = (h * f) & topMask, prr = (hrc * f) & topMask ;
	  uint64_t m = pf < prr ? pf : prr ;			/* canonical hash << shift */
	  bool sel = (first + j < nk) && (m * hp.wInv <= hp.wLim) ;
	  if (!WODD) sel = sel && ((m & tzMaskSh) == 0) ;
	  if (sel)
	    { uint64_t hash = m >> hp.shift ;
	      uint32_t pos = atomicAdd (&sCount, 1u) ;
	      if (pos < a.cap) A[pos] = (hash << rb) | pr ;
	      atomicAdd (&start[(uint32_t) (hash >> buckShift)], 1u) ;
	    }
	}
    }
  __syncthreads () ;

  uint32_t n = sCount ;
  bool bad = n > a.cap ;
  if (!bad && n == 0)		/* hash10x.c:167-168: the phantom entry of an empty block */
    { if (t == 0) { A[0] = 0 ; start[0] = 1 ; }
      n = 1 ;
      __syncthreads () ;
    }

  uint32_t U = 0 ;
  if (!bad)
    { /* ---- bucket offsets ---- */
      for (uint32_t i = t ; i < a.nbuck ; i += THREADS) if (start[i] > H10X_BUCKET_LIMIT) sBig = 1 ;
      __syncthreads () ;
      bad = sBig != 0 ;
    }
  if (!bad)
    { cta_exclusive_scan<THREADS> (start, a.nbuck + 1, warpTmp) ;
      for (uint32_t i = t ; i < a.nbuck ; i += THREADS) cur[i] = start[i] ;
      __syncthreads () ;
      /* ---- scatter to buckets ---- */
      for (uint32_t i = t ; i < n ; i += THREADS)
	{ uint64_t key = A[i] ;
	  uint32_t b = (uint32_t) ((key >> rb) >> buckShift) ;
	  S[atomicAdd (&cur[b], 1u)] = key ;
	}
      __syncthreads () ;
      /* ---- insertion sort inside each bucket: buckets are in hash order, so S ends up sorted ---- */
      for (uint32_t b = t ; b < a.nbuck ; b += THREADS)
	{ uint32_t lo = start[b], hi = start[b + 1] ;
	  for (uint32_t i = lo + 1 ; i < hi ; ++i)
	    { uint64_t key = S[i] ; uint32_t j = i ;
	      while (j > lo && S[j - 1] > key) { S[j] = S[j - 1] ; --j ; }
	      S[j] = key ;
	    }
	}
      __syncthreads () ;
      /* ---- dedup: keep the first key of every run of equal hashes (lowest read index) ---- */
      const uint32_t per = (n + THREADS - 1) / THREADS ;
      const uint32_t lo = min (t * per, n), hi = min (lo + per, n) ;
      uint32_t cnt = 0 ;
      for (uint32_t i = lo ; i < hi ; ++i) cnt += (i == 0 || (S[i] >> rb) != (S[i - 1] >> rb)) ? 1u : 0u ;
      uint32_t *perThread = cur ;		/* reuse: nbuck >= THREADS is guaranteed by the host */
      perThread[t] = cnt ;
      __syncthreads () ;
      U = cta_exclusive_scan<THREADS> (perThread, THREADS, warpTmp) ;
      uint32_t o = perThread[t] ;
      for (uint32_t i = lo ; i < hi ; ++i)
	if (i == 0 || (S[i] >> rb) != (S[i - 1] >> rb)) A[o++] = S[i] ;
      if (t == 0)
	{ unsigned long long b0 = atomicAdd (a.cursor, (unsigned long long) U) ;
	  sBase = b0 ;
	}
      __syncthreads () ;
      if (sBase + U > a.scratchCap) bad = true ;
    }
  if (bad)
    { if (t == 0) { a.blkCnt[blk] = H10X_BLK_FALLBACK ; a.srcOff[blk] = 0 ; }
      return ;
    }
  uint64_t *dst = a.scratch + sBase ;
  for (uint32_t i = t ; i < U ; i += THREADS) dst[i] = A[i] ;
  if (t == 0) { a.blkCnt[blk] = U ; a.srcOff[blk] = sBase | ((uint64_t) rb << 56) ; }
}

/* Moves every block's unique list to its final place, in block order, and splits it into the
   arrays the index stages use: eHash (hash value), eRead (read index, 16 bits: hash10x.c:37,180) and
   entryBlk (1-based block number).  One CTA per block.  kind 0: fused scratch (key = hash<<rb|read);
   kind 1: generic path arrays (gHash, gRec). */
__global__ void k_place (uint32_t nProcBlk, const uint64_t *__restrict__ srcOff, const uint32_t *__restrict__ blkCnt,
			 const uint64_t *__restrict__ blkOff, const uint64_t *__restrict__ scratch,
			 const uint64_t *__restrict__ gHash, const uint32_t *__restrict__ gRec,
			 const uint32_t *__restrict__ blkStart,
			 uint64_t *__restrict__ eHash, uint16_t *__restrict__ eRead, uint32_t *__restrict__ entryBlk)
{ for (uint32_t blk = blockIdx.x ; blk < nProcBlk ; blk += gridDim.x)
    { uint64_t so = srcOff[blk] ;
      uint32_t n = blkCnt[blk] ;
      uint64_t dst = blkOff[blk] ;
      if (so >> 63)	/* generic source */
	{ uint64_t off = so & 0x7fffffffffffffffull ;
	  uint32_t r0 = blkStart[blk] ;
	  for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	    { eHash[dst + i] = gHash[off + i] ; eRead[dst + i] = (uint16_t) (gRec[off + i] - r0) ; entryBlk[dst + i] = blk + 1 ; }
	}
      else
	{ uint32_t rb = (uint32_t) (so >> 56) ;
	  uint64_t off = so & 0x00ffffffffffffffull ;
	  uint64_t rmask = ((uint64_t) 1 << rb) - 1 ;
	  for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	    { uint64_t key = scratch[off + i] ;
	      eHash[dst + i] = key >> rb ; eRead[dst + i] = (uint16_t) (key & rmask) ; entryBlk[dst + i] = blk + 1 ;
	    }
	}
    }
}
