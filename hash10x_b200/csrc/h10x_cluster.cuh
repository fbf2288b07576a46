/* h10x_cluster.cuh - code->hash lists: every barcode block's (bin id, read16) entries sorted by bin id
 * (the qsort of hash10x.c:183) and packed as 8-byte ClusterHash (hash10x.c:35-43).
 *
 * One persistent CTA takes whole blocks.  A block's ids and reads are loaded into shared memory and
 * sorted by an LSD radix sort that never leaves the SM: per pass every warp walks its contiguous slice of
 * the keys in order, `__match_any_sync` groups equal digits inside a batch of 32, the group leader keeps a
 * per-warp digit counter row (no atomics: a row belongs to one warp), a (digit, warp)-ordered scan turns
 * the counters into bases, and the same walk scatters stably.  Distribution-independent: bin ids are not
 * uniform (a block's new hashes are one contiguous id range, the old ones a sample of all earlier ids).
 * Blocks with more entries than the largest class go to the library segmented sort.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ClusterArgs {
  const uint32_t *list ;	/* 0-based block numbers of this launch */
  const uint64_t *blkOff ;	/* entry offsets of the processed blocks */
  const uint32_t *entryId ;	/* bin id of every entry (entry order = block, hash) */
  const uint16_t *eRead ;	/* read index of every entry */
  uint64_t *clus ;		/* out: id | read << 32 */
  uint16_t *valsOut ;		/* or, when not NULL: only the 16-bit values, in key order (the good-hash lists of --hashDepthRange) */
  unsigned int *work ;		/* ticket counter */
  uint32_t nList ;
  uint32_t cap ;		/* entries per block this launch can hold */
  uint32_t digitBits ;		/* radix digit width */
  uint32_t passes ;		/* digitBits * passes >= bits of the largest id */
} ;

template <int THREADS>
__global__ void __launch_bounds__ (THREADS)
k_cluster_sort (ClusterArgs a)
{
  extern __shared__ __align__ (16) unsigned char smemRaw[] ;
  constexpr int NW = THREADS / 32 ;
  const uint32_t nd = 1u << a.digitBits ;
  uint32_t *kA = (uint32_t*) smemRaw ;
  uint32_t *kB = kA + a.cap ;
  uint16_t *vA = (uint16_t*) (kB + a.cap) ;
  uint16_t *vB = vA + a.cap ;
  uint32_t *digitBase = (uint32_t*) (vB + a.cap + (a.cap & 1)) ;	/* nd */
  uint16_t *hist = (uint16_t*) (digitBase + nd) ;			/* NW rows of nd counters */
  __shared__ uint32_t sTicket, warpTmp[33] ;

  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5 ;
  const uint32_t ltMask = (1u << lane) - 1u ;

  for (;;)
    { if (t == 0) sTicket = atomicAdd (a.work, 1u) ;
      __syncthreads () ;
      const uint32_t ticket = sTicket ;
      if (ticket >= a.nList) break ;
      const uint32_t blk = a.list[ticket] ;
      const uint64_t off = a.blkOff[blk] ;
      const uint32_t n = (uint32_t) (a.blkOff[blk + 1] - off) ;
      for (uint32_t i = t ; i < n ; i += THREADS) { kA[i] = a.entryId[off + i] ; vA[i] = a.eRead[off + i] ; }
      /* each warp owns a contiguous slice, a multiple of 32 long so batches stay aligned */
      const uint32_t per = ((n + NW - 1) / NW + 31) & ~31u ;
      const uint32_t w0 = min (wid * per, n), w1 = min (w0 + per, n) ;
      uint32_t *src = kA, *dst = kB ; uint16_t *vsrc = vA, *vdst = vB ;
      for (uint32_t pass = 0 ; pass < a.passes ; ++pass)
	{ const uint32_t shift = pass * a.digitBits ;
	  for (uint32_t i = t ; i < nd * NW ; i += THREADS) hist[i] = 0 ;
	  __syncthreads () ;
	  uint16_t *row = hist + wid * nd ;
	  /* count */
	  for (uint32_t b = w0 ; b < w1 ; b += 32)
	    { const uint32_t i = b + lane ;
	      const bool act = i < w1 ;
	      const uint32_t mask = __ballot_sync (0xffffffffu, act) ;
	      if (act)
		{ const uint32_t d = (src[i] >> shift) & (nd - 1) ;
		  const uint32_t peers = __match_any_sync (mask, d) ;
		  if ((peers & ltMask) == 0) row[d] = (uint16_t) (row[d] + __popc (peers)) ;
		}
	      __syncwarp () ;
	    }
	  __syncthreads () ;
	  /* bases in (digit, warp) order: totals per digit, scan over digits, then per-warp prefix */
	  for (uint32_t d = t ; d < nd ; d += THREADS)
	    { uint32_t s = 0 ;
#pragma unroll
	      for (int w = 0 ; w < NW ; ++w) s += hist[w * nd + d] ;
	      digitBase[d] = s ;
	    }
	  __syncthreads () ;
	  cta_exclusive_scan<THREADS> (digitBase, nd, warpTmp) ;
	  for (uint32_t d = t ; d < nd ; d += THREADS)
	    { uint32_t base = digitBase[d] ;
#pragma unroll
	      for (int w = 0 ; w < NW ; ++w) { uint32_t c = hist[w * nd + d] ; hist[w * nd + d] = (uint16_t) base ; base += c ; }
	    }
	  __syncthreads () ;
	  /* stable scatter: same walk, rank inside the batch from the match mask */
	  for (uint32_t b = w0 ; b < w1 ; b += 32)
	    { const uint32_t i = b + lane ;
	      const bool act = i < w1 ;
	      const uint32_t mask = __ballot_sync (0xffffffffu, act) ;
	      uint32_t d = 0, peers = 0 ;
	      if (act)
		{ const uint32_t key = src[i] ;
		  d = (key >> shift) & (nd - 1) ;
		  peers = __match_any_sync (mask, d) ;
		  const uint32_t pos = row[d] + __popc (peers & ltMask) ;
		  dst[pos] = key ; vdst[pos] = vsrc[i] ;
		}
	      __syncwarp () ;
	      if (act && (peers & ltMask) == 0) row[d] = (uint16_t) (row[d] + __popc (peers)) ;
	      __syncwarp () ;
	    }
	  __syncthreads () ;
	  uint32_t *tk = src ; src = dst ; dst = tk ;
	  uint16_t *tv = vsrc ; vsrc = vdst ; vdst = tv ;
	}
      if (a.valsOut) for (uint32_t i = t ; i < n ; i += THREADS) a.valsOut[off + i] = vsrc[i] ;
      else for (uint32_t i = t ; i < n ; i += THREADS) a.clus[off + i] = (uint64_t) src[i] | ((uint64_t) vsrc[i] << 32) ;
      __syncthreads () ;
    }
}
