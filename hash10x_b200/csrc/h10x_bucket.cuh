/* h10x_bucket.cuh - single-GPU hash grouping: bucket-major (key, value) pairs.
 *
 * Every block's unique list leaves the fused kernel sorted by hash, so splitting the entries by the top
 * bits of q = hash / w costs nothing: a bucket is a contiguous piece of every list.  The entries are
 * placed BUCKET-MAJOR (block-ascending inside a bucket) as
 *        key   = q mod 2^32                 (32 bits; the bucket number carries the bits above)
 *        value = block | read << 32         (64 bits)
 * and each bucket is radix-sorted, stable, on the key: 4 passes of 24 B per entry.  Ties keep the placement
 * order, i.e. ascending block, which is the order fillHashTable (hash10x.c:317-347) produces.  Because the
 * payload travels with the key, no later stage has to chase an entry index back into block-major arrays:
 * position i of bucket v holds q = v << 32 | sk[i] and (block, read) = sv[i].
 * (Round 1 sorted packed words `key << 31 | index in bucket`, keys only, and gathered (block, read) by index
 * afterwards: 244.9 ms per 200M-pair build against 213.2 ms for this layout - hash sort 58 -> 49 ms, codes
 * 26 -> 10 ms, bin ids 24 -> 19 ms.)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define H10X_MAX_BUCKETS 256

__device__ __forceinline__ uint32_t h10x_bucket_of (uint64_t i, const uint64_t *__restrict__ base, uint32_t nb)
{ uint32_t lo = 0, hi = nb ;		/* largest v with base[v] <= i (base[nb] = total) */
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1 ; if (base[mid] <= i) lo = mid ; else hi = mid ; }
  return lo ;
}

/* per (bucket, block) entry counts, bucket-major: cnt[v * nProcBlk + blk].  One warp per block; a lane
   binary-searches the block's sorted list for the first entry whose q reaches v << 32
   (q >= Q  <=>  hash >= Q * w, because the hashes are multiples of w). */
__device__ __forceinline__ uint32_t h10x_list_bound (const uint64_t *__restrict__ list, uint32_t n, uint32_t sh, uint64_t thr)
{ uint32_t lo = 0, hi = n ;
  while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1) ; if ((list[mid] >> sh) < thr) lo = mid + 1 ; else hi = mid ; }
  return lo ;
}

__global__ void k_bucket_count (uint32_t nProcBlk, uint32_t nb, const uint64_t *__restrict__ srcOff,
				const uint32_t *__restrict__ blkCnt, const uint64_t *__restrict__ scratch,
				const uint64_t *__restrict__ gHash, uint64_t wDiv, uint32_t *__restrict__ cnt)
{ const uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31 ;
  if (blk >= nProcBlk) return ;
  const uint64_t so = srcOff[blk] ;
  const uint32_t n = blkCnt[blk] ;
  const bool generic = (so >> 63) != 0 ;
  const uint32_t sh = generic ? 0 : (uint32_t) (so >> 56) ;
  const uint64_t *list = generic ? gHash + (so & 0x7fffffffffffffffull) : scratch + (so & 0x00ffffffffffffffull) ;
  for (uint32_t v = lane ; v < nb ; v += 32)
    { uint32_t a = v ? h10x_list_bound (list, n, sh, ((uint64_t) v << 32) * wDiv) : 0 ;
      uint32_t b = (v + 1 < nb) ? h10x_list_bound (list, n, sh, ((uint64_t) (v + 1) << 32) * wDiv) : n ;
      cnt[(size_t) v * nProcBlk + blk] = b - a ;
    }
}

__global__ void k_bucket_base (uint32_t nb, uint32_t nProcBlk, const uint64_t *__restrict__ off, uint64_t total,
			       uint64_t *__restrict__ base)
{ uint32_t v = blockIdx.x * blockDim.x + threadIdx.x ;
  if (v < nb) base[v] = off[(size_t) v * nProcBlk] ;
  if (v == nb) base[v] = total ;
}

/* bucket-major placement of (key, value).  One CTA per block. */
__global__ void k_place_bucketed_kv (uint32_t nProcBlk, uint32_t nb, const uint64_t *__restrict__ srcOff,
				     const uint32_t *__restrict__ blkCnt, const uint32_t *__restrict__ cnt,
				     const uint64_t *__restrict__ off,
				     const uint64_t *__restrict__ scratch, const uint64_t *__restrict__ gHash,
				     const uint32_t *__restrict__ gRec, const uint32_t *__restrict__ blkStart,
				     uint32_t blkBase, uint64_t wInvFull, uint32_t *__restrict__ key32, uint64_t *__restrict__ eBR)
{ __shared__ uint64_t sOff[H10X_MAX_BUCKETS] ;	/* where bucket v of this block starts, minus its first list index */
  __shared__ uint32_t sBnd[H10X_MAX_BUCKETS + 1] ;
  for (uint32_t blk = blockIdx.x ; blk < nProcBlk ; blk += gridDim.x)
    { const uint64_t so = srcOff[blk] ;
      const uint32_t n = blkCnt[blk] ;
      __syncthreads () ;
      /* exclusive scan of this block's nb bucket counts by warp 0 (nb <= 256: 8 per lane) */
      if (threadIdx.x < 32)
	{ const uint32_t lane = threadIdx.x, per = (nb + 31) / 32 ;
	  const uint32_t lo = min (lane * per, nb), hi = min (lo + per, nb) ;
	  uint32_t sum = 0 ;
	  for (uint32_t v = lo ; v < hi ; ++v) sum += cnt[(size_t) v * nProcBlk + blk] ;
	  uint32_t inc = sum ;
#pragma unroll
	  for (int d = 1 ; d < 32 ; d <<= 1) { uint32_t u = __shfl_up_sync (0xffffffffu, inc, d) ; if (lane >= d) inc += u ; }
	  uint32_t run = inc - sum ;
	  for (uint32_t v = lo ; v < hi ; ++v) { sBnd[v] = run ; run += cnt[(size_t) v * nProcBlk + blk] ; }
	  if (lane == 31) sBnd[nb] = inc ;
	}
      __syncthreads () ;
      for (uint32_t v = threadIdx.x ; v < nb ; v += blockDim.x) sOff[v] = off[(size_t) v * nProcBlk + blk] - sBnd[v] ;
      __syncthreads () ;
      const bool generic = (so >> 63) != 0 ;
      const uint32_t sh = generic ? 0 : (uint32_t) (so >> 56) ;
      const uint64_t o = generic ? (so & 0x7fffffffffffffffull) : (so & 0x00ffffffffffffffull) ;
      const uint64_t rmask = ((uint64_t) 1 << sh) - 1 ;
      const uint32_t r0 = blkStart[blk] ;
      for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	{ uint64_t hash ; uint16_t rd ;
	  if (generic) { hash = gHash[o + i] ; rd = (uint16_t) (gRec[o + i] - r0) ; }
	  else { uint64_t key = scratch[o + i] ; hash = key >> sh ; rd = (uint16_t) (key & rmask) ; }
	  const uint64_t q = hash * wInvFull ;
	  const uint64_t pos = sOff[(uint32_t) (q >> 32)] + i ;	/* off[] is global: bucket base + blocks before + index in chunk */
	  key32[pos] = (uint32_t) q ;
	  eBR[pos] = (uint64_t) (blkBase + blk + 1) | ((uint64_t) rd << 32) ;
	}
    }
}

/* selects the first position of every bin in the sorted order (for cub::DeviceSelect::If over a counting
   iterator): the key changes, or a new bucket starts.  Most positions repeat the key of their predecessor, so the
   bucket test runs for nearly every entry: hashes are uniform, the bucket holding i is almost always
   floor (i * nb / H) (scale = 2^64 * nb / H), and only a wrong guess pays for the binary search. */
struct HeadPredKV {
  const uint32_t *sk ; const uint64_t *base ; uint32_t nb ; uint64_t scale ;
  __device__ __forceinline__ bool operator() (uint32_t i) const
  { if (i == 0) return true ;
    if (sk[i] != sk[i-1]) return true ;
    if (nb == 1) return false ;
    uint32_t v = min ((uint32_t) __umul64hi ((uint64_t) i, scale), nb - 1) ;
    const uint64_t b0 = base[v] ;
    if (b0 <= i && i < base[v+1]) return b0 == i ;
    return base[h10x_bucket_of (i, base, nb)] == i ;
  }
} ;

/* key of a bin for the id order (hash10x.c:147): first block holding it (the first entry of the bin: the sort is
   stable and a bucket is filled block-ascending), then hash */
__global__ void k_first_key_kv (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint32_t *__restrict__ sk,
				const uint64_t *__restrict__ sv, const uint64_t *__restrict__ base, uint32_t nb, int qBits,
				uint64_t *__restrict__ key, uint32_t *__restrict__ segIdx)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ;
  if (s >= nSeg) return ;
  uint32_t i = segStart[s] ;
  uint64_t q = ((uint64_t) h10x_bucket_of (i, base, nb) << 32) | sk[i] ;
  key[s] = ((uint64_t) (uint32_t) sv[i] << qBits) | q ;
  segIdx[s] = s ;
}

/* rank r in that order is bin id r + 1; the hash comes back out of the sorted key */
__global__ void k_bins_by_rank_kv (uint32_t nSeg, const uint32_t *__restrict__ sortedSeg, const uint64_t *__restrict__ sortedKey,
				   const uint32_t *__restrict__ segStart, int qBits, uint64_t wMul,
				   uint32_t *__restrict__ idOfSeg, uint64_t *__restrict__ hashValue, uint32_t *__restrict__ hashDepth)
{ uint32_t r = blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= nSeg) return ;
  uint32_t s = sortedSeg[r], id = r + 1u ;
  idOfSeg[s] = id ;
  hashValue[id] = (sortedKey[r] & (((uint64_t) 1 << qBits) - 1)) * wMul ;
  hashDepth[id] = segStart[s+1] - segStart[s] ;
}

/* fillHashTable (hash10x.c:317-347) + the transposed view.  A bin's entries are contiguous in the sorted order and
   go, in the same (ascending block) order, to codes[codeOff[id] ..) and idRead[..].  A warp takes 32 consecutive
   bins: lane l fetches bin l's (start, length, id, destination) - coalesced, and one dependent gather for codeOff -
   then the warp walks the bins with shuffles, 32 entries at a time, loading the next bin's entries before storing
   the current one's so that the DRAM latency is not paid once per bin. */
__global__ void k_codes_seg_kv (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint32_t *__restrict__ idOfSeg,
				const uint64_t *__restrict__ sv, const uint64_t *__restrict__ codeOff,
				uint32_t *__restrict__ codes, uint64_t *__restrict__ idRead)
{ const uint32_t lane = threadIdx.x & 31 ;
  const uint64_t warp = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
  const uint64_t s0 = warp * 32 ;
  if (s0 >= nSeg) return ;
  const uint32_t cnt = (uint32_t) min ((uint64_t) 32, (uint64_t) nSeg - s0) ;
  uint32_t myI0 = 0, myN = 0, myId = 0 ; uint64_t myDst = 0 ;
  if (lane < cnt)
    { const uint32_t s = (uint32_t) s0 + lane ;
      myI0 = segStart[s] ; myN = segStart[s+1] - myI0 ; myId = idOfSeg[s] ; myDst = codeOff[myId] ;
    }
  uint32_t i0 = __shfl_sync (0xffffffffu, myI0, 0), n = __shfl_sync (0xffffffffu, myN, 0) ;
  uint64_t cur = (lane < n) ? sv[(uint64_t) i0 + lane] : 0 ;
  for (uint32_t t = 0 ; t < cnt ; ++t)
    { const uint32_t id = __shfl_sync (0xffffffffu, myId, t) ;
      const uint64_t dst = __shfl_sync (0xffffffffu, myDst, t) ;
      const uint32_t tn = (t + 1 < cnt) ? t + 1 : t ;
      const uint32_t i0n = __shfl_sync (0xffffffffu, myI0, tn), nn = __shfl_sync (0xffffffffu, myN, tn) ;
      uint64_t nxt = (t + 1 < cnt && lane < nn) ? sv[(uint64_t) i0n + lane] : 0 ;
      if (lane < n)
	{ codes[dst + lane] = (uint32_t) cur ;
	  idRead[dst + lane] = (uint64_t) id | (cur & 0xffff00000000ull) ;
	}
      for (uint32_t j = lane + 32 ; j < n ; j += 32)	/* deep bins */
	{ uint64_t br = sv[(uint64_t) i0 + j] ;
	  codes[dst + j] = (uint32_t) br ;
	  idRead[dst + j] = (uint64_t) id | (br & 0xffff00000000ull) ;
	}
      i0 = i0n ; n = nn ; cur = nxt ;
    }
}
