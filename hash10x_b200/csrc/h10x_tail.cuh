/* h10x_tail.cuh - the grouping tail of the single-GPU build, hand-written (round 2): no library sort.
 *
 * After the fused kernel every barcode block owns a sorted list of unique (hash, read) keys.  What the
 * reference does next, one hash at a time (hashIndexFind hash10x.c:139-152, ++hashDepth :178, the
 * ClusterHash sort :183, fillHashTable :317-347), is here five data movements over H entries:
 *
 *   P1  range partition   every block list is sorted, so the entries of one of ~2048 hash RANGES are a
 *                         contiguous piece of it: a warp walks the blocks of a tile in order and appends each
 *                         piece to the range's area (k_p1_count, scan, k_p1_place) - stable, no ranking at
 *                         all.  From here on an entry is ONE 64-bit word
 *                              E = (q mod 2^lowBits) << (blkBits+16) | block << 16 | read16,   q = hash / w
 *                         (the range number carries q's top bits), 8 bytes instead of the 12 the library
 *                         sort moved 7 times.
 *   P2  digit partition   inside each range, a stable partition on the next p2 bits of q (k_part_hist, scan,
 *                         k_part_scatter) cuts it into SUB-RANGES of ~2600 entries ...
 *   S   sub-range sort    ... which one CTA sorts in shared memory on the remaining q bits (k_sr_sort: LSD, 9 bits
 *                         per pass, stable, so inside a hash the entries keep ascending block order) and scans for
 *                         the first entry of every hash = the bins, in hash order.
 *   ids                   bin ids are the reference's insertion order = (first block, hash): the bins already
 *                         stand in hash order, so two stable partition passes on the first block finish it.
 *   C   codes             a warp per 32 bins copies each bin's entries to codes[] (fillHashTable's lists) and
 *                         (id, read) beside it, bin-major by id (k_codes_seg_e).
 *   T1,T2 transposition   two stable partition passes on the block number (low bits, then high bits) turn that
 *                         into every block's ClusterHash list sorted by bin id (hash10x.c:183).
 *
 * The partition kernel is one template (k_part_hist / k_part_scatter): a job is a contiguous piece of the
 * input handled by one CTA in sub-tiles; ranks inside a sub-tile come from __match_any_sync + warp-private
 * counters, the order of warps and sub-tiles is the order of the data, so every pass is stable.
 *
 * A sub-range that does not fit shared memory (skewed data: one hash held by very many blocks) is sorted by
 * the library on the full word and scanned by k_sr_heads_big; results never depend on which way was taken.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define H10X_SR_CAP 5120u		/* entries of a sub-range the shared-memory sort takes */
#define H10X_SR_THREADS_DEFAULT 512
#define H10X_SR_TARGET 1400.0		/* average sub-range size aimed at; min(hashF, hashR) is not uniform (density 2(1-x)), so the
					   low sub-ranges hold twice the average and the high ones next to nothing: k_sr_jobs joins
					   aligned pairs / quads of sub-ranges into sort jobs of up to H10X_SR_CAP entries */
#define H10X_SR_DIGIT 9			/* bits per shared-memory sort pass */
#define H10X_P1_MAX_RANGES 2304u		/* 64-byte ring + 12 bytes of cursors per range in shared memory */
#define H10X_PART_MAX_BINS 1024u

struct TailGeom {
  int sortBits ;	/* bits of q = hash / w */
  int blkBits ;		/* bits of a (global, 1-based) block number */
  int remBits ;		/* q bits sorted in shared memory */
  int p2 ;		/* q bits of the digit partition */
  int lowBits ;		/* remBits + p2: q bits kept inside E */
  int eShift ;		/* blkBits + 16: where they sit */
  uint32_t nRanges ;	/* (top >> lowBits) + 1 */
  uint32_t nSub ;	/* nRanges << p2 */
} ;

/* ---------------------------------------------------------------- P1: range partition of the sorted block lists */

/* cnt[v * nTiles + tile] = entries of range v in the blocks of a tile (G consecutive blocks) */
__global__ void k_p1_count (uint32_t nProcBlk, uint32_t G, uint32_t nTiles, uint32_t nRanges, int lowBits,
			    const uint64_t *__restrict__ srcOff, const uint32_t *__restrict__ blkCnt,
			    const uint64_t *__restrict__ scratch, const uint64_t *__restrict__ gHash, uint64_t wInvFull,
			    uint32_t *__restrict__ cnt)
{ extern __shared__ uint32_t p1hist[] ;
  for (uint32_t tile = blockIdx.x ; tile < nTiles ; tile += gridDim.x)
    { for (uint32_t v = threadIdx.x ; v < nRanges ; v += blockDim.x) p1hist[v] = 0 ;
      __syncthreads () ;
      const uint32_t b0 = tile * G, b1 = min (b0 + G, nProcBlk) ;
      for (uint32_t blk = b0 ; blk < b1 ; ++blk)
	{ const uint64_t so = srcOff[blk] ;
	  const uint32_t n = blkCnt[blk] ;
	  const bool generic = (so >> 63) != 0 ;
	  const uint32_t sh = generic ? 0 : (uint32_t) (so >> 56) ;
	  const uint64_t *list = generic ? gHash + (so & 0x7fffffffffffffffull) : scratch + (so & 0x00ffffffffffffffull) ;
	  for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	    { const uint64_t q = (list[i] >> sh) * wInvFull ;
	      atomicAdd (&p1hist[(uint32_t) (q >> lowBits)], 1u) ;
	    }
	}
      __syncthreads () ;
      for (uint32_t v = threadIdx.x ; v < nRanges ; v += blockDim.x) cnt[(size_t) v * nTiles + tile] = p1hist[v] ;
      __syncthreads () ;
    }
}

/* rangeStart[v] = first entry of range v in the range-major array (off = exclusive scan of cnt) */
__global__ void k_p1_range_start (uint32_t nRanges, uint32_t nTiles, const uint32_t *__restrict__ off, uint64_t total,
				  uint64_t *__restrict__ rangeStart)
{ uint32_t v = blockIdx.x * blockDim.x + threadIdx.x ;
  if (v < nRanges) rangeStart[v] = off[(size_t) v * nTiles] ;
  if (v == nRanges) rangeStart[v] = total ;
}

/* Write combining.  An entry's destination is one of thousands of output streams, each advancing a few entries at
   a time; stored straight to global memory those 8-byte pieces leave L2 as half-filled 32-byte sectors and DRAM
   pays a read-modify-write for each (ncu: 2.5x the bytes, both ways).  So every stream owns a ring of 8 entries
   in shared memory, slot = position & 7: entries collect there and an ALIGNED group of 4 (one full sector) is
   written by 4 adjacent lanes once it is complete.  `lo` = the stream's position at the last flush, rounded down
   to 4: an entry at position < lo + 8 goes to the ring, one beyond it (more than a ring's worth arrived between
   two flushes) is stored directly, and h10x_ring_flush then fetches the unfinished tail of those back into the
   ring.  first = the stream's first position in this job: the group holding it is shared with the stream's
   previous owner and only written from there on.  Called by all threads after a barrier; leaves posOld = posNew. */
#define H10X_RING_STRIDE 8u
__device__ __forceinline__ uint32_t h10x_ring_slot (uint32_t d, uint32_t s) { return d * H10X_RING_STRIDE + s ; }
__device__ __forceinline__ void h10x_ring_flush (uint32_t nStreams, uint32_t nThreads, uint32_t *posOld, const uint32_t *posNew,
						  const uint32_t *first, uint64_t *ring, uint64_t *__restrict__ out)
{ /* four adjacent lanes per stream: a complete group leaves as one 32-byte request.  (One thread per stream with two
     128-bit stores was tried: k_p1_place 14.6 -> 17.9 ms, the 1024-digit pass 19.7 -> 24.9 ms.) */
  for (uint32_t x0 = 0 ; x0 < nStreams * 4u ; x0 += nThreads)		/* whole warps stay in the loop: __syncwarp below */
    { const uint32_t x = x0 + threadIdx.x ;
      const bool in = x < nStreams * 4u ;
      const uint32_t d = in ? x >> 2 : 0u, sub = x & 3u ;
      const uint32_t c0 = posOld[d], hi = posNew[d] ;
      const bool act = in && hi != c0 ;
      if (act)
	{ const uint32_t lo = c0 & ~3u, f = first[d] ;
	  if (lo + 4u <= hi) { const uint32_t p = lo + sub ; if (p >= f) out[p] = ring[h10x_ring_slot (d, p & 7u)] ; }
	  if (lo + 8u <= hi) { const uint32_t p = lo + 4u + sub ; if (p >= f) out[p] = ring[h10x_ring_slot (d, p & 7u)] ; }
	  if (hi - lo > 8u) { const uint32_t p = (hi & ~3u) + sub ; if (p < hi) ring[h10x_ring_slot (d, p & 7u)] = out[p] ; }
	}
      __syncwarp () ;		/* the four lanes of a stream sit in one warp: all have read posOld */
      if (act && sub == 0) posOld[d] = hi ;
    }
}

/* end of a job: what is still in the rings */
__device__ __forceinline__ void h10x_ring_drain (uint32_t nStreams, uint32_t nThreads, const uint32_t *pos, const uint32_t *first,
						  const uint64_t *ring, uint64_t *__restrict__ out)
{ for (uint32_t d = threadIdx.x ; d < nStreams ; d += nThreads)
    { const uint32_t c = pos[d] ;
      for (uint32_t p = max (c & ~3u, first[d]) ; p < c ; ++p) out[p] = ring[h10x_ring_slot (d, p & 7u)] ;
    }
}

/* One CTA per tile, its blocks one after the other.  base[v] is the next free slot of range v for this tile; an
   entry takes its slot with a shared-memory atomic.  A barrier between blocks keeps a range block-ascending, which
   is all the later stable passes need: inside one block every hash occurs once, so the order of a block's
   entries inside a range never matters.  The first H10X_P1_PF * threads keys of the next block are fetched before
   the current block is placed. */
#define H10X_P1_THREADS 1024
#define H10X_P1_PF 3
__global__ void __launch_bounds__ (H10X_P1_THREADS, 1)
k_p1_place (uint32_t flushEvery, uint32_t nProcBlk, uint32_t G, uint32_t nTiles, uint32_t nRanges, int lowBits, int eShift,
	    const uint64_t *__restrict__ srcOff, const uint32_t *__restrict__ blkCnt, const uint32_t *__restrict__ off,
	    const uint64_t *__restrict__ scratch, const uint64_t *__restrict__ gHash, const uint32_t *__restrict__ gRec,
	    const uint32_t *__restrict__ blkStart, uint32_t blkBase, uint64_t wInvFull, uint64_t *__restrict__ out)
{ extern __shared__ __align__ (16) unsigned char p1raw[] ;
  uint64_t *ring = (uint64_t*) p1raw ;				/* nRanges rings */
  uint32_t *base = (uint32_t*) (ring + (size_t) nRanges * H10X_RING_STRIDE) ;	/* next free slot */
  uint32_t *baseOld = base + nRanges ;				/* ... at the last flush */
  uint32_t *first = baseOld + nRanges ;				/* ... at the start of the tile */
  const uint32_t t = threadIdx.x ;
  const uint64_t lowMask = (lowBits >= 64) ? ~(uint64_t) 0 : (((uint64_t) 1 << lowBits) - 1) ;
  for (uint32_t tile = blockIdx.x ; tile < nTiles ; tile += gridDim.x)
    { __syncthreads () ;
      for (uint32_t v = t ; v < nRanges ; v += H10X_P1_THREADS)
	{ const uint32_t o = off[(size_t) v * nTiles + tile] ; base[v] = o ; baseOld[v] = o ; first[v] = o ; }
      __syncthreads () ;
      const uint32_t b0 = tile * G, b1 = min (b0 + G, nProcBlk) ;
      uint64_t pre[H10X_P1_PF] ;
      { const uint64_t so = srcOff[b0] ; const uint32_t n = blkCnt[b0] ;
#pragma unroll
	for (int u = 0 ; u < H10X_P1_PF ; ++u)
	  { const uint32_t i = u * H10X_P1_THREADS + t ;
	    pre[u] = (!(so >> 63) && i < n) ? __ldcs (scratch + (so & 0x00ffffffffffffffull) + i) : 0 ;
	  }
      }
      for (uint32_t blk = b0 ; blk < b1 ; ++blk)
	{ const uint64_t so = srcOff[blk] ;
	  const uint32_t n = blkCnt[blk] ;
	  const bool generic = (so >> 63) != 0 ;
	  const uint32_t sh = generic ? 0 : (uint32_t) (so >> 56) ;
	  const uint64_t o = generic ? (so & 0x7fffffffffffffffull) : (so & 0x00ffffffffffffffull) ;
	  const uint64_t rmask = ((uint64_t) 1 << sh) - 1 ;
	  const uint32_t r0 = generic ? blkStart[blk] : 0 ;
	  const uint64_t blkWord = (uint64_t) (blkBase + blk + 1) << 16 ;
	  uint64_t cur[H10X_P1_PF] ;
#pragma unroll
	  for (int u = 0 ; u < H10X_P1_PF ; ++u) cur[u] = pre[u] ;
	  if (blk + 1 < b1)
	    { const uint64_t so2 = srcOff[blk + 1] ; const uint32_t n2 = blkCnt[blk + 1] ;
#pragma unroll
	      for (int u = 0 ; u < H10X_P1_PF ; ++u)
		{ const uint32_t i = u * H10X_P1_THREADS + t ;
		  pre[u] = (!(so2 >> 63) && i < n2) ? __ldcs (scratch + (so2 & 0x00ffffffffffffffull) + i) : 0 ;
		}
	    }
	  auto place = [&] (uint64_t hash, uint32_t rd)
	    { const uint64_t q = hash * wInvFull ;
	      const uint32_t v = (uint32_t) (q >> lowBits) ;
	      const uint32_t pos = atomicAdd (&base[v], 1u) ;
	      const uint64_t E = ((q & lowMask) << eShift) | blkWord | (uint64_t) (rd & 0xffffu) ;
	      if (pos - (baseOld[v] & ~3u) < 8u) ring[h10x_ring_slot (v, pos & 7u)] = E ;
	      else out[pos] = E ;
	    } ;
	  if (!generic)
	    {
#pragma unroll
	      for (int u = 0 ; u < H10X_P1_PF ; ++u)
		if ((uint32_t) u * H10X_P1_THREADS + t < n) place (cur[u] >> sh, (uint32_t) (cur[u] & rmask)) ;
	      for (uint32_t i = H10X_P1_PF * H10X_P1_THREADS + t ; i < n ; i += H10X_P1_THREADS)
		{ const uint64_t key = __ldcs (scratch + o + i) ; place (key >> sh, (uint32_t) (key & rmask)) ; }
	    }
	  else
	    for (uint32_t i = t ; i < n ; i += H10X_P1_THREADS) place (gHash[o + i], gRec[o + i] - r0) ;
	  __syncthreads () ;
	  if ((blk - b0) % flushEvery == flushEvery - 1 || blk + 1 == b1)
	    { h10x_ring_flush (nRanges, H10X_P1_THREADS, baseOld, base, first, ring, out) ;	/* also: baseOld = base */
	      __syncthreads () ;
	    }
	}
      h10x_ring_drain (nRanges, H10X_P1_THREADS, base, first, ring, out) ;
    }
}

/* ---------------------------------------------------------------- the stable partition pass (P2, ids, T1, T2) */

/* element i -> (digit, word written).  LoadWord: a field of a 64-bit word; LoadCodes: the block number of
   codes[] with (id, read) beside it - the high block bits ride in the word's free top 16 bits. */
struct LoadWord {
  const uint64_t *in ; int shift ; uint32_t mask ; uint64_t outMask ;
  __device__ __forceinline__ uint32_t digit (uint64_t i) const { return (uint32_t) (in[i] >> shift) & mask ; }
  __device__ __forceinline__ void get (uint64_t i, uint32_t &d, uint64_t &v) const
  { const uint64_t x = in[i] ; d = (uint32_t) (x >> shift) & mask ; v = x & outMask ; }
} ;
struct LoadCodes {
  const uint32_t *codes ; const uint64_t *idRead ; uint32_t mask ; int b1 ;
  __device__ __forceinline__ uint32_t digit (uint64_t i) const { return codes[i] & mask ; }
  __device__ __forceinline__ void get (uint64_t i, uint32_t &d, uint64_t &v) const
  { const uint32_t c = codes[i] ; d = c & mask ; v = idRead[i] | ((uint64_t) (c >> b1) << 48) ; }
} ;

/* hist[d * strideBin + job * strideJob] = elements of job `job` with digit d */
/* Jobs are taken round robin, or - ticket != NULL - in order from a counter: the jobs of P2 are the hash ranges, whose
   sizes fall linearly from twice the average to nothing (density 2 (1 - x) of min (hash, hashRC)), so a fixed assignment
   left the CTA holding ranges 0, 592, 1184, 1776 with 30 % more than the average; handing the ranges out largest first
   as CTAs become free evens that out. */
__device__ __forceinline__ uint32_t h10x_next_job (unsigned int *ticket, uint32_t prev, bool first, uint32_t *sJob)
{ if (!ticket) return first ? blockIdx.x : prev + gridDim.x ;
  __syncthreads () ;
  if (threadIdx.x == 0) *sJob = atomicAdd (ticket, 1u) ;
  __syncthreads () ;
  return *sJob ;
}

template <class L>
__global__ void k_part_hist (L ld, const uint64_t *__restrict__ jobStart, uint32_t nJobs, uint32_t nBins,
			     uint64_t strideBin, uint64_t strideJob, uint32_t *__restrict__ hist, unsigned int *ticket)
{ extern __shared__ uint32_t phist[] ;
  __shared__ uint32_t sJob ;
  for (uint32_t job = h10x_next_job (ticket, 0, true, &sJob) ; job < nJobs ; job = h10x_next_job (ticket, job, false, &sJob))
    { for (uint32_t d = threadIdx.x ; d < nBins ; d += blockDim.x) phist[d] = 0 ;
      __syncthreads () ;
      const uint64_t a = jobStart[job], b = jobStart[job + 1] ;
      for (uint64_t i0 = a ; i0 < b ; i0 += 4ull * blockDim.x)
	{ uint32_t dg[4] ;
#pragma unroll
	  for (int u = 0 ; u < 4 ; ++u)
	    { const uint64_t i = i0 + (uint64_t) u * blockDim.x + threadIdx.x ;
	      dg[u] = (i < b) ? ld.digit (i) : 0xffffffffu ;
	    }
#pragma unroll
	  for (int u = 0 ; u < 4 ; ++u) if (dg[u] != 0xffffffffu) atomicAdd (&phist[dg[u]], 1u) ;
	}
      __syncthreads () ;
      for (uint32_t d = threadIdx.x ; d < nBins ; d += blockDim.x) hist[(uint64_t) d * strideBin + (uint64_t) job * strideJob] = phist[d] ;
      __syncthreads () ;
    }
}

/* off = exclusive scan of hist in memory order = where (job, digit) starts in the output.  A CTA takes a job in
   sub-tiles of NW*32*ITEMS elements; warp w owns ITEMS*32 consecutive elements of a sub-tile and ranks them, 32 at
   a time, against its private counters (match_any: equal digits of a step get consecutive ranks in lane order);
   a scan over the warps' counters turns them into offsets inside the sub-tile.  Output goes through the
   write-combining rings (h10x_ring_flush). */
__host__ __device__ inline size_t h10x_part_smem (uint32_t nBins, int nw)
{ return (size_t) nBins * 8 * H10X_RING_STRIDE + (size_t) nBins * 8 + (size_t) ((nBins + 1u) & ~1u) * 4 + (size_t) nw * nBins * 2 + 32 ; }

template <class L, int NW, int ITEMS, int MINB>
__global__ void __launch_bounds__ (NW * 32, MINB)
k_part_scatter (L ld, const uint64_t *__restrict__ jobStart, uint32_t nJobs, uint32_t nBins, uint64_t strideBin,
		uint64_t strideJob, const uint32_t *__restrict__ off, uint64_t *__restrict__ out, unsigned int *ticket)
{ __shared__ uint32_t sJob ; extern __shared__ __align__ (16) unsigned char psmRaw[] ;
  uint64_t *ring = (uint64_t*) psmRaw ;				/* nBins rings */
  uint32_t *cursor = (uint32_t*) (ring + (size_t) nBins * H10X_RING_STRIDE) ;	/* nBins: next output slot of every digit */
  uint32_t *first = cursor + nBins ;				/* nBins: the digit's first slot in this job */
  uint32_t *cursorNew = first + nBins ;				/* nBins (padded to even): cursor after this sub-tile */
  uint16_t *wc = (uint16_t*) (cursorNew + ((nBins + 1u) & ~1u)) ;	/* NW * nBins warp-private counters, then offsets */
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5 ;
  const uint32_t ltMask = (1u << lane) - 1u ;
  uint16_t *wcw = wc + (size_t) w * nBins ;
  uint32_t *wcWords = (uint32_t*) wc ;
  const uint32_t nWcWords = (NW * nBins + 1u) / 2u ;
  constexpr uint32_t TILE = NW * 32 * ITEMS ;
  for (uint32_t job = h10x_next_job (ticket, 0, true, &sJob) ; job < nJobs ; job = h10x_next_job (ticket, job, false, &sJob))
    { const uint64_t a = jobStart[job], b = jobStart[job + 1] ;
      __syncthreads () ;
      for (uint32_t d = t ; d < nBins ; d += NW * 32)
	{ const uint32_t o = off[(uint64_t) d * strideBin + (uint64_t) job * strideJob] ; cursor[d] = o ; first[d] = o ; }
      for (uint32_t x = t ; x < nWcWords ; x += NW * 32) wcWords[x] = 0 ;
      __syncthreads () ;
      for (uint64_t t0 = a ; t0 < b ; t0 += TILE)
	{ uint32_t dr[ITEMS] ; uint64_t val[ITEMS] ;	/* dr: digit, later digit | rank << 16 */
	  const uint64_t mine = t0 + (uint64_t) w * (32 * ITEMS) + lane ;
#pragma unroll
	  for (int s = 0 ; s < ITEMS ; ++s)
	    { const uint64_t i = mine + 32u * s ;
	      dr[s] = 0xffffu ; val[s] = 0 ;
	      if (i < b) ld.get (i, dr[s], val[s]) ;
	    }
	  uint32_t mm[ITEMS] ;
#pragma unroll
	  for (int s = 0 ; s < ITEMS ; ++s) mm[s] = __match_any_sync (0xffffffffu, dr[s]) ;	/* independent: their latencies overlap */
#pragma unroll
	  for (int s = 0 ; s < ITEMS ; ++s)
	    { const uint32_t d = dr[s], m = mm[s] ;
	      const int leader = __ffs (m) - 1 ;
	      uint32_t old = 0 ;
	      if ((int) lane == leader && d != 0xffffu) { old = wcw[d] ; wcw[d] = (uint16_t) (old + __popc (m)) ; }
	      old = __shfl_sync (0xffffffffu, old, leader) ;
	      dr[s] = d | ((old + __popc (m & ltMask)) << 16) ;
	      __syncwarp () ;
	    }
	  __syncthreads () ;
	  for (uint32_t d = t ; d < nBins ; d += NW * 32)
	    { uint32_t run = 0 ;
#pragma unroll
	      for (int ww = 0 ; ww < NW ; ++ww) { const uint32_t c = wc[(size_t) ww * nBins + d] ; wc[(size_t) ww * nBins + d] = (uint16_t) run ; run += c ; }
	      cursorNew[d] = cursor[d] + run ;
	    }
	  __syncthreads () ;
#pragma unroll
	  for (int s = 0 ; s < ITEMS ; ++s)
	    { const uint32_t d = dr[s] & 0xffffu ;
	      if (d != 0xffffu)
		{ const uint32_t c0 = cursor[d], pos = c0 + wcw[d] + (dr[s] >> 16) ;
		  if (pos - (c0 & ~3u) < 8u) ring[h10x_ring_slot (d, pos & 7u)] = val[s] ;
		  else out[pos] = val[s] ;
		}
	    }
	  __syncthreads () ;
	  h10x_ring_flush (nBins, NW * 32, cursor, cursorNew, first, ring, out) ;	/* also: cursor = cursorNew */
	  for (uint32_t x = t ; x < nWcWords ; x += NW * 32) wcWords[x] = 0 ;
	  __syncthreads () ;
	}
      h10x_ring_drain (nBins, NW * 32, cursor, first, ring, out) ;
    }
}

/* ---------------------------------------------------------------- S: sub-range sort + first entry of every hash */

struct SrArgs {
  uint64_t *A ;			/* entries, sub-range-major; sorted in place */
  const uint32_t *srStart ;	/* nSub + 1 */
  uint32_t *stage ;		/* first entries (global positions) of sub-range j's hashes, at stage[srStart[j] ..) */
  uint32_t *nHeads ;		/* nSub + 1 (the last one stays 0 for the scan) */
  uint32_t *srCount ;		/* nSub: entries a job kept (a lean fused kernel leaves the duplicates of a (hash, block) in) */
  uint32_t *blkDup ;		/* per (global, 1-based) block number: entries dropped as duplicates */
  unsigned long long *nDup ;	/* ... in total */
  uint32_t blkMask ;
  uint32_t *overList ;		/* sub-ranges left to the library sort */
  uint32_t *jobList ;		/* first sub-range | log2 (sub-ranges) << 30 */
  unsigned int *nOver ;
  unsigned int *ticket ;
  unsigned int *nJobs ;
  uint32_t nSub, cap, groupCap, maxLog ;
  int eShift, remBits ;
} ;

/* sort jobs: an aligned quad of sub-ranges when it holds at most groupCap entries, else its pairs, else single
   sub-ranges (maxLog = min (2, p2) keeps a job inside one range: its words then differ only in the q bits kept
   in E).  A job of 2^m sub-ranges sorts on remBits + m bits.  The order of the list is irrelevant. */
__global__ void k_sr_jobs (SrArgs a)
{ const uint32_t quad = blockIdx.x * blockDim.x + threadIdx.x ;
  const uint32_t j0 = quad * 4 ;
  if (j0 >= a.nSub) return ;
  uint32_t b[5] ;
#pragma unroll
  for (int x = 0 ; x < 5 ; ++x) b[x] = a.srStart[min (j0 + x, a.nSub)] ;
  const uint32_t tot = b[4] - b[0] ;
  if (!tot) return ;
  if (a.maxLog >= 2 && tot <= a.groupCap) { a.jobList[atomicAdd (a.nJobs, 1u)] = j0 | (2u << 30) ; return ; }
#pragma unroll
  for (int h = 0 ; h < 2 ; ++h)
    { const uint32_t pt = b[2*h + 2] - b[2*h] ;
      if (!pt || j0 + 2*h >= a.nSub) continue ;
      if (a.maxLog >= 1 && pt <= a.groupCap) { a.jobList[atomicAdd (a.nJobs, 1u)] = (j0 + 2*h) | (1u << 30) ; continue ; }
#pragma unroll
      for (int x = 2*h ; x < 2*h + 2 ; ++x)
	if (b[x+1] != b[x] && j0 + x < a.nSub) a.jobList[atomicAdd (a.nJobs, 1u)] = j0 + x ;
    }
}

__device__ __forceinline__ uint32_t sr_cta_exclusive_scan (uint32_t v, uint32_t *warpTmp /* 33 */, uint32_t &total)
{ const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5 ;
  uint32_t inc = v ;
#pragma unroll
  for (int d = 1 ; d < 32 ; d <<= 1) { uint32_t u = __shfl_up_sync (0xffffffffu, inc, d) ; if (lane >= (uint32_t) d) inc += u ; }
  if (lane == 31) warpTmp[wid] = inc ;
  __syncthreads () ;
  if (wid == 0)
    { uint32_t x = (lane < blockDim.x / 32) ? warpTmp[lane] : 0, ix = x ;
#pragma unroll
      for (int d = 1 ; d < 32 ; d <<= 1) { uint32_t u = __shfl_up_sync (0xffffffffu, ix, d) ; if (lane >= (uint32_t) d) ix += u ; }
      if (lane < blockDim.x / 32) warpTmp[lane] = ix - x ;
      if (lane == 31) warpTmp[32] = ix ;
    }
  __syncthreads () ;
  const uint32_t r = warpTmp[wid] + inc - v ;
  total = warpTmp[32] ;
  __syncthreads () ;
  return r ;
}

template <int H10X_SR_THREADS>
__global__ void __launch_bounds__ (H10X_SR_THREADS, 1024 / H10X_SR_THREADS)
k_sr_sort (SrArgs a)
{ extern __shared__ __align__ (16) unsigned char srRaw[] ;
  constexpr int NW = H10X_SR_THREADS / 32 ;
  constexpr int STEPS = (H10X_SR_CAP / NW + 31) / 32 ;		/* 32-entry steps of a warp's piece */
  uint64_t *buf0 = (uint64_t*) srRaw, *buf1 = buf0 + a.cap ;
  uint32_t *binStart = (uint32_t*) (buf1 + a.cap) ;		/* up to 2^H10X_SR_DIGIT */
  uint16_t *wc = (uint16_t*) (binStart + (1u << H10X_SR_DIGIT)) ;	/* NW * nBins */
  __shared__ uint32_t warpTmp[33], sJob ;
  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5 ;
  const uint32_t ltMask = (1u << lane) - 1u ;
  uint32_t *wcWords = (uint32_t*) wc ;
  const uint32_t nJobs = *a.nJobs ;
  for (;;)
    { __syncthreads () ;
      if (t == 0) sJob = atomicAdd (a.ticket, 1u) ;
      __syncthreads () ;
      if (sJob >= nJobs) break ;
      const uint32_t jl = a.jobList[sJob] ;
      const uint32_t j = jl & 0x3fffffffu, lm = jl >> 30 ;
      const uint32_t st = a.srStart[j], n = a.srStart[min (j + (1u << lm), a.nSub)] - st ;
      if (n > a.cap)		/* only a single sub-range can be: left to the library sort */
	{ if (t == 0) a.overList[atomicAdd (a.nOver, 1u)] = j ;
	  continue ;
	}
      const int keyBits = a.remBits + (int) lm ;
      const int passes = (keyBits + H10X_SR_DIGIT - 1) / H10X_SR_DIGIT ;
      const int digitBits = passes ? (keyBits + passes - 1) / passes : 1 ;
      const uint32_t nBins = 1u << digitBits ;
      uint16_t *wcw = wc + (size_t) w * nBins ;
      const uint32_t nWcWords = NW * nBins / 2u ;
      uint64_t *g = a.A + st ;
      for (uint32_t i = t ; i < n ; i += H10X_SR_THREADS) buf0[i] = g[i] ;
      uint64_t *src = buf0, *dst = buf1 ;
      const uint32_t per = (((n + NW - 1) / NW) + 31u) & ~31u ;	/* a warp's piece: consecutive entries */
      const uint32_t lo = min (w * per, n), hi = min (lo + per, n) ;
      for (int p = 0 ; p < passes ; ++p)
	{ const int shift = a.eShift + p * digitBits ;
	  const uint32_t mask = (uint32_t) min (digitBits, keyBits - p * digitBits) ;
	  const uint32_t dmask = (1u << mask) - 1u ;
	  for (uint32_t x = t ; x < nWcWords ; x += H10X_SR_THREADS) wcWords[x] = 0 ;
	  __syncthreads () ;		/* also: buf0 loaded / previous scatter done */
	  uint64_t val[STEPS] ; uint32_t dr[STEPS] ;
	  uint32_t mm[STEPS] ;
#pragma unroll
	  for (int s = 0 ; s < STEPS ; ++s)
	    { if (lo + 32u * s >= hi) break ;
	      const uint32_t i = lo + 32u * s + lane ;
	      val[s] = 0 ; dr[s] = 0xffffu ;
	      if (i < hi) { val[s] = src[i] ; dr[s] = (uint32_t) (val[s] >> shift) & dmask ; }
	      mm[s] = __match_any_sync (0xffffffffu, dr[s]) ;
	    }
#pragma unroll
	  for (int s = 0 ; s < STEPS ; ++s)
	    { if (lo + 32u * s >= hi) break ;
	      const uint32_t d = dr[s] ;
	      const uint32_t m = mm[s] ;
	      const int leader = __ffs (m) - 1 ;
	      uint32_t old = 0 ;
	      if ((int) lane == leader && d != 0xffffu) { old = wcw[d] ; wcw[d] = (uint16_t) (old + __popc (m)) ; }
	      old = __shfl_sync (0xffffffffu, old, leader) ;
	      dr[s] = d | ((old + __popc (m & ltMask)) << 16) ;
	      __syncwarp () ;
	    }
	  __syncthreads () ;
	  /* per digit: offsets of the warps' pieces, and the digit's total */
	  uint32_t myTot = 0 ;
	  if (t < nBins)
	    { uint32_t run = 0 ;
#pragma unroll
	      for (int ww = 0 ; ww < NW ; ++ww) { const uint32_t c = wc[(size_t) ww * nBins + t] ; wc[(size_t) ww * nBins + t] = (uint16_t) run ; run += c ; }
	      myTot = run ;
	    }
	  uint32_t total ;
	  const uint32_t ex = sr_cta_exclusive_scan (myTot, warpTmp, total) ;	/* nBins <= threads */
	  if (t < nBins) binStart[t] = ex ;
	  __syncthreads () ;
#pragma unroll
	  for (int s = 0 ; s < STEPS ; ++s)
	    { if (lo + 32u * s >= hi) break ;
	      const uint32_t d = dr[s] & 0xffffu ;
	      if (d != 0xffffu) dst[binStart[d] + wcw[d] + (dr[s] >> 16)] = val[s] ;
	    }
	  uint64_t *x = src ; src = dst ; dst = x ;
	  __syncthreads () ;		/* the counters are cleared again by the next pass */
	}
      __syncthreads () ;
      /* one entry per (hash, block): equal ones stand side by side (the passes above are stable and the entries came
	 block-ascending), the lowest read index is kept (hash10x.c:166-172: the stable qsort leaves the first-generated
	 mosh of a hash first).  First entry of every hash: q's kept bits change (the sub-range fixes the bits above). */
      uint32_t cntFH = 0 ;		/* kept entries | heads << 16 (n <= cap < 2^16) */
      const uint32_t perT = (n + H10X_SR_THREADS - 1) / H10X_SR_THREADS ;
      const uint32_t l0 = min (t * perT, n), l1 = min (l0 + perT, n) ;
      for (uint32_t i = l0 ; i < l1 ; ++i)
	{ const uint64_t e = src[i], p = i ? src[i - 1] : 0 ;
	  if (i == 0 || (e >> 16) != (p >> 16)) cntFH += 1u + ((i == 0 || (e >> a.eShift) != (p >> a.eShift)) ? 0x10000u : 0u) ;
	}
      uint32_t total ;
      const uint32_t at = sr_cta_exclusive_scan (cntFH, warpTmp, total) ;
      uint32_t atF = at & 0xffffu, atH = at >> 16 ;
      for (uint32_t i = l0 ; i < l1 ; ++i)
	{ const uint64_t e = src[i], p = i ? src[i - 1] : 0 ;
	  if (i == 0 || (e >> 16) != (p >> 16))
	    { uint32_t r = (uint32_t) e & 0xffffu ;
	      for (uint32_t k = i + 1 ; k < n && (src[k] >> 16) == (e >> 16) ; ++k) r = min (r, (uint32_t) src[k] & 0xffffu) ;
	      dst[atF] = (e & ~(uint64_t) 0xffffu) | r ;
	      if (i == 0 || (e >> a.eShift) != (p >> a.eShift)) a.stage[st + atH++] = st + atF ;
	      ++atF ;
	    }
	  else atomicAdd (&a.blkDup[(uint32_t) (e >> 16) & a.blkMask], 1u) ;
	}
      __syncthreads () ;
      const uint32_t nF = total & 0xffffu ;
      if (t == 0) { a.nHeads[j] = total >> 16 ; a.srCount[j] = nF ; if (nF != n) atomicAdd (a.nDup, (unsigned long long) (n - nF)) ; }
      for (uint32_t i = t ; i < nF ; i += H10X_SR_THREADS) g[i] = dst[i] ;
    }
}

/* sub-ranges too large for shared memory (one hash held by thousands of blocks: repeats) are sorted by the library,
   all of them in one segmented sort over a compact copy: segment x of the copy is sub-range list[x] */
__global__ void k_over_copy (uint64_t *__restrict__ A, const uint32_t *__restrict__ srStart, const uint32_t *__restrict__ list,
			     const uint64_t *__restrict__ segOff, uint64_t *__restrict__ compact, int toCompact)
{ const uint32_t j = list[blockIdx.x] ;
  const uint32_t st = srStart[j], n = srStart[j + 1] - st ;
  uint64_t *c = compact + segOff[blockIdx.x] ;
  if (toCompact) for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x) c[i] = A[st + i] ;
  else for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x) A[st + i] = c[i] ;
}

/* the same scan for a sub-range the library sorted in global memory: one CTA per listed sub-range, in chunks */
__global__ void k_sr_heads_big (SrArgs a, const uint32_t *__restrict__ list)
{ __shared__ uint32_t warpTmp[33] ;
  const uint32_t j = list[blockIdx.x] ;
  const uint32_t st = a.srStart[j], n = a.srStart[j + 1] - st ;
  uint64_t *g = a.A + st ;		/* sorted on the whole word: (hash, block, read) */
  uint32_t doneF = 0, doneH = 0 ;
  for (uint32_t c0 = 0 ; c0 < n ; c0 += blockDim.x)
    { const uint32_t i = c0 + threadIdx.x ;
      const uint64_t e = i < n ? g[i] : 0, p = (i && i < n) ? g[i - 1] : 0 ;
      const bool first = i < n && (i == 0 || (e >> 16) != (p >> 16)) ;
      const bool head = first && (i == 0 || (e >> a.eShift) != (p >> a.eShift)) ;
      uint32_t total ;
      const uint32_t at = sr_cta_exclusive_scan ((first ? 1u : 0u) | (head ? 0x10000u : 0u), warpTmp, total) ;	/* barriers: all have read */
      if (first) g[doneF + (at & 0xffffu)] = e ;	/* at or before i; the first of a group carries the lowest read index */
      else if (i < n) atomicAdd (&a.blkDup[(uint32_t) (e >> 16) & a.blkMask], 1u) ;
      if (head) a.stage[st + doneH + (at >> 16)] = st + doneF + (at & 0xffffu) ;
      doneF += total & 0xffffu ; doneH += total >> 16 ;
      __syncthreads () ;
    }
  if (threadIdx.x == 0) { a.nHeads[j] = doneH ; a.srCount[j] = doneF ; if (doneF != n) atomicAdd (a.nDup, (unsigned long long) (n - doneF)) ; }
}

/* ---------------------------------------------------------------- bins in hash order -> ids */

/* one warp per sub-range: its hashes' first positions move to their place in the global (hash-ordered) bin list;
   the bin's hash value and its sort key (first block << 32 | bin) come from the first entry */
__global__ void k_heads_compact (uint32_t nSub, const uint32_t *__restrict__ srStart, const uint32_t *__restrict__ nHeads,
				 const uint32_t *__restrict__ srCount, const uint32_t *__restrict__ binBase, const uint32_t *__restrict__ stage,
				 const uint64_t *__restrict__ A, int eShift, int lowBits, int p2, uint32_t blkMask, uint64_t wMul,
				 uint32_t *__restrict__ segStart, uint32_t *__restrict__ segLen, uint64_t *__restrict__ fkey,
				 uint64_t *__restrict__ hvHash, uint32_t *__restrict__ firstBlk /* or NULL: the first block alone (multi-GPU) */)
{ const uint32_t lane = threadIdx.x & 31 ;
  const uint64_t nw = ((uint64_t) gridDim.x * blockDim.x) >> 5 ;
  for (uint64_t j = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ; j < nSub ; j += nw)
    { const uint32_t nh = nHeads[j] ;
      if (!nh) continue ;
      const uint32_t bb = binBase[j], st = srStart[j], end = st + srCount[j] ;
      const uint64_t qTop = (lowBits >= 64) ? 0 : ((uint64_t) (j >> p2) << lowBits) ;
      for (uint32_t x = lane ; x < nh ; x += 32)
	{ const uint32_t pos = stage[st + x] ;
	  const uint64_t E = A[pos] ;
	  const uint32_t s = bb + x ;
	  segStart[s] = pos ;
	  segLen[s] = ((x + 1 < nh) ? stage[st + x + 1] : end) - pos ;	/* the job's kept entries are contiguous from st */
	  fkey[s] = ((uint64_t) ((uint32_t) (E >> 16) & blkMask) << 32) | s ;
	  if (firstBlk) firstBlk[s] = (uint32_t) (E >> 16) & blkMask ;
	  hvHash[s] = (qTop | (E >> eShift)) * wMul ;
	}
    }
}

/* rank r in (first block, hash) order is bin id r + 1 (hash10x.c:147) */
__global__ void k_bins_by_rank_e (uint32_t nSeg, const uint64_t *__restrict__ sortedKey, const uint32_t *__restrict__ segLen,
				  const uint64_t *__restrict__ hvHash, uint32_t *__restrict__ idOfSeg,
				  uint64_t *__restrict__ hashValue, uint32_t *__restrict__ hashDepth)
{ uint32_t r = blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= nSeg) return ;
  const uint32_t s = (uint32_t) sortedKey[r], id = r + 1u ;
  idOfSeg[s] = id ;
  hashValue[id] = hvHash[s] ;
  hashDepth[id] = segLen[s] ;	/* one entry per (block, hash): hash10x.c:178 */
}

/* multi-GPU: the global bin ids came back per local bin (hash order); the local lists are laid out in id order so that
   the transposition passes leave every block's list sorted by id.  w = id << 32 | local bin */
__global__ void k_id_words (uint32_t n, const uint32_t *__restrict__ gid, uint64_t *__restrict__ w)
{ uint32_t s = blockIdx.x * blockDim.x + threadIdx.x ; if (s < n) w[s] = ((uint64_t) gid[s] << 32) | s ; }

__global__ void k_local_ranks (uint32_t n, const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ segLen,
			       uint32_t *__restrict__ rankOfSeg, uint32_t *__restrict__ idByRank, uint32_t *__restrict__ depthByRank)
{ uint32_t r = blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= n) return ;
  const uint64_t w = sorted[r] ;
  const uint32_t s = (uint32_t) w ;
  rankOfSeg[s] = r + 1u ; idByRank[r] = (uint32_t) (w >> 32) ; depthByRank[r + 1] = segLen[s] ;
}

__global__ void k_u64_to_u32_from (const uint64_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ out)
{ uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ; if (i < n) out[i] = (uint32_t) in[i] ; }

/* fillHashTable (hash10x.c:317-347) + the transposed view: as k_codes_seg_kv, on E words */
__global__ void k_codes_seg_e (uint32_t nSeg, const uint32_t *__restrict__ segStart, const uint32_t *__restrict__ segLen,
			       const uint32_t *__restrict__ idOfSeg, const uint32_t *__restrict__ payloadId /* or NULL = idOfSeg */,
			       const uint64_t *__restrict__ A, uint32_t blkMask, const uint64_t *__restrict__ codeOff,
			       uint32_t *__restrict__ codes, uint64_t *__restrict__ idRead)
{ const uint32_t lane = threadIdx.x & 31 ;
  const uint64_t warp = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
  const uint64_t s0 = warp * 32 ;
  if (s0 >= nSeg) return ;
  const uint32_t cnt = (uint32_t) min ((uint64_t) 32, (uint64_t) nSeg - s0) ;
  uint32_t myI0 = 0, myN = 0, myId = 0 ; uint64_t myDst = 0 ;
  if (lane < cnt)
    { const uint32_t s = (uint32_t) s0 + lane ;
      myI0 = segStart[s] ; myN = segLen[s] ; myId = idOfSeg[s] ; myDst = codeOff[myId] ;
      if (payloadId) myId = payloadId[s] ;	/* multi-GPU: lists are placed by the rank-local order, entries carry the global bin id */
    }
  uint32_t i0 = __shfl_sync (0xffffffffu, myI0, 0), n = __shfl_sync (0xffffffffu, myN, 0) ;
  uint64_t cur = (lane < n) ? A[(uint64_t) i0 + lane] : 0 ;
  for (uint32_t t = 0 ; t < cnt ; ++t)
    { const uint32_t id = __shfl_sync (0xffffffffu, myId, t) ;
      const uint64_t dst = __shfl_sync (0xffffffffu, myDst, t) ;
      const uint32_t tn = (t + 1 < cnt) ? t + 1 : t ;
      const uint32_t i0n = __shfl_sync (0xffffffffu, myI0, tn), nn = __shfl_sync (0xffffffffu, myN, tn) ;
      uint64_t nxt = (t + 1 < cnt && lane < nn) ? A[(uint64_t) i0n + lane] : 0 ;
      if (lane < n)
	{ codes[dst + lane] = (uint32_t) (cur >> 16) & blkMask ;
	  idRead[dst + lane] = (uint64_t) id | ((cur & 0xffffull) << 32) ;
	}
      for (uint32_t x = lane + 32 ; x < n ; x += 32)	/* deep bins */
	{ const uint64_t E = A[(uint64_t) i0 + x] ;
	  codes[dst + x] = (uint32_t) (E >> 16) & blkMask ;
	  idRead[dst + x] = (uint64_t) id | ((E & 0xffffull) << 32) ;
	}
      i0 = i0n ; n = nn ; cur = nxt ;
    }
}

/* ---------------------------------------------------------------- index digest (h10x_digest.h) */
#include "h10x_digest.h"

template <class T>
__global__ void k_digest (const T *__restrict__ a, uint64_t n, uint64_t base, uint64_t mask, unsigned long long *__restrict__ out)
{ uint64_t d = 0 ;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; i < n ; i += (uint64_t) gridDim.x * blockDim.x)
    d += h10x_dg_term (base + i, (uint64_t) a[i] & mask) ;
#pragma unroll
  for (int o = 16 ; o ; o >>= 1) d += __shfl_xor_sync (0xffffffffu, d, o) ;
  if ((threadIdx.x & 31) == 0 && d) atomicAdd (out, (unsigned long long) d) ;
}
