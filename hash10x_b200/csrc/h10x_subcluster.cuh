/* h10x_subcluster.cuh - `--cluster codeMin codeMax` on the resident index ("next" row f2 of SURVEY.md section 8):
 * codeClusterFind (hash10x.c:770-835) + codeClusterReadMerge (hash10x.c:837-868), one persistent 1024-thread CTA per
 * SM taking whole barcode blocks.
 *
 * The reference walks a block's good hashes i = 1 .. n-1 in order and keeps, per other barcode cj, the first step
 * at which cj shared a hash with this block (minShare[cj] = i + 1).  Nothing in that walk is really sequential:
 *
 *   A  minShare is a minimum, so every (good hash i, barcode cj holding it) pair does an atomic min into an
 *      open-addressing table cj -> min (i + 1), all 32 warps in parallel.  The table lives in SHARED memory (16K
 *      entries); a block that shares hashes with more barcodes than fit is redone with a global-memory table sized
 *      for all blocks.  Entries carry a 16-bit stamp (one per block handled by the CTA), so neither table is ever
 *      cleared between blocks: an entry with another stamp is empty.
 *   B  a warp per step i counts, over the barcodes of hash i, how many have minShare - 1 == j for every j < i
 *      (barcodes first seen at step i sit at index i and never enter, hash10x.c:793-806), and keeps the largest
 *      count (first j wins ties, :803) and the total: (msBest, msMax, msTot) of step i.
 *   C  the label logic (:807-824) is a forest: a step with msMax >= clusterThreshold points at the earlier hash
 *      msBest and takes its cluster; a hash that is pointed at without having a cluster of its own founds one.  So a
 *      step's cluster is the ROOT of its chain of msBest pointers, roots are exactly the hashes that are not such
 *      steps themselves, cluster numbers are handed out in the order in which roots are first pointed at, and the
 *      hash whose count enters pointToMin (clusterMin[], :820-823) is that root.  Pointer jumping finds the roots, an
 *      atomic min per root its first direct child, a scan over the steps the numbering; the 256th founding step
 *      abandons the block (:810-817: every label 0, pointToMin keeps the terms of the earlier steps).  All warps then
 *      compute the steps' terms count / msTot (a recount only where the root is not msBest) and one warp adds them
 *      in step order, i.e. in the reference's order of double additions.
 *   E  codeClusterReadMerge is connected components (sub-clusters joined through shared reads, the smallest label
 *      wins, :848-858): min-label propagation over (sub-cluster, read) edges with pointer jumping, then the
 *      reference's renumbering of the surviving labels (:862-865).
 *
 * Per-step arrays sit in shared memory for blocks of up to 8192 good hashes and in global memory above that; bins
 * deeper than 128 barcodes count through per-warp global counters instead of the warp's shared-memory buffer.
 * Labels left by an earlier --cluster on entries outside the current good lists that exceed the block's new
 * nSubCluster count as 0 (the reference reads past trueCluster[] there: undefined behaviour; same rule as the oracle).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define H10X_SC_THREADS 1024
#define H10X_SC_WARPS (H10X_SC_THREADS / 32)
#define H10X_SC_STEPS_SMEM 8192u	/* good hashes of a block whose per-step arrays fit in shared memory */
#define H10X_SC_READ_SMEM 4096u		/* read labels kept in shared memory up to this many read pairs */
#define H10X_SC_VBUF 128		/* per-warp buffer of a hash's barcodes' minShare values */
#define H10X_SC_GROUP 4			/* steps whose barcode lists a warp loads together */
#define H10X_SC_SMALL_CAP 16384u	/* entries of the shared-memory table (128 KB) */
#define H10X_SC_SMALL_SHIFT 18u		/* 32 - log2 (H10X_SC_SMALL_CAP) */
#define H10X_SC_SMALL_LIMIT 9800u	/* barcodes it takes before the block is redone with the global table */
/* dynamic shared memory: table | root pointers u16 (read labels later) | first child u32 | founder scan u16 | labels u8 */
#define H10X_SC_DYN_SMEM (H10X_SC_SMALL_CAP * 8 + H10X_SC_STEPS_SMEM * (2 + 4 + 2 + 1))

struct SubClusterArgs {
  unsigned long long *clus ;		/* ClusterHash as one word: bin id | read << 32 | subCluster << 48 */
  const uint64_t *blkOff ; const uint32_t *blkNHash, *blkNRead ;
  const uint32_t *hashDepth ; const uint64_t *codeOff ; const uint32_t *codes ;
  const uint64_t *goodOff ; const uint16_t *good ;
  uint32_t *nSub ; double *pointToMin ;
  uint32_t codeMin, codeMax ; int threshold ;
  uint32_t codeBase ;			/* codes[] holds global block numbers, this context's block b is codeBase + b (0 on one GPU) */
  unsigned int *work ;			/* ticket counter */
  unsigned long long *table ;		/* per CTA: tableCap entries (barcode << 32 | stamp << 16 | minShare), zeroed once */
  uint32_t tableCap, tableShift ;	/* power of two >= 2 * blocks; shift = 32 - log2 (tableCap) */
  uint32_t *pre ;			/* per CTA: 2 * 65536 words: depth and codes offset of every good hash's bin */
  uint32_t *cnt ;			/* per warp: 65536 counters, all zero between uses (bins deeper than H10X_SC_VBUF) */
  uint32_t *res ;			/* per CTA: 6 * 65536 words: msBest, msMax, msTot, root and (double) term of every step */
  /* per CTA, for blocks whose per-step arrays do not fit in shared memory: 65536 entries each */
  uint16_t *parG ; uint32_t *firstG ; uint16_t *fscanG ; uint8_t *gsubG ;
  int *readLabG ;			/* per CTA: 65536 ints, read labels when they do not fit in smem */
} ;

/* table words are read past L1 when the table is in global memory (they change under atomics of other warps) and
   as volatile shared loads otherwise */
__device__ __forceinline__ unsigned long long sc_table_load (const unsigned long long *tab, uint32_t h, bool inSmem)
{ return inSmem ? ((const volatile unsigned long long*) tab)[h] : __ldcg (tab + h) ; }

/* claims, overflow: shared-memory words of the CTA.  Every new barcode counts; past `limit` the overflow flag goes
   up and everybody leaves (the caller redoes the block with the global table, whose limit is never reached). */
__device__ __forceinline__ void sc_table_min (unsigned long long *tab, uint32_t mask, uint32_t shift, bool inSmem, uint32_t stamp,
					      uint32_t cj, uint32_t val, uint32_t *claims, volatile uint32_t *overflow, uint32_t limit)
{ const unsigned long long tag = ((unsigned long long) cj << 32) | ((unsigned long long) stamp << 16) ;
  const unsigned long long cand = tag | val ;
  uint32_t h = (cj * 0x9E3779B1u) >> shift ;
  for (;;)
    { unsigned long long old = sc_table_load (tab, h, inSmem) ;
      if ((old & ~0xffffull) == tag)			/* this barcode, this block */
	{ if (val < (uint32_t) (old & 0xffffull)) atomicMin (tab + h, cand) ;
	  return ;
	}
      if ((uint32_t) ((old >> 16) & 0xffffull) != stamp)	/* left by an earlier block, or never used: claim it */
	{ if (*overflow) return ;
	  if (atomicCAS (tab + h, old, cand) == old)
	    { if (atomicAdd (claims, 1u) >= limit) *overflow = 1u ;
	      return ;
	    }
	  continue ;					/* someone else took the slot: look at it again */
	}
      h = (h + 1) & mask ;
    }
}

/* minShare of a barcode that step A has entered (0 if it is not there, which cannot happen for a barcode of a
   scanned hash) */
__device__ __forceinline__ uint32_t sc_table_get (const unsigned long long *tab, uint32_t mask, uint32_t shift, bool inSmem,
						  uint32_t stamp, uint32_t cj)
{ const unsigned long long tag = ((unsigned long long) cj << 32) | ((unsigned long long) stamp << 16) ;
  uint32_t h = (cj * 0x9E3779B1u) >> shift ;
  for (uint32_t probes = 0 ; probes <= mask ; ++probes)
    { unsigned long long old = sc_table_load (tab, h, inSmem) ;
      if ((old & ~0xffffull) == tag) return (uint32_t) (old & 0xffffull) ;
      if ((uint32_t) ((old >> 16) & 0xffffull) != stamp) return 0 ;
      h = (h + 1) & mask ;
    }
  return 0 ;
}

__global__ void __launch_bounds__ (H10X_SC_THREADS, 1)
k_subcluster (SubClusterArgs a)
{
  extern __shared__ __align__ (16) unsigned char scSmem[] ;
  unsigned long long *const tabSmall = (unsigned long long*) scSmem ;
  unsigned char *const stepSmem = scSmem + (size_t) H10X_SC_SMALL_CAP * 8 ;
  __shared__ uint16_t vbuf[H10X_SC_WARPS][H10X_SC_VBUF] ;
  __shared__ int label[257], newLab[257] ;
  __shared__ uint32_t warpSum[H10X_SC_WARPS] ;
  __shared__ uint32_t sTicket, sNs, sChanged, sClaims, sOverflow, sTotal, sAbandonAt ;

  const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5 ;
  unsigned long long *const tabBig = a.table + (size_t) blockIdx.x * a.tableCap ;
  uint32_t *preNc = a.pre + (size_t) blockIdx.x * 2 * 65536, *preOff = preNc + 65536 ;
  uint32_t *cnt = a.cnt + ((size_t) blockIdx.x * H10X_SC_WARPS + w) * 65536 ;
  uint32_t *resBest = a.res + (size_t) blockIdx.x * 6 * 65536, *resMax = resBest + 65536, *resTot = resMax + 65536,
    *resCm = resTot + 65536 ;
  double *resTerm = (double*) (resCm + 65536) ;
  uint32_t stamp = 0 ;

  for (uint32_t i = t ; i < H10X_SC_SMALL_CAP ; i += H10X_SC_THREADS) tabSmall[i] = 0 ;	/* stamp 0 = never used */
  __syncthreads () ;

  for (;;)
    { if (t == 0) sTicket = atomicAdd (a.work, 1u) ;
      __syncthreads () ;
      const uint32_t code = a.codeMin + sTicket ;
      __syncthreads () ;
      if (code >= a.codeMax) break ;
      const uint32_t self = code + a.codeBase ;	/* this block's number in codes[]: global after a multi-GPU build */

      unsigned long long *ch = a.clus + a.blkOff[code] ;
      const uint16_t *g = a.good + a.goodOff[code] ;
      const uint32_t n = (uint32_t) (a.goodOff[code + 1] - a.goodOff[code]) ;
      const uint32_t nHash = a.blkNHash[code] ;

      if (n)	/* ---------------- codeClusterFind ---------------- */
	{ /* per-step arrays: par = msBest pointer, later the root; first = first direct child of a root; fscan = number of
	     founding steps up to here; gsub = the labels (ClusterHash.subCluster of the good entries) */
	  const bool stepsInSmem = n <= H10X_SC_STEPS_SMEM ;
	  volatile uint16_t *par = stepsInSmem ? (uint16_t*) stepSmem : a.parG + (size_t) blockIdx.x * 65536 ;
	  volatile uint32_t *first = stepsInSmem ? (uint32_t*) (stepSmem + H10X_SC_STEPS_SMEM * 2) : a.firstG + (size_t) blockIdx.x * 65536 ;
	  volatile uint16_t *fscan = stepsInSmem ? (uint16_t*) (stepSmem + H10X_SC_STEPS_SMEM * 6) : a.fscanG + (size_t) blockIdx.x * 65536 ;
	  volatile uint8_t *gsub = stepsInSmem ? (uint8_t*) (stepSmem + H10X_SC_STEPS_SMEM * 8) : a.gsubG + (size_t) blockIdx.x * 65536 ;
	  for (uint32_t i = t ; i < n ; i += H10X_SC_THREADS)
	    { par[i] = (uint16_t) i ; first[i] = 0xffffffffu ;
	      const uint32_t x = (uint32_t) ch[g[i]] ;				/* bin of good hash i: its depth and barcode list */
	      preNc[i] = a.hashDepth[x] ; preOff[i] = (uint32_t) a.codeOff[x] ;	/* fewer than 2^32 entries on a device */
	    }

	  /* A: minShare of every barcode that shares a good hash i >= 1 with this block; first in the shared-memory table */
	  unsigned long long *tab = tabSmall ;
	  uint32_t mask = H10X_SC_SMALL_CAP - 1, shift = H10X_SC_SMALL_SHIFT, limit = H10X_SC_SMALL_LIMIT ;
	  bool inSmem = true ;
	  for (int attempt = 0 ; attempt < 2 ; ++attempt)
	    { if (++stamp == 0x10000u)		/* 16-bit stamps used up: start over with clean tables */
		{ for (size_t i = t ; i < a.tableCap ; i += H10X_SC_THREADS) tabBig[i] = 0 ;
		  for (uint32_t i = t ; i < H10X_SC_SMALL_CAP ; i += H10X_SC_THREADS) tabSmall[i] = 0 ;
		  stamp = 1 ;
		}
	      if (t == 0) { sClaims = 0 ; sOverflow = 0 ; }
	      __syncthreads () ;
	      /* a warp takes 8 consecutive steps at a time: lanes 0-7 fetch their bin depth / list offset (one L2 round trip),
		 then the first 64 barcodes of 4 steps are loaded together (8 independent DRAM loads in flight per lane) before
		 any of them is worked on - the kernel is bound by the latency of these loads, not by their bytes */
	      for (uint32_t c0 = 1 + 8 * w ; c0 < n ; c0 += 8 * H10X_SC_WARPS)
		{ uint32_t myNc = 0, myOff = 0 ;
		  if (lane < 8 && c0 + lane < n) { myNc = __ldcg (preNc + c0 + lane) ; myOff = __ldcg (preOff + c0 + lane) ; }
		  for (uint32_t q0 = 0 ; q0 < 8 ; q0 += H10X_SC_GROUP)
		    { uint32_t nc[H10X_SC_GROUP], off[H10X_SC_GROUP], cja[H10X_SC_GROUP], cjb[H10X_SC_GROUP] ;
#pragma unroll
		      for (int q = 0 ; q < H10X_SC_GROUP ; ++q)
			{ nc[q] = __shfl_sync (0xffffffffu, myNc, q0 + q) ; off[q] = __shfl_sync (0xffffffffu, myOff, q0 + q) ; }
#pragma unroll
		      for (int q = 0 ; q < H10X_SC_GROUP ; ++q)	/* `self` stands for "no barcode here": it is skipped anyway */
			{ cja[q] = lane < nc[q] ? a.codes[(size_t) off[q] + lane] : self ;
			  cjb[q] = lane + 32 < nc[q] ? a.codes[(size_t) off[q] + 32 + lane] : self ;
			}
#pragma unroll
		      for (int q = 0 ; q < H10X_SC_GROUP ; ++q)
			{ const uint32_t i = c0 + q0 + q ;
			  if (cja[q] != self) sc_table_min (tab, mask, shift, inSmem, stamp, cja[q], i + 1, &sClaims, &sOverflow, limit) ;
			  if (cjb[q] != self) sc_table_min (tab, mask, shift, inSmem, stamp, cjb[q], i + 1, &sClaims, &sOverflow, limit) ;
			  for (uint32_t j = lane + 64 ; j < nc[q] ; j += 32)
			    { const uint32_t cj = a.codes[(size_t) off[q] + j] ;
			      if (cj != self) sc_table_min (tab, mask, shift, inSmem, stamp, cj, i + 1, &sClaims, &sOverflow, limit) ;
			    }
			}
		    }
		}
	      __syncthreads () ;
	      if (!sOverflow) break ;
	      __syncthreads () ;
	      tab = tabBig ; mask = a.tableCap - 1 ; shift = a.tableShift ; limit = 0xffffffffu ; inSmem = false ;
	    }

	  /* B: per step i the best earlier step, its count and the total (hash10x.c:793-806); the step points at it when
	     the count reaches the threshold (:807) */
	  for (uint32_t c0 = 1 + 8 * w ; c0 < n ; c0 += 8 * H10X_SC_WARPS)
	    { uint32_t myNc = 0, myOff = 0 ;
	      if (lane < 8 && c0 + lane < n) { myNc = __ldcg (preNc + c0 + lane) ; myOff = __ldcg (preOff + c0 + lane) ; }
	      for (uint32_t q0 = 0 ; q0 < 8 ; q0 += H10X_SC_GROUP)
		{ uint32_t ncs[H10X_SC_GROUP], off[H10X_SC_GROUP], cja[H10X_SC_GROUP], cjb[H10X_SC_GROUP] ;
#pragma unroll
		  for (int q = 0 ; q < H10X_SC_GROUP ; ++q)
		    { ncs[q] = __shfl_sync (0xffffffffu, myNc, q0 + q) ; off[q] = __shfl_sync (0xffffffffu, myOff, q0 + q) ; }
#pragma unroll
		  for (int q = 0 ; q < H10X_SC_GROUP ; ++q)
		    { cja[q] = lane < ncs[q] ? a.codes[(size_t) off[q] + lane] : self ;
		      cjb[q] = lane + 32 < ncs[q] ? a.codes[(size_t) off[q] + 32 + lane] : self ;
		    }
#pragma unroll
		  for (int q = 0 ; q < H10X_SC_GROUP ; ++q)
		    { const uint32_t i = c0 + q0 + q ;
		      if (i >= n) continue ;			/* warp-uniform */
		      const uint32_t nc = ncs[q] ;
		      const uint32_t *cl = a.codes + off[q] ;
		      uint32_t tot = 0, bMax = 0, bBest = 0xffffffffu ;
		      if (nc <= H10X_SC_VBUF)		/* the usual case: count equal values among the hash's barcodes in shared memory */
			{ for (uint32_t j = lane ; j < nc ; j += 32)
			    { const uint32_t cj = j < 32 ? cja[q] : j < 64 ? cjb[q] : cl[j] ;
			      uint32_t v = 0xffffu ;
			      if (cj != self) { v = sc_table_get (tab, mask, shift, inSmem, stamp, cj) - 1u ; if (v >= i) v = 0xffffu ; }
			      vbuf[w][j] = (uint16_t) v ;
			    }
			  __syncwarp () ;
			  for (uint32_t j = lane ; j < nc ; j += 32)
			    { const uint32_t v = vbuf[w][j] ;
			      const uint32_t r0 = j & ~31u ;				/* this round of 32 barcodes: equal values by one match ... */
			      const uint32_t live = nc - r0 >= 32u ? 0xffffffffu : (1u << (nc - r0)) - 1u ;
			      uint32_t c = __popc (__match_any_sync (live, v)) ;
			      if (v == 0xffffu) continue ;
			      ++tot ;
			      for (uint32_t k = 0 ; k < r0 ; ++k) c += (vbuf[w][k] == v) ? 1u : 0u ;	/* ... the other rounds one by one */
			      for (uint32_t k = r0 + 32 ; k < nc ; ++k) c += (vbuf[w][k] == v) ? 1u : 0u ;
			      if (c > bMax || (c == bMax && v < bBest)) { bMax = c ; bBest = v ; }
			    }
			}
		      else				/* a deep bin: per-warp counters in global memory */
			{ for (uint32_t j = lane ; j < nc ; j += 32)
			    { const uint32_t cj = cl[j] ;
			      uint32_t v = 0xffffu ;
			      if (cj != self) { v = sc_table_get (tab, mask, shift, inSmem, stamp, cj) - 1u ; if (v >= i) v = 0xffffu ; }
			      if (v != 0xffffu) { atomicAdd (cnt + v, 1u) ; ++tot ; }
			    }
			  __syncwarp () ;
			  for (uint32_t j = lane ; j < nc ; j += 32)
			    { const uint32_t cj = cl[j] ;
			      uint32_t v = 0xffffu ;
			      if (cj != self) { v = sc_table_get (tab, mask, shift, inSmem, stamp, cj) - 1u ; if (v >= i) v = 0xffffu ; }
			      if (v != 0xffffu)
				{ const uint32_t c = __ldcg (cnt + v) ;
				  if (c > bMax || (c == bMax && v < bBest)) { bMax = c ; bBest = v ; }
				}
			    }
			  __syncwarp () ;
			  for (uint32_t j = lane ; j < nc ; j += 32)	/* leave the counters zero for the next step */
			    { const uint32_t cj = cl[j] ;
			      uint32_t v = 0xffffu ;
			      if (cj != self) { v = sc_table_get (tab, mask, shift, inSmem, stamp, cj) - 1u ; if (v >= i) v = 0xffffu ; }
			      if (v != 0xffffu) cnt[v] = 0 ;
			    }
			}
#pragma unroll
		      for (int d = 16 ; d ; d >>= 1)
			{ const uint32_t oMax = __shfl_xor_sync (0xffffffffu, bMax, d), oBest = __shfl_xor_sync (0xffffffffu, bBest, d) ;
			  if (oMax > bMax || (oMax == bMax && oBest < bBest)) { bMax = oMax ; bBest = oBest ; }
			  tot += __shfl_xor_sync (0xffffffffu, tot, d) ;
			}
		      __syncwarp () ;
		      if (lane == 0)
			{ resBest[i] = bMax ? bBest : 0u ; resMax[i] = bMax ; resTot[i] = tot ;
			  if ((long long) bMax >= (long long) a.threshold) par[i] = (uint16_t) bBest ;
			}
		    }
		}
	    }
	  __syncthreads () ;

	  /* C: the labels (hash10x.c:807-824) as a forest - see the head of this file.
	     first[r] = the first step that points straight at root r */
	  for (uint32_t i = 1 + t ; i < n ; i += H10X_SC_THREADS)
	    { const uint32_t b = par[i] ;
	      if (b != i && par[b] == b) atomicMin ((uint32_t*) first + b, i) ;
	    }
	  if (t == 0) sAbandonAt = n ;
	  __syncthreads () ;
	  /* founding steps, counted in step order */
	  { const uint32_t per = (n + H10X_SC_THREADS - 1) / H10X_SC_THREADS ;
	    const uint32_t lo = min (t * per, n), hi = min (lo + per, n) ;
	    uint32_t sum = 0 ;
	    for (uint32_t i = lo ; i < hi ; ++i)
	      { const uint32_t b = par[i] ; sum += (b != i && par[b] == b && first[b] == i) ? 1u : 0u ; }
	    uint32_t inc = sum ;
#pragma unroll
	    for (int d = 1 ; d < 32 ; d <<= 1) { const uint32_t u = __shfl_up_sync (0xffffffffu, inc, d) ; if (lane >= d) inc += u ; }
	    if (lane == 31) warpSum[w] = inc ;
	    __syncthreads () ;
	    if (w == 0)
	      { const uint32_t v = warpSum[lane] ; uint32_t iv = v ;
#pragma unroll
		for (int d = 1 ; d < 32 ; d <<= 1) { const uint32_t u = __shfl_up_sync (0xffffffffu, iv, d) ; if (lane >= d) iv += u ; }
		warpSum[lane] = iv - v ;
		if (lane == 31) sTotal = iv ;
	      }
	    __syncthreads () ;
	    uint32_t run = warpSum[w] + inc - sum ;
	    for (uint32_t i = lo ; i < hi ; ++i)
	      { const uint32_t b = par[i] ;
		const uint32_t f = (b != i && par[b] == b && first[b] == i) ? 1u : 0u ;
		run += f ;
		fscan[i] = (uint16_t) run ;
		if (f && run == 256u) sAbandonAt = i ;		/* the step that would found cluster 256 (:810) */
	      }
	  }
	  __syncthreads () ;
	  /* roots by pointer jumping: pointers only ever move to an ancestor, so reading a half-updated one is fine */
	  for (int round = 0 ; round < 17 ; ++round)
	    { if (t == 0) sChanged = 0 ;
	      __syncthreads () ;
	      for (uint32_t i = t ; i < n ; i += H10X_SC_THREADS)
		{ const uint32_t p0 = par[i], pp = par[p0] ;
		  if (pp != p0) { par[i] = (uint16_t) pp ; sChanged = 1 ; }
		}
	      __syncthreads () ;
	      const bool again = sChanged != 0 ;
	      __syncthreads () ;
	      if (!again) break ;
	    }
	  { const uint32_t total = sTotal, abandonAt = sAbandonAt ;
	    const bool abandoned = total > 255u ;
	    for (uint32_t i = t ; i < n ; i += H10X_SC_THREADS)
	      { const uint32_t r = par[i] ;
		uint32_t lab = 0 ;
		if (!abandoned) { const uint32_t f = first[r] ; if (f != 0xffffffffu) lab = fscan[f] ; }
		gsub[i] = (uint8_t) lab ;
		resCm[i] = (r != i && i < abandonAt) ? r : 0xffffffffu ;	/* the step's clusterMin[], if it joined anything */
	      }
	    if (t == 0) a.nSub[code] = abandoned ? 0u : total ;
	  }
	  __syncthreads () ;
	  /* the terms minShareCount[clusterMin] / msTot (:823) in parallel ... */
	  for (uint32_t i = 1 + w ; i < n ; i += H10X_SC_WARPS)
	    { const uint32_t cm = __ldcg (resCm + i) ;
	      if (cm == 0xffffffffu) { if (lane == 0) resTerm[i] = 0.0 ; continue ; }
	      uint32_t cAt = __ldcg (resMax + i) ;
	      if (cm != __ldcg (resBest + i))		/* the cluster was founded by another hash: count that one at this step */
		{ const uint32_t nc = __ldcg (preNc + i) ;
		  const uint32_t *cl = a.codes + __ldcg (preOff + i) ;
		  cAt = 0 ;
		  for (uint32_t j = lane ; j < nc ; j += 32)
		    { const uint32_t cj = cl[j] ;
		      if (cj != self && sc_table_get (tab, mask, shift, inSmem, stamp, cj) - 1u == cm) ++cAt ;
		    }
#pragma unroll
		  for (int d = 16 ; d ; d >>= 1) cAt += __shfl_xor_sync (0xffffffffu, cAt, d) ;
		}
	      if (lane == 0) resTerm[i] = (double) (int) cAt / (double) (int) __ldcg (resTot + i) ;
	    }
	  __syncthreads () ;
	  /* ... and one warp adds them in step order, as the reference's double accumulation does (x + 0.0 == x for
	     the steps that joined nothing) */
	  if (w == 0)
	    { double ptm = 0.0 ;
	      for (uint32_t i0 = 1 ; i0 < n ; i0 += 32)
		{ const uint32_t mine = i0 + lane ;
		  const double v = mine < n ? __ldcg (resTerm + mine) : 0.0 ;
		  const uint32_t steps = min (32u, n - i0) ;
		  for (uint32_t k = 0 ; k < steps ; ++k) ptm += __shfl_sync (0xffffffffu, v, k) ;
		}
	      if (lane == 0) a.pointToMin[code] = ptm ;
	    }
	  for (uint32_t i = t ; i < n ; i += H10X_SC_THREADS) ((uint8_t*) (ch + g[i]))[6] = gsub[i] ;
	  __syncthreads () ;
	}

      /* ---------------- codeClusterReadMerge ---------------- */
      if (t == 0) sNs = a.nSub[code] ;
      __syncthreads () ;
      const uint32_t ns = sNs ;
      if (ns)
	{ const uint32_t nRead = a.blkNRead[code] ;
	  const uint32_t nR = min (nRead, 65536u) ;
	  int *readLab = (nR <= H10X_SC_READ_SMEM) ? (int*) stepSmem : a.readLabG + (size_t) blockIdx.x * 65536 ;
	  for (uint32_t s = t ; s <= 256 ; s += H10X_SC_THREADS) label[s] = s <= ns ? (int) s : 0 ;
	  __syncthreads () ;
	  for (;;)
	    { if (t == 0) sChanged = 0 ;
	      for (uint32_t r = t ; r < nR ; r += H10X_SC_THREADS) readLab[r] = 0x7fffffff ;
	      __syncthreads () ;
	      for (uint32_t e = t ; e < nHash ; e += H10X_SC_THREADS)
		{ const unsigned long long wd = __ldcg (ch + e) ;	/* byte 6 was just rewritten: read past L1 */
		  const uint32_t s = (uint32_t) (wd >> 48) & 0xffu ;
		  if (s && s <= ns) atomicMin (readLab + ((uint32_t) (wd >> 32) & 0xffffu), label[s]) ;
		}
	      __syncthreads () ;
	      for (uint32_t e = t ; e < nHash ; e += H10X_SC_THREADS)
		{ const unsigned long long wd = __ldcg (ch + e) ;	/* byte 6 was just rewritten: read past L1 */
		  const uint32_t s = (uint32_t) (wd >> 48) & 0xffu ;
		  if (s && s <= ns)
		    { const int m = readLab[(uint32_t) (wd >> 32) & 0xffffu] ;
		      if (m < label[s]) { atomicMin (label + s, m) ; sChanged = 1 ; }
		    }
		}
	      __syncthreads () ;
	      for (uint32_t s = 1 + t ; s <= ns ; s += H10X_SC_THREADS)	/* pointer jumping: labels only ever decrease */
		{ const int l = label[s], ll = label[l] ;
		  if (ll < l) { atomicMin (label + s, ll) ; sChanged = 1 ; }
		}
	      __syncthreads () ;
	      const bool again = sChanged != 0 ;
	      __syncthreads () ;
	      if (!again) break ;
	    }
	  /* surviving labels are the component minima; renumber them 1.. in order (hash10x.c:862-865) */
	  if (t == 0)
	    { int live = 0 ;
	      newLab[0] = 0 ;
	      for (uint32_t s = 1 ; s <= ns ; ++s) { if (label[s] == (int) s) ++live ; newLab[s] = live ; }	/* = deadCluster[] */
	      for (uint32_t s = 1 ; s <= ns ; ++s) label[s] = newLab[label[s]] ;
	      label[0] = 0 ;
	      a.nSub[code] = (uint32_t) live ;
	    }
	  __syncthreads () ;
	  for (uint32_t e = t ; e < nHash ; e += H10X_SC_THREADS)
	    { const uint32_t s = (uint32_t) (__ldcg (ch + e) >> 48) & 0xffu ;
	      ((uint8_t*) (ch + e))[6] = (uint8_t) (s <= ns ? label[s] : 0) ;
	    }
	  __syncthreads () ;
	}
    }
}
