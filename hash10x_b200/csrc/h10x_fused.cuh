/* h10x_fused.cuh - the hot kernel: persistent CTAs, each taking whole barcode blocks and doing
 * the first half of processBlock (hash10x.c:154-172) on chip:
 *
 *   ingest     one warp per read pair; lane l loads word l of the 30-word FQB record (coalesced,
 *              prefetched one pair ahead);
 *   windows    the pair's k-mers are split into chunks of 8 consecutive start positions, one chunk
 *              per lane (14 chunks of read 1, 17 of read 2 for k=21; the last chunk of a read is
 *              slid back to end at the last k-mer, the k-mers it shares with the chunk before are never selected).  The three
 *              32-bit words holding a chunk's 8+k-1 bases come from the owning lanes by warp
 *              shuffle and are funnel-shifted into one 64-bit window W; its 2-bit-reversed
 *              complement WR yields the reverse-complement k-mers, so there is no per-base rolling
 *              state and no unpacked base array;
 *   hash       h_j = (W >> (64-2k-2j)) & mask, hRC_j = (WR >> 2j) & mask, each times factor1
 *              (seqhash.c:58-59); canonical = the smaller top-2k-bit product (seqhash.c:67); a
 *              "mosh" iff divisible by w (seqhash.c:171,189), tested without division on the
 *              unshifted product m = hash << (64-2k);
 *   collect    a selected key (m | readIndex) is stored, predicated, to the lane's private column
 *              of a per-CTA staging area in global memory that the persistent CTA reuses for every
 *              block (so it lives in L2): no atomics and no divergence in the hash loop;
 *   sort       keys are gathered into shared-memory buckets on their top hash bits (histogram,
 *              scan, scatter) and each bucket is insertion-sorted: expected O(n), hashes are uniform;
 *   dedup      first key of each run of equal hashes = lowest read index, which is what the
 *              reference's stable qsort + dedup loop keeps (hash10x.c:166-172); an empty block
 *              yields the phantom {hash 0, read 0} entry (hash10x.c:167-168);
 *   output     the block's sorted unique list goes to a global scratch slab at an atomically
 *              reserved offset; k_place later moves it to its final place in block order.
 *
 * A block whose keys do not fit (more moshes than `cap`, a lane column or a bucket over-full from
 * low-complexity reads, scratch exhausted) is flagged H10X_BLK_FALLBACK and re-done by the generic
 * global-memory path, so results never depend on which path ran.
 */
#pragma once
#include "h10x_common.cuh"

#define H10X_BLK_FALLBACK 0xffffffffu
#define H10X_BUCKET_LIMIT 48u		/* longest bucket the in-bucket ordering will take (its cost is quadratic) */

/* complement of x with the two bits of every base swapped, ~(((x >> 1) & 0x5555..) | ((x << 1) & 0xaaaa..)): the bit
   select and the inversion are one LOP3 (lut = ~((a & c) | (b & ~c)) = 0x1b) */
__device__ __forceinline__ uint32_t h10x_swap_pairs_not (uint32_t x)
{ uint32_t r ;
  asm ("lop3.b32 %0, %1, %2, %3, 0x1b;" : "=r" (r) : "r" (x >> 1), "r" (x << 1), "r" (0x55555555u)) ;
  return r ;
}

/* low 64 bits of (xhi:xlo) * (fhi:flo) in three multiply instructions */
__device__ __forceinline__ uint64_t h10x_mul64 (uint32_t xlo, uint32_t xhi, uint32_t flo, uint32_t fhi)
{ uint32_t plo, phi ;
  asm ("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %4;\n\tmov.b64 {%0, %1}, t;\n\t"
       "mad.lo.u32 %1, %2, %5, %1;\n\tmad.lo.u32 %1, %3, %4, %1;\n\t}"
       : "=r" (plo), "=r" (phi) : "r" (xlo), "r" (xhi), "r" (flo), "r" (fhi)) ;
  return ((uint64_t) phi << 32) | plo ;
}

/* exclusive scan in place of a[0..n) by the whole CTA; returns the total to every thread.
   warpTmp: 33 words of shared memory. */
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
  uint64_t *stage ;		/* per-CTA staging columns: gridDim * THREADS * rowCap keys */
  uint64_t *scratch ;		/* unique-key slab */
  unsigned long long *cursor ;	/* next free scratch slot */
  unsigned int *work ;		/* ticket counter of this launch */
  uint64_t scratchCap ;
  uint64_t *srcOff ;		/* per block: where its list went */
  uint32_t *blkCnt ;		/* per block: unique count, or H10X_BLK_FALLBACK */
  uint32_t nList ;
  uint32_t cap ;		/* key capacity of the shared-memory bucket array */
  uint32_t nbuck ;		/* buckets (power of two, >= THREADS) */
  uint32_t lb ;			/* log2 (nbuck) */
  uint32_t rowCap ;		/* keys per lane column of the staging area */
  uint32_t c1, c2 ;		/* 8-k-mer chunks of read 1 / read 2 */
  uint32_t n1, n2 ;		/* k-mers of read 1 / read 2 */
  unsigned long long *moshCount ;	/* selected k-mers of the blocks this path completed (seqhash.c:171,189) */
} ;

#ifndef H10X_MULSHIFT_R
#define H10X_MULSHIFT_R 0xffu
#endif
#ifndef H10X_MULSHIFT_F
#define H10X_MULSHIFT_F 0xffu
#endif

/* CTAs per SM that the shared memory of a size class allows (h10x_gpu.cu, kClasses): the register budget follows it */
__host__ __device__ constexpr int h10x_fused_ctas_per_sm (int threads)
{ return threads <= 256 ? 5 : threads <= 384 ? 4 : threads <= 512 ? 2 : 1 ; }

/* K > 0: k fixed at compile time (shifts and masks become immediates); K == 0: k from hp.
   WODD: w is odd and at least 3 (the host sends w = 1 and even w to the other instantiation).
   LEAN: the block's selected keys leave unsorted and with their duplicates - the single-GPU tail (h10x_tail.cuh)
   sorts every key once, globally, and drops the duplicates of a (hash, block) there, keeping the lowest read index
   (k_sr_sort), so the in-CTA bucket sort + dedup below would be a second sort of the same keys. */
template <int THREADS, bool WODD, int K, bool LEAN>
__global__ void __launch_bounds__ (THREADS, h10x_fused_ctas_per_sm (THREADS))
k_fused_block (FusedArgs a, HashParams hp)
{
  extern __shared__ __align__ (16) unsigned char smemRaw[] ;
  uint64_t *S = (uint64_t*) smemRaw ;			/* cap keys, bucketed */
  uint32_t *start = (uint32_t*) (S + a.cap) ;		/* nbuck + 1: bucket counts, then starts, then ends */
  uint32_t *cur = start ;				/* the dedup scan reuses it */
  __shared__ uint32_t sCount, sBad, sTicket, warpTmp[33] ;
  __shared__ unsigned long long sBase ;

  const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5 ;
  const int kk = K ? K : hp.k ;
  const int SH = 64 - 2 * kk ;				/* low zero bits of a masked product */
  const uint32_t HMASK = (kk > 16) ? ((1u << (2 * kk - 32)) - 1u) : 0u ;	/* high word of the k-mer mask */
  const uint32_t LMASK = (kk >= 16) ? 0xffffffffu : ((1u << (2 * kk)) - 1u) ;	/* low word of the k-mer mask */
  const uint32_t TOPLO = (SH >= 32) ? 0u : ~((1u << SH) - 1u) ;		/* low word of the product mask */
  const uint32_t TOPHI = (SH > 32) ? ~((1u << (SH - 32)) - 1u) : 0xffffffffu ;
  const uint32_t flo = (uint32_t) hp.factor1, fhi = (uint32_t) (hp.factor1 >> 32) ;
  const uint32_t ilo = (uint32_t) hp.wInv, ihi = (uint32_t) (hp.wInv >> 32) ;
  const uint64_t wLim = hp.wLim ;
  const uint64_t tzMaskSh = hp.wTzMask << SH ;
  const int buckShift = 64 - (int) a.lb ;

  /* ---- lane's chunk: which read, first k-mer start position, source lanes of its 3 words ---- */
  const bool isR2 = lane >= a.c1 ;
  const bool laneActive = lane < a.c1 + a.c2 ;
  const uint32_t nk = isR2 ? a.n2 : a.n1 ;			/* k-mers of this read */
  const uint32_t chunk = isR2 ? lane - a.c1 : lane ;
  const uint32_t first = min (8u * chunk, nk - 8u) ;		/* the last chunk slides back: all 8 valid */
  /* ... and its first dupSkip k-mers belong to the chunk before: they are never selected, so every selected k-mer is
     stored once and the block's key count is its mosh count (seqhash.c:171,189).  The read ranges are fixed
     (hash10x.c:162-163), so for a compile-time K the two tail lanes and their dupSkip are constants and the test folds
     into the predicate input of the selection compare: no instruction in the hash loop, where a compare + add per
     k-mer on the integer ALU pipe - the kernel's limiter - cost 14 ms of 79 at the 1 Gb workload */
  const uint32_t dupSkip = 8u * chunk - first ;
  constexpr int KN1 = K ? H10X_R1_LEN - K + 1 : 8, KN2 = K ? H10X_R2_LEN - K + 1 : 8 ;
  constexpr int KC1 = (KN1 + 7) / 8, KC2 = (KN2 + 7) / 8 ;
  constexpr int KD1 = 8 * (KC1 - 1) - (KN1 - 8), KD2 = 8 * (KC2 - 1) - (KN2 - 8) ;
  const bool notTail1 = lane != (uint32_t) (KC1 - 1), notTail2 = lane != (uint32_t) (KC1 + KC2 - 1) ;
  const bool notTail = notTail1 && notTail2 ;
  /* WODD: the hash loop tests only the high word, q_hi < limHi1 = wLim_hi + 1 - every multiple of w passes, and so does
     one other k-mer in 2^32; each lane then re-tests its few stored keys exactly before they are counted.  One
     compare per k-mer instead of two, and the tail lanes' limit for their shared k-mers is simply 0. */
  /* ... and, for K > 16 (SH < 32), on an estimate of that word that is at most 2^(32-SH) - 1 below it: with
     H = m_hi 2^tb + t (t = the tb = 32 - SH hash bits of the low word), q_hi = bits tb .. tb+31 of H w^-1
     = lo32 (m_hi ilo + t (ihi << SH) + floor (t ilo / 2^tb)), and floor (t ilo / 2^tb) = t (ilo >> tb) + e, 0 <= e < t:
     two plain multiply-adds instead of a wide one and two - the 64-bit multiplies are what saturates the SM here
     (ncu: fmaheavy pipe 86 %, IMAD.WIDE / IMAD.HI take 4 of its cycles, IMAD 2).  The estimate gets the bias 2^tb so that
     e never wraps it below zero, and so does the limit. */
  constexpr bool QFAST = WODD && K > 16 ;
  constexpr int TB = QFAST ? 2 * K - 32 : 1 ;
  const uint32_t qK = QFAST ? (ihi << (32 - TB)) + (ilo >> TB) : 0u ;
  const uint32_t qBias = QFAST ? (1u << TB) : 0u ;
  const uint32_t limHi1 = (uint32_t) (wLim >> 32) + 1u + qBias ;
  const uint32_t limT = notTail ? limHi1 : 0u, limT1 = notTail1 ? limHi1 : 0u, limT2 = notTail2 ? limHi1 : 0u ;
  const uint32_t p0 = (isR2 ? H10X_R2_START : H10X_R1_START) + first ;	/* unpacked position of k-mer 0 */
  const uint32_t wi = p0 >> 4, sh = 2 * (p0 & 15) ;
  const uint32_t base = isR2 ? 15u : 0u ;
  const uint32_t src0 = base + wi, src1 = base + min (wi + 1, 9u), src2 = base + min (wi + 2, 9u) ;
  const bool has1 = wi + 1 <= 9, has2 = wi + 2 <= 9 ;
  uint64_t *const G = a.stage + ((size_t) blockIdx.x * THREADS) * a.rowCap ;	/* this CTA's staging area */
  uint64_t *const col = G + t ;							/* the lane's column: slot i at col[i*THREADS] */

  for (;;)
    { if (t == 0) { sTicket = atomicAdd (a.work, 1u) ; sCount = 0 ; sBad = 0 ; }
      for (uint32_t i = t ; i <= a.nbuck ; i += THREADS) start[i] = 0 ;
      __syncthreads () ;
      const uint32_t ticket = sTicket ;
      if (ticket >= a.nList) break ;
      const uint32_t blk = a.list[ticket] ;
      const uint32_t r0 = a.blkStart[blk], nRead = a.blkStart[blk + 1] - r0 ;
      const uint32_t *rec0 = a.fqb + (size_t) H10X_REC_WORDS * r0 ;

      /* ---- hash loop: no atomics, no divergence; selected keys go to the lane's column ---- */
      uint32_t off = 0 ; bool over = false ;	/* byte offset of the next free slot of the lane's column (slot i at col[i * THREADS]): the
						   address is then an add with carry, where index * stride would be one more wide multiply */
      uint32_t wordNext = (wid < nRead && lane < H10X_REC_WORDS) ? __ldcs (rec0 + (size_t) H10X_REC_WORDS * wid + lane) : 0u ;
      for (uint32_t pr = wid ; pr < nRead ; pr += THREADS / 32)
	{ const uint32_t word = wordNext ;
	  const uint32_t nx = pr + THREADS / 32 ;
	  if (nx < nRead && lane < H10X_REC_WORDS) wordNext = __ldcs (rec0 + (size_t) H10X_REC_WORDS * nx + lane) ;
	  uint32_t w0 = __shfl_sync (0xffffffffu, word, src0) ;
	  uint32_t w1 = __shfl_sync (0xffffffffu, word, src1) ;
	  uint32_t w2 = __shfl_sync (0xffffffffu, word, src2) ;
	  if (!has1) w1 = 0 ;
	  if (!has2) w2 = 0 ;
	  if (!laneActive) continue ;
	  const uint32_t Whi = __funnelshift_l (w1, w0, sh), Wlo = __funnelshift_l (w2, w1, sh) ;	/* bases p0..p0+31 */
	  const uint32_t WRhi = h10x_swap_pairs_not (__brev (Wlo)), WRlo = h10x_swap_pairs_not (__brev (Whi)) ;
	  if (off + 8u * 8u * THREADS > a.rowCap * (8u * THREADS)) { over = true ; off = 0 ; }
#pragma unroll
	  for (int j = 0 ; j < 8 ; ++j)
	    { const int c = SH - 2 * j ;		/* (W >> c) & mask: k-mer j, first base on top */
	      uint32_t hlo = __funnelshift_r (Wlo, Whi, c) & LMASK, hhi ;
	      uint32_t rlo = __funnelshift_r (WRlo, WRhi, 2 * j) & LMASK, rhi ;
	      /* the high fields: shift + mask (two ALU-pipe instructions) or a multiply by a power of two + one shift (one on
		 the multiply pipe, one on the ALU pipe).  Which pipe has room decides, per k-mer slot j: H10X_MULSHIFT_R /
		 H10X_MULSHIFT_F are the slots whose reverse / forward field takes the multiply (round 1: all of them, the ALU
		 pipe was the limiter; since the selection got cheaper the multiply pipe is, ncu 86 % against 60 %).  Measured at
		 the 1 Gb workload: all multiply 60.6 ms, forward only 61.3, reverse only 61.0, alternate slots 61.2, all shift +
		 mask 62.6 - the two pipes and the issue slots (72 %) are within a few per cent of each other. */
	      if constexpr (K > 16)
		{ if ((H10X_MULSHIFT_R >> j) & 1)
		    { uint32_t tt ;
		      asm ("mul.lo.u32 %0, %1, %2;" : "=r" (tt) : "r" (WRhi), "r" (1u << (64 - 2 * K - 2 * j))) ;
		      rhi = tt >> (64 - 2 * K) ;
		    }
		  else rhi = (WRhi >> (2 * j)) & HMASK ;
		  if (j == 0) hhi = Whi >> c ;	/* c = 64-2K: nothing above the field */
		  else if ((H10X_MULSHIFT_F >> j) & 1)
		    { uint32_t uu ; asm ("mul.lo.u32 %0, %1, %2;" : "=r" (uu) : "r" (Whi), "r" (1u << (2 * j))) ; hhi = uu >> (64 - 2 * K) ; }
		  else hhi = (Whi >> c) & HMASK ;
		}
	      else { hhi = (Whi >> c) & HMASK ; rhi = (WRhi >> (2 * j)) & HMASK ; }
	      if (kk <= 16) { hhi = 0 ; rhi = 0 ; if (c >= 32) hlo = (Whi >> (c - 32)) & LMASK ; }
	      /* comparing the whole products orders them by their top 2k bits; when those are equal the
		 two hashes are equal and either may be taken, so one mask after the min is enough */
	      uint64_t pf = h10x_mul64 (hlo, hhi, flo, fhi) ;
	      uint64_t pq = h10x_mul64 (rlo, rhi, flo, fhi) ;
	      if constexpr (QFAST)
		{ const bool lt = pf < pq ;
		  const uint32_t mhi = lt ? (uint32_t) (pf >> 32) : (uint32_t) (pq >> 32), mlo = lt ? (uint32_t) pf : (uint32_t) pq ;
		  const uint32_t tq = mlo >> (32 - TB) ;
		  uint32_t est ;
		  asm ("mad.lo.u32 %0, %1, %2, %3;" : "=r" (est) : "r" (mhi), "r" (ilo), "r" (qBias)) ;
		  asm ("mad.lo.u32 %0, %1, %2, %0;" : "+r" (est) : "r" (tq), "r" (qK)) ;
		  if (est < ((j < KD1 && j < KD2) ? limT : (j < KD1) ? limT1 : (j < KD2) ? limT2 : limHi1))
		    { *(uint64_t*) ((char*) col + off) = ((uint64_t) mhi << 32) | ((tq << (32 - TB)) + pr) ;	/* hash << SH | read index */
		      off += 8u * THREADS ;
		    }
		}
	      else
		{ uint64_t m = (pf < pq ? pf : pq) & (((uint64_t) TOPHI << 32) | TOPLO) ;	/* canonical hash << SH */
		  uint64_t q = h10x_mul64 ((uint32_t) m, (uint32_t) (m >> 32), ilo, ihi) ;
		  bool sel ;
		  if constexpr (WODD && K > 0)
		    sel = (uint32_t) (q >> 32) < ((j < KD1 && j < KD2) ? limT : (j < KD1) ? limT1 : (j < KD2) ? limT2 : limHi1) ;
		  else
		    { sel = q <= wLim ;
		      if (!WODD) sel = sel && ((m & tzMaskSh) == 0) ;
		      if constexpr (K > 0)
			{ if (j < KD1 && j < KD2) sel = sel && notTail ;
			  else if (j < KD1) sel = sel && notTail1 ;
			  else if (j < KD2) sel = sel && notTail2 ;
			}
		      else sel = sel && (uint32_t) j >= dupSkip ;
		    }
		  if (sel) { *(uint64_t*) ((char*) col + off) = (m & 0xffffffff00000000ull) | ((uint32_t) m + pr) ; off += 8u * THREADS ; }
		}
	    }
	}
      uint32_t cnt = off / (8u * THREADS) ;
      if constexpr (WODD && K > 0)	/* the exact test of the lane's stored keys; a failing one (1 in 2^32 k-mers) is squeezed out */
	{ uint32_t v = 0 ;
	  for (uint32_t i = 0 ; i < cnt ; ++i)
	    { const uint64_t key = col[(size_t) i * THREADS] ;
	      const uint64_t m = key & (((uint64_t) TOPHI << 32) | TOPLO) ;
	      if (h10x_mul64 ((uint32_t) m, (uint32_t) (m >> 32), ilo, ihi) <= wLim)
		{ if (v != i) col[(size_t) v * THREADS] = key ;
		  ++v ;
		}
	    }
	  cnt = v ;
	}
      if (over) sBad = 1 ;
      if (cnt) atomicAdd (&sCount, cnt) ;
      __syncthreads () ;

      uint32_t n = sCount ;
      bool bad = sBad != 0 || n > a.cap ;
      if constexpr (LEAN)
	{ /* ---- the keys as they are: lane columns -> shared memory -> one contiguous piece of the scratch slab ---- */
	  if (!bad)
	    { cur[t] = cnt ;		/* nbuck >= THREADS */
	      __syncthreads () ;
	      cta_exclusive_scan<THREADS> (cur, THREADS, warpTmp) ;
	      uint64_t *dst = S + cur[t] ;
	      for (uint32_t i = 0 ; i < cnt ; ++i) dst[i] = col[(size_t) i * THREADS] ;
	      if (n == 0) { if (t == 0) S[0] = 0 ; n = 1 ; }	/* hash10x.c:167-168: the phantom entry of an empty block */
	      if (t == 0) sBase = atomicAdd (a.cursor, (unsigned long long) n) ;
	      __syncthreads () ;
	      if (sBase + n > a.scratchCap) bad = true ;
	      else for (uint32_t i = t ; i < n ; i += THREADS) a.scratch[sBase + i] = S[i] ;
	    }
	  if (t == 0)
	    { if (bad) { a.blkCnt[blk] = H10X_BLK_FALLBACK ; a.srcOff[blk] = 0 ; }
	      else { a.blkCnt[blk] = n ; a.srcOff[blk] = sBase | ((uint64_t) SH << 56) ; atomicAdd (a.moshCount, (unsigned long long) sCount) ; }
	    }
	  __syncthreads () ;
	  continue ;
	}
      if (!bad)
	{ /* ---- bucket histogram on the top hash bits ---- */
	  for (uint32_t i = 0 ; i < cnt ; ++i) atomicAdd (&start[(uint32_t) (col[(size_t) i * THREADS] >> buckShift)], 1u) ;
	  __syncthreads () ;
	  for (uint32_t i = t ; i < a.nbuck ; i += THREADS) if (start[i] > H10X_BUCKET_LIMIT) sBad = 1 ;
	  __syncthreads () ;
	  bad = sBad != 0 ;
	}
      uint32_t U = 0 ;
      if (!bad)
	{ cta_exclusive_scan<THREADS> (start, a.nbuck + 1, warpTmp) ;
	  /* ---- scatter to buckets; start[b] is the cursor, so afterwards it is the END of bucket b ---- */
	  for (uint32_t i = 0 ; i < cnt ; ++i)
	    { uint64_t key = col[(size_t) i * THREADS] ;
	      S[atomicAdd (&start[(uint32_t) (key >> buckShift)], 1u)] = key ;
	    }
	  __syncthreads () ;
	  /* ---- order inside the buckets by counting, one thread per key: its place is the bucket's start plus the
		 number of smaller keys in the bucket (equal keys - the same k-mer twice in one read pair - keep their
		 order).  Buckets are in hash order, so the places are the sorted order of the whole block.  The sorted
		 keys go to the CTA's staging area (every column has been read by now) and come back to S in order:
		 balanced work for all 32 lanes, where a per-bucket insertion sort kept 4-10 of them busy. ---- */
	  for (uint32_t i = t ; i < n ; i += THREADS)
	    { const uint64_t key = S[i] ;
	      const uint32_t b = (uint32_t) (key >> buckShift) ;
	      const uint32_t lo = b ? start[b - 1] : 0u, hi = start[b] ;
	      uint32_t r = lo ;
#pragma unroll 1		/* a bucket holds 1.5 keys on average: unrolled remainders only add branches */
	      for (uint32_t j = lo ; j < i ; ++j) r += (S[j] <= key) ? 1u : 0u ;
#pragma unroll 1
	      for (uint32_t j = i + 1 ; j < hi ; ++j) r += (S[j] < key) ? 1u : 0u ;
	      G[r] = key ;
	    }
	  if (n == 0) { if (t == 0) G[0] = 0 ; n = 1 ; }	/* hash10x.c:167-168: the phantom entry of an empty block */
	  __syncthreads () ;
	  for (uint32_t i = t ; i < n ; i += THREADS) S[i] = __ldcg (G + i) ;
	  __syncthreads () ;
	  /* ---- dedup: keep the first key of every run of equal hashes (lowest read index) ---- */
	  const uint32_t per = (n + THREADS - 1) / THREADS ;
	  const uint32_t lo = min (t * per, n), hi = min (lo + per, n) ;
	  uint32_t ucnt = 0 ;
	  for (uint32_t i = lo ; i < hi ; ++i) ucnt += (i == 0 || (S[i] >> SH) != (S[i - 1] >> SH)) ? 1u : 0u ;
	  cur[t] = ucnt ;			/* the bucket array is free again; nbuck >= THREADS */
	  __syncthreads () ;
	  U = cta_exclusive_scan<THREADS> (cur, THREADS, warpTmp) ;
	  if (t == 0) sBase = atomicAdd (a.cursor, (unsigned long long) U) ;
	  __syncthreads () ;
	  if (sBase + U > a.scratchCap) bad = true ;
	  else
	    { uint64_t *dst = a.scratch + sBase + cur[t] ;
	      for (uint32_t i = lo ; i < hi ; ++i)
		if (i == 0 || (S[i] >> SH) != (S[i - 1] >> SH)) *dst++ = S[i] ;
	    }
	}
      if (t == 0)
	{ if (bad) { a.blkCnt[blk] = H10X_BLK_FALLBACK ; a.srcOff[blk] = 0 ; }
	  else { a.blkCnt[blk] = U ; a.srcOff[blk] = sBase | ((uint64_t) SH << 56) ; atomicAdd (a.moshCount, (unsigned long long) sCount) ; }
	}
      __syncthreads () ;
    }
}

/* Moves every block's unique list to its final place, in block order, and splits it into the
   arrays the index stages use: eHash (hash value DIVIDED BY w - every mosh is a multiple of w, so the
   exact quotient hash * w^-1 mod 2^64 orders like the hash and needs log2(w) fewer radix bits), eRead (read index, 16 bits: hash10x.c:37,180) and
   entryBlk (1-based GLOBAL block number: blkBase is this rank's first block in a multi-GPU build); the
   single-GPU tail takes the last two packed in one word, eBR = block | read << 32.  One CTA per block.  Source kind by bit 63 of srcOff:
   0 = fused scratch (key = hash << sh | read, sh in bits 56..61); 1 = generic path arrays. */
__global__ void k_place (uint32_t nProcBlk, const uint64_t *__restrict__ srcOff, const uint32_t *__restrict__ blkCnt,
			 const uint64_t *__restrict__ blkOff, const uint64_t *__restrict__ scratch,
			 const uint64_t *__restrict__ gHash, const uint32_t *__restrict__ gRec,
			 const uint32_t *__restrict__ blkStart, uint32_t blkBase, uint64_t wInvFull,
			 uint64_t *__restrict__ eHash, uint16_t *__restrict__ eRead, uint32_t *__restrict__ entryBlk,
			 uint64_t *__restrict__ eBR)
{ for (uint32_t blk = blockIdx.x ; blk < nProcBlk ; blk += gridDim.x)
    { uint64_t so = srcOff[blk] ;
      uint32_t n = blkCnt[blk] ;
      uint64_t dst = blkOff[blk] ;
      if (so >> 63)
	{ uint64_t off = so & 0x7fffffffffffffffull ;
	  uint32_t r0 = blkStart[blk] ;
	  for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	    { eHash[dst + i] = gHash[off + i] * wInvFull ;
	      uint16_t rd = (uint16_t) (gRec[off + i] - r0) ;
	      if (eBR) eBR[dst + i] = (uint64_t) (blkBase + blk + 1) | ((uint64_t) rd << 32) ;
	      else { eRead[dst + i] = rd ; entryBlk[dst + i] = blkBase + blk + 1 ; }
	    }
	}
      else
	{ uint32_t sh = (uint32_t) (so >> 56) ;
	  uint64_t off = so & 0x00ffffffffffffffull ;
	  uint64_t rmask = ((uint64_t) 1 << sh) - 1 ;
	  for (uint32_t i = threadIdx.x ; i < n ; i += blockDim.x)
	    { uint64_t key = scratch[off + i] ;
	      eHash[dst + i] = (key >> sh) * wInvFull ;
	      uint16_t rd = (uint16_t) (key & rmask) ;
	      if (eBR) eBR[dst + i] = (uint64_t) (blkBase + blk + 1) | ((uint64_t) rd << 32) ;
	      else { eRead[dst + i] = rd ; entryBlk[dst + i] = blkBase + blk + 1 ; }
	    }
	}
    }
}
