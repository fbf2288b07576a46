/* h10x_crib.cuh - `--cribBuild genome1.fa genome2.fa` (hash10x.c:426-510) and the histograms behind --hashStats /
 * --codeStats (hash10x.c:351-402) on the resident index ("next" row f4 of SURVEY.md section 8).
 *
 * cribAddGenome walks a genome's sequences with the mosh iterator and looks every mosh up in the bin table
 * (hashIndexFind (hash, FALSE), hash10x.c:139-152); a bin remembers the first place it was seen and how often.  Here one
 * thread takes one k-mer start: canonical hash, the modulo test, the reference's own double-hash probe of hashIndex[],
 * then an atomic count and an atomic min of (sequence << 32 | position) - "first" in the reference's order of walking.
 * The CribInfo fields follow from (count, first): chr = the sequence for one hit, -(count - 1) as I16 for more
 * (the reference decrements an I16, wrap included), pos = first position >> 10 as U16 (hash10x.c:440-444).
 */
#pragma once
#include "h10x_common.cuh"

struct CribCounts { unsigned long long nPresent, nAbsent ; } ;

__global__ void k_crib_scan (const uint8_t *__restrict__ codes, uint64_t total, const uint64_t *__restrict__ seqOff, uint32_t nSeq,
			     HashParams hp, const uint32_t *__restrict__ table, int B, const uint64_t *__restrict__ hashValue,
			     uint32_t *__restrict__ cnt, unsigned long long *__restrict__ first, CribCounts *__restrict__ cc)
{ const uint64_t mask = ((uint64_t) 1 << B) - 1 ;
  for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; j < total ; j += (uint64_t) gridDim.x * blockDim.x)
    { uint32_t lo = 0, hi = nSeq ;		/* the sequence holding position j */
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1 ; if (seqOff[mid] <= j) lo = mid ; else hi = mid ; }
      if (j + (uint64_t) hp.k > seqOff[lo + 1]) continue ;
      uint64_t h = 0, hrc = 0 ;
      for (int i = 0 ; i < hp.k ; ++i)
	{ const uint64_t b = codes[j + i] & 3u ;
	  h = (h << 2) | b ;
	  hrc = (hrc >> 2) | ((3u - b) << hp.rcShift) ;
	}
      const uint64_t x = h10x_canonical (h, hrc, hp) ;
      if (!h10x_divisible (x, hp)) continue ;
      uint64_t off = x & mask ; const uint64_t diff = ((x >> B) & mask) | 1 ;
      uint32_t index ;
      while ((index = table[off]) && hashValue[index] != x) off = (off + diff) & mask ;
      if (index)
	{ atomicAdd (&cnt[index], 1u) ;
	  atomicMin (&first[index], ((unsigned long long) (lo + 1) << 32) | (unsigned long long) (uint32_t) (j - seqOff[lo])) ;
	  atomicAdd (&cc->nPresent, 1ull) ;
	}
      else atomicAdd (&cc->nAbsent, 1ull) ;
    }
}

__device__ __forceinline__ int crib_chr (uint32_t n, unsigned long long first)
{ if (!n) return 0 ;
  if (n == 1) return (int) (int16_t) (uint16_t) (first >> 32) ;
  return (int) (int16_t) (uint16_t) (0u - (n - 1u)) ;		/* I16 decremented n - 1 times from 0 */
}

/* the classification loop of cribBuild (hash10x.c:478-494); hist[t * histLen + depth] counts type t's bins */
__global__ void k_crib_classify (uint32_t hn, const uint32_t *__restrict__ cnt1, const unsigned long long *__restrict__ first1,
				 const uint32_t *__restrict__ cnt2, const unsigned long long *__restrict__ first2,
				 const uint32_t *__restrict__ hashDepth, uint32_t histLen, uint8_t *__restrict__ type,
				 int16_t *__restrict__ chr, uint16_t *__restrict__ pos, int *__restrict__ hist)
{ const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x ;
  if (i == 0 || i >= hn) { if (i == 0 && hn) { type[0] = 0 ; chr[0] = 0 ; pos[0] = 0 ; } return ; }
  int c1 = crib_chr (cnt1[i], first1[i]), c2 = crib_chr (cnt2[i], first2[i]) ;
  uint16_t p1 = cnt1[i] ? (uint16_t) ((uint32_t) first1[i] >> 10) : 0, p2 = cnt2[i] ? (uint16_t) ((uint32_t) first2[i] >> 10) : 0 ;
  uint32_t t ;
  if (c1 == 0 && c2 == 0) t = 0 ;						/* CRIB_ERR */
  else if (c1 > 0 && c2 > 0) t = 3 ;						/* CRIB_HOM */
  else if (c1 < 0 || c2 < 0) { t = 4 ; if (c2 < c1) c1 = c2 ; }			/* CRIB_MUL */
  else { t = c1 ? 1 : 2 ; if (!c1) { c1 = c2 ; p1 = p2 ; } }			/* CRIB_HTA / CRIB_HTB */
  type[i] = (uint8_t) t ; chr[i] = (int16_t) c1 ; pos[i] = p1 ;
  const uint32_t group = t == 0 ? 0u : t == 3 ? 2u : t == 4 ? 3u : 1u ;		/* err, het, hom, mul */
  const uint32_t d = hashDepth[i] ;
  if (d < histLen) atomicAdd (&hist[(size_t) group * histLen + d], 1) ;
}

/* ---- --hashStats / --codeStats: the count-per-value arrays of hashDepthHist / codeSizeHist (hash10x.c:377-402) ---- */

__global__ void k_max_u32 (const uint32_t *__restrict__ v, uint64_t n, uint32_t *__restrict__ out)
{ uint32_t m = 0 ;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; i < n ; i += (uint64_t) gridDim.x * blockDim.x) m = max (m, v[i]) ;
#pragma unroll
  for (int o = 16 ; o ; o >>= 1) m = max (m, __shfl_xor_sync (0xffffffffu, m, o)) ;
  if ((threadIdx.x & 31) == 0 && m) atomicMax (out, m) ;
}

/* values below H10X_HIST_SMEM are counted in shared memory first (bin depths pile up on a few dozen values) */
#define H10X_HIST_SMEM 4096u
__global__ void k_hist_u32 (const uint32_t *__restrict__ v, uint64_t n, uint32_t histLen, int *__restrict__ hist)
{ __shared__ int sh[H10X_HIST_SMEM] ;
  for (uint32_t i = threadIdx.x ; i < H10X_HIST_SMEM ; i += blockDim.x) sh[i] = 0 ;
  __syncthreads () ;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ; i < n ; i += (uint64_t) gridDim.x * blockDim.x)
    { const uint32_t x = v[i] ;
      if (x < H10X_HIST_SMEM) atomicAdd (&sh[x], 1) ; else atomicAdd (&hist[x], 1) ;
    }
  __syncthreads () ;
  for (uint32_t i = threadIdx.x ; i < H10X_HIST_SMEM && i < histLen ; i += blockDim.x) if (sh[i]) atomicAdd (&hist[i], sh[i]) ;
}
