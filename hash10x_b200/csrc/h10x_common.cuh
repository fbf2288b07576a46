/* h10x_common.cuh - shared device helpers for the --readFQB build kernels (sm_100a).
 *
 * Hash arithmetic follows seqhash.c:58-80 of the reference: a k-mer is the 2k-bit number
 * h = sum s[j+i]*4^(k-1-i); its reverse complement hRC = sum (3-s[j+i])*4^i; each is hashed as
 * (x * factor1 mod 2^64) >> (64-2k); the canonical hash is the smaller; a "mosh" is a k-mer whose
 * canonical hash is a multiple of w (seqhash.c:171,189).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct HashParams {
  uint64_t factor1 ;
  uint64_t kmask ;	/* 2^(2k) - 1 */
  uint64_t wInv ;	/* inverse of the odd part of w modulo 2^64 */
  uint64_t wLim ;	/* floor((2^64-1) / odd part of w) */
  uint64_t wTzMask ;	/* 2^tz - 1, tz = trailing zero bits of w */
  int k, w, shift, wTz ;
  int rcShift ;		/* 2(k-1) */
} ;

/* x % w == 0 without a division: w = 2^tz * wo, wo odd.  For odd wo, x is a multiple of wo
   iff x * wo^-1 (mod 2^64) <= floor((2^64-1)/wo).  */
__host__ __device__ __forceinline__ bool h10x_divisible (uint64_t x, const HashParams &hp)
{ if (x & hp.wTzMask) return false ;
  return ((x >> hp.wTz) * hp.wInv) <= hp.wLim ;
}

__host__ __device__ __forceinline__ uint64_t h10x_canonical (uint64_t h, uint64_t hrc, const HashParams &hp)
{ uint64_t hf = (h * hp.factor1) >> hp.shift ;
  uint64_t hr = (hrc * hp.factor1) >> hp.shift ;
  return hf < hr ? hf : hr ;
}

/* base p (0..159) of a read packed as 10 words, MSB first (hash10x.c:112) */
__device__ __forceinline__ uint32_t h10x_base (const uint32_t *u, int p)
{ return (u[p >> 4] >> (30 - 2*(p & 15))) & 3u ; }

/* Generic (any k, any w) scan of bases [start, start+len) of one read; emit(hash) is called for
   every mosh in left-to-right order (seqhash.c:154-195). */
template <class Emit>
__device__ __forceinline__ void h10x_scan_read (const uint32_t *u, int start, int len,
						const HashParams &hp, Emit emit)
{
  if (len < hp.k) return ;
  uint64_t h = 0, hrc = 0 ;
  for (int i = 0 ; i < hp.k - 1 ; ++i)
    { uint32_t b = h10x_base (u, start + i) ;
      h = (h << 2) | b ;
      hrc = (hrc >> 2) | ((uint64_t)(3u - b) << hp.rcShift) ;
    }
  for (int j = 0 ; j + hp.k <= len ; ++j)
    { uint32_t b = h10x_base (u, start + j + hp.k - 1) ;
      h = ((h << 2) & hp.kmask) | b ;
      hrc = (hrc >> 2) | ((uint64_t)(3u - b) << hp.rcShift) ;
      uint64_t x = h10x_canonical (h, hrc, hp) ;
      if (h10x_divisible (x, hp)) emit (x) ;
    }
}

/* read 1 is hashed over unpacked positions [23,150), read 2 over [0,150) (hash10x.c:162-163) */
#define H10X_R1_START 23
#define H10X_R1_LEN 127
#define H10X_R2_START 0
#define H10X_R2_LEN 150
#define H10X_REC_WORDS 30
