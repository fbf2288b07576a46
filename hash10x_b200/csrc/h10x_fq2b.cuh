/* h10x_fq2b.cuh - FASTQ pair -> FQB records on the device: the reference's fq2b (fq2b.c:108-178) and the external
 * `bsort -k 4 -r 120` that follows it in the README pipeline (README.md:25-26), the stage that feeds --readFQB.
 *
 *   lines      a FASTQ record is four lines; the newline positions of a text (selected with the library's stream
 *              compaction) give record r its lines 4r .. 4r+3, whatever the length of the id line;
 *   check      gzReadFastq's rules (fq2b.c:180-208): the id line starts with '@', the sequence and quality lines have
 *              the length of the file's first sequence line, the third line is exactly "+"; the first offending
 *              record is reported with the reference's message;
 *   pack       seqPack / qualPack (fq2b.c:33-42, 52-61): 16 bases per word, first base on top, acgtACGT -> 0..3 and
 *              everything else (N) -> 0; 32 quality bits per word, 1 for q >= '$' + 20; the last, partial word of a
 *              line is right-aligned, and a line whose length is a multiple of 16 (32) has its last full group there;
 *   whitelist  read10xWhitelist / find10xBarcode (fq2b.c:71-104): a table over all 2^32 barcodes holding, for every
 *              whitelist entry and each of its 64 one-base variants, the code that resets the changed base.  The
 *              reference fills it line by line, so where two entries claim the same variant the later line wins, and
 *              for the entry itself the code of its last base; here every claim is an atomicMax of (line << 8 | code),
 *              which keeps exactly that winner;
 *   sort       records grouped by their first four BYTES as bsort compares them, i.e. by the byte-swapped first word,
 *              stable: the partition passes of h10x_tail.cuh on (key << 32 | index), then one gather of the records.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct FqIsNewline {
  const char *t ;
  __device__ __forceinline__ bool operator() (unsigned long long i) const { return t[i] == '\n' ; }
} ;

struct FqNlCount {
  const char *t ;
  __device__ __forceinline__ unsigned long long operator() (unsigned long long i) const { return t[i] == '\n' ? 1ull : 0ull ; }
} ;

/* error word: record << 3 | code, the smallest wins (the reference dies at the first one it meets) */
enum { FQ_ERR_ID = 1, FQ_ERR_SEQ = 2, FQ_ERR_PLUS = 3, FQ_ERR_QUAL = 4 } ;

__global__ void k_fq_check (const char *__restrict__ t, const unsigned long long *__restrict__ nl, uint64_t nRec, uint32_t L,
			    unsigned long long *__restrict__ firstErr)
{ const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= nRec) return ;
  const uint64_t idStart = r ? nl[4*r - 1] + 1 : 0 ;
  const uint64_t e0 = nl[4*r], e1 = nl[4*r + 1], e2 = nl[4*r + 2], e3 = nl[4*r + 3] ;
  uint32_t code = 0 ;
  if (t[idStart] != '@') code = FQ_ERR_ID ;
  else if (e1 - e0 - 1 != L) code = FQ_ERR_SEQ ;
  else if (e2 - e1 - 1 != 1 || t[e1 + 1] != '+') code = FQ_ERR_PLUS ;
  else if (e3 - e2 - 1 != L) code = FQ_ERR_QUAL ;
  if (code) atomicMin (firstErr, (unsigned long long) ((r << 3) | code)) ;
}

__device__ __forceinline__ uint32_t fq_base (char c)
{ switch (c) { case 'c': case 'C': return 1u ; case 'g': case 'G': return 2u ; case 't': case 'T': return 3u ; default: return 0u ; } }

/* one thread per output word of one read: words [0, ws) the bases, [ws, ws + wq) the quality bits */
__global__ void k_fq_pack (const char *__restrict__ t, const unsigned long long *__restrict__ nl, uint64_t nRec, uint32_t L,
			   uint32_t recWords, uint32_t wordOff, uint32_t *__restrict__ out)
{ const uint32_t ws = (L + 15) / 16, wq = (L + 31) / 32 ;
  const uint64_t x = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  const uint64_t r = x / (ws + wq) ;
  if (r >= nRec) return ;
  const uint32_t w = (uint32_t) (x - r * (ws + wq)) ;
  uint32_t u = 0 ;
  if (w < ws)
    { const char *s = t + nl[4*r] + 1 + 16u * w ;
      const uint32_t n = (w + 1 < ws) ? 16u : L - 16u * w ;		/* fq2b.c:36-41: `while (len > 16)`, then what is left */
      for (uint32_t i = 0 ; i < n ; ++i) u = (u << 2) | fq_base (s[i]) ;
    }
  else
    { const uint32_t v = w - ws ;
      const char *q = t + nl[4*r + 2] + 1 + 32u * v ;
      const uint32_t n = (v + 1 < wq) ? 32u : L - 32u * v ;
      for (uint32_t i = 0 ; i < n ; ++i) u = (u << 1) | (((unsigned char) q[i] >= (unsigned char) ('$' + 20)) ? 1u : 0u) ;
    }
  out[r * recWords + wordOff + w] = u ;
}

/* switchBase (fq2b.c:68-69) for code c = 1 + 4 i + j: base i (counted from the low end) becomes j */
__device__ __forceinline__ uint32_t fq_switch_base (uint32_t u, uint32_t c)
{ --c ; const uint32_t i = c >> 2, j = c & 3u ; return (u & ~(3u << (2*i))) | (j << (2*i)) ; }

/* one thread per (whitelist line, base, variant) */
__global__ void k_wl_build (const uint32_t *__restrict__ wl, uint64_t nWl, uint32_t *__restrict__ table)
{ const uint64_t x = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  const uint64_t n = x >> 6 ;
  if (n >= nWl) return ;
  const uint32_t i = (uint32_t) (x >> 2) & 15u, j = (uint32_t) x & 3u ;
  const uint32_t u = wl[n] ;
  const uint32_t ui = 1u + 4u * i + ((u >> (2*i)) & 3u) ;	/* the code that puts u's own base back */
  atomicMax (&table[fq_switch_base (u, 1u + 4u * i + j)], (uint32_t) ((n + 1) << 8) | ui) ;
}

struct FqStats { unsigned long long nBad, nFixed, nFixBase[16] ; } ;

/* find10xBarcode (fq2b.c:98-104) on word 0 of every record; keep[r] = 1 when it matched */
__global__ void k_wl_apply (uint32_t *__restrict__ rec, uint64_t nRec, uint32_t recWords, const uint32_t *__restrict__ table,
			    uint32_t *__restrict__ keep, FqStats *__restrict__ st)
{ const uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (r >= nRec) return ;
  const uint32_t u = rec[r * recWords] ;
  const uint32_t c = table[u] & 0xffu ;
  if (!c) { keep[r] = 0 ; atomicAdd (&st->nBad, 1ull) ; return ; }
  const uint32_t v = fq_switch_base (u, c) ;
  if (v != u)
    { rec[r * recWords] = v ;
      atomicAdd (&st->nFixed, 1ull) ; atomicAdd (&st->nFixBase[15 - (c - 1) / 4], 1ull) ;
    }
  keep[r] = 1 ;
}

/* sort word of kept record x (idx[x] = its number among all records): its first four bytes in memory order, then x */
__global__ void k_fq_sort_words (const uint32_t *__restrict__ rec, uint32_t recWords, const unsigned long long *__restrict__ idx,
				 uint64_t n, uint64_t *__restrict__ words)
{ const uint64_t x = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (x >= n) return ;
  const uint32_t w0 = rec[(idx ? idx[x] : x) * recWords] ;
  words[x] = ((uint64_t) __byte_perm (w0, 0, 0x0123) << 32) | x ;
}

/* out record x = in record (idx ? idx[order[x] & 0xffffffff] : ...): a warp per record */
__global__ void k_fq_gather (const uint32_t *__restrict__ rec, uint32_t recWords, const unsigned long long *__restrict__ idx,
			     const uint64_t *__restrict__ order, uint64_t n, uint32_t *__restrict__ out)
{ const uint64_t x = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
  const uint32_t lane = threadIdx.x & 31 ;
  if (x >= n) return ;
  uint64_t src = order ? (order[x] & 0xffffffffull) : x ;
  if (idx) src = idx[src] ;
  for (uint32_t w = lane ; w < recWords ; w += 32) out[x * recWords + w] = rec[src * recWords + w] ;
}
