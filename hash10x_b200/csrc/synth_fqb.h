/* synth_fqb.h - deterministic synthetic 10X linked-read FQB generator.
 *
 * Test / benchmark infrastructure, NOT part of the product path.  It produces the
 * 120-byte FQB records that the reference's fq2b writes (fq2b.c:33-61,159-160:
 * 30 little-endian U32 = 10 words read-1 bases, 5 words read-1 1-bit quals,
 * 10 words read-2 bases, 5 words read-2 quals; 16 bases per word, first base in
 * bits 31..30, the last partial word right-aligned), grouped by barcode as the
 * external `bsort` leaves them (README.md:26 of the reference).
 *
 * Every record is a closed-form function of (params, barcode index, pair index),
 * so the same header compiles as plain C for the CPU tools and as __device__
 * code for the on-GPU generator used by bench.py at sizes (24 GB) that cannot
 * be produced on the host in reasonable time.
 *
 * Model: a random diploid genome of haploid length G
 * (haplotype B = A with a substitution every ~snpPeriod bases), nBarcodes
 * barcodes each holding molPerBarcode molecules of molLen bases drawn from a
 * random haplotype, a uniform number of read pairs per barcode, 151+151 bp
 * pairs by default (read 1 = 16 bp barcode + 7 bp spacer + insert, read 2 = the
 * reverse-complement end of a 300-699 bp fragment), substitution errors with
 * probability errThresh / 2^32 per base.  readLen = 160 fills the last packed
 * word: with 151 bp the reference hashes the zero padding of word 9 (SURVEY.md
 * Appendix C), which adds ~0.39 never-seen hashes per pair and overflows the
 * reference's -B 28 table at the 200M-pair scale.
 */
#ifndef H10X_SYNTH_FQB_H
#define H10X_SYNTH_FQB_H

#include <stdint.h>

#ifdef __CUDACC__
#define SY_HD __host__ __device__ __forceinline__
#else
#define SY_HD static inline
#endif

typedef struct {
  uint64_t seed;
  uint64_t genomeLen;      /* haploid genome length G (> molLen)                */
  uint32_t nBarcodes;      /* number of barcode runs                            */
  uint32_t pairsMin;       /* read pairs per barcode: uniform in [min,max]      */
  uint32_t pairsMax;
  uint32_t molPerBarcode;  /* molecules per barcode                             */
  uint32_t molLen;         /* molecule length in bases (> 700)                  */
  uint32_t snpPeriod;      /* hap B differs from hap A at ~1/snpPeriod bases; 0 = haploid */
  uint32_t errThresh;      /* per-base substitution probability * 2^32          */
  uint32_t readLen;        /* bases per read, 145..160 (0 = 151); 160 fills the last packed word */
} synth_params;

SY_HD uint64_t sy_mix (uint64_t x)        /* splitmix64 finaliser */
{ x += 0x9E3779B97F4A7C15ull ;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull ;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull ;
  return x ^ (x >> 31) ;
}

SY_HD uint64_t sy_h2 (uint64_t seed, uint64_t a, uint64_t b)
{ return sy_mix (sy_mix (seed ^ (a * 0xD6E8FEB86659FD93ull)) + b) ; }

/* distinct, non-zero 32-bit barcode word for barcode index b (odd multiplier = bijection) */
SY_HD uint32_t sy_barcode (uint32_t b) { return (uint32_t)((b + 1u) * 0x9E3779B1u) ; }

SY_HD uint32_t sy_pairs (const synth_params *p, uint32_t b)
{ uint32_t span = p->pairsMax - p->pairsMin + 1u ;
  return p->pairsMin + (uint32_t)(sy_h2 (p->seed, 0x7061697273ull, b) % span) ;
}

SY_HD uint32_t sy_genome_base (const synth_params *p, int hap, uint64_t x)
{ uint32_t a = (uint32_t)(sy_h2 (p->seed, 0x67656e6f6d65ull, x) >> 62) ;
  if (hap && p->snpPeriod)
    { uint64_t s = sy_h2 (p->seed, 0x736e70ull, x) ;
      if ((s % p->snpPeriod) == 0) a = (a + 1u + (uint32_t)((s >> 40) % 3u)) & 3u ;
    }
  return a ;
}

/* Fill rec[0..29] for pair j of barcode b; recGlobal only seeds the error stream. */
SY_HD void sy_record (const synth_params *p, uint32_t b, uint32_t j, uint64_t recGlobal,
		      uint32_t *rec)
{
  uint64_t r0 = sy_h2 (p->seed, ((uint64_t)b << 32) | j, 1) ;
  uint64_t r1 = sy_mix (r0) ;
  uint32_t mol = (uint32_t)(r0 % p->molPerBarcode) ;
  uint64_t rm = sy_h2 (p->seed, ((uint64_t)b << 32) | mol, 2) ;
  int hap = (int)(rm & 1u) ;
  uint64_t molStart = (rm >> 1) % (p->genomeLen - p->molLen) ;
  uint32_t fragLen = 300u + (uint32_t)((r0 >> 32) % 400u) ;
  uint64_t fragStart = molStart + (r1 % (uint64_t)(p->molLen - fragLen)) ;
  int strand = (int)((r1 >> 60) & 1u) ;
  uint64_t errKey = sy_mix (p->seed ^ (recGlobal * 0xA24BAED4963EE407ull)) ;
  uint32_t bc = sy_barcode (b) ;
  int rd, q, i ;
  const int L = p->readLen ? (int) p->readLen : 151 ;
  const int nFull = L / 16 ;

  for (i = 0 ; i < 30 ; ++i) rec[i] = 0 ;
  for (rd = 0 ; rd < 2 ; ++rd)
    { uint32_t *u = rec + 15*rd ;
      for (q = 0 ; q < L ; ++q)
	{ uint32_t base ;
	  if (rd == 0 && q < 16) base = (bc >> (2*(15-q))) & 3u ;             /* barcode */
	  else if (rd == 0 && q < 23) base = (uint32_t)((r1 >> (2*q)) & 3u) ;   /* spacer */
	  else
	    { /* t = offset along the sequenced strand of the fragment */
	      uint32_t t = rd ? (uint32_t)q : (uint32_t)(q - 23) ;
	      /* read 1 reads the fragment strand 5'->3', read 2 its reverse complement */
	      int rc = rd ^ strand ;
	      uint64_t x = rc ? fragStart + fragLen - 1u - t : fragStart + t ;
	      base = sy_genome_base (p, hap, x) ;
	      if (rc) base = 3u - base ;
	      if (p->errThresh)
		{ uint64_t e = sy_mix (errKey + (uint64_t)(rd*L + q)) ;
		  if ((uint32_t)e < p->errThresh) base = (base + 1u + (uint32_t)((e >> 40) % 3u)) & 3u ;
		}
	    }
	  /* fq2b.c:33-42 packing: full words MSB first, the last partial word right-aligned */
	  if (q < 16*nFull) u[q >> 4] |= base << (2*(15 - (q & 15))) ;
	  else u[nFull] |= base << (2*(L - 1 - q)) ;
	}
      /* fq2b.c:52-61: 1 bit per base, all "good" quality; 151 = 4 full words + 23 bits */
      for (i = 0 ; i < L/32 ; ++i) u[10 + i] = 0xFFFFFFFFu ;
      if (L % 32) u[10 + L/32] = (1u << (L % 32)) - 1u ;
    }
}

#endif
