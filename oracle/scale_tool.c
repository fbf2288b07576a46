/* scale_tool.c - host tools for parity at the sizes bench.py times.  TEST INFRASTRUCTURE ONLY (see h10x_oracle.c).
 *
 *   scale_tool gen <out.fqb> seed genomeLen nBarcodes pairsMin pairsMax molPerBarcode molLen snpPeriod errThresh readLen
 *       writes the synthetic FQB of those generator parameters (the same closed-form records the device
 *       generator of bench.py produces: hash10x_b200/csrc/synth_fqb.h), OpenMP over record ranges;
 *   scale_tool digest <file.hash>
 *       streams a `.hash` file (layout: hash10x.c:244-267 of the reference, array.h:41-50) and prints the
 *       counters and the position-salted sum digests (hash10x_b200/csrc/h10x_digest.h) of hashIndex, hashValue,
 *       hashDepth, the block table and the ClusterHash stream as one JSON object.
 *
 * tests/golden/make_golden_scale.py runs `gen`, then the UNMODIFIED reference binary (oracle/_ref/hash10x
 * --readFQB ... --writeHash), then `digest`, and commits the result as tests/golden/golden_scale.json.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../hash10x_b200/csrc/synth_fqb.h"
#include "../hash10x_b200/csrc/h10x_digest.h"

static void die (const char *m) { fprintf (stderr, "scale_tool: %s\n", m) ; exit (2) ; }

static int do_gen (int argc, char **argv)
{ if (argc < 13) die ("gen: 11 parameters expected") ;
  synth_params p ; memset (&p, 0, sizeof (p)) ;
  const char *path = argv[2] ;
  p.seed = strtoull (argv[3], 0, 10) ; p.genomeLen = strtoull (argv[4], 0, 10) ;
  p.nBarcodes = (uint32_t) strtoul (argv[5], 0, 10) ; p.pairsMin = (uint32_t) strtoul (argv[6], 0, 10) ;
  p.pairsMax = (uint32_t) strtoul (argv[7], 0, 10) ; p.molPerBarcode = (uint32_t) strtoul (argv[8], 0, 10) ;
  p.molLen = (uint32_t) strtoul (argv[9], 0, 10) ; p.snpPeriod = (uint32_t) strtoul (argv[10], 0, 10) ;
  p.errThresh = (uint32_t) strtoul (argv[11], 0, 10) ; p.readLen = (uint32_t) strtoul (argv[12], 0, 10) ;
  uint64_t *off = malloc (8 * ((size_t) p.nBarcodes + 1)) ;
  if (!off) die ("out of memory") ;
  uint64_t n = 0 ;
  for (uint32_t b = 0 ; b < p.nBarcodes ; ++b) { off[b] = n ; n += sy_pairs (&p, b) ; }
  off[p.nBarcodes] = n ;
  FILE *f = fopen (path, "wb") ;
  if (!f) die ("cannot open output") ;
  const uint32_t step = 4096 ;		/* barcodes per slab */
  uint32_t *buf = 0 ; size_t cap = 0 ;
  for (uint32_t b0 = 0 ; b0 < p.nBarcodes ; b0 += step)
    { uint32_t b1 = b0 + step < p.nBarcodes ? b0 + step : p.nBarcodes ;
      uint64_t r0 = off[b0], nr = off[b1] - r0 ;
      if (nr > cap) { free (buf) ; cap = nr ; buf = malloc (cap * 120) ; if (!buf) die ("out of memory") ; }
#pragma omp parallel for schedule(dynamic, 16)
      for (uint32_t b = b0 ; b < b1 ; ++b)
	for (uint64_t r = off[b] ; r < off[b+1] ; ++r)
	  sy_record (&p, b, (uint32_t) (r - off[b]), r, buf + 30 * (r - r0)) ;
      if (fwrite (buf, 120, nr, f) != nr) die ("write failed") ;
    }
  fclose (f) ;
  printf ("{\"records\": %llu, \"barcodes\": %u}\n", (unsigned long long) n, p.nBarcodes) ;
  return 0 ;
}

#define CHUNK (1u << 22)

static uint64_t dg_stream (FILE *f, uint64_t n, int elem, uint64_t mask, uint64_t *sumOut)
{ /* digest of n elements of `elem` bytes read from f; also their plain sum (used for nHashes) */
  unsigned char *buf = malloc ((size_t) CHUNK * elem) ;
  if (!buf) die ("out of memory") ;
  uint64_t dg = 0, sum = 0, pos = 0 ;
  while (pos < n)
    { size_t want = (size_t) (n - pos < CHUNK ? n - pos : CHUNK) ;
      if (fread (buf, elem, want, f) != want) die ("truncated .hash file") ;
      uint64_t d = 0, s = 0 ;
#pragma omp parallel for reduction(+:d,s)
      for (size_t i = 0 ; i < want ; ++i)
	{ uint64_t v = elem == 4 ? ((uint32_t*) buf)[i] : ((uint64_t*) buf)[i] ;
	  v &= mask ;
	  d += h10x_dg_term (pos + i, v) ; s += v ;
	}
      dg += d ; sum += s ; pos += want ;
    }
  free (buf) ;
  if (sumOut) *sumOut = sum ;
  return dg ;
}

typedef struct { int32_t magic, pad0 ; uint64_t base ; int32_t dim, size, max, pad1 ; } ArrayHdr ;	/* array.h:41-50 */

static int do_digest (int argc, char **argv)
{ if (argc < 3) die ("digest: file expected") ;
  FILE *f = fopen (argv[2], "rb") ;
  if (!f) die ("cannot open .hash file") ;
  char magic[4] ; uint32_t version ; uint16_t chSize, cbSize ; int32_t B ;
  if (fread (magic, 1, 4, f) != 4 || memcmp (magic, "10XH", 4)) die ("not a 10XH file") ;
  if (fread (&version, 4, 1, f) != 1 || fread (&chSize, 2, 1, f) != 1 || fread (&cbSize, 2, 1, f) != 1
      || fread (&B, 4, 1, f) != 1) die ("truncated header") ;
  if (version != 2 || chSize != 8 || cbSize != 32) die ("unexpected version / struct sizes") ;
  uint64_t dgIndex = dg_stream (f, (uint64_t) 1 << B, 4, 0xffffffffull, 0) ;
  uint32_t hashNumber ;
  if (fread (&hashNumber, 4, 1, f) != 1) die ("truncated") ;
  uint64_t dgValue = dg_stream (f, hashNumber, 8, ~(uint64_t) 0, 0) ;
  ArrayHdr ah ;
  if (fread (&ah, 32, 1, f) != 1 || ah.size != 4) die ("bad hashDepth array") ;
  uint64_t sumDepth = 0 ;
  uint64_t dgDepth = dg_stream (f, (uint64_t) ah.max, 4, 0xffffffffull, &sumDepth) ;
  if (fseeko (f, (off_t) 4 * (ah.dim - ah.max), SEEK_CUR)) die ("seek") ;	/* arrayWrite stores dim elements */
  if ((uint32_t) ah.max != hashNumber && !(hashNumber == 1 && ah.max == 0)) die ("hashDepth max != hashNumber") ;
  ArrayHdr bh ;
  if (fread (&bh, 32, 1, f) != 1 || bh.size != 32) die ("bad clusterBlocks array") ;
  uint32_t nb = (uint32_t) bh.max ;
  uint32_t *blk = malloc ((size_t) 32 * bh.dim) ;
  if (!blk || fread (blk, 32, bh.dim, f) != (size_t) bh.dim) die ("truncated block table") ;
  uint64_t dgNRead = 0, dgNHash = 0, nHashes = 0, nReads = 0 ;
  for (uint32_t b = 0 ; b < nb ; ++b)
    { dgNRead += h10x_dg_term (b, blk[8*b]) ; dgNHash += h10x_dg_term (b, blk[8*b + 1]) ;
      nReads += blk[8*b] ; if (b) nHashes += blk[8*b + 1] ;
    }
  uint64_t dgClus = dg_stream (f, nHashes, 8, H10X_DG_CLUS_MASK, 0) ;
  if (fgetc (f) != EOF) die ("trailing bytes") ;
  off_t size = ftello (f) ;
  fclose (f) ;
  printf ("{\"B\": %d, \"hashNumber\": %u, \"nBlocksMax\": %u, \"nReads\": %llu, \"nHashes\": %llu, \"sumDepth\": %llu, "
	  "\"fileSize\": %lld, \"dg_hashIndex\": \"%016llx\", \"dg_hashValue\": \"%016llx\", \"dg_hashDepth\": \"%016llx\", "
	  "\"dg_blkNRead\": \"%016llx\", \"dg_blkNHash\": \"%016llx\", \"dg_clusHash\": \"%016llx\"}\n",
	  B, hashNumber, nb, (unsigned long long) nReads, (unsigned long long) nHashes, (unsigned long long) sumDepth,
	  (long long) size, (unsigned long long) dgIndex, (unsigned long long) dgValue, (unsigned long long) dgDepth,
	  (unsigned long long) dgNRead, (unsigned long long) dgNHash, (unsigned long long) dgClus) ;
  return 0 ;
}

int main (int argc, char **argv)
{ if (argc >= 2 && !strcmp (argv[1], "gen")) return do_gen (argc, argv) ;
  if (argc >= 2 && !strcmp (argv[1], "digest")) return do_digest (argc, argv) ;
  die ("usage: scale_tool gen ... | digest file.hash") ;
  return 2 ;
}
