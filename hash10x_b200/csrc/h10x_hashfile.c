/* h10x_hashfile.c - `.hash` writer / reader for a host h10x_index (plain C, no CUDA).
 *
 * File layout: writeHashFile / readHashFile, hash10x.c:244-315 of the reference, with the two
 * Arrays serialised as array.c:213-238 does (raw 32-byte ArrayStruct, then `dim` elements).
 * SURVEY.md Appendix B lists every field.  What differs from a reference-written file, and why it
 * does not matter to its reader: the raw pointers inside ArrayStruct / ClusterBlock are written as
 * 0 (readHashFile overwrites them, hash10x.c:305, array.c:223) and ClusterHash bytes 6-7 are 0 until --cluster
 * sets byte 6 (the reference leaves malloc garbage there, hash10x.c:175).  `dim` follows the reference's growth rule
 * (array.c:144-170) so the file has exactly the size the reference would write.
 */
#include "../../include/h10x_gpu.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ARRAY_MAGIC 8918274	/* array.h:56 */

typedef struct {		/* array.h:41-50 on x86-64 */
  int32_t magic, pad0 ;
  uint64_t base ;
  int32_t dim, size, max, pad1 ;
} array_header ;

/* dim of an Array created with dim0 whose elements 0..max-1 were touched one after another */
static int grown_dim (int dim, int size, int max)
{ while (max > dim)
    { if ((long) dim * size < (1 << 23)) dim *= 2 ;
      else dim += 1024 + ((1 << 23) / size) ;
    }
  return dim ;
}

/* dim0 > 0: the Array grew from dim0 as its elements were touched; dim0 == 0: it was created with exactly max elements */
static int put_array (FILE *f, const void *data, int size, int max, int dim0)
{ array_header a ;
  memset (&a, 0, sizeof (a)) ;
  a.magic = ARRAY_MAGIC ; a.size = size ; a.max = max ; a.dim = dim0 ? grown_dim (dim0, size, max) : max ;
  if (fwrite (&a, sizeof (a), 1, f) != 1) return 0 ;
  if (max && fwrite (data, size, max, f) != (size_t) max) return 0 ;
  size_t rest = (size_t) (a.dim - max) * size ;
  while (rest)			/* the unused tail of the Array is calloc'ed zeros (array.c:63,163) */
    { static const char zeros[4096] ;
      size_t n = rest < sizeof (zeros) ? rest : sizeof (zeros) ;
      if (fwrite (zeros, 1, n, f) != n) return 0 ;
      rest -= n ;
    }
  return 1 ;
}

int h10x_write_hash (const h10x_index *ix, const char *path)
{
  if (!ix || !path || ix->onDevice || !ix->hashIndex) return H10X_ERR_BAD_PARAM ;
  FILE *f = fopen (path, "wb") ;
  if (!f) return H10X_ERR_IO ;
  uint32_t version = 2 ; uint16_t chSize = 8, cbSize = 32 ; int32_t B = ix->B ;
  uint64_t tableSize = (uint64_t) 1 << ix->B ;
  int ok = fwrite ("10XH", 4, 1, f) == 1 && fwrite (&version, 4, 1, f) == 1
    && fwrite (&chSize, 2, 1, f) == 1 && fwrite (&cbSize, 2, 1, f) == 1 && fwrite (&B, 4, 1, f) == 1 ;
  ok = ok && fwrite (ix->hashIndex, 4, tableSize, f) == tableSize ;
  ok = ok && fwrite (&ix->hashNumber, 4, 1, f) == 1 ;
  ok = ok && fwrite (ix->hashValue, 8, ix->hashNumber, f) == ix->hashNumber ;
  /* arrayMax(hashDepth) is hashNumber once any block was processed, else 0 (hash10x.c:178,1114) */
  ok = ok && put_array (f, ix->hashDepth, 4, ix->hashNumber > 1 ? (int) ix->hashNumber : 0, 1 << 20) ;
  if (ok)
    { /* ClusterBlock, hash10x.c:62-70: nRead, nHash, nSubCluster, clusterParent, pointer, double */
      uint32_t nb = ix->nBlocksMax, b ;
      uint32_t *cb = calloc ((size_t) nb * 8, sizeof (uint32_t)) ;
      if (!cb) { fclose (f) ; return H10X_ERR_NOMEM ; }
      for (b = 0 ; b < nb ; ++b)
	{ cb[8*b] = ix->blkNRead[b] ; cb[8*b + 1] = ix->blkNHash[b] ;
	  if (ix->blkNSubCluster) cb[8*b + 2] = ix->blkNSubCluster[b] ;
	  if (ix->blkPointToMin) memcpy (&cb[8*b + 6], &ix->blkPointToMin[b], 8) ;
	  if (ix->blkClusterParent) cb[8*b + 3] = ix->blkClusterParent[b] ;
	}
      /* dim0 = 1200: hash10x.c:1151; clusterSplitCodes makes a new Array of exactly the blocks it needs (:961), and
	 arrayRead keeps whatever dim a file had - H10X_INDEX_EXACT_BLOCKS in `reserved` remembers that */
      ok = put_array (f, cb, 32, (int) nb, (ix->reserved & H10X_INDEX_EXACT_BLOCKS) ? 0 : 1200) ;
      free (cb) ;
    }
  ok = ok && (!ix->nHashes || fwrite (ix->clusHash, 8, ix->nHashes, f) == ix->nHashes) ;
  if (fclose (f)) ok = 0 ;
  return ok ? H10X_OK : H10X_ERR_IO ;
}

static void seterr (char *err, size_t errlen, const char *msg)
{ if (err && errlen) { strncpy (err, msg, errlen - 1) ; err[errlen - 1] = 0 ; } }

/* readHashFile hash10x.c:269-315 (version 2 files).  B must match the current -B (hash10x.c:284).
   The hash->code lists are not in the file; the caller rebuilds them (fillHashTable). */
int h10x_read_hash (const char *path, int32_t wantB, h10x_index *out, char *err, size_t errlen)
{
  if (!path || !out) return H10X_ERR_BAD_PARAM ;
  memset (out, 0, sizeof (*out)) ;
  FILE *f = fopen (path, "rb") ;
  if (!f) { seterr (err, errlen, "failed to open hash file") ; return H10X_ERR_IO ; }
  char name[5] = "abcd" ; uint32_t version ; uint16_t chSize, cbSize ; int32_t B ;
  int st = H10X_ERR_IO ;
  array_header a ;
  uint32_t *cb = 0 ;
  if (fread (name, 4, 1, f) != 1 || fread (&version, 4, 1, f) != 1 || fread (&chSize, 2, 1, f) != 1
      || fread (&cbSize, 2, 1, f) != 1) { seterr (err, errlen, "read fail 0") ; goto fail ; }
  if (strcmp (name, "10XH")) { seterr (err, errlen, "not a 10X hash file") ; goto fail ; }
  if (version > 2)		/* hash10x.c:277-278; files of version 1 (hashValue as an Array) are read as the reference reads them */
    { char msg[96] ; snprintf (msg, sizeof (msg), "hash file version mismatch: file %d > code %d", (int) version, 2) ;
      seterr (err, errlen, msg) ; goto fail ;
    }
  if (chSize != 8)
    { char msg[96] ; snprintf (msg, sizeof (msg), "ClusterHash structure size mismatch: file %d != code %d", (int) chSize, 8) ;
      seterr (err, errlen, msg) ; goto fail ;
    }
  if (cbSize != 32)
    { char msg[96] ; snprintf (msg, sizeof (msg), "ClusterBlock structure size mismatch: file %d != code %d", (int) cbSize, 32) ;
      seterr (err, errlen, msg) ; goto fail ;
    }
  if (fread (&B, 4, 1, f) != 1) { seterr (err, errlen, "read fail 1") ; goto fail ; }
  if (B != wantB)
    { char msg[96] ; snprintf (msg, sizeof (msg), "incompatible hash table size: rerun with -B %d", B) ;
      seterr (err, errlen, msg) ; st = H10X_ERR_BAD_PARAM ; goto fail ;
    }
  out->B = B ;
  { uint64_t tableSize = (uint64_t) 1 << B ;
    if (!(out->hashIndex = malloc (tableSize * 4))) { st = H10X_ERR_NOMEM ; goto fail ; }
    if (fread (out->hashIndex, 4, tableSize, f) != tableSize) { seterr (err, errlen, "read fail 2") ; goto fail ; }
  }
  if (version == 1)		/* hash10x.c:285-291: an Array of U64, hashNumber = its max */
    { if (fread (&a, sizeof (a), 1, f) != 1 || a.size != 8 || a.dim < a.max || a.max < 0)
	{ seterr (err, errlen, "failed to read hashValue array") ; goto fail ; }
      out->hashNumber = (uint32_t) a.max ;
      if (!(out->hashValue = malloc ((size_t) a.dim * 8 + 8))) { st = H10X_ERR_NOMEM ; goto fail ; }
      if (fread (out->hashValue, 8, a.dim, f) != (size_t) a.dim) { seterr (err, errlen, "failed to read hashValue array") ; goto fail ; }
    }
  else
    { if (fread (&out->hashNumber, 4, 1, f) != 1) { seterr (err, errlen, "failed to read hashNumber") ; goto fail ; }
      if (!(out->hashValue = malloc ((size_t) out->hashNumber * 8 + 8))) { st = H10X_ERR_NOMEM ; goto fail ; }
      if (fread (out->hashValue, 8, out->hashNumber, f) != out->hashNumber)
	{ seterr (err, errlen, "failed to read hashValue") ; goto fail ; }
    }
  if (out->hashNumber < 1 || (uint64_t) out->hashNumber > ((uint64_t) 1 << B))
    { seterr (err, errlen, "hashNumber outside the table") ; goto fail ; }
  /* hashDepth Array */
  if (fread (&a, sizeof (a), 1, f) != 1 || a.size != 4 || a.dim < a.max || a.max < 0)
    { seterr (err, errlen, "failed to read hashDepth array") ; goto fail ; }
  { size_t n = (size_t) a.dim > out->hashNumber ? (size_t) a.dim : out->hashNumber ;
    if (!(out->hashDepth = calloc (n + 1, 4))) { st = H10X_ERR_NOMEM ; goto fail ; }
    if (fread (out->hashDepth, 4, a.dim, f) != (size_t) a.dim) { seterr (err, errlen, "failed to read hashDepth array") ; goto fail ; }
  }
  /* clusterBlocks Array */
  if (fread (&a, sizeof (a), 1, f) != 1 || a.size != 32 || a.dim < a.max || a.max < 1)
    { seterr (err, errlen, "failed to read clusterBlocks array") ; goto fail ; }
  out->nBlocksMax = (uint32_t) a.max ;
  if (a.dim == a.max) out->reserved |= H10X_INDEX_EXACT_BLOCKS ;
  if (!(cb = malloc ((size_t) a.dim * 32 + 32))) { st = H10X_ERR_NOMEM ; goto fail ; }
  if (fread (cb, 32, a.dim, f) != (size_t) a.dim) { seterr (err, errlen, "failed to read clusterBlocks array") ; goto fail ; }
  { uint32_t nb = out->nBlocksMax, b ;
    out->blkNRead = calloc (nb, 4) ; out->blkNHash = calloc (nb, 4) ; out->blkOff = calloc ((size_t) nb + 1, 8) ;
    out->blkNSubCluster = calloc (nb, 4) ; out->blkPointToMin = calloc (nb, 8) ; out->blkClusterParent = calloc (nb, 4) ;
    if (!out->blkNRead || !out->blkNHash || !out->blkOff || !out->blkNSubCluster || !out->blkPointToMin || !out->blkClusterParent)
      { st = H10X_ERR_NOMEM ; goto fail ; }
    for (b = 0 ; b < nb ; ++b)
      { out->blkNRead[b] = cb[8*b] ; out->blkNHash[b] = b ? cb[8*b + 1] : 0 ;
	out->blkNSubCluster[b] = cb[8*b + 2] ; memcpy (&out->blkPointToMin[b], &cb[8*b + 6], 8) ;
	out->blkClusterParent[b] = cb[8*b + 3] ;
	out->blkOff[b] = out->nHashes ;
	if (b) { out->nReads += out->blkNRead[b] ; out->nHashes += out->blkNHash[b] ; }
      }
    out->blkOff[nb] = out->nHashes ;
  }
  if (!(out->clusHash = malloc (out->nHashes * 8 + 8))) { st = H10X_ERR_NOMEM ; goto fail ; }
  if (fread (out->clusHash, 8, out->nHashes, f) != out->nHashes) { seterr (err, errlen, "read fail 3") ; goto fail ; }
  /* The reference trusts the file.  Here its arrays go on to index device memory (k_good_mark by bin, k_subcluster's
     256 labels per block, the reads of a block), so what the commands rely on is checked once: every entry names an
     existing bin and a read of its block, no block claims more than 255 sub-clusters, and the depths add up to the entries. */
  { uint64_t depthSum = 0, e ; uint32_t b ;
    for (b = 1 ; b < out->hashNumber ; ++b) depthSum += out->hashDepth[b] ;
    if (depthSum != out->nHashes) { seterr (err, errlen, "inconsistent hash file: bin depths do not add up to the entries") ; goto fail ; }
    for (b = 1 ; b < out->nBlocksMax ; ++b)
      { uint32_t nr = out->blkNRead[b] ? out->blkNRead[b] : 1 ;
	if (out->blkNSubCluster[b] > 255) { seterr (err, errlen, "inconsistent hash file: more than 255 sub-clusters in a block") ; goto fail ; }
	for (e = out->blkOff[b] ; e < out->blkOff[b + 1] ; ++e)
	  { const h10x_cluster_hash *ch = &out->clusHash[e] ;
	    if (ch->hash >= out->hashNumber || ch->read >= nr)	/* a stale subCluster label above nSubCluster is legal: it counts as 0 */
	      { seterr (err, errlen, "inconsistent hash file: a ClusterHash entry is out of range") ; goto fail ; }
	  }
      }
  }
  free (cb) ;
  fclose (f) ;
  return H10X_OK ;
 fail:
  free (cb) ;
  fclose (f) ;
  free (out->hashIndex) ; free (out->hashValue) ; free (out->hashDepth) ; free (out->blkNRead) ;
  free (out->blkNHash) ; free (out->blkOff) ; free (out->clusHash) ; free (out->blkNSubCluster) ; free (out->blkPointToMin) ; free (out->blkClusterParent) ;
  memset (out, 0, sizeof (*out)) ;
  return st ;
}
