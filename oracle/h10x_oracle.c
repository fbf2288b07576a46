/* h10x_oracle.c - CPU restatement of hash10x's `--readFQB` index build.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: it may be
 * loaded by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * `--impl reference` legs, always as the checker or the timed CPU baseline, never as
 * the thing shipped.  The product (hash10x_b200/) fails loudly without its CUDA library.
 *
 * Parity pinning: the reference ships no golden vectors.  This
 * restatement is pinned (tests/test_oracle_vs_reference.py, tests/test_golden.py)
 * against (i) the reference binary itself, compiled unmodified from /root/reference by
 * oracle/Makefile into oracle/_ref/, run on the same FQB files, (ii) the known-answer
 * vectors of SURVEY.md Appendix E produced by the reference seqhash.c, and (iii)
 * fixtures under tests/golden/ generated from the reference binary by
 * tests/golden/make_golden.py.
 *
 * Each function cites the reference lines (in /root/reference) whose behaviour it states.
 * It is written as straight array code, not in the reference's iterator/Array idiom.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t U8 ; typedef uint16_t U16 ; typedef uint32_t U32 ; typedef uint64_t U64 ;

enum { ORC_OK = 0, ORC_TABLE_TOO_SMALL = 1, ORC_CHUNK_TOO_SMALL = 2, ORC_BAD_PARAM = 3,
       ORC_NOMEM = 4, ORC_IO = 5 } ;

typedef struct {
  int32_t k, w, B, status ;
  U64 factor1 ;
  U32 hashNumber ;		/* bins are 1..hashNumber-1 (hash10x.c:113,147) */
  U32 nBlocksMax ;		/* arrayMax(clusterBlocks) = runs + 1 (block 0 is a dummy) */
  U64 nReads ;			/* records consumed, including the unhashed last run */
  U64 nHashes ;			/* sum of nHash over processed blocks */
  U32 *hashIndex ;		/* 2^B open-addressing table of bin ids */
  U64 *hashValue ;		/* hashNumber values, [0] unused */
  U32 *hashDepth ;		/* hashNumber depths, [0] = 0 */
  U32 *blkNRead ;		/* nBlocksMax */
  U32 *blkNHash ;		/* nBlocksMax */
  U64 *blkOff ;			/* nBlocksMax+1 offsets into clus */
  U64 *clus ;			/* nHashes ClusterHash as idx | (U64)read16 << 32 (bytes 6,7 zero) */
  U64 *codeOff ;		/* hashNumber+1 offsets into codes (fillHashTable as CSR) */
  U32 *codes ;			/* nHashes block numbers, ascending within each bin */
} orc_index ;

/* ---- seqhash.c:20-35 with initialise() hash10x.c:1101: factor1 from glibc random() ----
   gcc evaluates the left random() first; seed 17 -> 0x49308bb9003cb3ad */
U64 orc_factor1 (int seed)
{ srandom ((unsigned) seed) ;
  U64 hi = (U64) random () ; U64 lo = (U64) random () ;
  return (hi << 32) | lo | 1 ;
}

/* ---- seqhash.c:58-69: multiplicative hash of a k-mer and of its reverse complement ---- */
static inline U64 kmer_hash (U64 x, U64 factor1, int k) { return (x * factor1) >> (64 - 2*k) ; }

void orc_kmer_hashes (U64 h, U64 hRC, U64 factor1, int k, U64 *hashF, U64 *hashR)
{ *hashF = kmer_hash (h, factor1, k) ; *hashR = kmer_hash (hRC, factor1, k) ; }

/* ---- seqhash.c:154-195 (moshRCiterator/moshRCnext) over one base string ----
   every k-mer start j in 0..len-k whose canonical hash is a multiple of w, left to right.
   s[] holds 2-bit codes.  Returns the number written (at most cap). */
int orc_seq_moshes (const U8 *s, int len, int k, int w, U64 factor1,
		    U64 *outHash, int *outPos, U8 *outFwd, int cap)
{
  if (len < k) return 0 ;	/* seqhash.c:162 */
  U64 mask = (k == 32) ? ~(U64)0 : (((U64)1 << (2*k)) - 1) ;
  U64 h = 0, hRC = 0 ;
  int n = 0, j, i ;
  for (i = 0 ; i < k-1 ; ++i)	/* seqhash.c:165-168 */
    { h = (h << 2) | s[i] ; hRC = (hRC >> 2) | ((U64)(3 - s[i]) << (2*(k-1))) ; }
  for (j = 0 ; j + k <= len ; ++j)
    { U8 b = s[j+k-1] ;
      h = ((h << 2) & mask) | b ;			/* seqhash.c:74 */
      hRC = (hRC >> 2) | ((U64)(3 - b) << (2*(k-1))) ;	/* seqhash.c:75,33 */
      U64 hf = kmer_hash (h, factor1, k), hr = kmer_hash (hRC, factor1, k) ;
      U64 x = (hf < hr) ? hf : hr ;			/* seqhash.c:67-68 */
      if (x % (U64) w == 0)				/* seqhash.c:171,189 */
	{ if (n < cap)
	    { outHash[n] = x ; if (outPos) outPos[n] = j ; if (outFwd) outFwd[n] = (hf < hr) ; }
	  ++n ;
	}
    }
  return n ;
}

/* ---- hash10x.c:108-119 unpackFQB (bases only; quals are never used, hash10x.c:161) ---- */
static void unpack_bases (const U32 *u, U8 *s)	/* 10 words -> 160 codes */
{ int i, j ;
  for (i = 0 ; i < 10 ; ++i) for (j = 16 ; j-- ; ) *s++ = (u[i] >> (2*j)) & 3 ;
}

/* moshes of one record in the order processBlock generates them (hash10x.c:160-164):
   read 1 bases [23,150) then read 2 bases [0,150).  which[] = 0 for read 1, 1 for read 2. */
int orc_record_moshes (const U32 *rec, int k, int w, U64 factor1,
		       U64 *outHash, int *outPos, U8 *outWhich, int cap)
{
  U8 s1[160], s2[160] ;
  unpack_bases (rec, s1) ; unpack_bases (rec + 15, s2) ;
  int n1 = orc_seq_moshes (s1 + 23, 127, k, w, factor1, outHash, outPos, 0, cap) ;
  int m1 = n1 < cap ? n1 : cap, i ;
  for (i = 0 ; i < m1 ; ++i) if (outWhich) outWhich[i] = 0 ;
  int n2 = orc_seq_moshes (s2, 150, k, w, factor1, outHash + m1, outPos ? outPos + m1 : 0, 0, cap - m1) ;
  int m2 = n2 < cap - m1 ? n2 : cap - m1 ;
  for (i = 0 ; i < m2 ; ++i) if (outWhich) outWhich[m1+i] = 1 ;
  return n1 + n2 ;
}

/* ------------------------------------------------------------------------------------- */

typedef struct { U64 hash ; U32 read ; U32 seq ; } Mosh ;

static int cmp_mosh (const void *a, const void *b)
{ const Mosh *x = a, *y = b ;
  if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1 ;
  return x->seq < y->seq ? -1 : (x->seq > y->seq) ;	/* = glibc's stable qsort, SURVEY D4 */
}

static int cmp_u64 (const void *a, const void *b)
{ U64 x = *(const U64*)a & 0xffffffffu, y = *(const U64*)b & 0xffffffffu ; /* by bin id */
  return x < y ? -1 : (x > y) ;
}

typedef struct {
  orc_index *ix ;
  U64 tableSize, tableMask ;
  Mosh *m ; size_t mCap ;
  size_t clusCap, blkCap ;
} Build ;

/* ---- hash10x.c:139-152 hashIndexFind(hash, TRUE) ---- */
static int bin_find_add (Build *bd, U64 hash, U32 *out)
{ orc_index *ix = bd->ix ;
  U64 offset = hash & bd->tableMask ;
  U64 diff = ((hash >> ix->B) & bd->tableMask) | 1 ;
  U32 idx ;
  while ((idx = ix->hashIndex[offset]) && ix->hashValue[idx] != hash)
    offset = (offset + diff) & bd->tableMask ;
  if (!idx)
    { idx = ix->hashIndex[offset] = ix->hashNumber++ ;
      ix->hashValue[idx] = hash ;
      if (ix->hashNumber > (bd->tableSize >> 2) - 2) return ORC_TABLE_TOO_SMALL ; /* :149 */
    }
  *out = idx ;
  return ORC_OK ;
}

/* ---- hash10x.c:154-186 processBlock ---- */
static int process_block (Build *bd, const U32 *recs, U32 nRead, U32 blk)
{
  orc_index *ix = bd->ix ;
  size_t need = (size_t) nRead * 260 + 1 ;
  if (need > bd->mCap)
    { bd->mCap = need * 2 ; free (bd->m) ;
      if (!(bd->m = malloc (bd->mCap * sizeof (Mosh)))) return ORC_NOMEM ;
    }
  Mosh *m = bd->m ;
  U64 hs[260] ;
  size_t n = 0 ; U32 i ; int j, c ;
  for (i = 0 ; i < nRead ; ++i)
    { c = orc_record_moshes (recs + 30*(size_t)i, ix->k, ix->w, ix->factor1, hs, 0, 0, 260) ;
      for (j = 0 ; j < c ; ++j) { m[n].hash = hs[j] ; m[n].read = i ; m[n].seq = (U32) n ; ++n ; }
    }
  size_t nu ;
  if (!n)			/* hash10x.c:167-168 on an empty zero-filled Array: phantom {0,0} */
    { m[0].hash = 0 ; m[0].read = 0 ; m[0].seq = 0 ; nu = 1 ; }
  else
    { qsort (m, n, sizeof (Mosh), cmp_mosh) ;
      size_t t ;
      for (nu = 1, t = 1 ; t < n ; ++t)	/* hash10x.c:168-172: keep the first of each run */
	if (m[t].hash != m[nu-1].hash) m[nu++] = m[t] ;
    }

  if (ix->nHashes + nu > bd->clusCap)
    { bd->clusCap = (ix->nHashes + nu) * 2 ;
      if (!(ix->clus = realloc (ix->clus, bd->clusCap * sizeof (U64)))) return ORC_NOMEM ;
    }
  U64 *ch = ix->clus + ix->nHashes ;
  size_t t ;
  for (t = 0 ; t < nu ; ++t)	/* hash10x.c:176-181 */
    { U32 idx ; int st = bin_find_add (bd, m[t].hash, &idx) ;
      if (st) return st ;
      ++ix->hashDepth[idx] ;
      ch[t] = (U64) idx | ((U64)(U16) m[t].read << 32) ;	/* read is U16, hash10x.c:37,180 */
    }
  qsort (ch, nu, sizeof (U64), cmp_u64) ;	/* hash10x.c:183; ids within a block are distinct */
  ix->blkNHash[blk] = (U32) nu ;
  ix->blkOff[blk] = ix->nHashes ;
  ix->nHashes += nu ;
  return ORC_OK ;
}

static int grow_blocks (Build *bd, U32 need)
{ orc_index *ix = bd->ix ;
  if (need < bd->blkCap) return ORC_OK ;
  size_t cap = bd->blkCap ? bd->blkCap * 2 : 1200 ;
  while (cap <= need) cap *= 2 ;
  ix->blkNRead = realloc (ix->blkNRead, cap * sizeof (U32)) ;
  ix->blkNHash = realloc (ix->blkNHash, cap * sizeof (U32)) ;
  ix->blkOff = realloc (ix->blkOff, (cap + 1) * sizeof (U64)) ;
  if (!ix->blkNRead || !ix->blkNHash || !ix->blkOff) return ORC_NOMEM ;
  size_t i ;
  for (i = bd->blkCap ; i < cap ; ++i) { ix->blkNRead[i] = 0 ; ix->blkNHash[i] = 0 ; ix->blkOff[i] = 0 ; }
  bd->blkCap = cap ;
  return ORC_OK ;
}

void orc_free (orc_index *ix)
{ if (!ix) return ;
  free (ix->hashIndex) ; free (ix->hashValue) ; free (ix->hashDepth) ;
  free (ix->blkNRead) ; free (ix->blkNHash) ; free (ix->blkOff) ;
  free (ix->clus) ; free (ix->codeOff) ; free (ix->codes) ;
  free (ix) ;
}

/* ---- hash10x.c:1099-1118 initialise + :188-236 readFQB + :317-347 fillHashTable ----
   recs: the FQB "file" (nFile whole records; a trailing partial record is not passed in,
   as fread drops it).  The chunk loop is kept because three behaviours depend on it:
   the final run is never processed (:209,216), "chunkSize too small" (:206), and the
   `if (!barcode)` re-seed at each chunk start (:212) which glues an all-A (word 0)
   barcode run that ends exactly on a chunk boundary onto the run that follows.
   minB/maxB: the reference accepts 20..30 (:1107); tests may pass a wider range to model
   the "only the B bound relaxed" oracle of SURVEY.md 8c. */
orc_index *orc_build (const U32 *recs, U64 nFile, int k, int w, U64 factor1, int B,
		      int64_t N, int chunkSize, int minB, int maxB)
{
  orc_index *ix = calloc (1, sizeof (orc_index)) ;
  if (!ix) return 0 ;
  Build bd ; memset (&bd, 0, sizeof (bd)) ; bd.ix = ix ;
  ix->k = k ; ix->w = w ; ix->B = B ; ix->factor1 = factor1 ;
  if (k < 1 || k >= 32 || w < 1 || B < minB || B > maxB) { ix->status = ORC_BAD_PARAM ; return ix ; }
  bd.tableSize = (U64)1 << B ; bd.tableMask = bd.tableSize - 1 ;
  ix->hashIndex = calloc (bd.tableSize, sizeof (U32)) ;
  ix->hashValue = calloc (bd.tableSize >> 2, sizeof (U64)) ;
  ix->hashDepth = calloc (bd.tableSize >> 2, sizeof (U32)) ;
  if (!ix->hashIndex || !ix->hashValue || !ix->hashDepth) { ix->status = ORC_NOMEM ; return ix ; }
  ix->hashNumber = 1 ;
  if ((ix->status = grow_blocks (&bd, 2))) return ix ;

  U64 nReads = 0, pos = 0, start = 0 ;	/* start = file index of the open run's first record */
  U32 cur = 1, barcode = 0 ;
  ix->blkNRead[1] = 0 ;
  while (!N || nReads < (U64) N)
    { int64_t thisChunk = (int64_t) chunkSize - (int64_t) ix->blkNRead[cur] ;
      if (thisChunk <= 0) { ix->status = ORC_CHUNK_TOO_SMALL ; break ; }
      if (N && nReads + (U64) thisChunk > (U64) N) thisChunk = N - (int64_t) nReads ;
      U64 nRec = (nFile - pos < (U64) thisChunk) ? nFile - pos : (U64) thisChunk ;
      if (!nRec) break ;
      if (!barcode) barcode = recs[30*pos] ;
      U64 i ;
      for (i = 0 ; i < nRec ; ++i)
	{ U32 w0 = recs[30*(pos+i)] ;
	  if (w0 == barcode) ++ix->blkNRead[cur] ;
	  else
	    { if ((ix->status = process_block (&bd, recs + 30*start, ix->blkNRead[cur], cur))) goto done ;
	      ++cur ;
	      if ((ix->status = grow_blocks (&bd, cur + 1))) goto done ;
	      ix->blkNRead[cur] = 1 ; barcode = w0 ; start = pos + i ;
	    }
	}
      nReads += nRec ; pos += nRec ;
    }
 done:
  free (bd.m) ;
  ix->nReads = nReads ;
  ix->nBlocksMax = cur + 1 ;
  ix->blkOff[cur] = ix->nHashes ;		/* the last run: nHash 0, never processed */
  ix->blkOff[cur + 1] = ix->nHashes ;
  if (ix->status) return ix ;

  /* fillHashTable (hash10x.c:317-347) as CSR: bin i lists the blocks that hold it, ascending */
  U32 hn = ix->hashNumber, i ;
  ix->codeOff = calloc ((size_t) hn + 1, sizeof (U64)) ;
  ix->codes = malloc ((ix->nHashes ? ix->nHashes : 1) * sizeof (U32)) ;
  U64 *fill = calloc ((size_t) hn + 1, sizeof (U64)) ;
  if (!ix->codeOff || !ix->codes || !fill) { free (fill) ; ix->status = ORC_NOMEM ; return ix ; }
  for (i = 0 ; i < hn ; ++i) ix->codeOff[i+1] = ix->codeOff[i] + ix->hashDepth[i] ;
  U32 b ;
  for (b = 1 ; b < ix->nBlocksMax ; ++b)
    { U64 e ;
      for (e = ix->blkOff[b] ; e < ix->blkOff[b] + ix->blkNHash[b] ; ++e)
	{ U32 idx = (U32) ix->clus[e] ;
	  ix->codes[ix->codeOff[idx] + fill[idx]++] = b ;
	}
    }
  free (fill) ;
  return ix ;
}

/* ---- array.c:144-170 arrayExtend growth rule, replayed for sequential access 0..max-1 ---- */
static int array_dim_after (int dim, int size, int max)
{ int i = dim ;
  while (max > dim)		/* the access that triggers the extension is index i == dim */
    { i = dim ;
      if ((long) dim * size < (1 << 23)) dim *= 2 ; else dim += 1024 + ((1 << 23) / size) ;
      if (i >= dim) dim = i + 1 ;
    }
  return dim ;
}

static int write_array (FILE *f, const void *data, int size, int max, int dim0)
{ /* array.h:41-50 ArrayStruct on x86-64: int magic; pad; char *base; int dim,size,max; pad */
  int dim = array_dim_after (dim0, size, max) ;
  struct { int32_t magic, pad0 ; U64 base ; int32_t dim, size, max, pad1 ; } a =
    { 8918274, 0, 0, dim, size, max, 0 } ;
  if (fwrite (&a, 32, 1, f) != 1) return 0 ;
  if (max && fwrite (data, size, max, f) != (size_t) max) return 0 ;
  size_t rest = (size_t)(dim - max) * size ;
  if (rest) { void *z = calloc (1, rest) ; int ok = z && fwrite (z, 1, rest, f) == rest ; free (z) ; if (!ok) return 0 ; }
  return 1 ;
}

/* ---- hash10x.c:244-267 writeHashFile, SURVEY.md Appendix B.  Raw pointers are written as 0;
   Array dims follow the reference's growth so that the file has the reference's size. ---- */
int orc_write_hash_clustered (const orc_index *ix, const U32 *nSub, const double *pointToMin, const char *path)
{
  FILE *f = fopen (path, "wb") ; if (!f) return ORC_IO ;
  U32 version = 2 ; U16 chSize = 8, cbSize = 32 ; int32_t B = ix->B ;
  U64 tableSize = (U64)1 << ix->B ;
  int ok = fwrite ("10XH", 4, 1, f) == 1 && fwrite (&version, 4, 1, f) == 1 &&
    fwrite (&chSize, 2, 1, f) == 1 && fwrite (&cbSize, 2, 1, f) == 1 && fwrite (&B, 4, 1, f) == 1 ;
  ok = ok && fwrite (ix->hashIndex, 4, tableSize, f) == tableSize ;
  ok = ok && fwrite (&ix->hashNumber, 4, 1, f) == 1 ;
  ok = ok && fwrite (ix->hashValue, 8, ix->hashNumber, f) == ix->hashNumber ;
  /* arrayMax(hashDepth) stays 0 when no block was ever processed (hash10x.c:178 never ran) */
  ok = ok && write_array (f, ix->hashDepth, 4, ix->hashNumber > 1 ? (int) ix->hashNumber : 0, 1 << 20) ;
  /* ClusterBlock (hash10x.c:62-70): U32 nRead,nHash,nSubCluster,clusterParent; ptr; double */
  U32 nb = ix->nBlocksMax, b ;
  U32 *cb = calloc ((size_t) nb * 8, sizeof (U32)) ;
  if (!cb) { fclose (f) ; return ORC_NOMEM ; }
  for (b = 0 ; b < nb ; ++b)
    { cb[8*b] = ix->blkNRead[b] ; cb[8*b+1] = ix->blkNHash[b] ;
      if (nSub) cb[8*b+2] = nSub[b] ;
      if (pointToMin) memcpy (&cb[8*b+6], &pointToMin[b], 8) ;
    }
  ok = ok && write_array (f, cb, 32, (int) nb, 1200) ;
  free (cb) ;
  ok = ok && (!ix->nHashes || fwrite (ix->clus, 8, ix->nHashes, f) == ix->nHashes) ;
  if (fclose (f)) ok = 0 ;
  return ok ? ORC_OK : ORC_IO ;
}

int orc_write_hash (const orc_index *ix, const char *path)
{ return orc_write_hash_clustered (ix, NULL, NULL, path) ; }

/* ---- hash10x.c:528-539 hashWithinRangeBuild + :738-766 goodHashesBuild ("next" row f1) ----
   within[] (hashNumber bytes) is in/out: the reference only ever SETS flags (:535), so ranges accumulate
   over successive --hashDepthRange commands.  good[] receives, block after block, the indices (into the
   block's ClusterHash list) of the entries whose bin is within range, ordered by increasing bin depth;
   glibc's qsort is a stable merge sort, so ties keep list order.  Blocks with more than 65535 hashes
   get an empty list (:748).  PARITY: the reference never prints goodHashes, but its --cluster walks them in
   order, so this restatement is pinned together with orc_cluster below (tests/test_oracle.py::
   test_cluster_equals_reference_binary). */
typedef struct { U32 depth ; U16 idx ; } GoodKey ;
static int cmp_good (const void *a, const void *b)
{ const GoodKey *x = a, *y = b ;
  if (x->depth != y->depth) return x->depth < y->depth ? -1 : 1 ;
  return (int) x->idx - (int) y->idx ;
}

int orc_good_hashes (U32 hashNumber, const U32 *hashDepth, U32 nBlocksMax, const U32 *blkNHash, const U64 *blkOff,
		     const U64 *clus, int min, int max, U8 *within, U64 *goodOff, U16 *good)
{ U32 i, c ;
  struct { U32 hashNumber, nBlocksMax ; const U32 *hashDepth, *blkNHash ; const U64 *blkOff, *clus ; } x =
    { hashNumber, nBlocksMax, hashDepth, blkNHash, blkOff, clus }, *ix = &x ;
  for (i = 0 ; i < ix->hashNumber ; ++i)
    { int n = (int) ix->hashDepth[i] ; if (n >= min && n < max) within[i] = 1 ; }
  GoodKey *keys = malloc (65536 * sizeof (GoodKey)) ;
  if (!keys) return ORC_NOMEM ;
  U64 out = 0 ;
  for (c = 0 ; c < ix->nBlocksMax ; ++c)
    { goodOff[c] = out ;
      U32 nh = c ? ix->blkNHash[c] : 0 ;
      if (nh > 65535) continue ;
      const U64 *ch = ix->clus + ix->blkOff[c] ;
      int n = 0 ;
      for (i = 0 ; i < nh ; ++i)
	{ U32 id = (U32) ch[i] ;
	  if (within[id]) { keys[n].depth = ix->hashDepth[id] ; keys[n].idx = (U16) i ; ++n ; }
	}
      qsort (keys, n, sizeof (GoodKey), cmp_good) ;
      for (i = 0 ; i < (U32) n ; ++i) good[out++] = keys[i].idx ;
    }
  goodOff[ix->nBlocksMax] = out ;
  free (keys) ;
  return ORC_OK ;
}

/* ---- hash10x.c:770-835 codeClusterFind + :837-868 codeClusterReadMerge ("next" row f2), as run by the
   `--cluster codeMin codeMax` branch of the command loop (hash10x.c:1241-1256) ----
   In/out: clus (byte 6 of every ClusterHash = subCluster), nSub and pointToMin (ClusterBlock.nSubCluster /
   .pointToMin, hash10x.c:62-70).  goodOff/good are the goodHashes lists of orc_good_hashes.  Written as the
   reference runs it: good hash 0 is never scanned (the loop starts at i = 1, :789), codes first seen at step i
   are counted at index i and therefore never enter that step's maximum (:793-806), a 256th sub-cluster abandons
   the block without resetting pointToMin (:810-817), and the read merge relabels by connected components with
   the smallest label winning (:848-858).  PARITY: pinned against the reference binary by
   tests/test_oracle.py::test_cluster_equals_reference_binary (`--readHash --hashDepthRange --cluster --writeHash`
   on a file whose subCluster bytes are zero, so the reference's uninitialised malloc bytes, hash10x.c:175, do
   not enter). */
/* ClusterHash entries outside the current good lists keep the labels of an earlier --cluster command; when such a
   label exceeds the block's new nSubCluster the reference indexes past trueCluster[] (hash10x.c:843,846: undefined
   behaviour).  Here, and in the CUDA path, such a label counts as 0 = unclustered; this counter lets the tests
   check that a pinned case never depends on it. */
static U64 orcStaleLabels = 0 ;
U64 orc_cluster_stale_labels (void) { return orcStaleLabels ; }

int orc_cluster (const U32 *hashDepth, const U64 *codeOff, const U32 *codes, U32 nBlocksMax,
		 const U32 *blkNRead, const U32 *blkNHash, const U64 *blkOff, U64 *clus,
		 const U64 *goodOff, const U16 *good, int codeMin, int codeMax, int clusterThreshold,
		 U32 *nSub, double *pointToMin)
{ if (!codeMin) codeMin = 1 ;
  if (!codeMax) codeMax = (int) nBlocksMax ;
  if (codeMax > (int) nBlocksMax) return ORC_BAD_PARAM ;	/* the reference would index past clusterBlocks */
  int *minShare = calloc (nBlocksMax ? nBlocksMax : 1, sizeof (int)) ;
  int *minShareCount = malloc (65536 * sizeof (int)) ;
  int clusterMin[257] ;
  if (!minShare || !minShareCount) { free (minShare) ; free (minShareCount) ; return ORC_NOMEM ; }
  int code ;
  for (code = codeMin ; code < codeMax ; ++code)
    { U8 *cb = (U8*) (clus + blkOff[code]) ;		/* ClusterHash i is bytes 8i..8i+7; subCluster is byte 6 */
      const U64 *ch = clus + blkOff[code] ;
      const U16 *g = good + goodOff[code] ;
      int n = (int) (goodOff[code+1] - goodOff[code]) ;
      int i, j ;
      /* ---- codeClusterFind ---- */
      if (n)
	{ for (i = 0 ; i < n ; ++i) cb[8 * (size_t) g[i] + 6] = 0 ;
	  nSub[code] = 0 ; pointToMin[code] = 0.0 ;
	  memset (minShare, 0, (size_t) nBlocksMax * sizeof (int)) ;
	  for (i = 1 ; i < n ; ++i)
	    { U32 x = (U32) ch[g[i]] ;
	      U32 nc = hashDepth[x] ;
	      memset (minShareCount, 0, (size_t) n * sizeof (int)) ;
	      for (j = 0 ; j < (int) nc ; ++j)
		{ int cj = (int) codes[codeOff[x] + j] ;
		  if (cj == code) continue ;
		  if (!minShare[cj]) minShare[cj] = i + 1 ;
		  ++minShareCount[minShare[cj] - 1] ;
		}
	      int msMax = 0, msTot = 0, msBest = 0 ;
	      for (j = 0 ; j < i ; ++j)
		{ if (minShareCount[j] > msMax) { msBest = j ; msMax = minShareCount[j] ; }
		  msTot += minShareCount[j] ;
		}
	      if (msMax >= clusterThreshold)
		{ U8 *sub = &cb[8 * (size_t) g[msBest] + 6] ;
		  if (!*sub)
		    { if (++nSub[code] > 255)
			{ nSub[code] = 0 ;
			  for (j = 0 ; j < i ; ++j) cb[8 * (size_t) g[j] + 6] = 0 ;
			  break ;
			}
		      *sub = (U8) nSub[code] ;
		      clusterMin[nSub[code]] = msBest ;
		    }
		  cb[8 * (size_t) g[i] + 6] = *sub ;
		  pointToMin[code] += minShareCount[clusterMin[*sub]] / (double) msTot ;
		}
	    }
	}
      /* ---- codeClusterReadMerge ---- */
      if (!nSub[code]) continue ;
      { int ns = (int) nSub[code] ;
	U32 nRead = blkNRead[code], nHash = blkNHash[code] ;
	int *readMap = calloc (nRead ? nRead : 1, sizeof (int)) ;
	int trueCluster[257], deadCluster[257] ;
	if (!readMap) { free (minShare) ; free (minShareCount) ; return ORC_NOMEM ; }
	for (i = 0 ; i <= 256 ; ++i) { trueCluster[i] = i <= ns ? i : 0 ; deadCluster[i] = 0 ; }
	for (i = 0 ; i < (int) nHash ; ++i)
	  { if (cb[8 * (size_t) i + 6] > ns) ++orcStaleLabels ;
	    int hashCluster = trueCluster[cb[8 * (size_t) i + 6]] ; if (!hashCluster) continue ;
	    U16 rd = (U16) (ch[i] >> 32) ;
	    int readCluster = trueCluster[readMap[rd]] ;
	    if (hashCluster == readCluster) continue ;
	    else if (!readCluster) readMap[rd] = hashCluster ;
	    else
	      { if (hashCluster > readCluster) { int t = hashCluster ; hashCluster = readCluster ; readCluster = t ; }
		for (j = 1 ; j <= ns ; ++j) if (trueCluster[j] == readCluster) trueCluster[j] = hashCluster ;
		deadCluster[readCluster] = 1 ;
	      }
	  }
	for (j = 1 ; j <= ns ; ++j) deadCluster[j] = deadCluster[j-1] + 1 - deadCluster[j] ;
	for (j = 1 ; j <= ns ; ++j) trueCluster[j] = deadCluster[trueCluster[j]] ;
	nSub[code] = (U32) deadCluster[ns] ;
	for (i = 0 ; i < (int) nHash ; ++i) cb[8 * (size_t) i + 6] = (U8) trueCluster[cb[8 * (size_t) i + 6]] ;
	free (readMap) ;
      }
    }
  free (minShare) ; free (minShareCount) ;
  return ORC_OK ;
}
