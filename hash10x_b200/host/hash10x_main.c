/* hash10x_main.c - `hash10x-b200`: hash10x's command loop over the B200 index build.
 *
 * Keeps the reference program's interface for the --readFQB pipeline (hash10x.c:1122-1305):
 * positional, order-sensitive command chaining; -k -w -r -B -N -c -o set globals that the next
 * operation uses; every command is echoed as "COMMAND ..." and followed by a resource line; fatal
 * errors are "FATAL ERROR: ...\n" on stderr with exit(-1) (utils.c:18-29).  --readFQB runs on the
 * GPU through libh10xgpu.so (include/h10x_gpu.h) - there is no CPU build in this program - and
 * leaves the same state behind (hash table, values, depths, block table, ClusterHash lists,
 * hash->code lists) for --writeHash, --hashStats and --codeStats, which are O(bins) / O(blocks) host report
 * passes exactly as in the reference.  --hashDepthRange and --cluster (with -ct) run on the GPU on the resident
 * index and have no CPU version here.  --readHash loads a .hash file written by either program and, when a CUDA
 * device is present, moves the index onto it so that the two commands serve read-then-cluster sessions too.  The
 * research probes (--clusterReport, --cribBuild, --hashExplore, ...) are out of scope and
 * die with a message saying so.
 */
#define _GNU_SOURCE
#include "h10x_gpu.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>

static struct { int k, w, r, B, N, chunkSize, clusterThreshold, gpus, wideB ; } params ;
static FILE *outFile ;
static h10x_index ix ;		/* the state --readFQB / --readHash leave behind */
static int haveIndex = 0, indexFromGpu = 0 ;
static h10x_ctx *ctx = 0 ;
static h10x_multi *multi = 0 ;	/* --gpus N: one context per GPU, kept for --hashDepthRange / --cluster */
static long totalAllocated = 0 ;

static void resetDerived (void) ;	/* drops what --hashDepthRange built on the previous index */

static void die (const char *format, ...)
{ va_list args ;
  va_start (args, format) ;
  fprintf (stderr, "FATAL ERROR: ") ; vfprintf (stderr, format, args) ; fprintf (stderr, "\n") ;
  va_end (args) ;
  exit (-1) ;
}

/* utils.c:122-150: rusage deltas since the previous call */
static struct rusage rOld, rFirst ;
static void timeUpdate (FILE *f)
{ static int isFirst = 1 ;
  struct rusage rNew ;
  getrusage (RUSAGE_SELF, &rNew) ;
  if (!isFirst)
    { long secs = rNew.ru_utime.tv_sec - rOld.ru_utime.tv_sec, usecs = rNew.ru_utime.tv_usec - rOld.ru_utime.tv_usec ;
      if (usecs < 0) { usecs += 1000000 ; secs -= 1 ; }
      fprintf (f, "user\t%d.%06d", (int) secs, (int) usecs) ;
      secs = rNew.ru_stime.tv_sec - rOld.ru_stime.tv_sec ; usecs = rNew.ru_stime.tv_usec - rOld.ru_stime.tv_usec ;
      if (usecs < 0) { usecs += 1000000 ; secs -= 1 ; }
      fprintf (f, "\tsystem\t%d.%06d", (int) secs, (int) usecs) ;
      fprintf (f, "\tmax_RSS\t%ld", rNew.ru_maxrss - rOld.ru_maxrss) ;
      fprintf (f, "\tmemory\t%li", totalAllocated) ;
      fputc ('\n', f) ;
    }
  else { rFirst = rNew ; isFirst = 0 ; }
  rOld = rNew ;
}
static void timeTotal (FILE *f) { rOld = rFirst ; timeUpdate (f) ; }

static void usage (void)
{ fprintf (stderr, "Usage: hash10x-b200 <commands>\n") ;
  fprintf (stderr, "Commands can be parameter settings with -x, or operations:\n") ;
  fprintf (stderr, "Be sure to set relevant parameters before invoking an operation!\n") ;
  fprintf (stderr, "   -k <kmer size> [%d]\n", params.k) ;
  fprintf (stderr, "   -w <window> [%d]\n", params.w) ;
  fprintf (stderr, "   -r <random number seed> [%d]\n", params.r) ;
  fprintf (stderr, "   -B <hash index table bitcount> [%d]\n", params.B) ;
  fprintf (stderr, "   -N <num records to read: 0 for all> [%d]\n", params.N) ;
  fprintf (stderr, "   -c <file chunkSize in readPairs> [%d]\n", params.chunkSize) ;
  fprintf (stderr, "   --gpus <number of GPUs for --readFQB> [%d]\n", params.gpus) ;
  fprintf (stderr, "   -o | --output <output filename> : '-' for stdout\n") ;
  fprintf (stderr, "   --readFQB <sorted fqb input file name>: must have this or readHash (runs on the GPU)\n") ;
  fprintf (stderr, "   --readHash <hash input file name>\n") ;
  fprintf (stderr, "   --writeHash <hash output file name>\n") ;
  fprintf (stderr, "   --hashDepthRange <min> <max>: set limits for hash counts\n") ;
  fprintf (stderr, "   -ct | --clusterThreshold <clusterThreshold> [%d]\n", params.clusterThreshold) ;
  fprintf (stderr, "   --cluster <codeMin> <codeMax> : cluster this range of barcodes; 0 for codeMax means to end (runs on the GPU)\n") ;
  fprintf (stderr, "   --clusterSplit : every sub-cluster becomes a barcode block of its own (runs on the GPU)\n") ;
  fprintf (stderr, "   --hashStats : distribution of hash counts and summary info\n") ;
  fprintf (stderr, "   --codeStats : distribution of barcode/cluster sizes and summary info\n") ;
  fprintf (stderr, "   --gpuStats : per-stage device times and roofline bytes of the last --readFQB\n") ;
  fprintf (stderr, "   --cribBuild <genome1.fa> <genome2.fa>: match to genomic hashes (runs on the GPU)\n") ;
  fprintf (stderr, "   --wideB : accept -B 31 to 34 (hash10x-b200 only; README.md:55)\n") ;
  fprintf (stderr, "   --help : print this usage message\n") ;
}

/* ---- initialise() hash10x.c:1099-1118 ---- */
static void initialise (int k, int w, int r, int B)
{ if (k <= 0 || w <= 0) die ("k %d, w %d must be > 0; run without args for usage", k, w) ;
  if (k >= 32) die ("seqhash k %d must be between 1 and 32\n", k) ;
  /* hash10x.c:1107 stops at 30; README.md:55 and moshset.c:17 speak of up to 34: --wideB opts into 31..34 (a 8..64 GiB
     hashIndex), with the same table algorithm, which the reference's own binary cannot read back */
  if (B < 20 || B > (params.wideB ? 34 : 30)) die ("hashTableBits %d out of range 20-%d", B, params.wideB ? 34 : 30) ;
  if (haveIndex) { h10x_index_free (&ix) ; memset (&ix, 0, sizeof (ix)) ; haveIndex = 0 ; }
  fprintf (outFile, "hash10x initialised with k = %d, w = %d, random seed = %d, hashtable bits = %d\n", k, w, r, B) ;
}

/* ---- --readFQB: readFQB() + fillHashTable(), hash10x.c:188-236,317-347, on the GPU ---- */
static void readFQB (const char *path)
{ char err[512] ;
  h10x_params p ;
  memset (&p, 0, sizeof (p)) ;
  p.k = params.k ; p.w = params.w ; p.B = params.B ; p.N = params.N ; p.chunkSize = params.chunkSize ;
  p.factor1 = h10x_factor1_from_seed (params.r) ;
  p.device = getenv ("H10X_DEVICE") ? atoi (getenv ("H10X_DEVICE")) : 0 ;
  p.flags = H10X_FLAG_LAZY_CODES | (params.wideB ? H10X_FLAG_WIDE_B : 0) ;	/* no command below reads hashCodes on the host: --cluster runs where they are */
  printf ("  reading and processing sorted fqb file with chunkSize %d", params.chunkSize) ;
  if (params.N) printf (", first %d records", params.N) ;
  printf ("\n") ;
  if (params.chunkSize <= 0) die ("chunkSize too small") ;
  resetDerived () ;		/* the good lists may point into the context's memory */
  if (ctx) { h10x_gpu_destroy (ctx) ; ctx = 0 ; }
  int st ;
  if (multi) { h10x_multi_destroy (multi) ; multi = 0 ; }
  if (params.gpus > 1)		/* one thread and one context per GPU, NCCL inside the library; the contexts stay for the commands below */
    st = h10x_multi_build_file (&p, params.gpus, path, &multi, &ix, err, sizeof (err)) ;
  else
    { if (!(ctx = h10x_gpu_create (&p, err, sizeof (err)))) die ("%s", err) ;
      st = h10x_gpu_build_file (ctx, path, &ix, err, sizeof (err)) ;
    }
  if (st == H10X_ERR_TABLE_TOO_SMALL) die ("hashTableSize is too small") ;
  else if (st == H10X_ERR_CHUNK_TOO_SMALL) die ("chunkSize too small") ;
  else if (st == H10X_ERR_IO) die ("file read problem") ;
  else if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
  haveIndex = 1 ; indexFromGpu = 1 ;
  totalAllocated += ((long) 4 << ix.B) + 12L * ix.hashNumber + 12L * (long) ix.nHashes ;

  int nBarcodes = (int) ix.nBlocksMax - 1, i ;
  /* the reference prints these twice when the output is stdout (hash10x.c:230 tests `!= stdin`) */
  for (i = 0 ; i < 2 ; ++i)
    { FILE *f = i ? stdout : outFile ;
      fprintf (f, "  read %d read pair records for %d barcodes, mean %.2f read pairs per barcode\n",
	       (int) ix.nReads, nBarcodes, (int) ix.nReads / (double) nBarcodes) ;
      fprintf (f, "  created %ld hashes, mean %.2f hashes per read pair, %.2f per barcode\n",
	       (long) ix.nHashes, ix.nHashes / (double) (int) ix.nReads, ix.nHashes / (double) nBarcodes) ;
    }
  fprintf (outFile, "  filled hash table: %ld hashes from %d barcodes in %d bins\n",
	   (long) ix.nHashes, (int) ix.nBlocksMax, (int) ix.hashNumber) ;
}

/* ---- --readHash: readHashFile() hash10x.c:269-315; hash->code lists rebuilt as in fillHashTable ---- */
static void readHash (const char *path)
{ char err[512] ;
  resetDerived () ;
  int st = h10x_read_hash (path, params.B, &ix, err, sizeof (err)) ;
  if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
  haveIndex = 1 ; indexFromGpu = 0 ;
  fprintf (outFile, "  read %ld hashes for %ld reads in %d barcode blocks\n", (long) ix.nHashes, (long) ix.nReads, (int) ix.nBlocksMax) ;
  if (outFile != stdout)
    printf ("  read %ld hashes for %ld reads in %d barcode blocks\n", (long) ix.nHashes, (long) ix.nReads, (int) ix.nBlocksMax) ;
  /* fillHashTable: counting pass over the ClusterHash lists in block order keeps lists ascending */
  uint32_t hn = ix.hashNumber, b ; uint64_t e ;
  ix.codeOff = calloc ((size_t) hn + 1, 8) ; ix.codes = malloc ((ix.nHashes ? ix.nHashes : 1) * 4) ;
  uint64_t *fill = calloc ((size_t) hn + 1, 8) ;
  if (!ix.codeOff || !ix.codes || !fill) die ("myalloc failure") ;
  long nHashes = 0 ;
  for (b = 1 ; b < hn ; ++b) nHashes += ix.hashDepth[b] ;
  for (b = 0 ; b < hn ; ++b) ix.codeOff[b+1] = ix.codeOff[b] + ix.hashDepth[b] ;
  for (b = 1 ; b < ix.nBlocksMax ; ++b)
    for (e = ix.blkOff[b] ; e < ix.blkOff[b] + ix.blkNHash[b] ; ++e)
      { uint32_t x = ix.clusHash[e].hash ; ix.codes[ix.codeOff[x] + fill[x]++] = b ; }
  free (fill) ;
  fprintf (outFile, "  filled hash table: %ld hashes from %d barcodes in %d bins\n", nHashes, (int) ix.nBlocksMax, (int) ix.hashNumber) ;
  /* with a GPU the index moves there as well, so that --hashDepthRange and --cluster run on it as after --readFQB */
  if (ctx) { h10x_gpu_destroy (ctx) ; ctx = 0 ; }
  if (multi) { h10x_multi_destroy (multi) ; multi = 0 ; }
  if (h10x_gpu_device_count () > 0)
    { h10x_params p ;
      memset (&p, 0, sizeof (p)) ;
      p.k = params.k ; p.w = params.w ; p.B = params.B ; p.N = params.N ; p.chunkSize = params.chunkSize > 0 ? params.chunkSize : 1 ;
      p.factor1 = h10x_factor1_from_seed (params.r) ;
      p.device = getenv ("H10X_DEVICE") ? atoi (getenv ("H10X_DEVICE")) : 0 ;
      if (!(ctx = h10x_gpu_create (&p, err, sizeof (err)))) die ("%s", err) ;
      st = h10x_gpu_load_index (ctx, &ix, err, sizeof (err)) ;
      if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
      indexFromGpu = 1 ;
    }
}

static void writeHash (const char *path)
{ if (!haveIndex) die ("write fail 1") ;
  int st = h10x_write_hash (&ix, path) ;
  if (st) die ("failed to open hash file %s", path) ;
  fprintf (outFile, "  wrote %lld hash table entries and %d barcode blocks\n", 1LL << ix.B, (int) ix.nBlocksMax) ;
  if (outFile != stdout) printf ("  wrote %lld hash table entries and %d barcode blocks\n", 1LL << ix.B, (int) ix.nBlocksMax) ;
}

/* ---- histogramReport hash10x.c:351-375: same arithmetic, including the int thresholds ---- */
static void histogramReport (const char *prefix, const int *a, int max)
{ uint64_t sum = 0, total = 0, partSum = 0, partTotal = 0, best = 0, massBest = 0 ;
  int i, median = 0, massMedian = 0, n99 = 0, nMass99 = 0, mode = 0, massMode = 0 ;
  for (i = 0 ; i < max ; ++i) { sum += a[i] ; total += i * a[i] ; }
  int t50 = sum * 0.5, tMass50 = total * 0.5, t99 = sum * 0.99, tMass99 = total * 0.99 ;
  for (i = 0 ; i < max ; ++i)
    { int n = a[i] ;
      partSum += n ; partTotal += i * n ;
      fprintf (outFile, "%s_HIST %6d %d %.4f %.4f\n", prefix, i, n, partSum / (double) sum, partTotal / (double) total) ;
      if ((uint64_t) n > best) { mode = i ; best = n ; }
      if ((uint64_t) (i * n) > massBest) { massMode = i ; massBest = i * n ; }
      if (partSum > (uint64_t) t50 && !median) median = i ;
      if (partTotal > (uint64_t) tMass50 && !massMedian) massMedian = i ;
      if (partSum > (uint64_t) t99 && !n99) n99 = i ;
      if (partTotal > (uint64_t) tMass99 && !nMass99) nMass99 = i ;
    }
  fprintf (outFile, "%s_STATS MEAN %.1f", prefix, total / (double) sum) ;
  fprintf (outFile, "  MODE %d  MEDIAN %d  PERCENT99 %d", mode, median, n99) ;
  fprintf (outFile, "  MASS_MODE %d  N50 %d  N99 %d\n", massMode, massMedian, nMass99) ;
}

static int *countHist (const uint32_t *v, uint32_t n, int *maxOut)
{ uint32_t i, top = 0 ;
  for (i = 0 ; i < n ; ++i) if (v[i] > top) top = v[i] ;
  int *a = calloc ((size_t) top + 1, sizeof (int)) ;
  if (!a) die ("myalloc failure") ;
  for (i = 0 ; i < n ; ++i) ++a[v[i]] ;
  *maxOut = (int) top + 1 ;
  return a ;
}

/* hashDepthHist hash10x.c:377-386: over all of hashDepth, dummy bin 0 included */
static void hashStats (void)
{ if (!haveIndex || ix.hashNumber <= 1) { fprintf (stderr, "  no hash list to print stats for\n") ; return ; }
  int max ; const int *d ; char err[512] ;
  if (ctx && indexFromGpu && !h10x_gpu_histogram (ctx, 0, &d, &max, err, sizeof (err)))	/* counted where the index is */
    { histogramReport ("HASH_COUNT", d, max) ; return ; }
  int *a = countHist (ix.hashDepth, ix.hashNumber, &max) ;
  histogramReport ("HASH_COUNT", a, max) ;
  free (a) ;
}

/* codeSizeHist hash10x.c:388-402: over all blocks, dummy block 0 and the unhashed last one included */
static void codeStats (void)
{ if (!haveIndex || !ix.nBlocksMax) { fprintf (stderr, "  no barcodes to print stats for\n") ; return ; }
  int max ; int *a ; const int *d ; char err[512] ;
  if (ctx && indexFromGpu && !h10x_gpu_histogram (ctx, 1, &d, &max, err, sizeof (err)))
    { histogramReport ("CODE_SIZE", d, max) ;
      if (ix.blkNSubCluster && !h10x_gpu_histogram (ctx, 2, &d, &max, err, sizeof (err)) && max > 1)
	histogramReport ("CODE_CLUSTER", d, max) ;
      return ;
    }
  a = countHist (ix.blkNHash, ix.nBlocksMax, &max) ;
  histogramReport ("CODE_SIZE", a, max) ;
  free (a) ;
  if (ix.blkNSubCluster)	/* only once some block has sub-clusters: arrayMax(clusterHist) > 1, hash10x.c:401 */
    { a = countHist (ix.blkNSubCluster, ix.nBlocksMax, &max) ;
      if (max > 1) histogramReport ("CODE_CLUSTER", a, max) ;
      free (a) ;
    }
}

/* ---- --hashDepthRange: hashWithinRangeBuild + goodHashesBuild, hash10x.c:528-539,738-766 ---- */
static unsigned char *hashWithinRange = 0 ;
static int hashRangeMin = 0, hashRangeMax = 0 ;
static uint16_t **goodHashes = 0 ;
static int *nGoodHashes = 0 ;

static void resetDerived (void)	/* a new index: the reference's initialise() starts over too (and leaks, hash10x.c:1099-1118) */
{ free (hashWithinRange) ; hashWithinRange = 0 ; hashRangeMin = hashRangeMax = 0 ;
  free (goodHashes) ; goodHashes = 0 ; free (nGoodHashes) ; nGoodHashes = 0 ;	/* the lists themselves: context slab or small leaks */
}

static void hashDepthRange (int min, int max)
{ if (!haveIndex) die ("cluster code called without setting hashDepthRange") ;
  uint32_t c ;
  if ((ctx || multi) && indexFromGpu)	/* the index is resident on the GPU(s): build the lists there */
    { char err[512] ; h10x_good_hashes g ; uint64_t nGood = 0 ;
      /* one GPU: the lists stay on the device, where --cluster reads them (no command of this program reads them on the
	 host; copying them would first pin 2 bytes per good hash of host memory, a second at the 1 Gb scale) */
      int st = multi ? h10x_multi_depth_range (multi, min, max, &g, err, sizeof (err))
	: h10x_gpu_depth_range_device (ctx, min, max, &nGood, err, sizeof (err)) ;
      if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
      if (!hashWithinRange) hashWithinRange = calloc (ix.hashNumber, 1) ;
      if (!hashWithinRange) die ("myalloc failure") ;
      if (multi) memcpy (hashWithinRange, g.within, ix.hashNumber) ;
      hashRangeMin = min ; hashRangeMax = max ;
      free (goodHashes) ; free (nGoodHashes) ;		/* the tables of an earlier --hashDepthRange (the lists are the context's) */
      goodHashes = calloc (ix.nBlocksMax, sizeof (uint16_t*)) ;		/* non-NULL = "--hashDepthRange has run" (hash10x.c:1258) */
      nGoodHashes = calloc (ix.nBlocksMax, sizeof (int)) ;
      if (!hashWithinRange || !goodHashes || !nGoodHashes) die ("myalloc failure") ;
      for (c = 0 ; c < ix.nBlocksMax ; ++c)
	{ if (c && ix.blkNHash[c] > 65535)
	    fprintf (stderr, "ignoring barcode %d - too many hashes %d > %d\n", (int) c, (int) ix.blkNHash[c], 65535) ;
	  if (multi)
	    { goodHashes[c] = g.good + g.goodOff[c] ;		/* stitched lists owned by the session */
	      nGoodHashes[c] = (int) (g.goodOff[c+1] - g.goodOff[c]) ;
	    }
	}
      printf ("  made goodHashes arrays for hash range %d to %d\n  ", hashRangeMin, hashRangeMax) ;
      timeUpdate (outFile) ; fflush (outFile) ;
      return ;
    }
  /* no GPU-resident index (no CUDA device after --readHash): this program has no CPU version */
  die ("--hashDepthRange runs on the GPU-resident index (--readFQB, or --readHash with a GPU present)") ;
}

/* ---- --cluster codeMin codeMax: hash10x.c:1241-1261, on the GPU (h10x_gpu_cluster) ---- */
static void clusterCodes (int codeMin, int codeMax)
{ if (!goodHashes)
    { fprintf (outFile, "!! you must set hashDepthRange before cluster\n") ;
      if (outFile != stdout) fprintf (stderr, "!! you must set hashDepthRange before cluster\n") ;
      return ;
    }
  if (!((ctx || multi) && indexFromGpu))
    die ("--cluster runs on the GPU-resident index (--readFQB, or --readHash with a GPU present): there is no CPU clustering in this program") ;
  if (!codeMin) codeMin = 1 ;
  if (!codeMax) codeMax = (int) ix.nBlocksMax ;
  char err[512] ; h10x_clusters cl ;
  int st = multi ? h10x_multi_cluster (multi, codeMin, codeMax, params.clusterThreshold, &cl, err, sizeof (err))
    : h10x_gpu_cluster (ctx, codeMin, codeMax, params.clusterThreshold, &cl, err, sizeof (err)) ;
  if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
  if (ix.pinned == 0)		/* an index h10x_read_hash malloc'ed: keep its own arrays (h10x_index_free frees them) */
    { size_t nb = ix.nBlocksMax ;
      memcpy (ix.clusHash, cl.clusHash, 8 * (size_t) ix.nHashes) ;
      if (!ix.blkNSubCluster) ix.blkNSubCluster = malloc (4 * nb) ;
      if (!ix.blkPointToMin) ix.blkPointToMin = malloc (8 * nb) ;
      if (!ix.blkNSubCluster || !ix.blkPointToMin) die ("myalloc failure") ;
      memcpy (ix.blkNSubCluster, cl.nSubCluster, 4 * nb) ; memcpy (ix.blkPointToMin, cl.pointToMin, 8 * nb) ;
    }
  else
    { ix.clusHash = cl.clusHash ;		/* the context's pinned copy, subCluster bytes set */
      ix.blkNSubCluster = cl.nSubCluster ; ix.blkPointToMin = cl.pointToMin ;
    }
  fprintf (outFile, "  clustered codes %d to %d\n", codeMin, codeMax) ;
  if (outFile != stdout) printf ("  clustered codes %d to %d\n", codeMin, codeMax) ;
}

/* ---- --clusterSplit: clusterSplitCodes hash10x.c:956-1013, on the GPU (h10x_gpu_cluster_split) ---- */
static void clusterSplit (void)
{ if (!haveIndex) die ("no index to split") ;
  if (multi || !(ctx && indexFromGpu))
    die ("--clusterSplit runs on the index resident on one GPU (--readFQB, or --readHash with a GPU present)") ;
  char err[512] ; h10x_index nx ; uint32_t nNew = 0 ;
  const int nOld = (int) ix.nBlocksMax ;
  int st = h10x_gpu_cluster_split (ctx, &nx, &nNew, err, sizeof (err)) ;
  if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
  if (ix.pinned == 0) h10x_index_free (&ix) ;		/* an index h10x_read_hash malloc'ed */
  ix = nx ;
  /* the reference keeps its goodHashes arrays, indexed by the OLD block numbers (a later --cluster without a new
     --hashDepthRange walks stale lists there); here they are dropped, hashWithinRange stays (flags only accumulate) */
  free (goodHashes) ; goodHashes = 0 ; free (nGoodHashes) ; nGoodHashes = 0 ;
  fprintf (outFile, "  made %d additional new barcodes from clusters in %d original barcodes\n", (int) nNew, nOld) ;
  if (outFile != stdout) printf ("  made %d additional new barcodes from clusters in %d original barcodes\n", (int) nNew, nOld) ;
  printf ("  cluster timepoint: ") ; timeUpdate (stdout) ;
  fprintf (outFile, "  filled hash table: %ld hashes from %d barcodes in %d bins\n",
	   (long) ix.nHashes, (int) ix.nBlocksMax, (int) ix.hashNumber) ;
}

/* ---- --cribBuild genome1.fa genome2.fa: cribBuild hash10x.c:426-510, on the GPU (h10x_gpu_crib_build) ---- */

/* readSequence (readseq.c:63-157) as cribAddGenome calls it (dna2indexConv with N -> 0, no id, hash10x.c:433-434),
   sequence after sequence until one comes back empty; the codes of all sequences back to back */
static void readGenome (FILE *f, uint8_t **codesOut, uint64_t **offOut, uint32_t *nSeqOut)
{ size_t cap = (size_t) 1 << 24, n = 0, offCap = 1024 ;
  uint8_t *codes = malloc (cap) ; uint64_t *off = malloc (8 * offCap) ;
  uint32_t nSeq = 0 ; int line = 1, c ;
  if (!codes || !off) die ("myalloc failure") ;
  off[0] = 0 ;
  for (;;)
    { size_t start = n ; int bad = 0 ;
      c = getc (f) ;
      if (c == '>') { while ((c = getc (f)) != EOF && c != '\n') ; ++line ; }
      else if (c != EOF) ungetc (c, f) ;
      while ((c = getc (f)) != EOF)
	{ int v ;
	  if (c == '>') { ungetc (c, f) ; break ; }
	  switch (c)
	    { case 'A': case 'a': case 'N': case 'n': v = 0 ; break ;
	      case 'C': case 'c': v = 1 ; break ;
	      case 'G': case 'g': v = 2 ; break ;
	      case 'T': case 't': v = 3 ; break ;
	      case '\n': ++line ; v = -1 ; break ;
	      case ' ': case '\t': v = -1 ; break ;
	      default: v = -2 ;
	    }
	  if (v == -2)
	    { fprintf (stderr, "Bad char 0x%x = '%c' at line %d, base %d\n", c, c, line, (int) (n - start)) ; bad = 1 ; break ; }
	  if (v < 0) continue ;
	  if (n == cap) { cap *= 2 ; if (!(codes = realloc (codes, cap))) die ("myalloc failure") ; }
	  codes[n++] = (uint8_t) v ;
	}
      if (bad) { n = start ; break ; }		/* readSequence returns 0: the walk over this genome ends */
      if (n == start) break ;
      if (nSeq + 2 > offCap) { offCap *= 2 ; if (!(off = realloc (off, 8 * offCap))) die ("myalloc failure") ; }
      off[++nSeq] = n ;
    }
  fclose (f) ;
  *codesOut = codes ; *offOut = off ; *nSeqOut = nSeq ;
}

static void printCribStats (const int *a, int max)		/* printArrayStats hash10x.c:457-468 */
{ int i, sum = 0, min = -1 ;
  double total = 0 ;
  for (i = 0 ; i < max ; ++i)
    if (a[i]) { sum += a[i] ; total += a[i] * i ; if (min == -1) min = i ; }
  fprintf (outFile, "  %d mean %.1f min %d max %d\n", sum, total / sum, min, max - 1) ;
}

static void cribBuild (FILE *f1, FILE *f2)
{ if (!(ctx && indexFromGpu)) die ("--cribBuild runs on the GPU-resident index (--readFQB on one GPU, or --readHash with a GPU present)") ;
  uint8_t *g[2] ; uint64_t *off[2] ; uint32_t nSeq[2] ; int i ;
  readGenome (f1, &g[0], &off[0], &nSeq[0]) ;
  readGenome (f2, &g[1], &off[1], &nSeq[1]) ;
  char err[512] ; h10x_crib cb ;
  int st = h10x_gpu_crib_build (ctx, g[0], off[0], nSeq[0], g[1], off[1], nSeq[1], &cb, err, sizeof (err)) ;
  if (st) die ("%s", *err ? err : h10x_strerror (st)) ;
  for (i = 0 ; i < 2 ; ++i)
    { fprintf (outFile, "  read %d known and %d unknown hashes from %d sequences in crib genome\n", cb.nPresent[i], cb.nAbsent[i], cb.nSeq[i]) ;
      if (outFile != stdout)
	printf ("  read %d known and %d unknown hashes from %d sequences in crib genome\n", cb.nPresent[i], cb.nAbsent[i], cb.nSeq[i]) ;
    }
  fprintf (outFile, "  crib matches\n") ;
  fprintf (outFile, "    hom  ") ; printCribStats (cb.hist[2], cb.histMax[2]) ;
  fprintf (outFile, "    het  ") ; printCribStats (cb.hist[1], cb.histMax[1]) ;
  fprintf (outFile, "    mul ") ; printCribStats (cb.hist[3], cb.histMax[3]) ;
  fprintf (outFile, "    err ") ; printCribStats (cb.hist[0], cb.histMax[0]) ;
  for (i = 0 ; i < 2 ; ++i) { free (g[i]) ; free (off[i]) ; }
}

static void gpuStats (void)
{ h10x_stats s ; int i ;
  if (!ctx || !indexFromGpu || h10x_gpu_stats (ctx, &s)) { fprintf (stderr, "  no GPU build to report\n") ; return ; }
  fprintf (outFile, "GPU_BUILD records %llu moshes %llu hashes %llu bins %llu blocks %llu\n",
	   (unsigned long long) s.nRecords, (unsigned long long) s.nMoshes, (unsigned long long) s.nHashes,
	   (unsigned long long) s.nBins, (unsigned long long) s.nBlocks) ;
  fprintf (outFile, "GPU_BUILD device_ms %.3f algorithmic_bytes %llu achieved_GBps %.1f launches %llu peak_device_bytes %llu\n",
	   s.msTotal, (unsigned long long) s.algorithmicBytes, s.algorithmicBytes / (s.msTotal * 1e6),
	   (unsigned long long) s.kernelLaunches, (unsigned long long) s.peakDeviceBytes) ;
  for (i = 0 ; i < H10X_NSTAGES ; ++i)
    if (s.msStage[i] > 0) fprintf (outFile, "GPU_STAGE %-10s %.3f ms\n", h10x_stage_name (i), s.msStage[i]) ;
}

int main (int argc, char *argv[])
{
  --argc ; ++argv ;
  outFile = stdout ;
  timeUpdate (stdout) ;
  params.k = 21 ; params.w = 31 ; params.r = 17 ; params.B = 28 ; params.N = 0 ;
  params.chunkSize = 100000 ; params.clusterThreshold = 5 ; params.gpus = 1 ;
  if (!argc) usage () ;

  while (argc)
    { if (**argv != '-') die ("option/command %s does not start with '-': run without arguments for usage", *argv) ;
      { int i ;
	fprintf (outFile, "COMMAND %s", *argv) ;
	for (i = 1 ; i < argc && *argv[i] != '-' ; ++i) fprintf (outFile, " %s", argv[i]) ;
	fputc ('\n', outFile) ;
	if (outFile != stdout)
	  { printf ("COMMAND %s", *argv) ;
	    for (i = 1 ; i < argc && *argv[i] != '-' ; ++i) fprintf (outFile, " %s", argv[i]) ;	/* sic: hash10x.c:1169 */
	    putchar ('\n') ;
	  }
      }
#define ARGMATCH(x,n)	(!strcmp (*argv, x) && argc >= n && (argc -= n, argv += n))
      if (ARGMATCH ("-k", 2)) params.k = atoi (argv[-1]) ;
      else if (ARGMATCH ("-w", 2)) params.w = atoi (argv[-1]) ;
      else if (ARGMATCH ("-r", 2)) params.r = atoi (argv[-1]) ;
      else if (ARGMATCH ("-B", 2)) params.B = atoi (argv[-1]) ;
      else if (ARGMATCH ("-N", 2)) params.N = atoi (argv[-1]) ;
      else if (ARGMATCH ("-c", 2)) params.chunkSize = atoi (argv[-1]) ;
      else if (ARGMATCH ("--gpus", 2)) params.gpus = atoi (argv[-1]) ;	/* new: GPUs used by --readFQB */
      else if (ARGMATCH ("--wideB", 1)) params.wideB = 1 ;			/* new: accept -B 31..34 */
      else if (ARGMATCH ("-t", 2) || ARGMATCH ("--threads", 2))
	fprintf (stderr, "  can't set thread number - not compiled with OMP\n") ;
      else if (ARGMATCH ("-o", 2) || ARGMATCH ("--output", 2))
	{ if (!strcmp (argv[-1], "-")) outFile = stdout ;
	  else if (!(outFile = fopen (argv[-1], "w")))
	    { fprintf (stderr, "can't open output file %s\n", argv[-1]) ; outFile = stdout ; }
	}
      else if (ARGMATCH ("--readFQB", 2))
	{ FILE *f = fopen (argv[-1], "r") ;
	  if (!f) die ("failed to open fqb file %s", argv[-1]) ;
	  fclose (f) ;
	  initialise (params.k, params.w, params.r, params.B) ;
	  readFQB (argv[-1]) ;
	}
      else if (ARGMATCH ("--readHash", 2))
	{ FILE *f = fopen (argv[-1], "r") ;
	  if (!f) die ("failed to open hash file %s", argv[-1]) ;
	  fclose (f) ;
	  initialise (params.k, params.w, params.r, params.B) ;
	  readHash (argv[-1]) ;
	}
      else if (ARGMATCH ("--writeHash", 2)) writeHash (argv[-1]) ;
      else if (ARGMATCH ("--hashDepthRange", 3)) hashDepthRange (atoi (argv[-2]), atoi (argv[-1])) ;
      else if (ARGMATCH ("-ct", 2) || ARGMATCH ("--clusterThreshold", 2)) params.clusterThreshold = atoi (argv[-1]) ;
      else if (ARGMATCH ("--cluster", 3)) clusterCodes (atoi (argv[-2]), atoi (argv[-1])) ;
      else if (ARGMATCH ("--clusterSplit", 1)) clusterSplit () ;
      else if (ARGMATCH ("--cribBuild", 3))
	{ FILE *f1, *f2 ;
	  if (!(f1 = fopen (argv[-2], "r"))) die ("failed to open .fa file %s", argv[-2]) ;
	  if (!(f2 = fopen (argv[-1], "r"))) die ("failed to open .fa file %s", argv[-1]) ;
	  cribBuild (f1, f2) ;
	}
      else if (ARGMATCH ("--hashStats", 1)) hashStats () ;
      else if (ARGMATCH ("--codeStats", 1)) codeStats () ;
      else if (ARGMATCH ("--gpuStats", 1)) gpuStats () ;
      else if (ARGMATCH ("--help", 1)) usage () ;
      else if (ARGMATCH ("--quit", 1) || ARGMATCH ("--exit", 1)) break ;
      else if (!strcmp (*argv, "--clusterReport")
	       || !strcmp (*argv, "--cribSummary") || !strcmp (*argv, "--hashInfo")
	       || !strcmp (*argv, "--hashExplore") || !strcmp (*argv, "--doubleShared") || !strcmp (*argv, "--codeExplore")
	       || !strcmp (*argv, "--errorFix") || !strcmp (*argv, "--shareScan") || !strcmp (*argv, "--interactive"))
	die ("command %s is outside the scope of hash10x-b200: write the index with --writeHash and run it in hash10x --readHash", *argv) ;
      else die ("unknown option/command %s; run without arguments for usage", *argv) ;

      printf ("  ") ; timeUpdate (stdout) ; fflush (stdout) ;
    }

  fprintf (outFile, "total resources used: ") ; timeTotal (outFile) ;
  if (outFile != stdout) { printf ("total resources used: ") ; timeTotal (stdout) ; }
  if (ctx) h10x_gpu_destroy (ctx) ;
  if (multi) h10x_multi_destroy (multi) ;
  return 0 ;
}
