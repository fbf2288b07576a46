/* h10x_gpu.h - C ABI of libh10xgpu.so: hash10x's `--readFQB` minhash index build on one B200.
 *
 * The reference (richarddurbin/hash10x, plain C, /root/reference) has no plugin or FFI seam;
 * the seam this library replaces is the `--readFQB` branch of its command loop
 * (hash10x.c:1200-1205: initialise(); readFQB(); fillHashTable();) and the global state that
 * branch leaves behind for every later command (hash10x.c:85-96).  Each entry
 * point below names the reference code it stands in for.  Plain C types only; no CUDA, C++ or
 * torch types cross this boundary (streams and device pointers travel as void* / integers).
 *
 * There is no CPU fallback: every build entry point returns H10X_ERR_NO_DEVICE when no CUDA
 * device is usable.  INTEGRATION.md shows the change a hash10x maintainer would make to call it.
 */
#ifndef H10X_GPU_H
#define H10X_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H10X_ABI_VERSION 3

/* return codes; the host turns 1 and 2 into the reference's die() texts
   ("hashTableSize is too small" hash10x.c:149, "chunkSize too small" hash10x.c:206) */
enum {
  H10X_OK = 0,
  H10X_ERR_TABLE_TOO_SMALL = 1,
  H10X_ERR_CHUNK_TOO_SMALL = 2,
  H10X_ERR_BAD_PARAM = 3,
  H10X_ERR_NOMEM = 4,
  H10X_ERR_IO = 5,
  H10X_ERR_CUDA = 6,
  H10X_ERR_NO_DEVICE = 7,
  H10X_ERR_UNSUPPORTED = 8
};

/* the `params` globals of hash10x.c:25-33 that the --readFQB path reads, plus the hasher
   constant.  factor1 is computed by the HOST exactly as seqhash.c:29 does after srandom(r)
   (hash10x.c:1101), so the libc RNG never enters device code; h10x_factor1_from_seed() does it. */
typedef struct h10x_params {
  int32_t k;			/* -k, 1..31 (seqhash.c:24) */
  int32_t w;			/* -w, >= 1 (seqhash.c:25) */
  uint64_t factor1;		/* Seqhash.factor1 (seqhash.h:20) */
  int32_t B;			/* -B, hashTableBits; 20..30 unless H10X_FLAG_WIDE_B (hash10x.c:1107) */
  int32_t chunkSize;		/* -c, only for the reference's chunk-boundary behaviour (hash10x.c:197-209) */
  int64_t N;			/* -N, 0 = all records (hash10x.c:202,207) */
  int32_t device;		/* CUDA device ordinal */
  uint32_t flags;		/* H10X_FLAG_* */
} h10x_params ;

#define H10X_FLAG_WIDE_B	1u	/* accept B up to 34 as README.md:55 / moshset.c:17 describe */
#define H10X_FLAG_NO_TABLE	2u	/* skip hashIndex[] materialisation (hashIndex stays NULL) */
#define H10X_FLAG_NO_CODES	4u	/* skip the hash->code CSR (fillHashTable) */
#define H10X_FLAG_GENERIC_ONLY	8u	/* force the generic (global-memory) sort path for every block */
#define H10X_FLAG_LEGACY_TAIL	16u	/* group / transpose with the library radix sort (round-1 tail) instead of h10x_tail.cuh */
#define H10X_FLAG_LAZY_CODES	32u	/* the hash->code CSR is built and stays resident (--cluster reads it there), but
					   h10x_gpu_build_host / h10x_gpu_download leave codes / codeOff NULL in the host index:
					   no host command of the --readFQB ... --writeHash chain reads them (the reference's
					   writeHashFile, hash10x.c:244-267, does not either); h10x_gpu_download_codes fetches them */

/* ClusterHash of hash10x.c:35-43, 8 bytes; subCluster and flags are written as 0 */
typedef struct h10x_cluster_hash {
  uint32_t hash ;		/* bin id, not the hash value */
  uint16_t read ;		/* read-pair index within the barcode block, truncated to 16 bits */
  uint8_t subCluster ;
  uint8_t flags ;
} h10x_cluster_hash ;

/* The state --readFQB leaves behind (hash10x.c:85-96), as flat arrays.  Block numbering is the
   reference's: entry 0 is the dummy, blocks 1..nBlocksMax-1 are the barcode runs in file order and
   the last run has nHash 0 (hash10x.c:209,216).  Block i's ClusterHash list is
   clusHash[blkOff[i] .. blkOff[i]+blkNHash[i]) sorted by bin id; bin x's barcode list
   (hashCodes[x], hash10x.c:317-338) is codes[codeOff[x] .. codeOff[x+1]) ascending. */
#define H10X_INDEX_EXACT_BLOCKS 1u
typedef struct h10x_index {
  int32_t B ;
  uint32_t hashNumber ;		/* bins are 1..hashNumber-1 */
  uint32_t nBlocksMax ;		/* arrayMax(clusterBlocks) */
  uint32_t reserved ;		/* H10X_INDEX_EXACT_BLOCKS: clusterBlocks holds exactly nBlocksMax elements (after --clusterSplit,
				   hash10x.c:961), which is the Array size --writeHash puts into the file */
  uint64_t nReads ;		/* records consumed (readFQB's nReads) */
  uint64_t nHashes ;		/* sum of nHash */
  uint32_t *hashIndex ;		/* 2^B, layout identical to the reference's sequential insertion */
  uint64_t *hashValue ;		/* hashNumber, [0] = 0 */
  uint32_t *hashDepth ;		/* hashNumber, [0] = 0 */
  uint32_t *blkNRead ;		/* nBlocksMax */
  uint32_t *blkNHash ;		/* nBlocksMax */
  uint64_t *blkOff ;		/* nBlocksMax + 1 */
  h10x_cluster_hash *clusHash ;	/* nHashes */
  uint64_t *codeOff ;		/* hashNumber + 1 */
  uint32_t *codes ;		/* nHashes */
  int32_t onDevice ;		/* 1: the pointers above are device pointers owned by the context */
  int32_t pinned ;		/* 0 malloc, 1 cudaHostAlloc (both freed by h10x_index_free), 2 = the
				   context's reusable pinned arena: valid until its next download */
  /* ClusterBlock.nSubCluster / .pointToMin (hash10x.c:62-70), nBlocksMax each; NULL = all zero (never clustered).
     Not owned by the index: h10x_read_hash allocates them with the block table and h10x_index_free releases
     them only then (pinned == 0); after h10x_gpu_cluster the host points them at h10x_clusters' arrays. */
  uint32_t *blkNSubCluster ;
  double *blkPointToMin ;
  uint32_t *blkClusterParent ;	/* ClusterBlock.clusterParent (hash10x.c:66), nBlocksMax; NULL = all zero (no --clusterSplit yet);
				   ownership as the two arrays above */
} h10x_index ;

/* per-build measurements for the roofline report */
#define H10X_NSTAGES 12
typedef struct h10x_stats {
  double msTotal ;		/* CUDA-event time of the whole device build */
  double msStage[H10X_NSTAGES] ;	/* per stage, names from h10x_stage_name() */
  uint64_t nRecords, nMoshes, nHashes, nBins, nBlocks ;
  uint64_t algorithmicBytes ;	/* 120 R + 12 H + 12 D + 4*2^B + 32 (nB+1)   */
  uint64_t kernelLaunches ;	/* kernels launched by the last build (ours + CUB's) */
  uint64_t fusedBlocks, genericBlocks ;	/* barcode blocks taken by each mosh path */
  uint64_t peakDeviceBytes ;
  uint64_t tailPath ;		/* 2 = hand-written tail (h10x_tail.cuh), 1 = library-sort tail */
} h10x_stats ;

typedef struct h10x_ctx h10x_ctx ;

int h10x_abi_version (void) ;
int h10x_gpu_device_count (void) ;
const char *h10x_strerror (int code) ;
const char *h10x_stage_name (int stage) ;

/* seqhash.c:29 under srandom(seed) (hash10x.c:1101) - glibc TYPE_3 random(); host only */
uint64_t h10x_factor1_from_seed (int seed) ;

/* initialise() hash10x.c:1099-1118: checks k, w, B; binds the device */
h10x_ctx *h10x_gpu_create (const h10x_params *p, char *err, size_t errlen) ;
void h10x_gpu_destroy (h10x_ctx *ctx) ;

/* readFQB() + fillHashTable() (hash10x.c:188-236, 317-347) on records already resident in device
   memory: d_fqb = nRecords * 30 little-endian U32 (16-byte aligned).  `stream` is a cudaStream_t
   (NULL = the context's own stream).  The index stays resident on the device, owned by ctx, until
   the next build or destroy.  Asynchronous only up to the host decisions it needs; returns after
   the last kernel is enqueued and counters are known. */
int h10x_gpu_build_device (h10x_ctx *ctx, const void *d_fqb, uint64_t nRecords, void *stream,
			   char *err, size_t errlen) ;

/* device-pointer view of the resident index (onDevice = 1); valid until the next build */
int h10x_gpu_index_device (h10x_ctx *ctx, h10x_index *out) ;

/* copy the resident index to pinned host memory owned by the context (pinned = 2): the arrays stay
   valid until the next download/build_host/build_file on this context or its destruction */
int h10x_gpu_download (h10x_ctx *ctx, h10x_index *out, char *err, size_t errlen) ;

/* ---- --hashStats / --codeStats / --cribBuild on the resident index ("next" row f4; hash10x.c:351-402, 426-510) ----
   h10x_gpu_histogram: the count-per-value array that hashDepthHist (which = 0: over hashDepth[0 .. hashNumber), dummy bin
   included), codeSizeHist (1: over the blocks' nHash) and its cluster part (2: over nSubCluster, after --cluster) hand to
   histogramReport; *n = largest value + 1.  The array is owned by the context, valid until the next call. */
int h10x_gpu_histogram (h10x_ctx *ctx, int which, const int **hist, int *n, char *err, size_t errlen) ;

/* cribBuild (hash10x.c:470-510): g1 / g2 = the two genomes as base codes 0..3, one byte per base, sequences back to back
   (off[s] .. off[s+1]), as readSequence + cribAddGenome's conversion deliver them (N -> 0).  Per bin: cribType
   (0 err, 1 htA, 2 htB, 3 hom, 4 mul), CribInfo.chr / .pos; per genome the counters of the report line; hist[g] =
   bins of group g (0 err, 1 het, 2 hom, 3 mul) by bin depth, histMax[g] = the reference's arrayMax of that array.
   Arrays are owned by the context, valid until the next call or build. */
typedef struct h10x_crib {
  uint32_t hashNumber ;
  int32_t histLen ;
  uint8_t *type ; int16_t *chr ; uint16_t *pos ;
  int32_t nPresent[2], nAbsent[2], nSeq[2] ;
  const int *hist[4] ; int32_t histMax[4] ;
} h10x_crib ;
int h10x_gpu_crib_build (h10x_ctx *ctx, const uint8_t *g1, const uint64_t *off1, uint32_t nSeq1,
			 const uint8_t *g2, const uint64_t *off2, uint32_t nSeq2, h10x_crib *out, char *err, size_t errlen) ;

/* ---- fq2b + bsort on the device (fq2b.c:108-178; README.md:25-26): the stage that produces the FQB file ----
   fq1 / fq2: the two UNCOMPRESSED FASTQ texts in host memory (fq2 may be NULL: single reads); the reference reads
   them through zlib, which stays a host matter.  whitelist: the 10x barcode list as packed 16-mers in file order
   (h10x_pack_barcode), NULL = no barcode correction (fq2b without -10x).  Records whose barcode has no whitelist entry
   within one mismatch are dropped, the others get the corrected barcode (fq2b.c:98-104, 159).  The records stay on the
   device (out->d_fqb, valid until the next h10x_gpu_fq2b or destroy) so that h10x_gpu_build_device can follow, and are
   copied to pinned host memory owned by ctx (out->fqb) unless H10X_FQ2B_NO_HOST.  A malformed entry gives H10X_ERR_IO
   with gzReadFastq's message (fq2b.c:180-208). */
typedef struct h10x_fq2b_out {
  void *fqb ;			/* nRecords * recWords little-endian U32, host */
  void *d_fqb ;			/* the same, device */
  uint64_t nRecords ;		/* records written */
  uint64_t nRead ;		/* entries read from fq1 */
  uint32_t recWords ;		/* (s1Len+15)/16 + (s1Len+31)/32 [+ the same for s2Len]: 30 for 145..160-bp pairs */
  uint32_t s1Len, s2Len, reserved ;
  uint64_t nBad, nFixed, nFixBase[16] ;	/* the counters fq2b prints (fq2b.c:168-177) */
} h10x_fq2b_out ;
#define H10X_FQ2B_SORT		1u	/* group the records by barcode as `bsort -k 4 -r <record bytes>` does: by their first 4 bytes, stable */
#define H10X_FQ2B_NO_HOST	2u	/* leave out->fqb NULL */
int h10x_gpu_fq2b (h10x_ctx *ctx, const char *fq1, uint64_t n1, const char *fq2, uint64_t n2,
		   const uint32_t *whitelist, uint64_t nWhitelist, uint32_t flags, h10x_fq2b_out *out, char *err, size_t errlen) ;
/* seqPack (fq2b.c:33-42) of one 16-base barcode line of the whitelist file; -1 if it is not 16 characters */
int h10x_pack_barcode (const char *s16, uint32_t *out) ;

/* fillHashTable()'s lists (hash10x.c:317-347) of the resident index -> out->codeOff / out->codes, in the context's
   pinned arena like the arrays of h10x_gpu_download; for contexts created with H10X_FLAG_LAZY_CODES */
int h10x_gpu_download_codes (h10x_ctx *ctx, h10x_index *out, char *err, size_t errlen) ;

/* the whole seam with HOST buffers: H2D of the FQB records, build, D2H of the index */
int h10x_gpu_build_host (h10x_ctx *ctx, const void *fqb, uint64_t nRecords, h10x_index *out,
			 char *err, size_t errlen) ;

/* same, reading `path` as readFQB's fread loop does (whole 120-byte records only) */
int h10x_gpu_build_file (h10x_ctx *ctx, const char *path, h10x_index *out, char *err, size_t errlen) ;

/* readHashFile()'s counterpart for the device (hash10x.c:269-315): a HOST index (h10x_read_hash + the hash->code lists
   the caller rebuilt as fillHashTable does, hash10x.c:317-347) becomes the context's resident index, so that
   h10x_gpu_depth_range / h10x_gpu_cluster serve a session that starts with --readHash (README.md:29,50-51).
   ctx must have been created with the file's B; nSubCluster / pointToMin of an earlier --cluster travel too. */
int h10x_gpu_load_index (h10x_ctx *ctx, const h10x_index *host, char *err, size_t errlen) ;

int h10x_gpu_stats (h10x_ctx *ctx, h10x_stats *out) ;

/* Verification of the resident index at sizes where copying it out would dominate: position-salted sum digests
   (hash10x_b200/csrc/h10x_digest.h: sum over i of mix(mix(base+i) ^ A[i]) mod 2^64) of every array that
   writeHashFile stores (hash10x.c:244-267), computed on the device.  A rank of a multi-GPU build passes the global
   number of its block 0 (blockBase) and of its first ClusterHash entry (entryBase) and withBlockZero = 0 on all
   ranks but the first; the per-rank digests then simply add up to the digest of the stitched arrays.  On a
   single-GPU index codesMissing / codesUnordered also count (block, bin) pairs of the ClusterHash lists absent from
   the bin's barcode list, and bins whose list is not strictly ascending or not hashDepth long: both 0 means
   codes[] is exactly what fillHashTable (hash10x.c:317-347) builds. */
typedef struct h10x_digest {
  uint64_t hashIndex, hashValue, hashDepth, blkNRead, blkNHash, clusHash, codes, codeOff ;
  uint64_t codesMissing, codesUnordered ;
  int32_t haveTable, haveBins, haveCodes, reserved ;
} h10x_digest ;
int h10x_gpu_index_digest (h10x_ctx *ctx, uint64_t blockBase, uint64_t entryBase, int withBlockZero, h10x_digest *out,
			   char *err, size_t errlen) ;

/* "next" row: --hashDepthRange on the index resident after a single-GPU build.
   hashWithinRangeBuild (hash10x.c:528-539): within[bin] is SET when min <= depth < max (flags accumulate over
   calls until the next build); goodHashesBuild (hash10x.c:738-766): for every block the indices into its
   ClusterHash list of the entries whose bin is within range, by increasing bin depth, ties in list order
   (glibc's stable qsort); blocks with more than 65535 hashes get an empty list (:748).
   The arrays are pinned host memory owned by the context, valid until the next call or build. */
typedef struct h10x_good_hashes {
  uint64_t nGood ;
  uint32_t hashNumber, nBlocksMax ;
  uint8_t *within ;		/* hashNumber */
  uint64_t *goodOff ;		/* nBlocksMax + 1: block c's list is good[goodOff[c] .. goodOff[c+1]) */
  uint16_t *good ;		/* nGood */
} h10x_good_hashes ;
int h10x_gpu_depth_range (h10x_ctx *ctx, int min, int max, h10x_good_hashes *out, char *err, size_t errlen) ;
/* the same without the host copies: the lists stay on the device, where h10x_gpu_cluster reads them; *nGood = their total length */
int h10x_gpu_depth_range_device (h10x_ctx *ctx, int min, int max, uint64_t *nGood, char *err, size_t errlen) ;
void h10x_index_free (h10x_index *ix) ;

/* "next" row: `--cluster codeMin codeMax` (hash10x.c:1241-1256) on the index and the goodHashes
   lists resident after h10x_gpu_depth_range: codeClusterFind (hash10x.c:770-835) assigns the good entries of
   every block codeMin <= code < codeMax (0 0 = all blocks, as in the reference) to sub-clusters of hashes that
   are first shared with the same other barcodes, clusterThreshold = -ct (hash10x.c:32,1137, default 5, >= 1);
   codeClusterReadMerge (hash10x.c:837-868) merges sub-clusters joined by a read pair and renumbers them.
   Results are the reference's, bit for bit: ClusterHash.subCluster of every entry (the resident ClusterHash array
   is updated in place and copied out), ClusterBlock.nSubCluster and .pointToMin per block.  State accumulates over
   calls like the reference's globals, until the next build.  Arrays are pinned host memory owned by the context,
   valid until the next call, download or build; clusHash is the same buffer h10x_gpu_download hands out. */
typedef struct h10x_clusters {
  uint32_t nBlocksMax, reserved ;
  uint64_t nHashes ;
  uint32_t *nSubCluster ;	/* nBlocksMax */
  double *pointToMin ;		/* nBlocksMax */
  h10x_cluster_hash *clusHash ;	/* nHashes, subCluster set */
  double msKernel ;		/* CUDA-event time of the clustering kernel */
} h10x_clusters ;
int h10x_gpu_cluster (h10x_ctx *ctx, int codeMin, int codeMax, int clusterThreshold, h10x_clusters *out,
		      char *err, size_t errlen) ;

/* `--clusterSplit` (clusterSplitCodes, hash10x.c:956-1013) on the resident single-GPU index after h10x_gpu_cluster: every
   sub-cluster j of block i becomes a barcode block of its own behind the original ones (the clusters of block i at
   nBlocksMax + sum of nSubCluster of the blocks before i, in label order) holding the entries with that label in list
   order, subCluster bytes wiped, reads renumbered in order of first appearance, clusterParent = i + 1; the parent keeps
   its unclustered entries and its nRead; blocks without sub-clusters are carried over whole.  The hash->code lists are
   rebuilt over the new blocks (fillHashTable, :1012).  The resident index is replaced and comes back in *out like
   h10x_gpu_download's (pinned arena of the context; blkNSubCluster is all zero afterwards, blkPointToMin survives only
   for the blocks that were not split, as in the reference).  *nNew = the number of blocks added.  The good-hash lists are
   dropped: the reference leaves its own indexed by the old block numbers; run h10x_gpu_depth_range again. */
int h10x_gpu_cluster_split (h10x_ctx *ctx, h10x_index *out, uint32_t *nNew, char *err, size_t errlen) ;

/* pinned host staging for callers that want the H2D copy to run at full PCIe rate */
void *h10x_host_alloc (size_t bytes) ;
void h10x_host_free (void *p) ;

/* ---- multi-GPU (new; the reference is single-process): one context per GPU, NCCL over NVLink ----
   Ranks own consecutive barcode-run ranges of the data set, cut at run boundaries.  After
   h10x_gpu_build_device_dist every rank holds its blocks' table and ClusterHash lists with GLOBAL bin
   ids and its part of every bin's barcode list; rank 0 also holds hashValue / hashDepth / hashIndex.
   -N must be 0 and no record may carry the all-A barcode word 0 (the chunk-boundary quirk of
   hash10x.c:212 depends on global record positions). */
#define H10X_DIST_ID_BYTES 128
typedef struct h10x_dist_info {
  int32_t rank, nranks ;
  uint32_t blockBase ;		/* global number of this rank's local block b is blockBase + b */
  uint32_t nBlocksGlobal ;	/* barcode runs over all ranks */
  uint64_t nReadsGlobal, nHashesGlobal ;
  uint32_t nLocalBins, reserved ;
  const uint32_t *localBinId ;	/* device: bin id of each rank-distinct hash (ascending hash order) */
  const uint32_t *localCodeOff ;	/* device: nLocalBins + 1 offsets into localCodes */
  const uint32_t *localCodes ;	/* device: this rank's (global) block numbers of each bin, ascending */
} h10x_dist_info ;

/* rank 0 creates the NCCL unique id; the caller ships the 128 bytes to the other ranks (torch.distributed,
   MPI, a file ...); every rank then joins with h10x_dist_init (collective) */
/* The hash ranges of the owners in a distributed build (host arithmetic, no device needed): thr[o] = first hash owned by
   rank o, thr[nranks] = 2^(2k).  flat = 0: cut at the quantiles of the density 2 (1 - x) of min (hash, hashRC), so that
   every owner numbers about the same count of bins; flat = 1: equal widths.  Any monotone cut gives the reference's ids. */
int h10x_dist_owner_thresholds (int k, int nranks, int flat, uint64_t *thr) ;
int h10x_dist_unique_id (void *id128, char *err, size_t errlen) ;
int h10x_dist_init (h10x_ctx *ctx, int rank, int nranks, const void *id128, char *err, size_t errlen) ;
/* collective over all ranks; d_fqb holds this rank's records */
int h10x_gpu_build_device_dist (h10x_ctx *ctx, const void *d_fqb, uint64_t nRecords, void *stream,
				char *err, size_t errlen) ;
/* the same with this rank's records in HOST memory: H2D, collective build, D2H of this rank's arrays */
int h10x_gpu_build_host_dist (h10x_ctx *ctx, const void *fqb, uint64_t nRecords, h10x_index *out,
			      char *err, size_t errlen) ;

/* Collective, after a distributed build: every rank receives hashDepth[] and the WHOLE hash -> code lists
   (fillHashTable, hash10x.c:317-347, over all ranks' barcode blocks: codes[] then holds global block numbers), so that
   h10x_gpu_depth_range and h10x_gpu_cluster (hash10x.c:528-539, 738-868) run on each rank for its own blocks - codeMin /
   codeMax and the arrays they return are in the rank's local block numbering, block b of rank r being global block
   blockBase + b (h10x_gpu_dist_info).  Needs fewer than 2^32 (block, hash) pairs over all ranks. */
int h10x_gpu_dist_global_codes (h10x_ctx *ctx, char *err, size_t errlen) ;
int h10x_gpu_dist_info (h10x_ctx *ctx, h10x_dist_info *out) ;
/* the whole seam on nGpus GPUs of this node from ONE process (a thread and a context per GPU, devices
   p->device .. p->device+nGpus-1): the file is cut at barcode-run boundaries, every GPU builds its range and
   the pieces are stitched into one host index (malloc'ed; free with h10x_index_free).  codes / codeOff stay
   NULL: the hash->code lists remain distributed and no command of the --readFQB ... --writeHash chain reads them. */
int h10x_gpu_build_file_multi (const h10x_params *p, int nGpus, const char *path, h10x_index *out,
			       char *err, size_t errlen) ;

/* The same build as a SESSION: the per-GPU contexts stay alive (with hashDepth and the whole hash->code CSR on every
   GPU, h10x_gpu_dist_global_codes), so that the commands hash10x chains after --readFQB run on all GPUs too, each on
   its own barcode blocks: --hashDepthRange (hash10x.c:528-539, 738-766) and --cluster (hash10x.c:770-868).  Block
   numbers, good lists and ClusterHash entries come back stitched in the global numbering of the one-GPU calls; the
   arrays belong to the session and are valid until its next call of the same kind. */
typedef struct h10x_multi h10x_multi ;
int h10x_multi_build_file (const h10x_params *p, int nGpus, const char *path, h10x_multi **session, h10x_index *out,
			   char *err, size_t errlen) ;
int h10x_multi_depth_range (h10x_multi *session, int dmin, int dmax, h10x_good_hashes *out, char *err, size_t errlen) ;
int h10x_multi_cluster (h10x_multi *session, int codeMin, int codeMax, int clusterThreshold, h10x_clusters *out,
			char *err, size_t errlen) ;
void h10x_multi_destroy (h10x_multi *session) ;
int h10x_gpu_memcpy_d2h (h10x_ctx *ctx, void *dst, const void *src, size_t bytes) ;

/* moshes of each record without the index build (the K1 stage alone), for parity tests of
   seqhash.c:154-195: outCount[i] = number of moshes of record i in generation order, written to
   outHash[outOff[i]..]; arrays are HOST memory, outOff has nRecords+1 entries. */
/* tests: the keys the FUSED kernel stored per processed block of a host FQB, before any grouping (see the definition):
   outOff[nBlocks+1], outHash / outRead [cap].  *lean = 1 when the kernel ran lean (every mosh, unsorted, duplicates in). */
int h10x_gpu_block_keys (h10x_ctx *ctx, const void *fqb, uint64_t nRecords, uint64_t *outOff, uint64_t *outHash,
			 uint32_t *outRead, uint64_t cap, uint32_t *nBlocks, int *lean, char *err, size_t errlen) ;
int h10x_gpu_record_moshes (h10x_ctx *ctx, const void *fqb, uint64_t nRecords, uint64_t *outOff,
			    uint64_t *outHash, uint64_t cap, char *err, size_t errlen) ;

/* writeHashFile()/readHashFile() hash10x.c:244-315 on a host index (host code, no CUDA) */
int h10x_write_hash (const h10x_index *ix, const char *path) ;
int h10x_read_hash (const char *path, int32_t B, h10x_index *out, char *err, size_t errlen) ;

#ifdef __cplusplus
}
#endif
#endif
