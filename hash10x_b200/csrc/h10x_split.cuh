/* h10x_split.cuh - `--clusterSplit` on the resident index: clusterSplitCodes, hash10x.c:956-1013.
 *
 * Every sub-cluster j of a barcode block i becomes a block of its own behind the original ones (at
 * nCodes + sum of nSubCluster of the blocks before i + j - 1, the reference's new2[] indexing :963-964) holding the
 * block's entries with that label in list order, labels wiped (:981), reads renumbered in order of first appearance
 * (:983-985: ONE readMap per original block - the number a read gets comes from the counter of the cluster of its FIRST
 * clustered entry); clusterParent = i + 1 (:976).  The parent keeps the unclustered entries.  fillHashTable (:1012) is
 * then run again over the new blocks.
 *
 * One WARP per original block walks its ClusterHash list 32 entries at a time - the list order is the only order that
 * matters and a warp step keeps it: equal labels of a step are ranked in lane order with __match_any_sync, running
 * counts per label live in shared memory.  Two kernels around one scan: k_split_count (entries and first-seen reads
 * per label -> the new blocks' nHash / nRead), k_split_place (the same walk, now writing).  Per-warp tables indexed
 * by the 16-bit read index live in global memory (L2): first[] = the read's first clustered entry, number[] = its new
 * read index + 1.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define H10X_SPLIT_WARPS 8

struct SplitArgs {
  const unsigned long long *clus ;	/* old ClusterHash words: bin | read << 32 | subCluster << 48 | flags << 56 */
  const uint64_t *blkOff ; const uint32_t *blkNHash, *nSub ;
  const uint32_t *clusterBase ;		/* new block number of block i's cluster 1 */
  uint32_t nBlocksOld ;
  uint32_t *first, *number ;		/* per resident warp: tblSize entries each */
  uint32_t tblSize ;			/* reads a block can name: min (65536, largest nRead), the read index is 16 bits wide */
  uint32_t *newNHash, *newNRead ;	/* count kernel: per NEW block (parents keep their old nRead, set by the host) */
  const uint64_t *newOff ;		/* place kernel: first entry of every new block */
  unsigned long long *newClus ;
  uint64_t *idBlock ;			/* place kernel: bin << 32 | new block of every entry, at its new position */
  unsigned int *ticket ;
} ;

/* first[read] = index of the read's first clustered entry in block `blk` (0xffffffff: none) */
__device__ __forceinline__ void split_first_entries (const SplitArgs &a, uint32_t *first, uint64_t o, uint32_t n, uint32_t lane)
{ for (uint32_t x = lane ; x < a.tblSize ; x += 32) first[x] = 0xffffffffu ;
  __syncwarp () ;
  for (uint32_t e0 = 0 ; e0 < n ; e0 += 32)
    { const uint32_t e = e0 + lane ;
      if (e < n)
	{ const unsigned long long w = a.clus[o + e] ;
	  if ((w >> 48) & 0xffu) atomicMin (&first[(uint32_t) (w >> 32) & 0xffffu], e) ;
	}
    }
  __syncwarp () ;
  __threadfence_block () ;
}

template <bool PLACE>
__global__ void __launch_bounds__ (H10X_SPLIT_WARPS * 32)
k_split (SplitArgs a)
{ __shared__ uint32_t cntE[H10X_SPLIT_WARPS][257], cntR[H10X_SPLIT_WARPS][257] ;	/* per label: entries, first-seen reads so far */
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5 ;
  const uint32_t ltMask = (1u << lane) - 1u ;
  const uint32_t gw = blockIdx.x * H10X_SPLIT_WARPS + w ;
  uint32_t *first = a.first + (size_t) gw * a.tblSize, *number = a.number + (size_t) gw * a.tblSize ;
  for (;;)
    { uint32_t blk = 0 ;
      if (lane == 0) blk = atomicAdd (a.ticket, 1u) ;
      blk = __shfl_sync (0xffffffffu, blk, 0) ;
      if (blk >= a.nBlocksOld) break ;
      const uint32_t ns = a.nSub[blk], n = blk ? a.blkNHash[blk] : 0u ;
      const uint64_t o = a.blkOff[blk] ;
      if (!ns)
	{ /* copied as it is (:998) */
	  if (PLACE)
	    { const uint64_t d = a.newOff[blk] ;
	      for (uint32_t e = lane ; e < n ; e += 32)
		{ const unsigned long long x = a.clus[o + e] ;
		  a.newClus[d + e] = x ; a.idBlock[d + e] = ((x & 0xffffffffull) << 32) | blk ;
		}
	    }
	  else if (lane == 0) a.newNHash[blk] = n ;
	  continue ;
	}
      split_first_entries (a, first, o, n, lane) ;
      for (uint32_t x = lane ; x <= ns ; x += 32) { cntE[w][x] = 0 ; cntR[w][x] = 0 ; }
      __syncwarp () ;
      const uint32_t cb = a.clusterBase[blk] ;
      for (uint32_t e0 = 0 ; e0 < n ; e0 += 32)
	{ const uint32_t e = e0 + lane ;
	  const bool in = e < n ;
	  const unsigned long long x = in ? a.clus[o + e] : 0ull ;
	  const uint32_t lab = in ? (uint32_t) (x >> 48) & 0xffu : 0xffffu ;	/* padding lanes match each other only */
	  const uint32_t rd = (uint32_t) (x >> 32) & 0xffffu ;
	  const bool isFirst = in && lab && first[rd] == e ;
	  /* rank among the entries of the step with the same label, and among those of them that open a read */
	  const uint32_t m = __match_any_sync (0xffffffffu, lab) ;
	  const uint32_t mf = m & __ballot_sync (0xffffffffu, isFirst) ;
	  const uint32_t rankE = in ? cntE[w][lab] + __popc (m & ltMask) : 0u ;
	  const uint32_t rankR = (in && lab) ? cntR[w][lab] + __popc (mf & ltMask) : 0u ;
	  __syncwarp () ;
	  if (in && (m & ltMask) == 0) { cntE[w][lab] += __popc (m) ; if (lab) cntR[w][lab] += __popc (mf) ; }	/* the group's lowest lane */
	  if (PLACE)
	    { if (isFirst) number[rd] = rankR + 1u ;	/* readMap[read] = ++new2[clus].nRead (:983) */
	      __syncwarp () ;
	      __threadfence_block () ;
	      if (in)
		{ const uint32_t nb = lab ? cb + lab - 1u : blk ;
		  const uint64_t d = a.newOff[nb] + rankE ;
		  unsigned long long y = x & 0xff00ffffffffffffull ;				/* c->subCluster = 0 (:981) */
		  if (lab) y = (y & ~(0xffffull << 32)) | ((unsigned long long) ((number[rd] - 1u) & 0xffffu) << 32) ;	/* :985 */
		  a.newClus[d] = y ; a.idBlock[d] = ((x & 0xffffffffull) << 32) | nb ;
		}
	    }
	  __syncwarp () ;
	}
      if (!PLACE)
	{ for (uint32_t x = lane ; x <= ns ; x += 32)
	    { const uint32_t nb = x ? cb + x - 1u : blk ;
	      a.newNHash[nb] = cntE[w][x] ;
	      if (x) a.newNRead[nb] = cntR[w][x] ;
	    }
	}
      __syncwarp () ;
    }
}

/* hashCodes again (fillHashTable :317-347 over the new blocks): the entries arrive sorted by (bin, new block) */
__global__ void k_split_codes (uint64_t n, const uint64_t *__restrict__ sorted, uint32_t *__restrict__ codes)
{ uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (i < n) codes[i] = (uint32_t) sorted[i] ;
}
