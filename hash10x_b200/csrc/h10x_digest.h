/* h10x_digest.h - position-salted 64-bit sum digest of an index array.
 *
 *     digest (A, base) = sum over i of mix (mix (base + i + SALT) ^ A[i])   (mod 2^64)
 *
 * It is a sum, so pieces of an array held by different ranks (or streamed from a file in chunks) are
 * digested independently with their global position as `base` and simply added; and it is salted by the
 * position, so a permutation of the values changes it.  The same header compiles as plain C for the host
 * tools (oracle/scale_tool.c digests the reference's own `.hash` files) and as __device__ code for
 * h10x_gpu_index_digest, which is how bench.py checks the index it has just timed - at sizes where
 * copying 18 GB to the host and comparing arrays would take longer than the whole benchmark.
 */
#ifndef H10X_DIGEST_H
#define H10X_DIGEST_H

#include <stdint.h>

#ifdef __CUDACC__
#define H10X_DG_HD __host__ __device__ __forceinline__
#else
#define H10X_DG_HD static inline
#endif

H10X_DG_HD uint64_t h10x_dg_mix (uint64_t x)	/* splitmix64 finaliser */
{ x += 0x9E3779B97F4A7C15ull ;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull ;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull ;
  return x ^ (x >> 31) ;
}

H10X_DG_HD uint64_t h10x_dg_term (uint64_t pos, uint64_t val)
{ return h10x_dg_mix (h10x_dg_mix (pos + 0x243F6A8885A308D3ull) ^ val) ; }

/* a ClusterHash entry is digested without its subCluster / flags bytes: --readFQB leaves them
   uninitialised in the reference (hash10x.c:175,179-180) */
#define H10X_DG_CLUS_MASK 0x0000ffffffffffffull

#endif
