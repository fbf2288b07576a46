// micro-benchmark: __match_any_sync against a ballot-per-bit peer mask, by number of distinct values in the warp
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_match mb_match.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int BITS>
__global__ void k (uint32_t *out, int iters, uint32_t distinct)
{ const uint32_t lane = threadIdx.x & 31 ;
  uint32_t acc = 0, v = (lane % distinct) * 2654435761u ;
  for (int i = 0 ; i < iters ; ++i)
    { uint32_t d = (v >> 7) & ((1u << BITS) - 1u) ;
      uint32_t m ;
      if (MODE == 0) m = __match_any_sync (0xffffffffu, d) ;
      else
	{ m = 0xffffffffu ;
#pragma unroll
	  for (int b = 0 ; b < BITS ; ++b)
	    { const uint32_t bal = __ballot_sync (0xffffffffu, (d >> b) & 1u) ;
	      m &= ((d >> b) & 1u) ? bal : ~bal ;
	    }
	}
      acc += __popc (m) + (__ffs (m) - 1) ;
      v = v * 1664525u + 1013904223u * (lane % distinct + 1) ;	/* stays a function of lane % distinct */
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ;
}

template <int MODE, int BITS> float run (uint32_t *d, int iters, uint32_t distinct)
{ cudaEvent_t a, b ; cudaEventCreate (&a) ; cudaEventCreate (&b) ;
  k<MODE, BITS><<<148 * 8, 256>>> (d, iters, distinct) ;
  cudaEventRecord (a) ;
  k<MODE, BITS><<<148 * 8, 256>>> (d, iters, distinct) ;
  cudaEventRecord (b) ; cudaEventSynchronize (b) ;
  float ms ; cudaEventElapsedTime (&ms, a, b) ; return ms ;
}

int main ()
{ uint32_t *d ; cudaMalloc (&d, 148 * 8 * 256 * 4) ;
  const int iters = 4096 ;
  const double warps = 148.0 * 8 * 8 ;
  for (uint32_t distinct : { 1u, 2u, 4u, 8u, 16u, 32u })
    { float m0 = run<0, 10> (d, iters, distinct), m1 = run<1, 10> (d, iters, distinct), m2 = run<1, 8> (d, iters, distinct) ;
      printf ("distinct %2u: match_any %.3f ms (%.1f ns per warp-op per SM-slot) | ballot x10 %.3f ms | ballot x8 %.3f ms  [warp-ops %.0f]\n",
	      distinct, m0, m0 * 1e6 / (iters * 8.0 * 8), m1, m2, warps * iters) ;
    }
  return 0 ;
}
