/* synth_gpu.cu - on-device synthetic FQB generator (bench / test infrastructure, not product).
 * Emits byte-for-byte what oracle/synth_cpu.c emits on the host (same synth_fqb.h record function);
 * tests/test_synth.py checks that.  Lets bench.py build the 24 GB "1 Gb genome" input in HBM. */
#include "synth_fqb.h"
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void k_synth (synth_params p, const uint64_t *__restrict__ recOff, uint64_t r0, uint64_t n,
			 uint32_t *__restrict__ out)
{ uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x ;
  if (t >= n) return ;
  uint64_t r = r0 + t ;
  uint32_t lo = 0, hi = p.nBarcodes ;
  while (hi - lo > 1) { uint32_t mid = lo + (hi - lo) / 2 ; if (recOff[mid] <= r) lo = mid ; else hi = mid ; }
  uint32_t rec[30] ;
  sy_record (&p, lo, (uint32_t) (r - recOff[lo]), r, rec) ;
  uint32_t *dst = out + 30 * t ;
#pragma unroll
  for (int i = 0 ; i < 30 ; ++i) dst[i] = rec[i] ;
}

extern "C" {

/* total records; recOff (host, nBarcodes+1) may be NULL */
uint64_t synth_layout_host (const synth_params *p, uint64_t *recOff)
{ uint64_t n = 0 ;
  for (uint32_t b = 0 ; b < p->nBarcodes ; ++b) { if (recOff) recOff[b] = n ; n += sy_pairs (p, b) ; }
  if (recOff) recOff[p->nBarcodes] = n ;
  return n ;
}

/* records r0..r1-1 into device memory d_out (30*(r1-r0) words); returns a cudaError_t value */
int synth_fqb_device (const synth_params *p, const uint64_t *recOffHost, uint64_t r0, uint64_t r1,
		      void *d_out, void *stream)
{ cudaStream_t s = (cudaStream_t) stream ;
  uint64_t *dOff = nullptr ;
  size_t bytes = 8 * ((size_t) p->nBarcodes + 1) ;
  cudaError_t e = cudaMalloc ((void**) &dOff, bytes) ;
  if (e != cudaSuccess) return (int) e ;
  e = cudaMemcpyAsync (dOff, recOffHost, bytes, cudaMemcpyHostToDevice, s) ;
  uint64_t n = r1 - r0 ;
  if (e == cudaSuccess && n)
    { k_synth<<<(unsigned) ((n + 127) / 128), 128, 0, s>>> (*p, dOff, r0, n, (uint32_t*) d_out) ;
      e = cudaGetLastError () ;
    }
  cudaError_t e2 = cudaStreamSynchronize (s) ;
  cudaFree (dOff) ;
  return (int) (e != cudaSuccess ? e : e2) ;
}

}
