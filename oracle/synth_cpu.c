/* synth_cpu.c - host build of the synthetic FQB generator (test/bench infrastructure).
 * The record function lives in hash10x_b200/csrc/synth_fqb.h so the CUDA generator used by
 * bench.py emits the same bytes; tests/test_synth.py checks the two against each other. */
#include "../hash10x_b200/csrc/synth_fqb.h"

/* number of records for the whole data set; recOff (nBarcodes+1) may be NULL */
uint64_t synth_layout (const synth_params *p, uint64_t *recOff)
{ uint64_t n = 0 ; uint32_t b ;
  for (b = 0 ; b < p->nBarcodes ; ++b) { if (recOff) recOff[b] = n ; n += sy_pairs (p, b) ; }
  if (recOff) recOff[p->nBarcodes] = n ;
  return n ;
}

/* fill out[30*(r1-r0)] with records r0..r1-1 of the data set (barcode-grouped order) */
void synth_fill (const synth_params *p, const uint64_t *recOff, uint64_t r0, uint64_t r1, uint32_t *out)
{ uint32_t b = 0 ; uint64_t r ;
  uint32_t lo = 0, hi = p->nBarcodes ;	/* find barcode of r0 */
  while (hi - lo > 1) { uint32_t mid = lo + (hi - lo)/2 ; if (recOff[mid] <= r0) lo = mid ; else hi = mid ; }
  b = lo ;
  for (r = r0 ; r < r1 ; ++r)
    { while (recOff[b+1] <= r) ++b ;
      sy_record (p, b, (uint32_t)(r - recOff[b]), r, out + 30*(r - r0)) ;
    }
}
