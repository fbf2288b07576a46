/* digest_cpu.c - host build of the index digest (test infrastructure): the checker's side of
 * h10x_gpu_index_digest.  Same header as the device code, hash10x_b200/csrc/h10x_digest.h. */
#include "../hash10x_b200/csrc/h10x_digest.h"
#include <stddef.h>

uint64_t orc_digest_u32 (const uint32_t *a, uint64_t n, uint64_t base)
{ uint64_t d = 0 ; for (uint64_t i = 0 ; i < n ; ++i) d += h10x_dg_term (base + i, a[i]) ; return d ; }

uint64_t orc_digest_u64 (const uint64_t *a, uint64_t n, uint64_t base, uint64_t mask)
{ uint64_t d = 0 ; for (uint64_t i = 0 ; i < n ; ++i) d += h10x_dg_term (base + i, a[i] & mask) ; return d ; }
