// Host-side check of the division-free arithmetic the kernels use (h10x_common.cuh):
//   h10x_divisible(x)  ==  (x % w == 0)          for odd and even w
//   exact quotient:     (x * wInv) == x / w       for multiples of odd w   (sort key of the hash sort)
//   h10x_canonical      ==  min of the two multiplicative hashes (seqhash.c:58-69)
// Compiled with nvcc, runs on the CPU only (no kernel is launched).
#include "../../hash10x_b200/csrc/h10x_common.cuh"
#include <cstdio>
#include <cstdlib>

static uint64_t inv64 (uint64_t a) { uint64_t x = a ; for (int i = 0 ; i < 6 ; ++i) x *= 2 - a * x ; return x ; }

static void setup (HashParams &hp, int k, int w, uint64_t f)
{ hp.k = k ; hp.w = w ; hp.factor1 = f ; hp.shift = 64 - 2 * k ; hp.kmask = (((uint64_t) 1) << (2 * k)) - 1 ; hp.rcShift = 2 * (k - 1) ;
  uint64_t wo = (uint64_t) w ; int tz = 0 ; while (!(wo & 1)) { wo >>= 1 ; ++tz ; }
  hp.wTz = tz ; hp.wTzMask = (((uint64_t) 1) << tz) - 1 ; hp.wInv = inv64 (wo) ; hp.wLim = ~(uint64_t) 0 / wo ;
}

static uint64_t rnd (uint64_t &s) { s ^= s << 13 ; s ^= s >> 7 ; s ^= s << 17 ; return s ; }

int main ()
{ uint64_t seed = 0x9E3779B97F4A7C15ull ; long checks = 0 ;
  for (int w = 1 ; w <= 300 ; ++w)
    for (int k : { 5, 13, 16, 21, 23, 31 })
      { HashParams hp ; setup (hp, k, w, 0x49308bb9003cb3adull) ;
	for (int t = 0 ; t < 4000 ; ++t)
	  { uint64_t x = rnd (seed) & hp.kmask ;
	    if (t % 3 == 0) x = (x / w) * w ;			// force multiples
	    if (t % 97 == 0) x = 0 ;
	    if (h10x_divisible (x, hp) != (x % (uint64_t) w == 0)) { printf ("divisible mismatch w=%d x=%llu\n", w, (unsigned long long) x) ; return 1 ; }
	    // the fused kernel tests the unshifted product m = hash << shift
	    uint64_t m = x << hp.shift ;
	    bool sel = ((m & (hp.wTzMask << hp.shift)) == 0) && (m * hp.wInv <= hp.wLim) ;
	    if (sel != (x % (uint64_t) w == 0)) { printf ("shifted divisible mismatch w=%d k=%d x=%llu\n", w, k, (unsigned long long) x) ; return 1 ; }
	    if ((w & 1) && x % (uint64_t) w == 0 && x * hp.wInv != x / (uint64_t) w) { printf ("quotient mismatch w=%d x=%llu\n", w, (unsigned long long) x) ; return 1 ; }
	    uint64_t h = rnd (seed) & hp.kmask, hrc = rnd (seed) & hp.kmask ;
	    uint64_t a = (h * hp.factor1) >> hp.shift, b = (hrc * hp.factor1) >> hp.shift ;
	    if (h10x_canonical (h, hrc, hp) != (a < b ? a : b)) { printf ("canonical mismatch\n") ; return 1 ; }
	    // comparing whole products orders them by their top 2k bits (or the hashes are equal)
	    uint64_t pf = h * hp.factor1, pq = hrc * hp.factor1 ;
	    uint64_t top = ~((((uint64_t) 1) << hp.shift) - 1) ;
	    if ((((pf < pq) ? pf : pq) & top) >> hp.shift != (a < b ? a : b)) { printf ("min-then-mask mismatch\n") ; return 1 ; }
	    ++checks ;
	  }
      }
  // k-mer 0 of the known-answer sequence
  HashParams hp ; setup (hp, 21, 31, 0x49308bb9003cb3adull) ;
  if (((0x154514a11fdull * hp.factor1) >> hp.shift) != 0x32db51c0bc3ull || ((0x202ed7aebaaull * hp.factor1) >> hp.shift) != 0x1ff0ab2c2aaull)
    { printf ("KAT mismatch\n") ; return 1 ; }
  printf ("ok %ld\n", checks) ;
  return 0 ;
}
