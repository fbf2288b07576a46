/* fq2b-b200 - the reference's fq2b command line (fq2b.c:108-178) over h10x_gpu_fq2b: same options, same .fqb bytes,
 * same report on stderr; `-sort` adds the `bsort -k 4 -r <record bytes>` step of the README pipeline (README.md:25-26),
 * so that the output can go straight into hash10x --readFQB.  Host code only: files (plain or gzip, through zlib as in
 * the reference) are read into memory, everything else happens on the GPU.
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "h10x_gpu.h"

static void die (const char *format, ...)
{ va_list args ;
  va_start (args, format) ;
  fprintf (stderr, "FATAL ERROR: ") ; vfprintf (stderr, format, args) ; fprintf (stderr, "\n") ;
  va_end (args) ;
  exit (-1) ;
}

static char *slurp (const char *path, uint64_t *n)
{ gzFile f = gzopen (path, "r") ;
  if (!f) die ("failed to open %s", path) ;
  size_t cap = (size_t) 64 << 20, len = 0 ;
  char *buf = malloc (cap) ;
  if (!buf) die ("myalloc failure") ;
  for (;;)
    { if (cap - len < ((size_t) 16 << 20)) { cap *= 2 ; if (!(buf = realloc (buf, cap))) die ("myalloc failure") ; }
      int got = gzread (f, buf + len, 16 << 20) ;
      if (got < 0) die ("read error in %s", path) ;
      if (!got) break ;
      len += (size_t) got ;
    }
  gzclose (f) ;
  *n = len ;
  return buf ;
}

/* the reference compares the id lines of the two files entry by entry and reports the first pair that differs
   (fq2b.c:152-157); ids are host business: four lines per entry, found with memchr */
static void checkIds (const char *t1, uint64_t n1, const char *t2, uint64_t n2)
{ const char *p1 = t1, *e1 = t1 + n1, *p2 = t2, *e2 = t2 + n2 ;
  while (p1 < e1 && p2 < e2)
    { const char *a = memchr (p1, '\n', (size_t) (e1 - p1)), *b = memchr (p2, '\n', (size_t) (e2 - p2)) ;
      if (!a || !b) return ;
      if (a - p1 != b - p2 || memcmp (p1, p2, (size_t) (a - p1)))
	{ fprintf (stderr, "proceeding despite paired read ids not matching, e.g. %.*s %.*s\n", (int) (a - p1), p1, (int) (b - p2), p2) ;
	  return ;
	}
      for (int k = 0 ; k < 3 ; ++k)
	{ a = memchr (a + 1, '\n', (size_t) (e1 - a - 1)) ; b = memchr (b + 1, '\n', (size_t) (e2 - b - 1)) ;
	  if (!a || !b) return ;
	}
      p1 = a + 1 ; p2 = b + 1 ;
    }
}

int main (int argc, char *argv[])
{ FILE *fout = stdout ;
  uint32_t *wl = 0 ; uint64_t nWl = 0 ; const char *wlName = 0 ;
  uint32_t flags = 0 ;
  --argc ; ++argv ;
  while (argc > 2 && *argv[0] == '-')
    if (!strcmp (*argv, "-10x"))
      { FILE *f = fopen (argv[1], "r") ; char s[64] ; size_t cap = 1 << 20 ;
	if (!f) die ("failed to open 10x whitelist file %s\n", argv[1]) ;
	wlName = argv[1] ;
	if (!(wl = malloc (4 * cap))) die ("can't allocate barcode table") ;
	while (fscanf (f, "%63s\n", s) == 1)
	  { if (nWl == cap) { cap *= 2 ; if (!(wl = realloc (wl, 4 * cap))) die ("can't allocate barcode table") ; }
	    if (h10x_pack_barcode (s, wl + nWl)) die ("bad barcode line %d in %s: %s", (int) nWl + 1, argv[1], s) ;
	    ++nWl ;
	  }
	fclose (f) ;
	fprintf (stderr, "read %d barcodes from file %s\n", (int) nWl, wlName) ;
	argc -= 2 ; argv += 2 ;
      }
    else if (!strcmp (*argv, "-checkId")) { --argc ; ++argv ; }
    else if (!strcmp (*argv, "-sort")) { flags |= H10X_FQ2B_SORT ; --argc ; ++argv ; }
    else if (!strcmp (*argv, "-o"))
      { if (!(fout = fopen (argv[1], "wb"))) die ("failed to open output file %s", argv[1]) ;
	argc -= 2 ; argv += 2 ;
      }
    else die ("Unknown arg %s for fq2b - run without args for usage", *argv) ;
  if (argc < 1 || argc > 2)
    die ("Usage: fq2fqb [opts] <fastq.gz> [<fastq.gz>]\n"
	 "  Converts fastq to binary with 2 bits per base, converting N to A (!).\n"
	 "  If two fastq files are given they are interleaved.\n"
	 "Opts: -10x <whitelist file>\n"
	 "      -checkId  checks whether id lines match in first and second files\n"
	 "      -sort     group the records by barcode, as `bsort -k 4 -r <record bytes>` does (on the GPU)\n"
	 "      -o <outfile> [standard output]\n"
	 "  10x option matches first16bp barcode of read 1 to whitelist.\n"
	 "  Only outputs an entry if there is a match after correcting for 1 mismatch\n") ;
  uint64_t n1 = 0, n2 = 0 ;
  char *t1 = slurp (argv[0], &n1), *t2 = argc == 2 ? slurp (argv[1], &n2) : 0 ;
  char err[512] ; err[0] = 0 ;
  h10x_params p ;
  memset (&p, 0, sizeof (p)) ;
  p.k = 21 ; p.w = 31 ; p.B = 20 ; p.chunkSize = 100000 ; p.factor1 = h10x_factor1_from_seed (17) ;
  p.device = getenv ("H10X_DEVICE") ? atoi (getenv ("H10X_DEVICE")) : 0 ;
  h10x_ctx *ctx = h10x_gpu_create (&p, err, sizeof (err)) ;
  if (!ctx) die ("%s", err) ;
  h10x_fq2b_out o ;
  int st = h10x_gpu_fq2b (ctx, t1, n1, t2, n2, wl, nWl, flags, &o, err, sizeof (err)) ;
  if (st) die ("%s", err) ;
  if (t2) checkIds (t1, n1, t2, n2) ;
  if (o.nRecords && fwrite (o.fqb, 4 * (size_t) o.recWords, o.nRecords, fout) != o.nRecords) die ("write error") ;
  if (fout != stdout) fclose (fout) ;
  int n = (int) o.nRecords ;
  if (t2)
    fprintf (stderr, "written %d read pairs %d + %d bp packed in %d word records\n", n, (int) o.s1Len, (int) o.s2Len, (int) o.recWords) ;
  else
    fprintf (stderr, "written %d reads %d bp packed in %d word records\n", n, (int) o.s1Len, (int) o.recWords) ;
  if (wl)
    { fprintf (stderr, "%d (%.1f%%) not matching barcodes were dropped\n", (int) o.nBad, 100.0 * o.nBad / (double) (o.nBad + n)) ;
      fprintf (stderr, "%d (%.1f%%) of those that matched were error corrected\n", (int) o.nFixed, 100.0 * o.nFixed / (double) n) ;
      fprintf (stderr, "by base position:") ;
      for (int i = 0 ; i < 16 ; ++i) fprintf (stderr, " %d", (int) o.nFixBase[i]) ;
      fprintf (stderr, "\n") ;
    }
  h10x_gpu_destroy (ctx) ;
  free (t1) ; free (t2) ; free (wl) ;
  return 0 ;
}
