/* Minimal GSL-compatible RNG shim (test infrastructure only; see oracle/README.md).
 * GSL is not installed in this image; this restates the published MT19937
 * generator as wrapped by gsl_rng_mt19937 (seed 0 -> 4357, uniform = u32/2^32). */
#ifndef SHIM_GSL_RNG_H
#define SHIM_GSL_RNG_H
#include <stddef.h>
typedef struct { const char *name; int kind; } gsl_rng_type;
typedef struct gsl_rng_s {
  const gsl_rng_type *type;
  unsigned long mt[624];
  int mti;
  /* shim extension: counts draws, lets the oracle driver inspect stream usage */
  unsigned long long ndraws;
  /* shim extension: counter-based substreams (type shim_rng_philox), see shim_philox_seek */
  unsigned int pkey[2], pctr[4], pbuf[4];
  unsigned int ppos;
  int pfirst_pending;        /* shim_philox_seek_cell: the first draw comes from a block shared by 4 cells */
  unsigned int pfirst;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_ranlux;
/* Philox4x32-10 (Salmon et al. 2011) in counter mode. Word j of substream (seed, stream, index)
 * is word j%4 of the block with counter {index_lo, index_hi, j/4, stream} and key {seed_lo, seed_hi}.
 * This is the definition the CUDA kernels implement (colore_b200/csrc/clr_rng.cuh). */
extern const gsl_rng_type *shim_rng_philox;
void shim_philox_seek(gsl_rng *r, unsigned long long seed, unsigned int stream, unsigned long long index);
void shim_philox4x32_10(const unsigned int ctr[4], const unsigned int key[2], unsigned int out[4]);
/* Per-cell Poisson substream: draw 0 = word (cell & 3) of the block {cell>>2 lo, cell>>2 hi, 0,
 * stream | 0x80000000} (one Philox call serves the first draw of 4 neighbouring cells -- the only draw
 * 97% of the cells ever need); draws j >= 1 = words j-1 of substream (seed, stream, cell). */
void shim_philox_seek_cell(gsl_rng *r, unsigned long long seed, unsigned int stream, unsigned long long cell);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);
void gsl_rng_free(gsl_rng *r);
#endif
