/* TEST INFRASTRUCTURE ONLY -- FFTW3 stand-in for building the unmodified reference.
 * Plain mixed radix-2 / naive-DFT transforms evaluated in double precision, OpenMP over lines.
 * Pass order of the 3-D c2r: axis 0 (complex), axis 1 (complex), axis 2 (half-complex -> real),
 * i.e. the imaginary parts of the x-DC and x-Nyquist lines are dropped exactly as FFTW's
 * rdft2 does ("Exact c2r semantics on non-Hermitian input"). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fftw3.h"
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex dc;

struct shim_fftw_plan_s {
  int kind;      /* 0: c2r 3d, 1: r2c 3d, 2: 1-D complex */
  int prec;      /* 4: float, 8: double */
  int n0, n1, n2, sign;
  void *in, *out;
};

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* 1-D complex DFT of length n, sign s (exp(s*2*pi*i*j*k/n)), in place on buf; tw = exp(s*2*pi*i*k/n) */
static void fft1d(dc *buf, dc *scratch, const dc *tw, int n)
{
  if (is_pow2(n)) {
    int i, j = 0, len;
    for (i = 1; i < n; i++) {
      int bit = n >> 1;
      for (; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if (i < j) { dc t = buf[i]; buf[i] = buf[j]; buf[j] = t; }
    }
    for (len = 2; len <= n; len <<= 1) {
      int half = len >> 1, step = n / len;
      for (i = 0; i < n; i += len) {
        int k;
        for (k = 0; k < half; k++) {
          dc w = tw[k * step];
          dc u = buf[i + k], v = buf[i + k + half] * w;
          buf[i + k] = u + v;
          buf[i + k + half] = u - v;
        }
      }
    }
  } else {
    int j, k;
    for (k = 0; k < n; k++) {
      dc acc = 0;
      for (j = 0; j < n; j++) acc += buf[j] * tw[(int)(((long)j * k) % n)];
      scratch[k] = acc;
    }
    memcpy(buf, scratch, n * sizeof(dc));
  }
}

static dc *make_tw(int n, int sign)
{
  dc *tw = malloc(n * sizeof(dc));
  int k;
  for (k = 0; k < n; k++) {
    double a = sign * 2.0 * M_PI * k / n;
    tw[k] = cos(a) + I * sin(a);
  }
  return tw;
}

#define LOADC(p, prec, idx) ((prec) == 4 ? (dc)(((float _Complex *)(p))[idx]) : ((double _Complex *)(p))[idx])
#define STOREC(p, prec, idx, v) do { if ((prec) == 4) ((float _Complex *)(p))[idx] = (float _Complex)(v); \
                                     else ((double _Complex *)(p))[idx] = (v); } while (0)
#define LOADR(p, prec, idx) ((prec) == 4 ? (double)(((float *)(p))[idx]) : ((double *)(p))[idx])
#define STORER(p, prec, idx, v) do { if ((prec) == 4) ((float *)(p))[idx] = (float)(v); \
                                     else ((double *)(p))[idx] = (v); } while (0)

/* complex transform of `nlines` lines of length n: element j of line l is at base[l*lstride + j*stride] */
static void strided_pass(void *data, int prec, int n, long stride, long nouter, long ostride,
                         long ninner, int sign)
{
  dc *tw = make_tw(n, sign);
#pragma omp parallel
  {
    dc *buf = malloc(n * sizeof(dc)), *scr = malloc(n * sizeof(dc));
    long t;
#pragma omp for schedule(static)
    for (t = 0; t < nouter * ninner; t++) {
      long o = t / ninner, in = t % ninner;
      long base = o * ostride + in;
      int j;
      for (j = 0; j < n; j++) buf[j] = LOADC(data, prec, base + j * stride);
      fft1d(buf, scr, tw, n);
      for (j = 0; j < n; j++) STOREC(data, prec, base + j * stride, buf[j]);
    }
    free(buf); free(scr);
  }
  free(tw);
}

static void exec_c2r(struct shim_fftw_plan_s *p)
{
  int n0 = p->n0, n1 = p->n1, n2 = p->n2, nc = n2 / 2 + 1;
  int inplace = ((void *)p->in == (void *)p->out);
  long rpitch = inplace ? 2 * nc : n2;
  /* axis 0: for each (i1,k2): stride n1*nc */
  strided_pass(p->in, p->prec, n0, (long)n1 * nc, 1, 0, (long)n1 * nc, +1);
  /* axis 1: for each i0, each k2: stride nc */
  strided_pass(p->in, p->prec, n1, nc, n0, (long)n1 * nc, nc, +1);
  /* axis 2: half-complex -> real */
  {
    dc *tw = make_tw(n2, +1);
#pragma omp parallel
    {
      dc *buf = malloc(n2 * sizeof(dc)), *scr = malloc(n2 * sizeof(dc));
      long l;
#pragma omp for schedule(static)
      for (l = 0; l < (long)n0 * n1; l++) {
        int k;
        for (k = 0; k < nc; k++) buf[k] = LOADC(p->in, p->prec, l * nc + k);
        buf[0] = creal(buf[0]);
        if ((n2 & 1) == 0) buf[n2 / 2] = creal(buf[n2 / 2]);
        for (k = 1; k < (n2 + 1) / 2; k++) buf[n2 - k] = conj(buf[k]);
        fft1d(buf, scr, tw, n2);
        for (k = 0; k < n2; k++) STORER(p->out, p->prec, l * rpitch + k, creal(buf[k]));
      }
      free(buf); free(scr);
    }
    free(tw);
  }
}

static void exec_r2c(struct shim_fftw_plan_s *p)
{
  int n0 = p->n0, n1 = p->n1, n2 = p->n2, nc = n2 / 2 + 1;
  int inplace = ((void *)p->in == (void *)p->out);
  long rpitch = inplace ? 2 * nc : n2;
  {
    dc *tw = make_tw(n2, -1);
#pragma omp parallel
    {
      dc *buf = malloc(n2 * sizeof(dc)), *scr = malloc(n2 * sizeof(dc));
      long l;
#pragma omp for schedule(static)
      for (l = 0; l < (long)n0 * n1; l++) {
        int k;
        for (k = 0; k < n2; k++) buf[k] = LOADR(p->in, p->prec, l * rpitch + k);
        fft1d(buf, scr, tw, n2);
        for (k = 0; k < nc; k++) STOREC(p->out, p->prec, l * nc + k, buf[k]);
      }
      free(buf); free(scr);
    }
    free(tw);
  }
  strided_pass(p->out, p->prec, n1, nc, n0, (long)n1 * nc, nc, -1);
  strided_pass(p->out, p->prec, n0, (long)n1 * nc, 1, 0, (long)n1 * nc, -1);
}

static void exec_1d(struct shim_fftw_plan_s *p)
{
  int n = p->n0, k;
  dc *tw = make_tw(n, p->sign);
  dc *buf = malloc(n * sizeof(dc)), *scr = malloc(n * sizeof(dc));
  for (k = 0; k < n; k++) buf[k] = ((dc *)p->in)[k];
  fft1d(buf, scr, tw, n);
  for (k = 0; k < n; k++) ((dc *)p->out)[k] = buf[k];
  free(buf); free(scr); free(tw);
}

static struct shim_fftw_plan_s *mkplan(int kind, int prec, int n0, int n1, int n2, int sign, void *in, void *out)
{
  struct shim_fftw_plan_s *p = calloc(1, sizeof(*p));
  p->kind = kind; p->prec = prec; p->n0 = n0; p->n1 = n1; p->n2 = n2; p->sign = sign; p->in = in; p->out = out;
  return p;
}

static void execute(struct shim_fftw_plan_s *p)
{
  if (p->kind == 0) exec_c2r(p);
  else if (p->kind == 1) exec_r2c(p);
  else exec_1d(p);
}

fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned f)
{ (void)f; return mkplan(0, 4, n0, n1, n2, +1, in, out); }
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned f)
{ (void)f; return mkplan(1, 4, n0, n1, n2, -1, in, out); }
void fftwf_execute(const fftwf_plan p) { execute(p); }
void fftwf_destroy_plan(fftwf_plan p) { free(p); }
fftwf_complex *fftwf_alloc_complex(size_t n) { void *q = NULL; if (posix_memalign(&q, 64, n * sizeof(fftwf_complex))) return NULL; return q; }
void fftwf_free(void *p) { free(p); }
int fftwf_init_threads(void) { return 1; }
void fftwf_plan_with_nthreads(int n) { (void)n; }
void fftwf_cleanup_threads(void) {}

fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in, double *out, unsigned f)
{ (void)f; return mkplan(0, 8, n0, n1, n2, +1, in, out); }
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned f)
{ (void)f; return mkplan(1, 8, n0, n1, n2, -1, in, out); }
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned f)
{ (void)f; return mkplan(2, 8, n, 1, 1, sign, in, out); }
void fftw_execute(const fftw_plan p) { execute(p); }
void fftw_destroy_plan(fftw_plan p) { free(p); }
fftw_complex *fftw_alloc_complex(size_t n) { void *q = NULL; if (posix_memalign(&q, 64, n * sizeof(fftw_complex))) return NULL; return q; }
void fftw_free(void *p) { free(p); }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int n) { (void)n; }
void fftw_cleanup_threads(void) {}
