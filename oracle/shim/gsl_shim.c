/* TEST INFRASTRUCTURE ONLY -- dependency shim so that the UNMODIFIED reference
 * sources under /root/reference/src compile in an image without GSL.
 *
 * Restates published algorithms of GSL (GNU Scientific Library, unpinned by the
 * reference; README says "tested for versions 3.*"):
 *   - MT19937 (Matsumoto & Nishimura 1998) with GSL's seeding (rng/mt.c: seed 0 ->
 *     4357, Knuth LCG 1812433253) and uniform = u32 / 2^32          [used by common.c:170-201]
 *   - gsl_ran_poisson: Knuth's product method for mu<=10, gamma/binomial reduction
 *     above (randist/poisson.c)                                      [used by common.c:185-189]
 *   - natural cubic spline / linear interpolation                   [cosmo.c, density.c:1304]
 *   - QUADPACK entry points (qng/qag/qagil) restated with ONE adaptive
 *     Gauss-Kronrod(7,15) scheme: same tolerances, not bit-identical nodes.
 *     They only build host-side tables (inputs of the hot path), never hot-path results.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_spline.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_sf_gamma.h>

/* ------------------------------------------------------------------ RNG */
static const gsl_rng_type mt_type = {"mt19937", 0};
static const gsl_rng_type rl_type = {"ranlux", 1};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_ranlux = &rl_type; /* only referenced in a comment of common.c */
static const gsl_rng_type px_type = {"philox4x32-10-counter", 2};
const gsl_rng_type *shim_rng_philox = &px_type;

void shim_philox4x32_10(const unsigned int ctr[4], const unsigned int key[2], unsigned int out[4])
{
  unsigned int c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  int r;
  for (r = 0; r < 10; r++) {
    unsigned long long p0 = 0xD2511F53ULL * c0, p1 = 0xCD9E8D57ULL * c2;
    unsigned int n0 = (unsigned int)(p1 >> 32) ^ c1 ^ k0;
    unsigned int n1 = (unsigned int)p1;
    unsigned int n2 = (unsigned int)(p0 >> 32) ^ c3 ^ k1;
    unsigned int n3 = (unsigned int)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9U; k1 += 0xBB67AE85U;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void shim_philox_seek(gsl_rng *r, unsigned long long seed, unsigned int stream, unsigned long long index)
{
  r->pkey[0] = (unsigned int)seed; r->pkey[1] = (unsigned int)(seed >> 32);
  r->pctr[0] = (unsigned int)index; r->pctr[1] = (unsigned int)(index >> 32);
  r->pctr[2] = 0; r->pctr[3] = stream;
  r->ppos = 0;
  r->pfirst_pending = 0;
}

void shim_philox_seek_cell(gsl_rng *r, unsigned long long seed, unsigned int stream, unsigned long long cell)
{
  unsigned int ctr[4], key[2], out[4];
  unsigned long long grp = cell >> 2;
  key[0] = (unsigned int)seed; key[1] = (unsigned int)(seed >> 32);
  ctr[0] = (unsigned int)grp; ctr[1] = (unsigned int)(grp >> 32); ctr[2] = 0; ctr[3] = stream | 0x80000000u;
  shim_philox4x32_10(ctr, key, out);
  shim_philox_seek(r, seed, stream, cell);
  r->pfirst = out[cell & 3];
  r->pfirst_pending = 1;
}

#define MT_N 624
#define MT_M 397

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
  gsl_rng *r = calloc(1, sizeof(gsl_rng));
  r->type = T;
  if (T->kind == 2) shim_philox_seek(r, 0, 0, 0);
  else gsl_rng_set(r, 0);
  return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long s)
{
  int i;
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  for (i = 1; i < MT_N; i++)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + i) & 0xffffffffUL;
  r->mti = MT_N;
  r->ndraws = 0;
}

unsigned long gsl_rng_get(gsl_rng *r)
{
  unsigned long k;
  unsigned long *mt = r->mt;
  if (r->type->kind == 2) {
    if (r->pfirst_pending) { r->pfirst_pending = 0; r->ndraws++; return r->pfirst; }
    if ((r->ppos & 3) == 0) {
      r->pctr[2] = r->ppos >> 2;
      shim_philox4x32_10(r->pctr, r->pkey, r->pbuf);
    }
    r->ndraws++;
    return r->pbuf[(r->ppos++) & 3];
  }
  if (r->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    for (; kk < MT_N - 1; kk++) {
      unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    {
      unsigned long y = (mt[MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
      mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    r->mti = 0;
  }
  k = mt[r->mti];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  r->mti++;
  r->ndraws++;
  return k & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

double gsl_rng_uniform_pos(gsl_rng *r)
{
  double x;
  do { x = gsl_rng_uniform(r); } while (x == 0);
  return x;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

/* ------------------------------------------------------------------ randist */
static double gamma_large(gsl_rng *r, double a)
{
  /* Ahrens' rejection method (Knuth vol.2, 3.4.1), as in GSL randist/gamma.c */
  double sqa, x, y, v;
  sqa = sqrt(2 * a - 1);
  do {
    do {
      y = tan(M_PI * gsl_rng_uniform(r));
      x = sqa * y + a - 1;
    } while (x <= 0);
    v = gsl_rng_uniform(r);
  } while (v > (1 + y * y) * exp((a - 1) * log(x / (a - 1)) - sqa * y));
  return x;
}

double gsl_ran_gamma_int(gsl_rng *r, unsigned int a)
{
  if (a < 12) {
    unsigned int i;
    double prod = 1;
    for (i = 0; i < a; i++) prod *= gsl_rng_uniform_pos(r);
    return -log(prod);
  }
  return gamma_large(r, (double)a);
}

static double stirling_corr(double y1)
{
  double y2 = y1 * y1;
  return (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / y2) / y2) / y2) / y2) / y1 / 166320.0;
}

unsigned int gsl_ran_binomial(gsl_rng *rng, double p, unsigned int n)
{
  /* Kachitvichyanukul & Schmeiser (1988): BINV for small mean, BTPE otherwise,
   * laid out as in GSL randist/binomial_tpe.c (SMALL_MEAN 14, BINV_CUTOFF 110,
   * FAR_FROM_MEAN 20). */
  int ix, flipped = 0;
  double q, s, np;
  if (n == 0) return 0;
  if (p > 0.5) { p = 1.0 - p; flipped = 1; }
  q = 1 - p;
  s = p / q;
  np = n * p;
  if (np < 14) {
    double f0 = pow(q, (double)n);
    while (1) {
      double f = f0;
      double u = gsl_rng_uniform(rng);
      for (ix = 0; ix <= 110; ++ix) {
        if (u < f) goto Finish;
        u -= f;
        f *= s * (n - ix) / (ix + 1);
      }
    }
  } else {
    int k;
    double ffm = np + p;
    int m = (int)ffm;
    double xm = m + 0.5;
    double npq = np * q;
    double p1 = floor(2.195 * sqrt(npq) - 4.6 * q) + 0.5;
    double xl = xm - p1;
    double xr = xm + p1;
    double c = 0.134 + 20.5 / (15.3 + m);
    double p2 = p1 * (1.0 + c + c);
    double al = (ffm - xl) / (ffm - xl * p);
    double lambda_l = al * (1.0 + 0.5 * al);
    double ar = (xr - ffm) / (xr * q);
    double lambda_r = ar * (1.0 + 0.5 * ar);
    double p3 = p2 + c / lambda_l;
    double p4 = p3 + c / lambda_r;
    double var, accept;
    double u, v;
  TryAgain:
    u = gsl_rng_uniform(rng) * p4;
    v = gsl_rng_uniform(rng);
    if (u <= p1) {
      ix = (int)(xm - p1 * v + u);
      goto Finish;
    } else if (u <= p2) {
      double x = xl + (u - p1) / c;
      v = v * c + 1.0 - fabs(x - xm) / p1;
      if (v > 1.0 || v <= 0.0) goto TryAgain;
      ix = (int)x;
    } else if (u <= p3) {
      ix = (int)(xl + log(v) / lambda_l);
      if (ix < 0) goto TryAgain;
      v *= ((u - p2) * lambda_l);
    } else {
      ix = (int)(xr - log(v) / lambda_r);
      if (ix > (double)n) goto TryAgain;
      v *= ((u - p3) * lambda_r);
    }
    k = abs(ix - m);
    if (k <= 20) {
      double g = (n + 1) * s;
      double f = 1.0;
      int i;
      var = v;
      if (m < ix) {
        for (i = m + 1; i <= ix; i++) f *= (g / i - s);
      } else if (m > ix) {
        for (i = ix + 1; i <= m; i++) f /= (g / i - s);
      }
      accept = f;
    } else {
      var = log(v);
      if (k < npq / 2 - 1) {
        double amaxp = k / npq * ((k * (k / 3.0 + 0.625) + (1.0 / 6.0)) / npq + 0.5);
        double ynorm = -(k * k / (2.0 * npq));
        if (var < ynorm - amaxp) goto Finish;
        if (var > ynorm + amaxp) goto TryAgain;
      }
      {
        double x1 = ix + 1.0;
        double w1 = n - ix + 1.0;
        double f1 = m + 1.0;
        double z1 = n + 1.0 - m;
        accept = xm * log(f1 / x1) + (n - m + 0.5) * log(z1 / w1) + (ix - m) * log(w1 * p / (x1 * q))
                 + stirling_corr(f1) + stirling_corr(z1) - stirling_corr(x1) - stirling_corr(w1);
      }
    }
    if (var <= accept) goto Finish;
    else goto TryAgain;
  }
Finish:
  return (flipped) ? (n - ix) : (unsigned int)ix;
}

unsigned int gsl_ran_poisson(gsl_rng *r, double mu)
{
  double emu;
  double prod = 1.0;
  unsigned int k = 0;
  while (mu > 10) {
    unsigned int m = mu * (7.0 / 8.0);
    double X = gsl_ran_gamma_int(r, m);
    if (X >= mu) {
      return k + gsl_ran_binomial(r, mu / X, m - 1);
    } else {
      k += m;
      mu -= X;
    }
  }
  emu = exp(-mu);
  do {
    prod *= gsl_rng_uniform(r);
    k++;
  } while (prod > emu);
  return k - 1;
}

/* ------------------------------------------------------------------ splines */
static const gsl_interp_type lin_type = {"linear", 0};
static const gsl_interp_type csp_type = {"cspline", 1};
const gsl_interp_type *gsl_interp_linear = &lin_type;
const gsl_interp_type *gsl_interp_cspline = &csp_type;

gsl_interp_accel *gsl_interp_accel_alloc(void) { return calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
  gsl_spline *s = calloc(1, sizeof(gsl_spline));
  s->type = T;
  s->size = size;
  s->x = malloc(size * sizeof(double));
  s->y = malloc(size * sizeof(double));
  s->y2 = calloc(size, sizeof(double));
  return s;
}

int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t n)
{
  size_t i;
  memcpy(s->x, xa, n * sizeof(double));
  memcpy(s->y, ya, n * sizeof(double));
  s->size = n;
  if (s->type->kind == 1 && n > 2) {
    /* natural cubic spline: tridiagonal solve for second derivatives, y2[0]=y2[n-1]=0 */
    double *u = malloc(n * sizeof(double));
    s->y2[0] = 0;
    u[0] = 0;
    for (i = 1; i < n - 1; i++) {
      double sig = (xa[i] - xa[i - 1]) / (xa[i + 1] - xa[i - 1]);
      double pp = sig * s->y2[i - 1] + 2.0;
      s->y2[i] = (sig - 1.0) / pp;
      u[i] = (ya[i + 1] - ya[i]) / (xa[i + 1] - xa[i]) - (ya[i] - ya[i - 1]) / (xa[i] - xa[i - 1]);
      u[i] = (6.0 * u[i] / (xa[i + 1] - xa[i - 1]) - sig * u[i - 1]) / pp;
    }
    s->y2[n - 1] = 0;
    for (i = n - 1; i-- > 0;) s->y2[i] = s->y2[i] * s->y2[i + 1] + u[i];
    free(u);
  }
  return 0;
}

double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
  size_t lo = 0, hi = s->size - 1;
  double h, A, B;
  (void)a;
  if (x < s->x[0] || x > s->x[s->size - 1]) {
    /* GSL raises GSL_EDOM and (handler off) returns NaN */
    return NAN;
  }
  while (hi - lo > 1) {
    size_t mid = (hi + lo) >> 1;
    if (s->x[mid] > x) hi = mid; else lo = mid;
  }
  h = s->x[hi] - s->x[lo];
  A = (s->x[hi] - x) / h;
  B = (x - s->x[lo]) / h;
  if (s->type->kind == 0)
    return A * s->y[lo] + B * s->y[hi];
  return A * s->y[lo] + B * s->y[hi]
         + ((A * A * A - A) * s->y2[lo] + (B * B * B - B) * s->y2[hi]) * (h * h) / 6.0;
}

void gsl_spline_free(gsl_spline *s)
{
  if (!s) return;
  free(s->x); free(s->y); free(s->y2); free(s);
}

/* ------------------------------------------------------------------ errno */
gsl_error_handler_t *gsl_set_error_handler_off(void) { return NULL; }

double gsl_sf_gamma(double x) { return tgamma(x); }

/* ------------------------------------------------------------------ integration */
static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
  0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
  0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
  0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
  0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
  0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
  0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
  0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

typedef double (*fn1)(double, void *);

static void gk15(fn1 f, void *p, double a, double b, double *res, double *err)
{
  double c = 0.5 * (a + b), h = 0.5 * (b - a);
  double fc = f(c, p);
  double rk = fc * wgk[7], rg = fc * wg[3];
  int j;
  for (j = 0; j < 7; j++) {
    double dx = h * xgk[j];
    double f1 = f(c - dx, p), f2 = f(c + dx, p);
    rk += wgk[j] * (f1 + f2);
    if (j & 1) rg += wg[j / 2] * (f1 + f2);
  }
  *res = rk * h;
  *err = fabs((rk - rg) * h);
}

typedef struct { double a, b, r, e; } seg_t;

static int adaptive_gk(fn1 f, void *p, double a, double b, double epsabs, double epsrel,
                       double *result, double *abserr)
{
  /* global adaptive bisection: always split the interval with the largest error */
  int nseg = 1, cap = 4096, it;
  seg_t *s = malloc(cap * sizeof(seg_t));
  double tot, err;
  s[0].a = a; s[0].b = b;
  gk15(f, p, a, b, &s[0].r, &s[0].e);
  for (it = 0; it < 200000; it++) {
    int i, imax = 0;
    double tol;
    tot = 0; err = 0;
    for (i = 0; i < nseg; i++) { tot += s[i].r; err += s[i].e; if (s[i].e > s[imax].e) imax = i; }
    tol = fmax(epsabs, 0.1 * epsrel * fabs(tot)); /* 10x tighter than asked: tables only */
    if (err <= tol || nseg + 1 >= cap) break;
    {
      seg_t o = s[imax];
      double m = 0.5 * (o.a + o.b);
      if (m <= o.a || m >= o.b) break;
      s[imax].a = o.a; s[imax].b = m;
      gk15(f, p, o.a, m, &s[imax].r, &s[imax].e);
      s[nseg].a = m; s[nseg].b = o.b;
      gk15(f, p, m, o.b, &s[nseg].r, &s[nseg].e);
      nseg++;
    }
  }
  *result = tot;
  *abserr = err;
  free(s);
  return 0;
}

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n)
{
  gsl_integration_workspace *w = malloc(sizeof(*w));
  w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }
gsl_integration_qawo_table *gsl_integration_qawo_table_alloc(double omega, double L,
    enum gsl_integration_qawo_enum sine, size_t n)
{
  (void)omega; (void)L; (void)sine; (void)n;
  return calloc(1, sizeof(gsl_integration_qawo_table));
}
void gsl_integration_qawo_table_free(gsl_integration_qawo_table *t) { free(t); }

int gsl_integration_qng(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        double *result, double *abserr, size_t *neval)
{
  if (neval) *neval = 0;
  return adaptive_gk(f->function, f->params, a, b, epsabs, epsrel, result, abserr);
}

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w,
                        double *result, double *abserr)
{
  (void)limit; (void)key; (void)w;
  return adaptive_gk(f->function, f->params, a, b, epsabs, epsrel, result, abserr);
}

typedef struct { fn1 f; void *p; double b; } il_t;
static double il_fn(double t, void *vp)
{
  /* x = b - (1-t)/t maps t in (0,1] onto (-inf,b] (QUADPACK qagi transformation) */
  il_t *q = vp;
  double x = q->b - (1 - t) / t;
  return q->f(x, q->p) / (t * t);
}

int gsl_integration_qagil(gsl_function *f, double b, double epsabs, double epsrel, size_t limit,
                          gsl_integration_workspace *w, double *result, double *abserr)
{
  il_t q;
  (void)limit; (void)w;
  q.f = f->function; q.p = f->params; q.b = b;
  return adaptive_gk(il_fn, &q, 0, 1, epsabs, epsrel, result, abserr);
}

int gsl_integration_qawf(gsl_function *f, double a, double epsabs, size_t limit,
                         gsl_integration_workspace *w, gsl_integration_workspace *cw,
                         gsl_integration_qawo_table *wf, double *result, double *abserr)
{
  /* Unreachable from CoLoRe: sigL2() always calls xi2p_L with r=0, which takes the
   * qagil branch (cosmo.c:382-414,441). Fail loudly if that ever changes. */
  (void)f; (void)a; (void)epsabs; (void)limit; (void)w; (void)cw; (void)wf; (void)result; (void)abserr;
  fprintf(stderr, "gsl shim: gsl_integration_qawf is not implemented\n");
  exit(1);
}
