#ifndef SHIM_GSL_SPLINE_H
#define SHIM_GSL_SPLINE_H
#include <stddef.h>
typedef struct { const char *name; int kind; } gsl_interp_type;
extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_cspline;
typedef struct { size_t cache; } gsl_interp_accel;
typedef struct {
  const gsl_interp_type *type;
  size_t size;
  double *x, *y, *y2; /* y2: second derivatives for the natural cubic spline */
} gsl_spline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size);
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *s);
#endif
