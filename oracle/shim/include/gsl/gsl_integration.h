/* Shim: same call signatures as GSL's QUADPACK wrappers, implemented with one
 * adaptive Gauss-Kronrod (7,15) bisection scheme. Only used by the reference's
 * host-side table construction (cosmo.c, cosmo_mad.c), never by the hot path. */
#ifndef SHIM_GSL_INTEGRATION_H
#define SHIM_GSL_INTEGRATION_H
#include <stddef.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F,x) (*((F)->function))(x,(F)->params)
typedef struct { size_t limit; } gsl_integration_workspace;
typedef struct { int dummy; } gsl_integration_qawo_table;
enum gsl_integration_qawo_enum { GSL_INTEG_COSINE, GSL_INTEG_SINE };
enum { GSL_INTEG_GAUSS15=1, GSL_INTEG_GAUSS21=2, GSL_INTEG_GAUSS31=3, GSL_INTEG_GAUSS41=4,
       GSL_INTEG_GAUSS51=5, GSL_INTEG_GAUSS61=6 };
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
gsl_integration_qawo_table *gsl_integration_qawo_table_alloc(double omega, double L,
    enum gsl_integration_qawo_enum sine, size_t n);
void gsl_integration_qawo_table_free(gsl_integration_qawo_table *t);
int gsl_integration_qng(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        double *result, double *abserr, size_t *neval);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w,
                        double *result, double *abserr);
int gsl_integration_qagil(gsl_function *f, double b, double epsabs, double epsrel, size_t limit,
                          gsl_integration_workspace *w, double *result, double *abserr);
int gsl_integration_qawf(gsl_function *f, double a, double epsabs, size_t limit,
                         gsl_integration_workspace *w, gsl_integration_workspace *cw,
                         gsl_integration_qawo_table *wf, double *result, double *abserr);
#endif
