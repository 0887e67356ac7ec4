#ifndef SHIM_GSL_RANDIST_H
#define SHIM_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
unsigned int gsl_ran_poisson(gsl_rng *r, double mu);
double gsl_ran_gamma_int(gsl_rng *r, unsigned int a);
unsigned int gsl_ran_binomial(gsl_rng *r, double p, unsigned int n);
#endif
