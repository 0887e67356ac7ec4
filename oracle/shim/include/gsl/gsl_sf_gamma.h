#ifndef SHIM_GSL_SF_GAMMA_H
#define SHIM_GSL_SF_GAMMA_H
double gsl_sf_gamma(double x);
#endif
