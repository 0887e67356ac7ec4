#ifndef SHIM_GSL_BLAS_H
#define SHIM_GSL_BLAS_H
/* not used by the reference beyond the include */
#endif
