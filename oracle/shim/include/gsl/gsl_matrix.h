#ifndef SHIM_GSL_MATRIX_H
#define SHIM_GSL_MATRIX_H
/* not used by the reference beyond the include */
#endif
