/* Minimal cfitsio-compatible shim (TEST INFRASTRUCTURE ONLY): writes standard-conforming FITS
 * binary tables / images for the calls made by healpix_extra.c:4-57 and io.c:1070-1205.
 * Reading entry points exist only so the reference links; they fail loudly. */
#ifndef SHIM_FITSIO_H
#define SHIM_FITSIO_H
#include <stdio.h>
typedef struct shim_fitsfile fitsfile;
typedef long long LONGLONG;
#define READONLY 0
#define BINARY_TBL 2
#define FLOAT_IMG (-32)
#define TSTRING 16
#define TINT 31
#define TLONG 41
#define TFLOAT 42
#define TDOUBLE 82
#define FLEN_VALUE 71
int fits_create_file(fitsfile **fptr, const char *fname, int *status);
int fits_create_tbl(fitsfile *fptr, int tbltype, LONGLONG naxis2, int tfields, char **ttype,
                    char **tform, char **tunit, const char *extname, int *status);
int fits_create_img(fitsfile *fptr, int bitpix, int naxis, long *naxes, int *status);
int fits_write_key(fitsfile *fptr, int datatype, const char *keyname, void *value,
                   const char *comment, int *status);
int fits_update_key(fitsfile *fptr, int datatype, const char *keyname, void *value,
                    const char *comment, int *status);
int fits_write_comment(fitsfile *fptr, const char *comment, int *status);
int fits_get_rowsize(fitsfile *fptr, long *nrows, int *status);
int fits_write_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                   LONGLONG nelements, void *array, int *status);
int fits_write_img(fitsfile *fptr, int datatype, LONGLONG firstelem, LONGLONG nelements,
                   void *array, int *status);
int fits_close_file(fitsfile *fptr, int *status);
int fits_open_file(fitsfile **fptr, const char *fname, int mode, int *status);
int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *hdutype, int *status);
int fits_read_key_lng(fitsfile *fptr, const char *key, long *value, char *comm, int *status);
int fits_read_keys_lng(fitsfile *fptr, const char *key, int nstart, int nmax, long *value,
                       int *nfound, int *status);
int fits_read_key(fitsfile *fptr, int datatype, const char *key, void *value, char *comm, int *status);
int fits_read_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                  LONGLONG nelements, void *nulval, void *array, int *anynul, int *status);
#endif
