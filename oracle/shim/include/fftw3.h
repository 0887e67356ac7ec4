/* Minimal FFTW3-compatible shim (TEST INFRASTRUCTURE ONLY). FFTW is not installed in
 * this image. Only the entry points the reference calls (fourier.c:81-283, fftlog.c:107-114)
 * are provided. Transform definition follows the FFTW manual: unnormalised, r2c = forward
 * (exp(-i..)), c2r = backward (exp(+i..)), last dimension halved (n/2+1 complex), in-place
 * real arrays padded to 2*(n/2+1); the multi-dimensional c2r runs complex transforms over the
 * leading dimensions first and the real (half-complex) transform over the last one. */
#ifndef SHIM_FFTW3_H
#define SHIM_FFTW3_H
#include <stddef.h>
#include <complex.h>
typedef float _Complex fftwf_complex;
typedef double _Complex fftw_complex;
typedef struct shim_fftw_plan_s *fftw_plan;
typedef struct shim_fftw_plan_s *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
fftwf_complex *fftwf_alloc_complex(size_t n);
void fftwf_free(void *p);
int fftwf_init_threads(void);
void fftwf_plan_with_nthreads(int n);
void fftwf_cleanup_threads(void);

fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in, double *out, unsigned flags);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
fftw_complex *fftw_alloc_complex(size_t n);
void fftw_free(void *p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int n);
void fftw_cleanup_threads(void);
#endif
