/* Minimal chealpix-compatible shim (TEST INFRASTRUCTURE ONLY): the HEALPix C entry points
 * the reference calls (srcs.c:268-269, common.c:536, healpix_extra.c, io.c). HEALPix is not
 * installed here; this restates the published pixelisation (Gorski et al. 2005) in the form
 * used by the HEALPix 3.x C library (ring/nest via (x,y,face) coordinates). */
#ifndef SHIM_CHEALPIX_H
#define SHIM_CHEALPIX_H
long nside2npix(long nside);
long npix2nside(long npix);
void vec2pix_ring(long nside, const double *vec, long *ipix);
void vec2pix_nest(long nside, const double *vec, long *ipix);
void ang2pix_ring(long nside, double theta, double phi, long *ipix);
void ang2pix_nest(long nside, double theta, double phi, long *ipix);
void pix2vec_ring(long nside, long ipix, double *vec);
void pix2vec_nest(long nside, long ipix, double *vec);
void pix2ang_ring(long nside, long ipix, double *theta, double *phi);
void pix2ang_nest(long nside, long ipix, double *theta, double *phi);
void ring2nest(long nside, long ipring, long *ipnest);
void nest2ring(long nside, long ipnest, long *ipring);
#endif
