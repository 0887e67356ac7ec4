/* TEST INFRASTRUCTURE ONLY -- HEALPix pixel arithmetic for building the unmodified reference. */
#include <math.h>
#include "chealpix.h"

static const double twothird = 2.0 / 3.0;
static const double pi = 3.141592653589793238462643383279502884197;
static const double twopi = 6.283185307179586476925286766559005768394;
static const double halfpi = 1.570796326794896619231321691639751442099;
static const double inv_halfpi = 0.6366197723675813430755350534900574;
static const int jrll[] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const int jpll[] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

static double fmodulo(double v1, double v2)
{
  if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
  {
    double tmp = fmod(v1, v2) + v2;
    return (tmp == v2) ? 0. : tmp;
  }
}
static int imodulo(int v1, int v2) { int v = v1 % v2; return (v >= 0) ? v : v + v2; }
static int isqrt(int v) { return (int)(sqrt(v + 0.5)); }

static int spread_bits(int v)
{
  /* interleave a zero bit between the low 16 bits of v */
  unsigned int x = (unsigned int)v & 0xffff;
  x = (x | (x << 8)) & 0x00ff00ff;
  x = (x | (x << 4)) & 0x0f0f0f0f;
  x = (x | (x << 2)) & 0x33333333;
  x = (x | (x << 1)) & 0x55555555;
  return (int)x;
}
static int compress_bits(int v)
{
  unsigned int x = (unsigned int)v & 0x55555555;
  x = (x | (x >> 1)) & 0x33333333;
  x = (x | (x >> 2)) & 0x0f0f0f0f;
  x = (x | (x >> 4)) & 0x00ff00ff;
  x = (x | (x >> 8)) & 0x0000ffff;
  return (int)x;
}

static int xyf2nest(int nside, int ix, int iy, int face)
{ return face * nside * nside + spread_bits(ix) + (spread_bits(iy) << 1); }

static void nest2xyf(int nside, int pix, int *ix, int *iy, int *face)
{
  int npface = nside * nside;
  int p = pix & (npface - 1);
  *face = pix / npface;
  *ix = compress_bits(p);
  *iy = compress_bits(p >> 1);
}

static int xyf2ring(int nside, int ix, int iy, int face)
{
  int nl4 = 4 * nside;
  int jr = jrll[face] * nside - ix - iy - 1;
  int nr, kshift, n_before, jp;
  if (jr < nside) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; n_before = 12 * nside * nside - 2 * (nr + 1) * nr; kshift = 0; }
  else { nr = nside; n_before = 2 * nside * (nside - 1) + (jr - nside) * nl4; kshift = (jr - nside) & 1; }
  jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  else if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

static void ring2xyf(int nside, int pix, int *ix, int *iy, int *face)
{
  int iring, iphi, kshift, nr, irt, ipt;
  int ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside, nl2 = 2 * nside;
  if (pix < ncap) {
    iring = (1 + isqrt(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0; nr = iring;
    *face = (iphi - 1) / nr;
  } else if (pix < (npix - ncap)) {
    int ip = pix - ncap, ire, irm, ifm, ifp;
    iring = ip / (4 * nside) + nside;
    iphi = ip % (4 * nside) + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    ire = iring - nside + 1;
    irm = nl2 + 2 - ire;
    ifm = (iphi - ire / 2 + nside - 1) / nside;
    ifp = (iphi - irm / 2 + nside - 1) / nside;
    if (ifp == ifm) *face = (ifp == 4) ? 4 : ifp + 4;
    else if (ifp < ifm) *face = ifp;
    else *face = ifm + 8;
  } else {
    int ip = npix - pix;
    iring = (1 + isqrt(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0; nr = iring;
    iring = 2 * nl2 - iring;
    *face = 8 + (iphi - 1) / nr;
  }
  irt = iring - jrll[*face] * nside + 1;
  ipt = 2 * iphi - jpll[*face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  *ix = (ipt - irt) >> 1;
  *iy = (-(ipt + irt)) >> 1;
}

static long ang2pix_ring_z_phi(long nside, double z, double phi)
{
  double za = fabs(z);
  double tt = fmodulo(phi, twopi) * inv_halfpi;
  if (za <= twothird) {
    double temp1 = nside * (0.5 + tt);
    double temp2 = nside * z * 0.75;
    int jp = (int)(temp1 - temp2);
    int jm = (int)(temp1 + temp2);
    int ir = nside + 1 + jp - jm;
    int kshift = 1 - (ir & 1);
    int ip = (jp + jm - nside + kshift + 1) / 2;
    ip = imodulo(ip, 4 * nside);
    return nside * (nside - 1) * 2 + (ir - 1) * 4 * nside + ip;
  } else {
    double tp = tt - (int)(tt);
    double tmp = nside * sqrt(3 * (1 - za));
    int jp = (int)(tp * tmp);
    int jm = (int)((1.0 - tp) * tmp);
    int ir = jp + jm + 1;
    int ip = (int)(tt * ir);
    ip = imodulo(ip, 4 * ir);
    if (z > 0) return 2 * ir * (ir - 1) + ip;
    else return 12 * nside * nside - 2 * ir * (ir + 1) + ip;
  }
}

static long ang2pix_nest_z_phi(long nside, double z, double phi)
{
  double za = fabs(z);
  double tt = fmodulo(phi, twopi) * inv_halfpi;
  int face, ix, iy;
  if (za <= twothird) {
    double temp1 = nside * (0.5 + tt);
    double temp2 = nside * (z * 0.75);
    int jp = (int)(temp1 - temp2);
    int jm = (int)(temp1 + temp2);
    int ifp = jp / nside;
    int ifm = jm / nside;
    face = (ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8));
    ix = jm & (nside - 1);
    iy = nside - (jp & (nside - 1)) - 1;
  } else {
    int ntt = (int)tt, jp, jm;
    double tp, tmp;
    if (ntt >= 4) ntt = 3;
    tp = tt - ntt;
    tmp = nside * sqrt(3 * (1 - za));
    jp = (int)(tp * tmp);
    jm = (int)((1.0 - tp) * tmp);
    if (jp >= nside) jp = nside - 1;
    if (jm >= nside) jm = nside - 1;
    if (z >= 0) { face = ntt; ix = nside - jm - 1; iy = nside - jp - 1; }
    else { face = ntt + 8; ix = jp; iy = jm; }
  }
  return xyf2nest(nside, ix, iy, face);
}

static void pix2ang_ring_z_phi(int nside, int pix, double *z, double *phi)
{
  long ncap = nside * (nside - 1) * 2;
  long npix = 12 * nside * nside;
  double fact2 = 4. / npix;
  if (pix < ncap) {
    int iring = (1 + isqrt(1 + 2 * pix)) >> 1;
    int iphi = (pix + 1) - 2 * iring * (iring - 1);
    *z = 1.0 - (iring * iring) * fact2;
    *phi = (iphi - 0.5) * halfpi / iring;
  } else if (pix < (npix - ncap)) {
    double fact1 = (nside << 1) * fact2;
    int ip = pix - ncap;
    int iring = ip / (4 * nside) + nside;
    int iphi = ip % (4 * nside) + 1;
    double fodd = ((iring + nside) & 1) ? 1 : 0.5;
    int nl2 = 2 * nside;
    *z = (nl2 - iring) * fact1;
    *phi = (iphi - fodd) * pi / nl2;
  } else {
    int ip = npix - pix;
    int iring = (1 + isqrt(2 * ip - 1)) >> 1;
    int iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    *z = -1.0 + (iring * iring) * fact2;
    *phi = (iphi - 0.5) * halfpi / iring;
  }
}

static void pix2ang_nest_z_phi(int nside, int pix, double *z, double *phi)
{
  int nl4 = nside * 4;
  int npix = 12 * nside * nside;
  double fact2 = 4. / npix;
  int face, ix, iy, nr, kshift, jr, jp;
  nest2xyf(nside, pix, &ix, &iy, &face);
  jr = jrll[face] * nside - ix - iy - 1;
  if (jr < nside) { nr = jr; *z = 1 - nr * nr * fact2; kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; *z = nr * nr * fact2 - 1; kshift = 0; }
  else { double fact1 = (nside << 1) * fact2; nr = nside; *z = (2 * nside - jr) * fact1; kshift = (jr - nside) & 1; }
  jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  if (jp < 1) jp += nl4;
  *phi = (jp - (kshift + 1) * 0.5) * (halfpi / nr);
}

long nside2npix(long nside) { return 12 * nside * nside; }
long npix2nside(long npix)
{
  long res = (long)(sqrt(npix / 12 + 0.5));
  return (res * res * 12 == npix) ? res : -1;
}
void vec2pix_ring(long nside, const double *vec, long *ipix)
{
  double vlen = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  *ipix = ang2pix_ring_z_phi(nside, vec[2] / vlen, atan2(vec[1], vec[0]));
}
void vec2pix_nest(long nside, const double *vec, long *ipix)
{
  double vlen = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  *ipix = ang2pix_nest_z_phi(nside, vec[2] / vlen, atan2(vec[1], vec[0]));
}
void ang2pix_ring(long nside, double theta, double phi, long *ipix)
{ *ipix = ang2pix_ring_z_phi(nside, cos(theta), phi); }
void ang2pix_nest(long nside, double theta, double phi, long *ipix)
{ *ipix = ang2pix_nest_z_phi(nside, cos(theta), phi); }
void pix2vec_ring(long nside, long ipix, double *vec)
{
  double z, phi, st;
  pix2ang_ring_z_phi(nside, ipix, &z, &phi);
  st = sqrt((1. - z) * (1. + z));
  vec[0] = st * cos(phi); vec[1] = st * sin(phi); vec[2] = z;
}
void pix2vec_nest(long nside, long ipix, double *vec)
{
  double z, phi, st;
  pix2ang_nest_z_phi(nside, ipix, &z, &phi);
  st = sqrt((1. - z) * (1. + z));
  vec[0] = st * cos(phi); vec[1] = st * sin(phi); vec[2] = z;
}
void pix2ang_ring(long nside, long ipix, double *theta, double *phi)
{ double z; pix2ang_ring_z_phi(nside, ipix, &z, phi); *theta = acos(z); }
void pix2ang_nest(long nside, long ipix, double *theta, double *phi)
{ double z; pix2ang_nest_z_phi(nside, ipix, &z, phi); *theta = acos(z); }
void ring2nest(long nside, long ipring, long *ipnest)
{
  int ix, iy, face;
  if ((nside & (nside - 1)) != 0) { *ipnest = -1; return; }
  ring2xyf(nside, ipring, &ix, &iy, &face);
  *ipnest = xyf2nest(nside, ix, iy, face);
}
void nest2ring(long nside, long ipnest, long *ipring)
{
  int ix, iy, face;
  if ((nside & (nside - 1)) != 0) { *ipring = -1; return; }
  nest2xyf(nside, ipnest, &ix, &iy, &face);
  *ipring = xyf2ring(nside, ix, iy, face);
}
