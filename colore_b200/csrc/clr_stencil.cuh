// Finite-difference stencils of the potential and grid-point lookups shared by the line-of-sight kernels
// (clr_maps.cu: kappa / ISW, clr_srcs.cu: RSD under beaming, clr_beam.cu: per-source lensing, skewers, custom maps).
// Restates get_element / interpolate_from_grid of beaming.c:31-268 for device grids with the internal row pitch.
#pragma once
#include "clr_internal.cuh"

// beaming.c:85-116: Hessian stencil of the potential at cell (ix,iy,iz_local), unnormalised
__device__ __forceinline__ void dev_tidal(const ClrDev &d, const float *__restrict__ g, int ix, int iy, int iz, float t[6])
{
  const long long ngx = d.pitch, plane = ngx * d.n;
  long long x0 = ix, xh = ix + 1 == d.n ? 0 : ix + 1, xl = ix == 0 ? d.n - 1 : ix - 1;
  long long y0 = (long long)iy * ngx, yh = (long long)(iy + 1 == d.n ? 0 : iy + 1) * ngx, yl = (long long)(iy == 0 ? d.n - 1 : iy - 1) * ngx;
  long long z0 = iz * plane;
  long long zh = ((iz == d.nz_here - 1) ? (long long)(d.nz_here + 1) : iz + 1) * plane;
  long long zl = ((iz == 0) ? (long long)d.nz_here : iz - 1) * plane;
  float c = g[x0 + y0 + z0];
  t[0] = (g[xh + y0 + z0] + g[xl + y0 + z0] - 2 * c);
  t[3] = (g[x0 + yh + z0] + g[x0 + yl + z0] - 2 * c);
  t[1] = (float)(0.25 * (double)(g[xh + yh + z0] + g[xl + yl + z0] - g[xh + yl + z0] - g[xl + yh + z0]));
  t[5] = (g[x0 + y0 + zh] + g[x0 + y0 + zl] - 2 * c);
  // the reference pairs the terms differently at the slab edges (beaming.c:92-107): same operands,
  // same left-to-right order hi,lo,-,- so one expression serves the three branches
  t[2] = (float)(0.25 * (double)(g[xh + y0 + zh] + g[xl + y0 + zl] - g[xh + y0 + zl] - g[xl + y0 + zh]));
  t[4] = (float)(0.25 * (double)(g[x0 + yh + zh] + g[x0 + yl + zl] - g[x0 + yh + zl] - g[x0 + yl + zh]));
}

// NGP cell of a sample (beaming.c:148-157); returns false when the plane is not in this slab
__device__ __forceinline__ bool dev_ngp(const ClrDev &d, const double xn[3], int c[3])
{
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    long v = (long)(xn[ax] + 0.5);
    if (v >= d.n) v -= d.n; else if (v < 0) v += d.n;
    c[ax] = (int)v;
  }
  c[2] -= d.iz0_here;
  return c[2] >= 0 && c[2] < d.nz_here;
}

// Plane index of LOCAL plane iz in [-2, nz_here+1]: the slab, then the halo planes stored behind it (clr_api.cu:
// [nz] = -1, [nz+1] = nz, [nz+2] = -2, [nz+3] = nz+1). On a single slab iz is periodic and always inside.
__device__ __forceinline__ long long dev_plane_index(const ClrDev &d, int iz)
{
  if (iz >= 0 && iz < d.nz_here) return iz;
  if (iz == -1) return d.nz_here;
  if (iz == d.nz_here) return d.nz_here + 1;
  if (iz == -2) return d.nz_here + 2;
  return d.nz_here + 3;
}
__device__ __forceinline__ void dev_vel_element(const ClrDev &d, const float *__restrict__ npot, int ix, int iy, int iz, bool whole_box,
                                                float v[3])
{
  const long long ngx = d.pitch, plane = ngx * d.n;
  int ix_hi = ix + 1 == d.n ? 0 : ix + 1, ix_lo = ix == 0 ? d.n - 1 : ix - 1;
  int iy_hi = iy + 1 == d.n ? 0 : iy + 1, iy_lo = iy == 0 ? d.n - 1 : iy - 1;
  long long pz, pz_hi, pz_lo;
  if (whole_box) {                               // one slab = the periodic box
    pz = iz;
    pz_hi = iz + 1 == d.n ? 0 : iz + 1;
    pz_lo = iz == 0 ? d.n - 1 : iz - 1;
  } else {
    pz = dev_plane_index(d, iz); pz_hi = dev_plane_index(d, iz + 1); pz_lo = dev_plane_index(d, iz - 1);
  }
  v[0] = npot[ix_hi + iy * ngx + pz * plane] - npot[ix_lo + iy * ngx + pz * plane];
  v[1] = npot[ix + iy_hi * ngx + pz * plane] - npot[ix + iy_lo * ngx + pz * plane];
  v[2] = npot[ix + iy * ngx + pz_hi * plane] - npot[ix + iy * ngx + pz_lo * plane];
}

