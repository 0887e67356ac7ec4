// fp32 Gaussian-mode arithmetic shared by the stand-alone fill (clr_fields.cu) and the fill fused into the z pass of
// the c2r (clr_fft.cu). create_grids_fourier (fourier.c:285-359) + rng_delta_gauss (common.c:191-201) + pk_linear0
// (cosmo.c:291-308), same draws and formulae as the double kernel, but:
//  * k^2/dk^2 = m is an integer: the P(k) table position of k = dk*sqrt(m) is
//    t = T[e] + c1*log2(f), m = f*2^e, with T[e] tabulated in double on the host and split into an
//    integer and a fraction, so the fp32 sum only ever carries the position INSIDE a few table bins;
//  * P(k)/dk^3 comes from an fp32 lerp table {p_i, p_{i+1}-p_i};
//  * the phase 2*pi*u1 (25 bits) = coarse angle from a 4096-entry table x a small-angle rotation;
//  * row invariants are hoisted: no per-mode 64-bit or double arithmetic is left.
#pragma once
#include "clr_internal.cuh"

struct FillFastK {
  int e_int[26];
  float e_frac[26];
  float c1, nscal_c, m3_c, tmax, p_first, p_last, neg_prefac_idk2, smooth_c, smooth_c2;
  int numk, do_smoothing, smooth_potential;
};
// host: constants + fp32 tables (c->d_pkt, c->d_sincos) of the fast fill; defined in clr_fields.cu
int clr_fill_fast_setup(clr_ctx *c, FillFastK *k);
// the fp32 fill needs k^2/dk^2 < 2^24 (an exact integer in a float) and exact_math = 0
bool clr_fill_fast_ok(const clr_ctx *c);

#ifdef __CUDACC__
// ln(1 - w*2^-32) for a 32-bit uniform word, relative error <~ 1.3e-6 (3e-7 typical), no branches:
//  * u < 1/8: the series -u(1 + u/2 + ... + u^7/8) (truncation 6e-8 relative);
//  * otherwise T = 2^32 - w is normalised to [0.5,1) (exact shift), ln = ln2*(lg2(T') - shift): lg2.approx carries
//    an absolute error of 1.6e-7, which is < 1.3e-6 relative once |ln| >= ln(8/7).
__device__ __forceinline__ float clr_log_one_minus_u(uint32_t w)
{
  const float u = __uint2float_rn(w) * 2.3283064365386963e-10f;
  float sr = fmaf(u, 0.125f, 0.14285715f);
  sr = fmaf(sr, u, 0.16666667f); sr = fmaf(sr, u, 0.2f); sr = fmaf(sr, u, 0.25f);
  sr = fmaf(sr, u, 0.33333334f); sr = fmaf(sr, u, 0.5f); sr = fmaf(sr, u, 1.f);
  const uint32_t T = 0u - w;
  const int cz = __clz(T);                      // w = 0 -> T = 0 -> cz = 32: the series branch is taken anyway
  const float Tf = __uint2float_rn(T << (cz & 31)) * 2.3283064365386963e-10f;
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(Tf));
  const float ll = 0.69314718f * (lg - (__int_as_float(0x4B000000 | cz) - 8388608.f));
  return w < 0x20000000u ? -u * sr : ll;
}

// One mode: m = kx^2 + ky^2 + kz^2 in units of dk^2, draws {phase word, modulus word}; `live` = false (k = 0 or a
// padding column) gives zeros. Branch free apart from the (never taken at usual sizes) power-law ends of the P(k) table.
// Every multiply-add is an explicit fmaf / __fmul_rn, so the result does not depend on the FMA-contraction setting of the
// translation unit (clr_fields.cu is built with -fmad=false, clr_fft.cu is not).
__device__ __forceinline__ void clr_fill_mode(const FillFastK &k, const float2 *__restrict__ pkt, const float2 *__restrict__ sct,
                                              int m, bool live, uint32_t w_phase, uint32_t w_mod, float2 &dk_out, float2 &pk_out)
{
  m = max(m, 1);
  const int e2 = 31 - __clz(m);
  const float mf = __int2float_rn(m);
  const float fm = __fmul_rn(mf, __int_as_float((127 - e2) << 23));     // m * 2^-e2 in [1,2)
  float lgm;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lgm) : "f"(fm));
  const float par = fmaf(k.c1, lgm, k.e_frac[e2]);                       // >= 0
  const float pm = clr_floor_magic(par);
  const float fl = pm - 8388608.f;
  const int ik = k.e_int[e2] + clr_magic_int(pm);
  const float2 t = __ldg(pkt + min(max(ik, 0), k.numk - 1));
  float sigma2 = fmaf(par - fl, t.y, t.x);
  if (ik < 0 || ik >= k.numk) {                                           // cosmo.c:295-304: k^ns below, k^-3 above the table
    const float tf = (float)k.e_int[e2] + par;
    sigma2 = ik < 0 ? k.p_first * exp10f(k.nscal_c * tf) : k.p_last * exp10f(k.m3_c * (tf - k.tmax));
  }
  const float delta_mod = clr_sqrt_fast(__fmul_rn(-sigma2, clr_log_one_minus_u(w_mod)));
  // phase = 2*pi*q/2^25, q = hi*2^13 + lo: table entry of the coarse angle, rotated by the small angle a < 1.6e-3
  // (sin a = a and cos a = 1 - a^2/2 to 7e-10 and 3e-13)
  const uint32_t q = w_phase >> 7;
  const float2 cs_h = __ldg(sct + (q >> 13));
  const float a = __fmul_rn(__int_as_float(0x4B000000 | (q & 8191u)) - 8388608.f, 1.872535141e-07f);   // 2*pi/2^25
  const float ca1 = __fmul_rn(__fmul_rn(0.5f, a), a);
  const float cs = fmaf(-cs_h.y, a, fmaf(-cs_h.x, ca1, cs_h.x));
  const float sn = fmaf(cs_h.x, a, fmaf(-cs_h.y, ca1, cs_h.y));
  const float amp = live ? delta_mod : 0.f;
  float dre = __fmul_rn(amp, cs), dim = __fmul_rn(amp, sn);
  const float pk2 = __fmul_rn(k.neg_prefac_idk2, clr_rcp_fast(mf));
  float pre = __fmul_rn(pk2, dre), pim = __fmul_rn(pk2, dim);
  if (k.do_smoothing) {
    const float sm = clr_ex2_fast(__fmul_rn(k.smooth_c2, mf));
    dre = __fmul_rn(dre, sm); dim = __fmul_rn(dim, sm);
    if (k.smooth_potential) { pre = __fmul_rn(pre, sm); pim = __fmul_rn(pim, sm); }
  }
  dk_out = make_float2(dre, dim);
  pk_out = make_float2(pre, pim);
}
#endif
