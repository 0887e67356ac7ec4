// Catalogue writer straight from the device-resident Src records: replaces the serial loops of write_catalog
// (io.c:1019-1236) for the ASCII and FITS formats without lensing / skewers (the populations the GPU path produces).
//
// Once the field -> catalogue path runs in tens of milliseconds, the reference's writer -- one fprintf per source, or
// 8192-row cfitsio column writes, all on one thread -- is >98 % of the wall time of a run (the reference's own timers at
// 512^3: 2.2 s of 4.2 s). Here the records leave the GPU in chunks through two pinned buffers (the copy of chunk k+1
// runs while chunk k is formatted), every chunk is formatted by all host threads into per-thread buffers, and one
// thread streams the buffers to the file in order. The bytes are the ones io.c produces:
//   ASCII (io.c:1208-1232): "#[1]type [2]RA, [3]dec, [4]z0, [5]dz_RSD \n", then "%d %E %E %E %E \n" per source
//         (formatted with the C library's own printf, so the digits cannot differ);
//   FITS  (io.c:1075-1120): empty primary HDU + one BINTABLE, columns TYPE 1J, RA / DEC / Z_COSMO / DZ_RSD 1E,
//         big-endian rows of 20 bytes, keyword CONTENTS = 'Source catalog'.
//
// HEALPix map writer (clr_write_healpix_map): replaces the per-shell loops of write_imap / write_kappa / write_isw
// (io.c:697-1017: scatter by listpix, divide by nadd) + he_write_healpix_map (healpix_extra.c:4-57: NEST -> RING
// reordering pixel by pixel, float conversion, one cfitsio column write) -- ~1 s per nside-1024 map on one thread,
// against ~40 ms for the rays that fill it. Here the RING-ordered big-endian column is gathered by all host threads.
#include "clr_internal.cuh"
#include <string.h>
#include <math.h>
#include <thread>
#include <memory>
#include <chrono>
#include <functional>
#include <algorithm>

namespace {

constexpr long long kChunk = 1 << 20;        // sources per device -> host chunk (36 MB of Src records)

void put_card(std::string &h, const char *txt)
{
  char buf[81];
  snprintf(buf, sizeof(buf), "%-80s", txt);
  h.append(buf, 80);
}
void pad_block(std::string &h, char fill)
{
  while (h.size() % 2880) h.push_back(fill);
}

// header of the file up to the first table row (io.c:1084-1088: fits_create_file, fits_create_tbl, fits_update_key)
std::string fits_header(long long nrows)
{
  std::string h;
  put_card(h, "SIMPLE  =                    T / file does conform to FITS standard");
  put_card(h, "BITPIX  =                    8 / number of bits per data pixel");
  put_card(h, "NAXIS   =                    0 / number of data axes");
  put_card(h, "EXTEND  =                    T / FITS dataset may contain extensions");
  put_card(h, "END");
  pad_block(h, ' ');
  char buf[128];
  put_card(h, "XTENSION= 'BINTABLE'           / binary table extension");
  put_card(h, "BITPIX  =                    8 / 8-bit bytes");
  put_card(h, "NAXIS   =                    2 / 2-dimensional binary table");
  snprintf(buf, sizeof(buf), "NAXIS1  = %20d / width of table in bytes", 20); put_card(h, buf);
  snprintf(buf, sizeof(buf), "NAXIS2  = %20lld / number of rows in table", nrows); put_card(h, buf);
  put_card(h, "PCOUNT  =                    0 / size of special data area");
  put_card(h, "GCOUNT  =                    1 / one data group");
  snprintf(buf, sizeof(buf), "TFIELDS = %20d / number of fields in each row", 5); put_card(h, buf);
  const char *ttype[5] = {"TYPE", "RA", "DEC", "Z_COSMO", "DZ_RSD"};
  const char *tunit[5] = {"NA", "DEG", "DEG", "NA", "NA"};
  for (int i = 0; i < 5; i++) {
    snprintf(buf, sizeof(buf), "TTYPE%-3d= '%-8s'", i + 1, ttype[i]); put_card(h, buf);
    snprintf(buf, sizeof(buf), "TFORM%-3d= '%d%c      '", i + 1, 1, i == 0 ? 'J' : 'E'); put_card(h, buf);
    snprintf(buf, sizeof(buf), "TUNIT%-3d= '%-8s'", i + 1, tunit[i]); put_card(h, buf);
  }
  snprintf(buf, sizeof(buf), "%-8.8s= '%-8s' / %s", "CONTENTS", "Source catalog", ""); put_card(h, buf);
  put_card(h, "END");
  pad_block(h, ' ');
  return h;
}

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

// format rows [r0, r1) of a chunk of Src records (9 floats each) into `out`
void format_rows(const float *srcs, long long r0, long long r1, int format, int type_id, std::string &out)
{
  out.clear();
  if (format == CLR_FORMAT_ASCII) {
    out.reserve((size_t)(r1 - r0) * 60);
    char line[128];
    for (long long i = r0; i < r1; i++) {
      const float *s = srcs + 9 * i;
      int n = snprintf(line, sizeof(line), "%d %E %E %E %E \n", type_id, s[0], s[1], s[2], s[3]);
      out.append(line, (size_t)n);
    }
  } else {
    out.resize((size_t)(r1 - r0) * 20);
    uint32_t *o = reinterpret_cast<uint32_t *>(&out[0]);
    const uint32_t t_be = bswap32((uint32_t)type_id);
    for (long long i = r0; i < r1; i++) {
      const uint32_t *s = reinterpret_cast<const uint32_t *>(srcs + 9 * i);
      uint32_t row[5] = {t_be, bswap32(s[0]), bswap32(s[1]), bswap32(s[2]), bswap32(s[3])};
      memcpy(o, row, 20);                       // rows are 20 bytes: keep the stores byte-wise (no alignment assumption)
      o += 5;
    }
  }
}

}  // namespace

namespace {

// ring index -> NEST index (the published HEALPix indexing, Gorski et al. 2005: ring -> (x, y, face) -> interleaved bits)
inline uint64_t spread_bits(uint32_t v)
{
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffULL;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffULL;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | (x << 2)) & 0x3333333333333333ULL;
  x = (x | (x << 1)) & 0x5555555555555555ULL;
  return x;
}
inline long long isqrt_ll(long long v)
{
  long long r = (long long)sqrt((double)v + 0.5);
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return r;
}
long long ring2nest_host(long long nside, long long pix)
{
  static const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4}, jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  const long long ncap = 2 * nside * (nside - 1), npix = 12 * nside * nside, nl2 = 2 * nside;
  long long iring, iphi, kshift, nr;
  int face;
  if (pix < ncap) {                                    // north polar cap
    iring = (1 + isqrt_ll(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0; nr = iring;
    face = (int)((iphi - 1) / nr);
  } else if (pix < npix - ncap) {                      // equatorial belt
    const long long ip = pix - ncap, tmp = ip / (4 * nside);
    iring = tmp + nside;
    iphi = ip - tmp * 4 * nside + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    const long long ire = tmp + 1, irm = nl2 + 2 - ire;
    const long long ifm = (iphi - ire / 2 + nside - 1) / nside, ifp = (iphi - irm / 2 + nside - 1) / nside;
    face = (int)(ifp == ifm ? (ifp | 4) : (ifp < ifm ? ifp : ifm + 8));
  } else {                                             // south polar cap
    const long long ip = npix - pix;
    iring = (1 + isqrt_ll(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0; nr = iring;
    iring = 2 * nl2 - iring;
    face = 8 + (int)((iphi - 1) / nr);
  }
  const long long irt = iring - jrll[face] * nside + 1;
  long long ipt = 2 * iphi - jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  const long long ix = (ipt - irt) >> 1, iy = (-ipt - irt) >> 1;
  return (long long)face * nside * nside + (long long)(spread_bits((uint32_t)ix) | (spread_bits((uint32_t)iy) << 1));
}

// header of he_write_healpix_map's file up to the first table row: empty primary HDU + BINTABLE with one 1E column
std::string healpix_header(long long npix, long nside)
{
  std::string h;
  put_card(h, "SIMPLE  =                    T / file does conform to FITS standard");
  put_card(h, "BITPIX  =                    8 / number of bits per data pixel");
  put_card(h, "NAXIS   =                    0 / number of data axes");
  put_card(h, "EXTEND  =                    T / FITS dataset may contain extensions");
  put_card(h, "END");
  pad_block(h, ' ');
  char buf[160];
  put_card(h, "XTENSION= 'BINTABLE'           / binary table extension");
  put_card(h, "BITPIX  =                    8 / 8-bit bytes");
  put_card(h, "NAXIS   =                    2 / 2-dimensional binary table");
  snprintf(buf, sizeof(buf), "NAXIS1  = %20d / width of table in bytes", 4); put_card(h, buf);
  snprintf(buf, sizeof(buf), "NAXIS2  = %20lld / number of rows in table", npix); put_card(h, buf);
  put_card(h, "PCOUNT  =                    0 / size of special data area");
  put_card(h, "GCOUNT  =                    1 / one data group");
  snprintf(buf, sizeof(buf), "TFIELDS = %20d / number of fields in each row", 1); put_card(h, buf);
  snprintf(buf, sizeof(buf), "TTYPE%-3d= '%-8s'", 1, "map 1"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "TFORM%-3d= '%d%c      '", 1, 1, 'E'); put_card(h, buf);
  snprintf(buf, sizeof(buf), "TUNIT%-3d= '%-8s'", 1, "uK"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "EXTNAME = '%-8s'", "BINTABLE"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "%-8.8s= '%-8s' / %s", "PIXTYPE", "HEALPIX", "HEALPIX Pixelisation"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "%-8.8s= '%-8s' / %s", "ORDERING", "RING", "Pixel ordering scheme, either RING or NESTED"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "%-8.8s= %20ld / %s", "NSIDE", nside, "Resolution parameter for HEALPIX"); put_card(h, buf);
  snprintf(buf, sizeof(buf), "%-8.8s= '%-8s' / %s", "COORDSYS", "G", "Pixelisation coordinate system"); put_card(h, buf);
  put_card(h, "COMMENT G = Galactic, E = ecliptic, C = celestial = equatorial");
  put_card(h, "END");
  pad_block(h, ' ');
  return h;
}

}  // namespace

// One HEALPix map as he_write_healpix_map (healpix_extra.c:4-57) writes it, with the shell loops of write_kappa /
// write_isw / write_imap (io.c:697-1017) in front when `nadd` / `listpix` are given:
//   map[listpix[i]] += data[i], hits[listpix[i]] += nadd[i] over the num_pix local pixels (listpix NULL: i itself),
//   map[p] /= hits[p] where hits[p] > 0, then -- isnest -- NEST -> RING, float32 big endian, one BINTABLE column.
// A leading '!' of fname (cfitsio: overwrite) is skipped. No device is involved: the maps are host arrays at the boundary.
extern "C" int clr_write_healpix_map(const float *data, const int *nadd, const long *listpix, long long num_pix, long nside,
                                     int isnest, const char *fname, int n_threads, double *seconds)
{
  CLR_CHECK(nside > 0 && (nside & (nside - 1)) == 0 && nside <= (1L << 15), "HEALPix nside %ld is not a power of two <= 32768", nside);
  CLR_CHECK(data && fname && num_pix >= 0, "clr_write_healpix_map: bad arguments");
  const long long npix = 12LL * nside * nside;
  CLR_CHECK(listpix || num_pix == npix, "clr_write_healpix_map: %lld pixels without a pixel list (the map has %lld)", num_pix, npix);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if (n_threads > 256) n_threads = 256;
  auto t_start = std::chrono::steady_clock::now();
  if (fname[0] == '!') fname++;
  // ---- the loops of io.c:820-845 (serial: a pixel list may name a pixel twice, the sums must not race)
  const float *map = data;
  std::vector<float> acc;
  if (nadd || listpix) {
    acc.assign((size_t)npix, 0.f);
    std::vector<int> hits(nadd ? (size_t)npix : 0, 0);
    for (long long i = 0; i < num_pix; i++) {
      const long long p = listpix ? listpix[i] : i;
      CLR_CHECK(p >= 0 && p < npix, "clr_write_healpix_map: pixel %lld outside the map", p);
      acc[(size_t)p] += data[i];
      if (nadd) hits[(size_t)p] += nadd[i];
    }
    if (nadd)
      for (long long p = 0; p < npix; p++) if (hits[(size_t)p] > 0) acc[(size_t)p] /= hits[(size_t)p];
    map = acc.data();
  }
  // ---- RING-ordered big-endian column, gathered by all threads
  std::unique_ptr<uint32_t[]> col(new uint32_t[(size_t)npix]);       // (not zero-filled: every entry is written below)
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(map);
    uint32_t *dst = col.get();
    auto work = [=](long long p0, long long p1) {
      if (isnest) for (long long p = p0; p < p1; p++) dst[p] = bswap32(src[ring2nest_host(nside, p)]);
      else for (long long p = p0; p < p1; p++) dst[p] = bswap32(src[p]);
    };
    std::vector<std::thread> th;
    const long long per = (npix + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
      const long long p0 = std::min(npix, t * per), p1 = std::min(npix, p0 + per);
      if (p0 < p1) th.emplace_back(work, p0, p1);
    }
    for (auto &x : th) x.join();
  }
  FILE *f = fopen(fname, "wb");
  CLR_CHECK(f, "Couldn't open file %s", fname);                  // common.c:58-62 error_open_file
  setvbuf(f, nullptr, _IOFBF, 8 << 20);
  int rc = 0;
  const std::string h = healpix_header(npix, nside);
  if (fwrite(h.data(), 1, h.size(), f) != h.size() || fwrite(col.get(), 4, (size_t)npix, f) != (size_t)npix) rc = 1;
  const long long r = (npix * 4) % 2880;
  if (!rc && r) { std::string z((size_t)(2880 - r), '\0'); if (fwrite(z.data(), 1, z.size(), f) != z.size()) rc = 1; }
  if (fclose(f) != 0) rc = 1;
  if (rc) clr_set_error("write error on %s", fname);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  return rc;
}

extern "C" int clr_write_catalog(clr_ctx *c, int ipop, const char *fname, int format, int type_id, int n_threads,
                                 double *seconds)
{
  CLR_CHECK(ipop >= 0 && ipop < CLR_NPOP_MAX && c->srcs[ipop].set, "population index %d out of range", ipop);
  CLR_CHECK(format == CLR_FORMAT_ASCII || format == CLR_FORMAT_FITS, "catalogue format %d not supported (0 ASCII, 1 FITS)", format);
  clr_ctx::Pop &P = c->srcs[ipop];
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  if (n_threads > 256) n_threads = 256;
  cudaEvent_t e0, e1;
  CLR_CUDA(cudaEventCreate(&e0)); CLR_CUDA(cudaEventCreate(&e1));
  auto t_start = std::chrono::steady_clock::now();
  CLR_CUDA(cudaStreamSynchronize(c->stream));                    // the Src records are final
  if (c->copy_pending) { CLR_CUDA(cudaStreamSynchronize(c->copy_stream)); c->copy_pending = false; }
  FILE *f = fopen(fname, "wb");
  CLR_CHECK(f, "Couldn't open file %s", fname);                  // common.c:58-62 error_open_file
  setvbuf(f, nullptr, _IOFBF, 8 << 20);
  const long long n = P.nsrc;
  int rc = 1;
  float *pin[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  do {
    if (format == CLR_FORMAT_ASCII) {
      const char *hdr = "#[1]type [2]RA, [3]dec, [4]z0, [5]dz_RSD \n";
      if (fwrite(hdr, 1, strlen(hdr), f) != strlen(hdr)) { clr_set_error("write error on %s", fname); break; }
    } else {
      std::string h = fits_header(n);
      if (fwrite(h.data(), 1, h.size(), f) != h.size()) { clr_set_error("write error on %s", fname); break; }
    }
    const long long n_chunks = (n + kChunk - 1) / kChunk;
    bool bad = false;
    for (int b = 0; b < 2 && !bad; b++) {
      if (cudaHostAlloc(&pin[b], (size_t)std::min(n > 0 ? n : 1, kChunk) * 9 * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess) {
        clr_set_error("clr_write_catalog: cannot allocate the pinned staging buffers");
        bad = true;
      }
    }
    if (bad) break;
    auto issue = [&](long long ch) {
      const long long r0 = ch * kChunk, cnt = std::min(kChunk, n - r0);
      cudaMemcpyAsync(pin[ch & 1], P.d_srcs + 9 * r0, (size_t)cnt * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream);
      cudaEventRecord(done[ch & 1], c->copy_stream);
    };
    if (n_chunks > 0) issue(0);
    std::vector<std::string> bufs(n_threads);
    long long bytes = 0;
    for (long long ch = 0; ch < n_chunks && !bad; ch++) {
      const long long cnt = std::min(kChunk, n - ch * kChunk);
      if (cudaEventSynchronize(done[ch & 1]) != cudaSuccess) { clr_set_error("clr_write_catalog: device -> host copy failed"); bad = true; break; }
      if (ch + 1 < n_chunks) issue(ch + 1);                        // flies under the formatting of this chunk
      const float *rows = pin[ch & 1];
      std::vector<std::thread> th;
      const long long per = (cnt + n_threads - 1) / n_threads;
      for (int t = 0; t < n_threads; t++) {
        const long long r0 = std::min(cnt, t * per), r1 = std::min(cnt, r0 + per);
        th.emplace_back(format_rows, rows, r0, r1, format, type_id, std::ref(bufs[t]));
      }
      for (auto &x : th) x.join();
      for (int t = 0; t < n_threads; t++) {
        if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) { clr_set_error("write error on %s", fname); bad = true; break; }
        bytes += (long long)bufs[t].size();
      }
    }
    if (bad) break;
    if (format == CLR_FORMAT_FITS) {                               // data area padded with zeros to whole 2880-byte blocks
      const long long r = (n * 20) % 2880;
      if (r) { std::string z((size_t)(2880 - r), '\0'); if (fwrite(z.data(), 1, z.size(), f) != z.size()) { clr_set_error("write error on %s", fname); break; } }
    }
    rc = 0;
  } while (0);
  if (fclose(f) != 0 && rc == 0) { clr_set_error("write error on %s", fname); rc = 1; }
  for (int b = 0; b < 2; b++) { if (pin[b]) cudaFreeHost(pin[b]); if (done[b]) cudaEventDestroy(done[b]); }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  return rc;
}
