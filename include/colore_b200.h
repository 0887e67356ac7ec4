/* colore_b200 -- C ABI of the B200-native density-field -> catalogue/maps path of CoLoRe.
 *
 * This header is the drop-in boundary: plain C, plain pointers and sizes. Each entry point
 * names the reference function(s) it replaces (file:line under damonge/CoLoRe src/). The
 * reference-side glue that maps `ParamCoLoRe *par` onto these calls is integration/colore_gpu_glue.c
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on failure; clr_last_error() describes the last
 *    failure of the calling thread's context. The glue maps failures to report_error(1,...)
 *    (common.c:290-306), i.e. the reference's "print and exit(1)" behaviour.
 *  - Host grids use the reference layout (fourier.c:46-51,328): real index
 *    ix + 2*(N/2+1)*(iy + N*iz_local), complex index kx + (N/2+1)*(ky + N*kz_local).
 *  - One context per process and per GPU (the reference is one `par` per MPI rank, common.c:216).
 *  - There is NO CPU fallback: every compute entry point fails if no CUDA device is present.
 */
#ifndef COLORE_B200_H
#define COLORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLR_NA 5001            /* common.h:132 */
#define CLR_NPOP_MAX 10        /* common.h:133 */
#define CLR_NPLANES_MAX 100    /* common.h:134 */

#define CLR_GRID_DENS 0        /* par->grid_dens / grid_dens_f (common.h:283-284) */
#define CLR_GRID_NPOT 1        /* par->grid_npot / grid_npot_f (common.h:285-286) */

#define CLR_DENS_TYPE_LGNR 0   /* common.h:115-118 */
#define CLR_DENS_TYPE_1LPT 1
#define CLR_DENS_TYPE_2LPT 2
#define CLR_DENS_TYPE_CLIP 3

typedef struct clr_ctx clr_ctx;

/* The scalar / table subset of ParamCoLoRe (common.h:220-381) that the path reads.
 * Tables are CLR_NA doubles, copied at clr_create time. */
typedef struct {
  int32_t n_grid;            /* common.h:273 */
  int32_t nz_here, iz0_here; /* common.h:275-276; init_fftw fourier.c:127-209 */
  int32_t dens_type;         /* common.h:267 */
  int32_t bias_model;        /* 1,2,3: _BIAS_MODEL_* of common.h:414-431 (compile-time in the reference) */
  int32_t do_smoothing;      /* common.h:265 */
  int32_t smooth_potential;  /* common.h:266 */
  int32_t nside_base;        /* common.h:378, io.c:224-244 */
  int32_t numk;              /* common.h:254 */
  uint32_t seed_rng;         /* common.h:271 */
  float l_box;               /* flouble, common.h:274 */
  float reserved_;
  double pos_obs[3];         /* common.h:294 */
  double r2_smooth;          /* common.h:264 */
  double prefac_lensing;     /* common.h:237 */
  double fgrowth_0, hubble_0, OmegaM, n_scal; /* common.h:228-236 */
  double r_max;              /* common.h:240 */
  double glob_idr;           /* common.h:251 */
  double logkmin, logkmax, idlogk; /* common.h:255-257 */
  const double *logkarr, *pkarr;   /* numk entries, common.h:258-259 */
  const double *r_arr_r2z, *z_arr_r2z, *growth_d_arr, *growth_d2_arr, *growth_v_arr,
               *growth_pd_arr, *ihub_arr;     /* common.h:244-250 */
  const double *a_arr_a2r, *r_arr_a2r;         /* common.h:242-243 */
} clr_params;

/* ---- lifetime ------------------------------------------------------------------------- */
int clr_version(void);
const char *clr_last_error(void);
/* number of visible CUDA devices (0 when none / driver missing); never fails */
int clr_device_count(void);
/* allocate_fftw + init_fftw (fourier.c:127-238): device grids for the slab, tables uploaded */
int clr_create(const clr_params *p, int device, clr_ctx **out);
/* end_fftw / param_colore_free (fourier.c:240-283, io.c:1247-1348) */
int clr_destroy(clr_ctx *ctx);
int clr_synchronize(clr_ctx *ctx);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
long long clr_launch_count(clr_ctx *ctx);

/* ---- multi-GPU (one process per GPU; replaces mpi_init common.c:216-274) ----------------- */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host (torch.distributed) */
int clr_comm_unique_id(void *id128);
int clr_comm_init(clr_ctx *ctx, int rank, int nranks, const void *id128);
/* 1 when the slab transpose of the distributed FFT runs as peer-memory stores fused into the producing pass
 * (every rank mapped every other rank's staging buffer over CUDA IPC), 0 when it runs as an NCCL all-to-all.
 * Option "p2p_fused" = 0 forces the NCCL path (clr_set_option). */
int clr_comm_p2p(clr_ctx *ctx);

/* ---- populations (cosmo.c:549-629 tables; one call per population) ---------------------- */
int clr_set_srcs(clr_ctx *ctx, int ipop, const double *nz_arr, const double *bz_arr);
int clr_set_imap(clr_ctx *ctx, int ipop, const double *tz_arr, const double *bz_arr,
                 int nside, int nr, const float *r0, const float *rf);

/* custom projected tracer (cstm.c): K(z) and b(z) tables of cosmo.c:631-717; takes part in the normalisation
 * (density.c:1177-1178) as kind 2 of clr_get_norm / clr_set_norm */
int clr_set_cstm(clr_ctx *ctx, int ipop, const double *kz_arr, const double *bz_arr);

/* ---- grids: host <-> device in the reference layout ------------------------------------- */
int clr_grid_put(clr_ctx *ctx, int which, const float *host_padded);  /* real or complex view */
int clr_grid_get(clr_ctx *ctx, int which, float *host_padded);
/* device pointer of a grid (for zero-copy callers that already live on the GPU) */
int clr_grid_device_ptr(clr_ctx *ctx, int which, void **dptr);
/* floats per row of the DEVICE grids: 2*ceil8(n_grid/2+1), rows 64-byte aligned (the host side of clr_grid_put / get is
 * the reference layout, 2*(n_grid/2+1) floats per row, fourier.c:46-51) */
int clr_grid_pitch(clr_ctx *ctx, long long *pitch_floats);

/* ---- Gaussian field (fourier.c) ---------------------------------------------------------- */
/* create_grids_fourier (fourier.c:285-359) with the counter-based RNG stream */
int clr_fill_modes(clr_ctx *ctx, uint32_t seed);
/* fftw_wrap_c2r / fftw_wrap_r2c (fourier.c:81-125), in place on a device grid */
int clr_fft_c2r(clr_ctx *ctx, int which);
int clr_fft_r2c(clr_ctx *ctx, int which);
/* normalisation loop + z-halo + compute_sigma_dens (fourier.c:381-416) on grids already in
 * real space; out2 = {mean, sigma2_gauss} */
int clr_normalize_fields(clr_ctx *ctx, double *out2);
/* the whole of create_cartesian_fields (fourier.c:361-423). inject=0: modes from clr_fill_modes
 * (seed); inject=1: Fourier modes already placed with clr_grid_put (the reference's own white
 * noise). out2 = {mean, sigma2_gauss}; sigma2 is also kept in the context. */
int clr_create_cartesian_fields(clr_ctx *ctx, uint32_t seed, int inject, double *out2);
int clr_set_sigma2_gauss(clr_ctx *ctx, double sigma2);
/* options: "exact_math" = 1 makes the streaming field kernels (mode fill, lognormal / clip, the
 * normalisation histogram) evaluate the reference's double-precision expressions verbatim
 * (fourier.c:337-353, density.c:1095-1098, 1166-1178); 0 (default) evaluates them in fp32 with
 * double only where it protects the result. Integer outputs (Poisson counts, pixel ids) are
 * exact in both modes. "lpt_interp_type" = 0/1/2 (NGP/CIC/TSC, field_par.lpt_interp_type),
 * "keep_particles" = 1 keeps the LPT particles resident for clr_lpt_get_particles.
 * "async_results" = 1 makes clr_srcs_get_local_properties return as soon as the device-to-host copy is
 * queued on a separate copy stream (the host buffer must be pinned and is valid after clr_synchronize);
 * the next run overlaps the copy and only waits for it before it reuses the catalogue buffers.
 * Kernel-path switches (all default to the fast path; the tests use them to compare one path against another):
 * "fill_fused" = 0: stand-alone mode fill instead of the fill fused into the z pass of the transforms (one GPU) /
 * into the peer-store z pass of the slab transpose (several GPUs); "fft_fused" = 0: three separate axis passes
 * instead of the fused y + x pass; "fill_w" = 4 / 8: kx lines per tile of the fused fill at n_grid = 1024;
 * "fill_cluster" = 1 / 0 / -1: fused fill + z pass on pairs of CTAs of a thread-block cluster (default -1: where one
 * SM cannot hold both fields' tiles, i.e. n_grid = 2048); "hist_fused" = 0: lognormal transform and normalisation
 * histogram as separate passes; "srcs_compact" = 0: dense per-cell counts; "los_precompute" = 0: kappa rays evaluate
 * the Hessian stencil per sample; "p2p_fused" = 0: NCCL all-to-all instead of peer-memory stores for the slab
 * transpose; "p2p_tiled" = 0 / 1: staging layout of that transpose; "fft_overlap" = 1: second transform pipeline. */
int clr_set_option(clr_ctx *ctx, const char *name, int value);
/* refresh the z-halo planes of the potential after clr_grid_put (fourier.c:401-414) */
int clr_update_halo(clr_ctx *ctx);

/* ---- physical density (density.c) -------------------------------------------------------- */
/* compute_physical_density_field (density.c:1105-1126): lognormal (lognormalize, 1070-1103), clipped
 * (densclip, 1034-1067), 1LPT (lpt_1, 376-644) and 2LPT (lpt_2, 646-1031) with the NGP / CIC / TSC mass
 * deposits (pos_2_*, 37-188; option "lpt_interp_type"). On several GPUs the particles whose deposit stencil
 * reaches another slab are exchanged over NCCL (share_particles, density.c:191-374), with buffers sized
 * from exact counts instead of field_par.lpt_buffer_fraction. */
int clr_compute_physical_density_field(clr_ctx *ctx);
/* particle positions of the last LPT call (for write_lpt, io.c:619-695); needs option
 * "keep_particles" = 1 before the density call. x, y, z: nz_here*n_grid^2 floats each. */
int clr_lpt_get_particles(clr_ctx *ctx, float *x, float *y, float *z);
/* particles this rank shipped to / received from other slabs in the last LPT density (0 on one GPU) */
int clr_lpt_exchange_counts(clr_ctx *ctx, long long *sent, long long *received);
/* compute_density_normalization (density.c:1227-1393). Afterwards the norm tables are resident;
 * clr_get_norm returns srcs (kind 0) / imap (kind 1) / custom (kind 2) tables: norm_arr[CLR_NA], ends[2] */
int clr_compute_density_normalization(clr_ctx *ctx);
int clr_get_norm(clr_ctx *ctx, int kind, int ipop, double *norm_arr, double *ends2, double *zends2);
int clr_set_norm(clr_ctx *ctx, int kind, int ipop, const double *norm_arr, const double *ends2);

/* ---- sources (srcs.c) -------------------------------------------------------------------- */
/* srcs_set_cartesian_single (srcs.c:120-283): Poisson counts, scan, positions/RSD/base pixel.
 * nsrc_out = sources found in this slab. */
int clr_srcs_set_cartesian(clr_ctx *ctx, int ipop, uint32_t seed, long long *nsrc_out);
/* per-cell counts of the last clr_srcs_set_cartesian, padded reference layout (srcs.c:125) */
int clr_srcs_get_counts(clr_ctx *ctx, int ipop, int32_t *nsources_padded);
/* CatalogCartesian (common.h:163-167): pos[4*n] floats, ipix[n] ints */
int clr_srcs_get_cartesian(clr_ctx *ctx, int ipop, float *pos4, int32_t *ipix);
/* srcs_get_local_properties_single (srcs.c:386-416): Src records, 9 floats each (common.h:169-179) */
int clr_srcs_get_local_properties(clr_ctx *ctx, int ipop, float *srcs9);
/* write_catalog (io.c:1019-1236) for the formats without lensing / skewers, straight from the device-resident Src
 * records: chunked device -> pinned host copies overlapped with multi-threaded formatting, one ordered file write.
 * format: CLR_FORMAT_ASCII ("%d %E %E %E %E \n" rows under io.c's header line) or CLR_FORMAT_FITS (BINTABLE TYPE 1J +
 * RA, DEC, Z_COSMO, DZ_RSD 1E, big endian). type_id = the population index io.c writes in column 1. n_threads <= 0: all
 * host cores. *seconds (may be NULL) receives the wall time. Same bytes as io.c for the same records. */
#define CLR_FORMAT_ASCII 0
#define CLR_FORMAT_FITS 1
int clr_write_catalog(clr_ctx *ctx, int ipop, const char *fname, int format, int type_id, int n_threads, double *seconds);
/* One HEALPix map file as he_write_healpix_map writes it (healpix_extra.c:4-57: BINTABLE column "map 1" 1E, ORDERING RING,
 * NSIDE, COORDSYS G), preceded -- when nadd / listpix are given -- by the shell loops of write_imap / write_kappa /
 * write_isw (io.c:697-1017): map[listpix[i]] += data[i], hits[listpix[i]] += nadd[i] over num_pix local pixels
 * (listpix NULL: pixel i, num_pix must then be 12 nside^2), map /= hits where hits > 0. isnest: the pixel indices are
 * NEST, the file is RING (he_nest2ring_inplace). The NEST -> RING gather and the big-endian conversion run on n_threads
 * host threads (<= 0: all cores). Host arrays in (data / nadd / listpix typed as in HealpixShells, common.h:207-218), no
 * device involved; a leading '!' of fname is skipped. */
int clr_write_healpix_map(const float *data, const int *nadd, const long *listpix, long long num_pix, long nside, int isnest,
                          const char *fname, int n_threads, double *seconds);
/* srcs_distribute_single (srcs.c:296-373), several GPUs: route every source to rank ipix % nranks, order preserved
 * (blocks received from rank-1, rank-2, ..., own sources last). beam_first != 0: evaluate the RSD-under-beaming
 * estimator (srcs.c:486-504) first, on the slab that holds the potential around each source, and carry it along.
 * Afterwards clr_srcs_get_cartesian / clr_srcs_get_local_properties return the redistributed catalogue. One GPU: no-op. */
int clr_srcs_distribute(clr_ctx *ctx, int ipop, int beam_first, long long *nsrc_out);
/* RSD under beaming: srcs_beams_preproc/get_beam_properties(lines 486-504)/postproc(656-662).
 * Updates dz_rsd (and e1=e2=0) of the resident catalogue; fetch with clr_srcs_get_local_properties */
int clr_srcs_beam_rsd(clr_ctx *ctx, int ipop);

/* The whole of srcs_beams_preproc / srcs_get_beam_properties / srcs_beams_postproc (srcs.c:425-744, default build
 * without _USE_FAST_LENSING): RSD under beaming as clr_srcs_beam_rsd, plus
 *   has_lensing: e1, e2, kappa, dra, ddec of every source from the NGP velocity / tidal stencils along its ray
 *                (srcs.c:531-614, 722-723); kappa / dra / ddec start from 0 (the reference accumulates into
 *                uninitialised my_malloc memory, common.c:375);
 *   has_skw:     density (skw_gauss = 0) or Gaussian-field (1, beaming.c:55-66) and radial-velocity skewers,
 *                nsrc x n_grid/2 samples (srcs.c:507-529, 725-733 including its overrun into the next skewer).
 * rsd_done != 0: dz_rsd already final (sources routed by clr_srcs_distribute with beam_first). Several GPUs: every
 * rank integrates its slab's part of every ray, the partial results are summed on the rank that holds the source.
 * Results: clr_srcs_get_local_properties (9-float Src records), clr_srcs_get_skewers. */
int clr_srcs_get_beam_properties(clr_ctx *ctx, int ipop, int has_lensing, int has_skw, int skw_gauss, int rsd_done);
/* Catalog.d_skw or g_skw, and v_skw (common.h:191-193): nsrc * (n_grid/2) floats each; either may be NULL */
int clr_srcs_get_skewers(clr_ctx *ctx, int ipop, float *dg_skw, float *v_skw);

/* ---- maps (imap.c, kappa.c, isw.c, cstm.c, beaming.c) --------------------------------------------- */
/* imap_set_cartesian_single (imap.c:135-245): data[nr*12*nside^2], nadd likewise (full sky) */
int clr_imap_set_cartesian(clr_ctx *ctx, int ipop, float *data, int32_t *nadd);
/* kappa_beams_preproc + kappa_get_beam_properties (kappa.c:39-175) for the pixels `pos`
 * (unit vectors, hp_shell_alloc common.c:505-552); rf sorted; data[nplanes*num_pix] */
int clr_kappa_get_beam_properties(clr_ctx *ctx, long long num_pix, const double *pos3, int nplanes,
                                  const float *rf, float *data);
/* isw_get_beam_properties (isw.c:78-147) */
int clr_isw_get_beam_properties(clr_ctx *ctx, long long num_pix, const double *pos3, int nplanes,
                                const float *rf, float *data);

/* cstm_beams_preproc + cstm_get_beam_properties (cstm.c:38-145): data[num_pix] = dr * sum_r K(r) (bias_model(delta_CIC,
 * b(r)) norm(r) - 1) for the pixels `pos` (unit vectors); several GPUs: summed over the slabs */
int clr_cstm_get_beam_properties(clr_ctx *ctx, int ipop, long long num_pix, const double *pos3, float *data);

/* Fast-lensing shells (reference builds with -D_USE_FAST_LENSING): lensing_beams_preproc + lensing_get_beam_properties
 * (lensing.c:39-250). nr_sh shells of radii r_sh (sorted ascending on entry, snapped to the radial sampling on exit,
 * lensing.c:233-236) with npp[ir] pixels per base pixel ("beam", hp_shell_adaptive_alloc common.c:452-504); pos3: unit
 * vectors of the finest shell, [nbeams][npp[nr_sh-1]][3]. data (may be NULL): shells concatenated, shell ir =
 * [nbeams][5 * npp[ir]] = {gamma1, gamma2, kappa, dx, dy} per pixel, summed over the slabs on several GPUs. The shells
 * also stay on the device for clr_srcs_lensing_from_shells. */
int clr_lensing_get_beam_properties(clr_ctx *ctx, int nbeams, int nr_sh, float *r_sh, const long long *npp,
                                    const double *pos3, float *data);
/* The lensing branch of srcs_beams_postproc under _USE_FAST_LENSING (srcs.c:666-723): e1, e2, kappa, dra, ddec of every
 * source interpolated in radius between the two shells that bracket it (including the reference's stride-2 read of the
 * upper shell, srcs.c:710-714). nside_sh: resolution of every shell; beam ib = base pixel ib * nnodes + node.
 * *n_bad: sources whose base pixel is not held (the reference stops with "Bad base"). */
int clr_srcs_lensing_from_shells(clr_ctx *ctx, int ipop, int nr_sh, const float *r_sh, const int32_t *nside_sh, int node,
                                 int nnodes, long long *n_bad);

/* ---- timing helpers for bench.py (CUDA events on the context's stream) -------------------- */
int clr_timer_start(clr_ctx *ctx);
int clr_timer_stop_ms(clr_ctx *ctx, float *ms);
/* last duration of a named kernel family, measured with CUDA events when profiling is on */
int clr_set_profiling(clr_ctx *ctx, int on);
int clr_get_stage_ms(clr_ctx *ctx, const char *stage, float *ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif /* COLORE_B200_H */
