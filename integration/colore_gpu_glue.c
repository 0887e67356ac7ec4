/* Reference-side binding of the colore_b200 C ABI.
 *
 * This ONE translation unit replaces fourier.c, density.c, srcs.c, imap.c, kappa.c, isw.c, cstm.c and
 * beaming.c of damonge/CoLoRe in the link line: it defines exactly the functions those files export
 * through common.h (464-543) and forwards them to include/colore_b200.h. main.c, io.c, cosmo.c,
 * cosmo_mad.c, common.c, healpix_extra.c, predictions.c, fftlog.c are linked UNCHANGED, so `./CoLoRe param.cfg`
 * stays the entry point (main.c:24-154). lensing.c: linked unchanged in the default build (its hooks are never
 * called); with -D_USE_FAST_LENSING (make FASTLENS=1) it is left out and its hooks are defined here as well.
 *
 * It is compiled against the reference's own common.h (-I<reference>/src); it contains no code of
 * the reference. One process drives one GPU; COLORE_B200_NGPUS=P makes the executable fork into P ranks (see
 * launch_ranks below), the slab decomposition itself lives below the C ABI (clr_comm_init).
 *
 * Host memory contract: par->cats_c / par->cats / shell data and nadd are
 * host-malloc'ed with the reference's own allocators because io.c writes and frees them;
 * par->grid_dens / grid_npot get host storage only when output_density asks for a dump.
 */
#include "common.h"
#include "colore_b200.h"
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#ifndef _SPREC
#error "colore_gpu_glue.c: the GPU grids are fp32 -- build the drop-in with -D_SPREC (flouble must be float)"
#endif

static clr_ctx *g_ctx = NULL;

/* ------------------------------------------------------------------ several GPUs without MPI
 * COLORE_B200_NGPUS=P ./CoLoRe_b200 param.cfg: the process forks into P ranks BEFORE main runs (nothing of CUDA is
 * initialised yet); rank r drives GPU r and owns the z slab [r n/P, (r+1) n/P) (fourier.c:172-181). This replaces
 * mpi_init (common.c:216-274): the unchanged host code sees NodeThis = r, NNodes = P from the first stage call on, so
 * io.c writes one catalogue / density file per rank exactly like an MPI run, and rank 0 writes the maps, which the
 * library returns already summed over the slabs. While read_run_params runs, NodeThis / NNodes still say (0, 1), so
 * that every rank allocates FULL-sky shells (hp_shell_alloc, common.c:505-552): the GPUs integrate every pixel over
 * their own slab and all-reduce, instead of the reference's pixel ownership + slab rotation (beaming.c:293-374).
 * The 128-byte NCCL id travels from rank 0 to the others through a file. */
static int g_rank = 0, g_nranks = 1;
static pid_t g_children[64];
static char g_idfile[256];

static void wait_for_ranks(void)
{
  int r, status;
  for (r = 1; r < g_nranks; r++)
    if (g_children[r] > 0) waitpid(g_children[r], &status, 0);
  unlink(g_idfile);
}

__attribute__((constructor)) static void launch_ranks(void)
{
  const char *s = getenv("COLORE_B200_NGPUS");
  int r, P = s ? atoi(s) : 1;
  if (P <= 1) return;
  if (P > 16) P = 16;
  g_nranks = P;
  snprintf(g_idfile, sizeof(g_idfile), "/tmp/colore_b200_nccl_id_%ld", (long)getpid());
  unlink(g_idfile);
  fflush(NULL);
  for (r = 1; r < P; r++) {
    pid_t pid = fork();
    if (pid == 0) {                       /* rank r: quiet stdout (print_info only gates on NodeThis later) */
      g_rank = r;
      if (!freopen("/dev/null", "w", stdout)) exit(1);
      return;
    }
    g_children[r] = pid;
  }
  atexit(wait_for_ranks);
}

static void exchange_nccl_id(unsigned char id[128])
{
  if (g_rank == 0) {
    char tmpn[300];
    FILE *f;
    snprintf(tmpn, sizeof(tmpn), "%s.tmp", g_idfile);
    f = fopen(tmpn, "wb");
    if (!f || fwrite(id, 1, 128, f) != 128) report_error(1, "cannot write %s\n", tmpn);
    fclose(f);
    if (rename(tmpn, g_idfile)) report_error(1, "cannot publish %s\n", g_idfile);
  } else {
    int tries;
    for (tries = 0; tries < 6000; tries++) {          /* up to 60 s */
      FILE *f = fopen(g_idfile, "rb");
      if (f) {
        size_t got = fread(id, 1, 128, f);
        fclose(f);
        if (got == 128) return;
      }
      usleep(10000);
    }
    report_error(1, "rank %d: no NCCL id from rank 0 (%s)\n", g_rank, g_idfile);
  }
}

static void chk(int status)
{
  if (status) report_error(1, "colore_b200: %s\n", clr_last_error());
}

/* first stage call after read_run_params: from here on the host code sees the rank layout */
static void adopt_rank_layout(void)
{
  if (g_nranks == 1 || NNodes == g_nranks) return;
  NodeThis = g_rank; NNodes = g_nranks;
  NodeLeft = (g_rank + g_nranks - 1) % g_nranks;
  NodeRight = (g_rank + 1) % g_nranks;
}

static size_t slab_floats(ParamCoLoRe *par)
{
  return (size_t)2 * (par->n_grid / 2 + 1) * par->n_grid * par->nz_here;
}

/* ------------------------------------------------------------------ fourier.c */
void init_fftw(ParamCoLoRe *par)
{ /* fourier.c:127-209: one z slab per rank (fourier.c:172-181); one rank = the whole box */
  int r;
  if (par->n_grid % g_nranks) report_error(1, "n_grid=%d is not divisible by %d GPUs\n", par->n_grid, g_nranks);
  par->nz_all = my_calloc(g_nranks, sizeof(int));
  par->iz0_all = my_calloc(g_nranks, sizeof(int));
  for (r = 0; r < g_nranks; r++) {
    par->nz_all[r] = par->n_grid / g_nranks;
    par->iz0_all[r] = r * (par->n_grid / g_nranks);
  }
  par->nz_here = par->nz_all[g_rank];
  par->iz0_here = par->iz0_all[g_rank];
  par->nz_max = par->nz_here;
  if (g_rank > 0) par->do_pred = 0;       /* the theory predictions are rank 0's job (main.c:58-59) */
}

static int g_ngrid = 0;
static int g_ctx_ngrid(void) { return g_ngrid; }

void allocate_fftw(ParamCoLoRe *par)
{
  g_ngrid = par->n_grid; /* fourier.c:211-238: grids live on the device; tables are uploaded once */
  clr_params p;
  int i;
  memset(&p, 0, sizeof(p));
  p.n_grid = par->n_grid; p.nz_here = par->nz_here; p.iz0_here = par->iz0_here;
  p.dens_type = par->dens_type;
#ifdef _BIAS_MODEL_2
  p.bias_model = 2;
#elif defined _BIAS_MODEL_3
  p.bias_model = 3;
#else
  p.bias_model = 1;
#endif
  p.do_smoothing = par->do_smoothing; p.smooth_potential = par->smooth_potential;
  p.nside_base = par->nside_base; p.numk = par->numk; p.seed_rng = par->seed_rng;
  p.l_box = par->l_box;
  for (i = 0; i < 3; i++) p.pos_obs[i] = par->pos_obs[i];
  p.r2_smooth = par->r2_smooth; p.prefac_lensing = par->prefac_lensing;
  p.fgrowth_0 = par->fgrowth_0; p.hubble_0 = par->hubble_0; p.OmegaM = par->OmegaM; p.n_scal = par->n_scal;
  p.r_max = par->r_max; p.glob_idr = par->glob_idr;
  p.logkmin = par->logkmin; p.logkmax = par->logkmax; p.idlogk = par->idlogk;
  p.logkarr = par->logkarr; p.pkarr = par->pkarr;
  p.r_arr_r2z = par->r_arr_r2z; p.z_arr_r2z = par->z_arr_r2z; p.growth_d_arr = par->growth_d_arr;
  p.growth_d2_arr = par->growth_d2_arr; p.growth_v_arr = par->growth_v_arr;
  p.growth_pd_arr = par->growth_pd_arr; p.ihub_arr = par->ihub_arr;
  p.a_arr_a2r = par->a_arr_a2r; p.r_arr_a2r = par->r_arr_a2r;
  {
    int ndev = clr_device_count();
    if (ndev < 1) report_error(1, "colore_b200: no CUDA device (there is no CPU fallback)\n");
    chk(clr_create(&p, g_rank % ndev, &g_ctx));
  }
  if (g_nranks > 1) {
    unsigned char id[128];
    if (g_rank == 0) chk(clr_comm_unique_id(id));
    exchange_nccl_id(id);
    chk(clr_comm_init(g_ctx, g_rank, g_nranks, id));
  }
  chk(clr_set_option(g_ctx, "lpt_interp_type", par->lpt_interp_type));
  chk(clr_set_option(g_ctx, "keep_particles", par->output_lpt));
  for (i = 0; i < par->n_srcs; i++) chk(clr_set_srcs(g_ctx, i, par->srcs_nz_arr[i], par->srcs_bz_arr[i]));
  for (i = 0; i < par->n_cstm; i++) chk(clr_set_cstm(g_ctx, i, par->cstm_kz_arr[i], par->cstm_bz_arr[i]));
  par->grid_dens_f = NULL; par->grid_dens = NULL;
  par->grid_npot_f = NULL; par->grid_npot = NULL;
  if (par->output_density) { /* io.c:565-595 reads par->grid_dens on the host */
    par->grid_dens = my_malloc(slab_floats(par) * sizeof(flouble));
    par->grid_dens_f = (dftw_complex *)par->grid_dens;
  }
}

void end_fftw(ParamCoLoRe *par)
{ /* fourier.c:240-283 */
  if (par->grid_dens != NULL) free(par->grid_dens);
  par->grid_dens = NULL; par->grid_dens_f = NULL;
  if (g_ctx) chk(clr_destroy(g_ctx));
  g_ctx = NULL;
}

/* fourier.c:81-125 on HOST arrays. Nothing of the host code that stays linked calls these (the transforms of the run
 * flow happen inside clr_create_cartesian_fields / the LPT kernels); a host round trip would have to borrow a device
 * grid and destroy the resident field, so they refuse instead of doing that silently. */
static int g_fields_resident = 0;
void fftw_wrap_c2r(int ng, dftw_complex *pin, flouble *pout)
{
  if (g_fields_resident || ng != g_ctx_ngrid()) report_error(1, "fftw_wrap_c2r on host arrays would overwrite the resident GPU fields\n");
  chk(clr_grid_put(g_ctx, CLR_GRID_DENS, (const float *)pin));
  chk(clr_fft_c2r(g_ctx, CLR_GRID_DENS));
  chk(clr_grid_get(g_ctx, CLR_GRID_DENS, (float *)pout));
}
void fftw_wrap_r2c(int ng, flouble *pin, dftw_complex *pout)
{
  if (g_fields_resident || ng != g_ctx_ngrid()) report_error(1, "fftw_wrap_r2c on host arrays would overwrite the resident GPU fields\n");
  chk(clr_grid_put(g_ctx, CLR_GRID_DENS, (const float *)pin));
  chk(clr_fft_r2c(g_ctx, CLR_GRID_DENS));
  chk(clr_grid_get(g_ctx, CLR_GRID_DENS, (float *)pout));
}

void create_cartesian_fields(ParamCoLoRe *par)
{ /* fourier.c:361-423 */
  double out[2];
  adopt_rank_layout();
  print_info("*** Creating Gaussian density field (GPU)\n");
  if (NodeThis == 0) timer(0);
  chk(clr_create_cartesian_fields(g_ctx, par->seed_rng, 0, out));
  g_fields_resident = 1;
  par->sigma2_gauss = out[1];
  if (NodeThis == 0) timer(2);
  print_info(" <d>=%.3lE, <d^2>=%.3lE\n", out[0], sqrt(par->sigma2_gauss));
  print_info("\n");
  if (par->output_density) {
    chk(clr_grid_get(g_ctx, CLR_GRID_DENS, (float *)par->grid_dens));
    write_density_grid(par, "gaussian");
  }
}

/* ------------------------------------------------------------------ density.c */
void compute_physical_density_field(ParamCoLoRe *par)
{ /* density.c:1105-1126 */
  print_info("*** Creating physical matter density (GPU)\n");
  if (NodeThis == 0) timer(0);
  chk(clr_compute_physical_density_field(g_ctx));
  chk(clr_synchronize(g_ctx));
  if (par->output_lpt && (par->dens_type == DENS_TYPE_1LPT || par->dens_type == DENS_TYPE_2LPT)) {
    unsigned long long np = par->nz_here * ((long)(par->n_grid * par->n_grid));   /* density.c:563 */
    flouble *x = my_malloc(np * sizeof(flouble)), *y = my_malloc(np * sizeof(flouble)), *z = my_malloc(np * sizeof(flouble));
    chk(clr_lpt_get_particles(g_ctx, x, y, z));
    write_lpt(par, np, x, y, z);
    free(x); free(y); free(z);
  }
  if (NodeThis == 0) timer(2);
  print_info("\n");
  if (par->output_density) {
    chk(clr_grid_get(g_ctx, CLR_GRID_DENS, (float *)par->grid_dens));
    write_density_grid(par, "lightcone");
  }
}

static void sorted_shells(HealpixShells *sh)
{ /* kappa.c:41-56, isw.c:41-56, imap.c:107-121: radii in ascending order of r0 */
  int i, *order = ind_sort(sh->nr, sh->r0);
  flouble *r0 = my_malloc(sh->nr * sizeof(flouble)), *rf = my_malloc(sh->nr * sizeof(flouble));
  memcpy(r0, sh->r0, sh->nr * sizeof(flouble));
  memcpy(rf, sh->rf, sh->nr * sizeof(flouble));
  for (i = 0; i < sh->nr; i++) { sh->r0[i] = r0[order[i]]; sh->rf[i] = rf[order[i]]; }
  free(r0); free(rf); free(order);
}

void compute_density_normalization(ParamCoLoRe *par)
{ /* density.c:1227-1393 */
  int i;
  double zends[2] = {0, 0};
  print_info("*** Computing normalization of density field (GPU)\n");
  if (NodeThis == 0) timer(0);
  /* intensity-map populations take part in the normalisation: their shells are known by now */
  for (i = 0; i < par->n_imap; i++) {
    sorted_shells(par->imap[i]);
    chk(clr_set_imap(g_ctx, i, par->imap_tz_arr[i], par->imap_bz_arr[i], par->imap[i]->nside, par->imap[i]->nr,
                     par->imap[i]->r0, par->imap[i]->rf));
  }
  chk(clr_compute_density_normalization(g_ctx));
  for (i = 0; i < par->n_srcs; i++) {
    double ends[2];
    par->srcs_norm_arr[i] = my_malloc(NA * sizeof(double));
    chk(clr_get_norm(g_ctx, 0, i, par->srcs_norm_arr[i], ends, zends));
    par->norm_srcs_0[i] = ends[0]; par->norm_srcs_f[i] = ends[1];
  }
  for (i = 0; i < par->n_imap; i++) {
    double ends[2];
    par->imap_norm_arr[i] = my_malloc(NA * sizeof(double));
    chk(clr_get_norm(g_ctx, 1, i, par->imap_norm_arr[i], ends, zends));
    par->norm_imap_0[i] = ends[0]; par->norm_imap_f[i] = ends[1];
  }
  par->z0_norm = zends[0]; par->zf_norm = zends[1];
  for (i = 0; i < par->n_cstm; i++) { /* density.c:1315-1354 */
    double ends[2];
    par->cstm_norm_arr[i] = my_malloc(NA * sizeof(double));
    chk(clr_get_norm(g_ctx, 2, i, par->cstm_norm_arr[i], ends, zends));
    par->norm_cstm_0[i] = ends[0]; par->norm_cstm_f[i] = ends[1];
  }
  if (NodeThis == 0) timer(2);
  print_info("\n");
}

/* ------------------------------------------------------------------ srcs.c */
void srcs_set_cartesian(ParamCoLoRe *par)
{ /* srcs.c:285-294 */
  int ipop;
  print_info("*** Getting point sources (GPU)\n");
  for (ipop = 0; ipop < par->n_srcs; ipop++) {
    long long n = 0;
    if (NodeThis == 0) timer(0);
    chk(clr_srcs_set_cartesian(g_ctx, ipop, par->seed_rng, &n));
    par->nsources_c_this[ipop] = (long)n;
    print_info("   There will be %ld objects in total \n", (long)n);
    par->cats_c[ipop] = catalog_cartesian_alloc((int)n);
    if (n > 0) chk(clr_srcs_get_cartesian(g_ctx, ipop, par->cats_c[ipop]->pos, par->cats_c[ipop]->ipix));
    if (NodeThis == 0) timer(2);
  }
  print_info("\n");
}

static int g_by_pixel = 0;   /* COLORE_B200_DISTRIBUTE=pixel: the reference's routing, source -> rank ipix % NNodes */
void srcs_distribute(ParamCoLoRe *par)
{ /* srcs.c:375-384. Default: every rank keeps (and writes) the sources of its own z slab -- the union of the per-rank
   * files is the same catalogue. With COLORE_B200_DISTRIBUTE=pixel the sources are routed like srcs.c:296-373 (on the
   * device, clr_srcs_distribute); the RSD under beaming is then evaluated BEFORE they leave the slab that holds the
   * potential around them. */
  int ipop;
  const char *mode = getenv("COLORE_B200_DISTRIBUTE");
  g_by_pixel = (NNodes > 1 && mode && !strcmp(mode, "pixel"));
  for (ipop = 0; ipop < par->n_srcs; ipop++) {
    if (g_by_pixel) {
      long long n = 0;
      if (NodeThis == 0) timer(0);
      chk(clr_srcs_distribute(g_ctx, ipop, par->need_beaming, &n));
      catalog_cartesian_free(par->cats_c[ipop]);
      par->cats_c[ipop] = catalog_cartesian_alloc((int)n);
      if (n > 0) chk(clr_srcs_get_cartesian(g_ctx, ipop, par->cats_c[ipop]->pos, par->cats_c[ipop]->ipix));
      par->nsources_c_this[ipop] = (long)n;
      if (NodeThis == 0) timer(2);
    }
    par->nsources_this[ipop] = par->nsources_c_this[ipop];
  }
}

void srcs_get_local_properties(ParamCoLoRe *par)
{ /* srcs.c:418-423 */
  int ipop;
  for (ipop = 0; ipop < par->n_srcs; ipop++) {
    par->cats[ipop] = catalog_alloc(par->cats_c[ipop]->nsrc, par->lensing_srcs[ipop], par->skw_srcs[ipop],
                                    par->skw_gauss[ipop], par->r_max, par->n_grid);
    if (par->cats[ipop]->nsrc > 0)
      chk(clr_srcs_get_local_properties(g_ctx, ipop, (float *)par->cats[ipop]->srcs));
  }
}

void srcs_beams_preproc(ParamCoLoRe *par) { (void)par; }
void srcs_get_beam_properties(ParamCoLoRe *par)
{ /* srcs.c:425-744: dz_rsd from the CIC-interpolated potential gradient, per-source lensing, skewers */
  int ipop;
  for (ipop = 0; ipop < par->n_srcs; ipop++) {
    Catalog *cat = par->cats[ipop];
    int lens_rays = par->lensing_srcs[ipop];
#ifdef _USE_FAST_LENSING
    lens_rays = 0;   /* srcs.c:531 is compiled out: the sources read the shells of lensing.c instead (srcs.c:666-721) */
#endif
    /* routed by pixel: dz_rsd was evaluated before the exchange, on the slab that holds the potential */
    chk(clr_srcs_get_beam_properties(g_ctx, ipop, lens_rays, par->skw_srcs[ipop], par->skw_gauss[ipop], g_by_pixel));
#ifdef _USE_FAST_LENSING
    if (par->lensing_srcs[ipop] && cat->nsrc > 0) {
      /* every rank holds full-sky shells (they were allocated while NNodes was still 1): beam ib = base pixel ib */
      HealpixShellsAdaptive *m = par->smap;
      long long bad = 0;
      chk(clr_srcs_lensing_from_shells(g_ctx, ipop, m->nr, m->r, m->nside, 0, 1, &bad));
      if (bad) report_error(1, "Bad base!!\n");      /* srcs.c:684-685 */
    }
#endif
    if (cat->nsrc > 0) {
      chk(clr_srcs_get_local_properties(g_ctx, ipop, (float *)cat->srcs));
      if (cat->has_skw) chk(clr_srcs_get_skewers(g_ctx, ipop, cat->skw_gauss ? cat->g_skw : cat->d_skw, cat->v_skw));
    }
  }
}
void srcs_beams_postproc(ParamCoLoRe *par) { (void)par; }

/* ------------------------------------------------------------------ imap.c */
void imap_set_cartesian(ParamCoLoRe *par)
{ /* imap.c:247-256; every rank stores the full sky (imap.c:123-132) */
  int ipop;
  print_info("*** Filling up intensity maps (GPU)\n");
  for (ipop = 0; ipop < par->n_imap; ipop++) {
    HealpixShells *im = par->imap[ipop];
    if (NodeThis == 0) timer(0);
    im->num_pix = he_nside2npix(im->nside);
    free(im->listpix); im->listpix = my_malloc(sizeof(long));
    free(im->pos); im->pos = my_malloc(sizeof(double));
    free(im->data); im->data = my_calloc(im->nr * im->num_pix, sizeof(flouble));
    free(im->nadd); im->nadd = my_calloc(im->nr * im->num_pix, sizeof(int));
    chk(clr_imap_set_cartesian(g_ctx, ipop, im->data, im->nadd));
    if (NodeThis == 0) timer(2);
  }
  print_info("\n");
}
void imap_distribute(ParamCoLoRe *par) { (void)par; }
void imap_get_local_properties(ParamCoLoRe *par) { (void)par; }
void imap_beams_preproc(ParamCoLoRe *par) { (void)par; }
void imap_get_beam_properties(ParamCoLoRe *par) { (void)par; }
void imap_beams_postproc(ParamCoLoRe *par) { (void)par; }

/* ------------------------------------------------------------------ kappa.c / isw.c */
void kappa_set_cartesian(ParamCoLoRe *par) { (void)par; }
void kappa_distribute(ParamCoLoRe *par) { (void)par; }
void kappa_get_local_properties(ParamCoLoRe *par) { (void)par; }
void kappa_beams_preproc(ParamCoLoRe *par)
{ /* kappa.c:39-76 */
  long i;
  sorted_shells(par->kmap);
  for (i = 0; i < par->kmap->num_pix * par->kmap->nr; i++) { par->kmap->data[i] = 0; par->kmap->nadd[i] = 1; }
}
void kappa_get_beam_properties(ParamCoLoRe *par)
{ /* kappa.c:78-175 */
  HealpixShells *k = par->kmap;
  chk(clr_kappa_get_beam_properties(g_ctx, k->num_pix, k->pos, k->nr, k->rf, k->data));
}
void kappa_beams_postproc(ParamCoLoRe *par) { (void)par; }

void isw_set_cartesian(ParamCoLoRe *par) { (void)par; }
void isw_distribute(ParamCoLoRe *par) { (void)par; }
void isw_get_local_properties(ParamCoLoRe *par) { (void)par; }
void isw_beams_preproc(ParamCoLoRe *par)
{ /* isw.c:39-76 */
  long i;
  sorted_shells(par->pd_map);
  for (i = 0; i < par->pd_map->num_pix * par->pd_map->nr; i++) { par->pd_map->data[i] = 0; par->pd_map->nadd[i] = 1; }
}
void isw_get_beam_properties(ParamCoLoRe *par)
{ /* isw.c:78-147 */
  HealpixShells *m = par->pd_map;
  chk(clr_isw_get_beam_properties(g_ctx, m->num_pix, m->pos, m->nr, m->rf, m->data));
}
void isw_beams_postproc(ParamCoLoRe *par) { (void)par; }

/* ------------------------------------------------------------------ cstm.c */
void cstm_set_cartesian(ParamCoLoRe *par) { (void)par; }
void cstm_distribute(ParamCoLoRe *par) { (void)par; }
void cstm_get_local_properties(ParamCoLoRe *par) { (void)par; }
void cstm_beams_preproc(ParamCoLoRe *par)
{ /* cstm.c:38-66 */
  int ipop;
  long i;
  for (ipop = 0; ipop < par->n_cstm; ipop++)
    for (i = 0; i < par->cstm[ipop]->num_pix; i++) { par->cstm[ipop]->data[i] = 0; par->cstm[ipop]->nadd[i] = 1; }
}
void cstm_get_beam_properties(ParamCoLoRe *par)
{ /* cstm.c:68-152; on several GPUs the map comes back summed over the slabs and rank 0 writes it (io.c:754-806
   * without _HAVE_MPI), like the kappa / ISW maps */
  int ipop;
  for (ipop = 0; ipop < par->n_cstm; ipop++) {
    HealpixShells *m = par->cstm[ipop];
    chk(clr_cstm_get_beam_properties(g_ctx, ipop, m->num_pix, m->pos, m->data));
  }
}
void cstm_beams_postproc(ParamCoLoRe *par) { (void)par; }

#ifdef _USE_FAST_LENSING
/* ------------------------------------------------------------------ lensing.c (this build leaves it out of the link) */
void lensing_set_cartesian(ParamCoLoRe *par) { (void)par; }
void lensing_distribute(ParamCoLoRe *par) { (void)par; }
void lensing_get_local_properties(ParamCoLoRe *par) { (void)par; }
void lensing_beams_preproc(ParamCoLoRe *par)
{ /* lensing.c:39-74: radii in ascending order; the library zeroes the maps */
  HealpixShellsAdaptive *m = par->smap;
  int ir, *order = ind_sort(m->nr, m->r);
  flouble *r = my_malloc(m->nr * sizeof(flouble));
  memcpy(r, m->r, m->nr * sizeof(flouble));
  for (ir = 0; ir < m->nr; ir++) m->r[ir] = r[order[ir]];
  free(r); free(order);
}
void lensing_get_beam_properties(ParamCoLoRe *par)
{ /* lensing.c:76-244 for all beams at once; on several GPUs the shells come back summed over the slabs */
  HealpixShellsAdaptive *m = par->smap;
  long ib, ir, npix_hi = m->num_pix_per_beam[m->nr - 1];
  long long total = 0, off = 0, *npp = my_malloc(m->nr * sizeof(long long));
  double *pos = my_malloc((size_t)m->nbeams * npix_hi * 3 * sizeof(double));
  float *data;
  for (ir = 0; ir < m->nr; ir++) { npp[ir] = m->num_pix_per_beam[ir]; total += 5LL * m->nbeams * npp[ir]; }
  for (ib = 0; ib < m->nbeams; ib++) memcpy(pos + ib * npix_hi * 3, m->pos[ib], npix_hi * 3 * sizeof(double));
  data = my_malloc(total * sizeof(float));
  chk(clr_lensing_get_beam_properties(g_ctx, m->nbeams, m->nr, m->r, npp, pos, data));
  for (ir = 0; ir < m->nr; ir++)
    for (ib = 0; ib < m->nbeams; ib++) {
      memcpy(m->data[ib][ir], data + off, 5 * npp[ir] * sizeof(float));
      off += 5 * npp[ir];
    }
  free(data); free(pos); free(npp);
}
void lensing_beams_postproc(ParamCoLoRe *par) { (void)par; }
#endif

/* ------------------------------------------------------------------ beaming.c */
int interpolate_from_grid(ParamCoLoRe *par, double *x, flouble *d, flouble v[3], flouble t[6], flouble *pd,
                          flouble *g, int flag_return, int interp_type)
{ /* beaming.c:120-268: every caller in the reference (kappa, isw, srcs, cstm, lensing) is replaced above */
  (void)par; (void)x; (void)d; (void)v; (void)t; (void)pd; (void)g; (void)flag_return; (void)interp_type;
  report_error(1, "interpolate_from_grid: the grids live on the GPU\n");
  return 0;
}

void get_beam_properties(ParamCoLoRe *par)
{ /* beaming.c:293-374 without the ring rotation: every GPU integrates inside its own slab (+ halo), maps are summed */
  print_info("*** Getting LOS information (GPU)\n");
  if (!par->need_beaming) { print_info("  No need!\n\n"); return; }
  if (par->do_kappa) kappa_beams_preproc(par);
  if (par->do_isw) isw_beams_preproc(par);
  if (par->do_srcs) srcs_beams_preproc(par);
  if (par->do_cstm) cstm_beams_preproc(par);
  if (NodeThis == 0) timer(0);
  if (par->do_kappa) kappa_get_beam_properties(par);
#ifdef _USE_FAST_LENSING
  if (par->do_lensing) { lensing_beams_preproc(par); lensing_get_beam_properties(par); }   /* before the sources read them */
#endif
  if (par->do_isw) isw_get_beam_properties(par);
  if (par->do_srcs) srcs_get_beam_properties(par);
  if (par->do_cstm) cstm_get_beam_properties(par);
  if (NodeThis == 0) timer(2);
  print_info("\n");
}
