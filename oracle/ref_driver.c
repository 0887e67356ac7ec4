/* TEST INFRASTRUCTURE ONLY.
 *
 * Driver that runs the UNMODIFIED reference translation units (compiled from
 * /root/reference/src against third_party/shim) in the order of main.c:50-147 and dumps the
 * state after every stage as .npy files. It is linked INSTEAD of the reference's main.c and
 * contains no reference code: it only calls the public functions declared in the reference's
 * common.h and reads fields of ParamCoLoRe.
 *
 * Usage: ref_driver param.cfg outdir
 */
#include "common.h"

static void npy_write(const char *dir, const char *name, const char *descr, int ndim,
                      const long *shape, const void *data, size_t elsize)
{
  char fname[512], hdr[256], shp[128] = "";
  size_t n = 1, hl;
  int i;
  FILE *f;
  for (i = 0; i < ndim; i++) {
    char t[32];
    sprintf(t, "%ld,", shape[i]);
    strcat(shp, t);
    n *= shape[i];
  }
  sprintf(hdr, "{'descr': '%s', 'fortran_order': False, 'shape': (%s), }", descr, shp);
  hl = strlen(hdr);
  while ((10 + hl + 1) % 64) hdr[hl++] = ' ';
  hdr[hl++] = '\n';
  sprintf(fname, "%s/%s.npy", dir, name);
  f = fopen(fname, "wb");
  if (!f) { fprintf(stderr, "cannot open %s\n", fname); exit(1); }
  fwrite("\x93NUMPY\x01\x00", 1, 8, f);
  fputc((int)(hl & 0xff), f);
  fputc((int)(hl >> 8), f);
  fwrite(hdr, 1, hl, f);
  if (n) fwrite(data, elsize, n, f);
  fclose(f);
}

static void dump_f64(const char *d, const char *nm, const double *x, long n)
{ npy_write(d, nm, "<f8", 1, &n, x, 8); }
static void dump_f32(const char *d, const char *nm, const float *x, long n)
{ npy_write(d, nm, "<f4", 1, &n, x, 4); }
static void dump_i32(const char *d, const char *nm, const int *x, long n)
{ npy_write(d, nm, "<i4", 1, &n, x, 4); }
static void dump_i64(const char *d, const char *nm, const long *x, long n)
{ npy_write(d, nm, "<i8", 1, &n, x, 8); }

static void dump_grid(const char *d, const char *nm, const flouble *g, ParamCoLoRe *par)
{
  long shape[3] = {par->nz_here, par->n_grid, 2 * (par->n_grid / 2 + 1)};
  npy_write(d, nm, "<f4", 3, shape, g, sizeof(flouble));
}

int main(int argc, char **argv)
{
  ParamCoLoRe *par;
  const char *d;
  char nm[128];
  int i;
  double sc[32];
  if (argc < 3) { fprintf(stderr, "usage: ref_driver param.cfg outdir\n"); return 1; }
  d = argv[2];
  mpi_init(&argc, &argv);
  setbuf(stdout, NULL);
  par = read_run_params(argv[1], 0);
  if (sizeof(flouble) != 4) { fprintf(stderr, "ref_driver expects -D_SPREC\n"); return 1; }

  /* ---- host tables = inputs of the hot path (cosmo.c:516-816) */
  dump_f64(d, "tab_z", par->z_arr_r2z, NA);
  dump_f64(d, "tab_r", par->r_arr_r2z, NA);
  dump_f64(d, "tab_d1", par->growth_d_arr, NA);
  dump_f64(d, "tab_d2", par->growth_d2_arr, NA);
  dump_f64(d, "tab_v1", par->growth_v_arr, NA);
  dump_f64(d, "tab_pd", par->growth_pd_arr, NA);
  dump_f64(d, "tab_ih", par->ihub_arr, NA);
  dump_f64(d, "tab_a2r_a", par->a_arr_a2r, NA);
  dump_f64(d, "tab_a2r_r", par->r_arr_a2r, NA);
  dump_f64(d, "pk_logk", par->logkarr, par->numk);
  dump_f64(d, "pk_pk", par->pkarr, par->numk);
  for (i = 0; i < par->n_srcs; i++) {
    sprintf(nm, "tab_srcs_nz_%d", i); dump_f64(d, nm, par->srcs_nz_arr[i], NA);
    sprintf(nm, "tab_srcs_bz_%d", i); dump_f64(d, nm, par->srcs_bz_arr[i], NA);
  }
  for (i = 0; i < par->n_imap; i++) {
    sprintf(nm, "tab_imap_tz_%d", i); dump_f64(d, nm, par->imap_tz_arr[i], NA);
    sprintf(nm, "tab_imap_bz_%d", i); dump_f64(d, nm, par->imap_bz_arr[i], NA);
  }
  for (i = 0; i < par->n_cstm; i++) {
    sprintf(nm, "tab_cstm_kz_%d", i); dump_f64(d, nm, par->cstm_kz_arr[i], NA);
    sprintf(nm, "tab_cstm_bz_%d", i); dump_f64(d, nm, par->cstm_bz_arr[i], NA);
  }
  sc[0] = par->l_box;          sc[1] = par->pos_obs[0];   sc[2] = par->glob_idr;
  sc[3] = par->prefac_lensing; sc[4] = par->fgrowth_0;    sc[5] = par->hubble_0;
  sc[6] = par->OmegaM;         sc[7] = par->n_scal;       sc[8] = par->r2_smooth;
  sc[9] = par->smooth_potential; sc[10] = par->do_smoothing; sc[11] = par->r_max;
  sc[12] = par->logkmin;       sc[13] = par->logkmax;     sc[14] = par->idlogk;
  sc[15] = par->numk;          sc[16] = par->n_grid;      sc[17] = par->seed_rng;
  sc[18] = par->nside_base;    sc[19] = par->dens_type;   sc[20] = par->sigma2_gauss; /* analytic */
  sc[21] = par->r_min;         sc[22] = par->z_min;       sc[23] = par->z_max;
  sc[24] = par->lpt_interp_type; sc[25] = par->lpt_buffer_fraction;
  dump_f64(d, "scalars", sc, 26);

  /* ---- stage 1: Gaussian fields (fourier.c:361-423) */
  create_cartesian_fields(par);
  dump_grid(d, "s1_dens_gauss", par->grid_dens, par);
  dump_grid(d, "s1_npot", par->grid_npot, par);
  dump_f64(d, "s1_sigma2_gauss", &par->sigma2_gauss, 1);

  /* ---- stage 2: physical density (density.c:1105-1126) */
  compute_physical_density_field(par);
  dump_grid(d, "s2_dens", par->grid_dens, par);

  /* ---- stage 3: normalisation (density.c:1227-1393) */
  compute_density_normalization(par);
  for (i = 0; i < par->n_srcs; i++) {
    double e[2] = {par->norm_srcs_0[i], par->norm_srcs_f[i]};
    sprintf(nm, "s3_srcs_norm_%d", i); dump_f64(d, nm, par->srcs_norm_arr[i], NA);
    sprintf(nm, "s3_srcs_norm_ends_%d", i); dump_f64(d, nm, e, 2);
  }
  for (i = 0; i < par->n_imap; i++) {
    double e[2] = {par->norm_imap_0[i], par->norm_imap_f[i]};
    sprintf(nm, "s3_imap_norm_%d", i); dump_f64(d, nm, par->imap_norm_arr[i], NA);
    sprintf(nm, "s3_imap_norm_ends_%d", i); dump_f64(d, nm, e, 2);
  }
  for (i = 0; i < par->n_cstm; i++) {
    double e[2] = {par->norm_cstm_0[i], par->norm_cstm_f[i]};
    sprintf(nm, "s3_cstm_norm_%d", i); dump_f64(d, nm, par->cstm_norm_arr[i], NA);
    sprintf(nm, "s3_cstm_norm_ends_%d", i); dump_f64(d, nm, e, 2);
  }
  { double e[2] = {par->z0_norm, par->zf_norm}; dump_f64(d, "s3_znorm_ends", e, 2); }

  /* ---- stage 4: tracers on the Cartesian grid (main.c:75-87) */
  if (par->do_srcs) {
    srcs_set_cartesian(par);
    for (i = 0; i < par->n_srcs; i++) {
      long n = par->nsources_c_this[i];
      sprintf(nm, "s4_srcs_pos_%d", i); dump_f32(d, nm, par->cats_c[i]->pos, NPOS_CC * n);
      sprintf(nm, "s4_srcs_ipix_%d", i); dump_i32(d, nm, par->cats_c[i]->ipix, n);
    }
  }
  if (par->do_imap) {
    imap_set_cartesian(par);
    for (i = 0; i < par->n_imap; i++) {
      HealpixShells *m = par->imap[i];
      sprintf(nm, "s4_imap_data_%d", i); dump_f32(d, nm, m->data, m->nr * m->num_pix);
      sprintf(nm, "s4_imap_nadd_%d", i); dump_i32(d, nm, m->nadd, m->nr * m->num_pix);
      sprintf(nm, "s4_imap_r0_%d", i); dump_f32(d, nm, m->r0, m->nr);
      sprintf(nm, "s4_imap_rf_%d", i); dump_f32(d, nm, m->rf, m->nr);
    }
  }
  /* ---- stage 5: distribute + local properties (main.c:89-117) */
  if (par->do_srcs) {
    srcs_distribute(par);
    srcs_get_local_properties(par);
    for (i = 0; i < par->n_srcs; i++) {
      sprintf(nm, "s5_srcs_cat_%d", i);
      dump_f32(d, nm, (float *)par->cats[i]->srcs, 9 * (long)par->cats[i]->nsrc);
    }
  }
  /* ---- stage 6: line-of-sight engine (beaming.c:293-374) */
  if (par->need_beaming) {
    get_beam_properties(par);
    if (par->do_kappa) {
      HealpixShells *m = par->kmap;
      dump_f32(d, "s6_kappa_data", m->data, m->nr * m->num_pix);
      dump_f32(d, "s6_kappa_rf", m->rf, m->nr);
      dump_i64(d, "s6_kappa_listpix", m->listpix, m->num_pix);
      dump_f64(d, "s6_kappa_pos", m->pos, 3 * m->num_pix);
    }
    if (par->do_isw) {
      HealpixShells *m = par->pd_map;
      dump_f32(d, "s6_isw_data", m->data, m->nr * m->num_pix);
      dump_f32(d, "s6_isw_rf", m->rf, m->nr);
    }
    if (par->do_srcs) {
      for (i = 0; i < par->n_srcs; i++) {
        Catalog *c = par->cats[i];
        double fl[3] = {c->has_lensing, c->has_skw, c->skw_gauss};
        sprintf(nm, "s6_srcs_cat_%d", i);
        dump_f32(d, nm, (float *)c->srcs, 9 * (long)c->nsrc);
        sprintf(nm, "s6_srcs_flags_%d", i); dump_f64(d, nm, fl, 3);
        if (c->nsrc > 0 && c->has_skw) { /* skewers of srcs.c:507-529, 725-733 */
          sprintf(nm, "s6_srcs_dgskw_%d", i); dump_f32(d, nm, c->skw_gauss ? c->g_skw : c->d_skw, (long)c->nsrc * c->nr);
          sprintf(nm, "s6_srcs_vskw_%d", i); dump_f32(d, nm, c->v_skw, (long)c->nsrc * c->nr);
        }
      }
    }
#ifdef _USE_FAST_LENSING
    if (par->do_lensing) { /* lensing.c:76-250: adaptive shells, 5 quantities per pixel */
      HealpixShellsAdaptive *m = par->smap;
      long nside[256], ir, ib, npix_hi = m->num_pix_per_beam[m->nr - 1];
      double *pos = my_malloc(m->nbeams * npix_hi * 3 * sizeof(double));
      for (ir = 0; ir < m->nr; ir++) nside[ir] = m->nside[ir];
      dump_f32(d, "s6_lens_r", m->r, m->nr);
      dump_i64(d, "s6_lens_nside", nside, m->nr);
      dump_i64(d, "s6_lens_npp", m->num_pix_per_beam, m->nr);
      for (ib = 0; ib < m->nbeams; ib++) memcpy(pos + ib * npix_hi * 3, m->pos[ib], npix_hi * 3 * sizeof(double));
      dump_f64(d, "s6_lens_pos", pos, m->nbeams * npix_hi * 3);
      free(pos);
      for (ir = 0; ir < m->nr; ir++) {
        long n5 = 5 * m->num_pix_per_beam[ir];
        flouble *buf = my_malloc(m->nbeams * n5 * sizeof(flouble));
        for (ib = 0; ib < m->nbeams; ib++) memcpy(buf + ib * n5, m->data[ib][ir], n5 * sizeof(flouble));
        sprintf(nm, "s6_lens_data_%03ld", ir); dump_f32(d, nm, buf, m->nbeams * n5);
        free(buf);
      }
    }
#endif
    if (par->do_cstm) { /* cstm.c:68-145 */
      for (i = 0; i < par->n_cstm; i++) {
        HealpixShells *m = par->cstm[i];
        sprintf(nm, "s6_cstm_data_%d", i); dump_f32(d, nm, m->data, m->num_pix);
        sprintf(nm, "s6_cstm_pos_%d", i); dump_f64(d, nm, m->pos, 3 * m->num_pix);
        sprintf(nm, "s6_cstm_listpix_%d", i); dump_i64(d, nm, m->listpix, m->num_pix);
      }
    }
  }
  param_colore_free(par);
  return 0;
}
