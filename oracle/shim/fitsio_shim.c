/* TEST INFRASTRUCTURE ONLY -- tiny FITS writer. Each HDU is buffered in memory and flushed
 * (header cards + big-endian data, 2880-byte blocks) when the next HDU starts or the file closes. */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "fitsio.h"

#define MAXCOL 64
typedef struct {
  int kind; /* 0 none, 1 bintable, 2 image */
  int ncol, width[MAXCOL], offset[MAXCOL];
  char code[MAXCOL];
  int rowbytes;
  long long nrows;
  unsigned char *data;
  size_t cap;
  char *cards; int ncards, capcards; /* user keywords */
  char tcards[3 * MAXCOL][81]; int ntcards;
  char extname[72];
  int bitpix, naxis; long naxes[8];
} hdu_t;

struct shim_fitsfile { FILE *f; int primary_written; hdu_t h; };

static void card(FILE *f, const char *txt, long *n)
{
  char buf[81];
  snprintf(buf, 81, "%-80s", txt);
  fwrite(buf, 1, 80, f);
  (*n)++;
}
static void pad_block(FILE *f, long nbytes, char fill)
{
  long r = nbytes % 2880;
  if (r) { long i; for (i = r; i < 2880; i++) fputc(fill, f); }
}
static void end_header(FILE *f, long ncards)
{
  long n = ncards;
  card(f, "END", &n);
  pad_block(f, n * 80, ' ');
}
static void write_primary(struct shim_fitsfile *ff)
{
  long n = 0;
  if (ff->primary_written) return;
  card(ff->f, "SIMPLE  =                    T / file does conform to FITS standard", &n);
  card(ff->f, "BITPIX  =                    8 / number of bits per data pixel", &n);
  card(ff->f, "NAXIS   =                    0 / number of data axes", &n);
  card(ff->f, "EXTEND  =                    T / FITS dataset may contain extensions", &n);
  end_header(ff->f, n);
  ff->primary_written = 1;
}

static void flush_hdu(struct shim_fitsfile *ff)
{
  hdu_t *h = &ff->h;
  char buf[128];
  long n = 0;
  int i;
  if (h->kind == 0) return;
  write_primary(ff);
  if (h->kind == 1) {
    card(ff->f, "XTENSION= 'BINTABLE'           / binary table extension", &n);
    card(ff->f, "BITPIX  =                    8 / 8-bit bytes", &n);
    card(ff->f, "NAXIS   =                    2 / 2-dimensional binary table", &n);
    snprintf(buf, 128, "NAXIS1  = %20d / width of table in bytes", h->rowbytes); card(ff->f, buf, &n);
    snprintf(buf, 128, "NAXIS2  = %20lld / number of rows in table", h->nrows); card(ff->f, buf, &n);
    card(ff->f, "PCOUNT  =                    0 / size of special data area", &n);
    card(ff->f, "GCOUNT  =                    1 / one data group", &n);
    snprintf(buf, 128, "TFIELDS = %20d / number of fields in each row", h->ncol); card(ff->f, buf, &n);
    for (i = 0; i < h->ntcards; i++) card(ff->f, h->tcards[i], &n);
    if (h->extname[0]) { snprintf(buf, 128, "EXTNAME = '%-8s'", h->extname); card(ff->f, buf, &n); }
  } else {
    long long tot = 1;
    card(ff->f, "XTENSION= 'IMAGE   '           / IMAGE extension", &n);
    snprintf(buf, 128, "BITPIX  = %20d / number of bits per data pixel", h->bitpix); card(ff->f, buf, &n);
    snprintf(buf, 128, "NAXIS   = %20d / number of data axes", h->naxis); card(ff->f, buf, &n);
    for (i = 0; i < h->naxis; i++) {
      snprintf(buf, 128, "NAXIS%-3d= %20ld", i + 1, h->naxes[i]); card(ff->f, buf, &n);
      tot *= h->naxes[i];
    }
    card(ff->f, "PCOUNT  =                    0", &n);
    card(ff->f, "GCOUNT  =                    1", &n);
    h->rowbytes = 4; h->nrows = tot;
  }
  for (i = 0; i < h->ncards; i++) card(ff->f, h->cards + 81 * i, &n);
  end_header(ff->f, n);
  {
    size_t nb = (size_t)h->rowbytes * h->nrows;
    if (nb > h->cap) { h->data = realloc(h->data, nb); memset(h->data + h->cap, 0, nb - h->cap); h->cap = nb; }
    fwrite(h->data, 1, nb, ff->f);
    pad_block(ff->f, (long)(nb % 2880), 0);
  }
  free(h->data); free(h->cards);
  memset(h, 0, sizeof(*h));
}

int fits_create_file(fitsfile **fptr, const char *fname, int *status)
{
  struct shim_fitsfile *ff = calloc(1, sizeof(*ff));
  if (*status) return *status;
  if (fname[0] == '!') fname++; /* cfitsio: leading '!' = clobber */
  ff->f = fopen(fname, "wb");
  if (!ff->f) { free(ff); *status = 105; return *status; }
  *fptr = ff;
  return 0;
}

static void add_user_card(hdu_t *h, const char *txt)
{
  if (h->ncards + 1 > h->capcards) { h->capcards = h->capcards ? 2 * h->capcards : 16; h->cards = realloc(h->cards, 81 * h->capcards); }
  snprintf(h->cards + 81 * h->ncards, 81, "%-80s", txt);
  h->ncards++;
}

int fits_create_tbl(fitsfile *ff, int tbltype, LONGLONG naxis2, int tfields, char **ttype,
                    char **tform, char **tunit, const char *extname, int *status)
{
  hdu_t *h = &ff->h;
  int i, off = 0;
  (void)tbltype; (void)naxis2;
  if (*status) return *status;
  flush_hdu(ff);
  h->kind = 1; h->ncol = tfields;
  for (i = 0; i < tfields; i++) {
    const char *t = tform[i];
    int rep = 1;
    char c;
    if (*t >= '0' && *t <= '9') rep = (int)strtol(t, (char **)&t, 10);
    c = *t;
    h->code[i] = c;
    h->width[i] = rep * ((c == 'D' || c == 'K') ? 8 : (c == 'I') ? 2 : (c == 'B' || c == 'A' || c == 'L') ? 1 : 4);
    h->offset[i] = off;
    off += h->width[i];
    snprintf(h->tcards[h->ntcards++], 81, "TTYPE%-3d= '%-8s'", i + 1, ttype[i]);
    snprintf(h->tcards[h->ntcards++], 81, "TFORM%-3d= '%d%c      '", i + 1, rep, c);
    if (tunit && tunit[i]) snprintf(h->tcards[h->ntcards++], 81, "TUNIT%-3d= '%-8s'", i + 1, tunit[i]);
  }
  h->rowbytes = off;
  if (extname) snprintf(h->extname, sizeof(h->extname), "%s", extname);
  return 0;
}

int fits_create_img(fitsfile *ff, int bitpix, int naxis, long *naxes, int *status)
{
  hdu_t *h = &ff->h;
  int i;
  if (*status) return *status;
  flush_hdu(ff);
  h->kind = 2; h->bitpix = bitpix; h->naxis = naxis;
  for (i = 0; i < naxis; i++) h->naxes[i] = naxes[i];
  h->rowbytes = 4;
  return 0;
}

int fits_write_key(fitsfile *ff, int datatype, const char *keyname, void *value, const char *comment, int *status)
{
  char buf[160];
  if (*status) return *status;
  if (datatype == TSTRING) snprintf(buf, 160, "%-8.8s= '%-8s' / %s", keyname, (char *)value, comment ? comment : "");
  else if (datatype == TLONG) snprintf(buf, 160, "%-8.8s= %20ld / %s", keyname, *(long *)value, comment ? comment : "");
  else if (datatype == TINT) snprintf(buf, 160, "%-8.8s= %20d / %s", keyname, *(int *)value, comment ? comment : "");
  else if (datatype == TFLOAT) snprintf(buf, 160, "%-8.8s= %20.8E / %s", keyname, *(float *)value, comment ? comment : "");
  else if (datatype == TDOUBLE) snprintf(buf, 160, "%-8.8s= %20.12E / %s", keyname, *(double *)value, comment ? comment : "");
  else { *status = 410; return *status; }
  add_user_card(&ff->h, buf);
  return 0;
}
int fits_update_key(fitsfile *ff, int datatype, const char *keyname, void *value, const char *comment, int *status)
{ return fits_write_key(ff, datatype, keyname, value, comment, status); }
int fits_write_comment(fitsfile *ff, const char *comment, int *status)
{
  char buf[160];
  if (*status) return *status;
  snprintf(buf, 160, "COMMENT %s", comment);
  add_user_card(&ff->h, buf);
  return 0;
}
int fits_get_rowsize(fitsfile *ff, long *nrows, int *status) { (void)ff; (void)status; *nrows = 8192; return 0; }

static void put_be(unsigned char *dst, const void *src, int nbytes)
{
  const unsigned char *s = src;
  int i;
  for (i = 0; i < nbytes; i++) dst[i] = s[nbytes - 1 - i];
}

static void ensure(hdu_t *h, long long nrows)
{
  size_t nb = (size_t)h->rowbytes * nrows;
  if (nb > h->cap) {
    size_t nc = h->cap ? h->cap : 4096;
    while (nc < nb) nc *= 2;
    h->data = realloc(h->data, nc);
    memset(h->data + h->cap, 0, nc - h->cap);
    h->cap = nc;
  }
  if (nrows > h->nrows) h->nrows = nrows;
}

int fits_write_col(fitsfile *ff, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                   LONGLONG nelements, void *array, int *status)
{
  hdu_t *h = &ff->h;
  int c = colnum - 1;
  LONGLONG i;
  (void)firstelem;
  if (*status) return *status;
  if (h->kind != 1 || c < 0 || c >= h->ncol) { *status = 302; return *status; }
  ensure(h, firstrow - 1 + nelements);
  for (i = 0; i < nelements; i++) {
    unsigned char *dst = h->data + (size_t)(firstrow - 1 + i) * h->rowbytes + h->offset[c];
    if (h->code[c] == 'E') {
      float v = (datatype == TFLOAT) ? ((float *)array)[i] : (datatype == TDOUBLE) ? (float)((double *)array)[i] : (float)((int *)array)[i];
      put_be(dst, &v, 4);
    } else if (h->code[c] == 'J') {
      int32_t v = (datatype == TINT) ? ((int *)array)[i] : (datatype == TLONG) ? (int32_t)((long *)array)[i] : (int32_t)((float *)array)[i];
      put_be(dst, &v, 4);
    } else if (h->code[c] == 'D') {
      double v = (datatype == TDOUBLE) ? ((double *)array)[i] : (double)((float *)array)[i];
      put_be(dst, &v, 8);
    } else { *status = 312; return *status; }
  }
  return 0;
}

int fits_write_img(fitsfile *ff, int datatype, LONGLONG firstelem, LONGLONG nelements, void *array, int *status)
{
  hdu_t *h = &ff->h;
  LONGLONG i;
  if (*status) return *status;
  if (h->kind != 2 || datatype != TFLOAT) { *status = 410; return *status; }
  h->rowbytes = 4;
  ensure(h, firstelem - 1 + nelements);
  for (i = 0; i < nelements; i++) put_be(h->data + 4 * (size_t)(firstelem - 1 + i), &((float *)array)[i], 4);
  return 0;
}

int fits_close_file(fitsfile *ff, int *status)
{
  (void)status;
  flush_hdu(ff);
  write_primary(ff);
  fclose(ff->f);
  free(ff);
  return 0;
}

static int unsupported(int *status)
{
  fprintf(stderr, "fitsio shim: reading FITS files is not implemented\n");
  *status = 104;
  exit(1);
  return *status;
}
int fits_open_file(fitsfile **f, const char *n, int m, int *s) { (void)f; (void)n; (void)m; return unsupported(s); }
int fits_movabs_hdu(fitsfile *f, int h, int *t, int *s) { (void)f; (void)h; (void)t; return unsupported(s); }
int fits_read_key_lng(fitsfile *f, const char *k, long *v, char *c, int *s) { (void)f; (void)k; (void)v; (void)c; return unsupported(s); }
int fits_read_keys_lng(fitsfile *f, const char *k, int a, int b, long *v, int *n, int *s) { (void)f; (void)k; (void)a; (void)b; (void)v; (void)n; return unsupported(s); }
int fits_read_key(fitsfile *f, int d, const char *k, void *v, char *c, int *s) { (void)f; (void)d; (void)k; (void)v; (void)c; return unsupported(s); }
int fits_read_col(fitsfile *f, int d, int c, LONGLONG a, LONGLONG b, LONGLONG n, void *nv, void *arr, int *an, int *s) { (void)f; (void)d; (void)c; (void)a; (void)b; (void)n; (void)nv; (void)arr; (void)an; return unsupported(s); }
