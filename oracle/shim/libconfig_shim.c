/* TEST INFRASTRUCTURE ONLY -- small recursive-descent parser for the libconfig
 * grammar subset found in CoLoRe parameter files: groups, scalar settings
 * (string / int / float / bool), arrays, '#', '//' and C comments, optional ';' or ','. */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "libconfig.h"

typedef struct { const char *s; int line; int err; } parser_t;

static config_setting_t *new_setting(const char *name, int type)
{
  config_setting_t *s = calloc(1, sizeof(*s));
  s->name = name ? strdup(name) : NULL;
  s->type = type;
  return s;
}

static void add_child(config_setting_t *p, config_setting_t *c)
{
  p->child = realloc(p->child, (p->nchild + 1) * sizeof(*p->child));
  p->child[p->nchild++] = c;
}

static void free_setting(config_setting_t *s)
{
  int i;
  if (!s) return;
  for (i = 0; i < s->nchild; i++) free_setting(s->child[i]);
  free(s->child); free(s->name); free(s->sval); free(s);
}

static void skip_ws(parser_t *p)
{
  for (;;) {
    while (*p->s && isspace((unsigned char)*p->s)) { if (*p->s == '\n') p->line++; p->s++; }
    if (*p->s == '#' || (p->s[0] == '/' && p->s[1] == '/')) {
      while (*p->s && *p->s != '\n') p->s++;
    } else if (p->s[0] == '/' && p->s[1] == '*') {
      p->s += 2;
      while (*p->s && !(p->s[0] == '*' && p->s[1] == '/')) { if (*p->s == '\n') p->line++; p->s++; }
      if (*p->s) p->s += 2;
    } else break;
  }
}

static config_setting_t *parse_value(parser_t *p, const char *name);

static void parse_settings(parser_t *p, config_setting_t *group, int until_brace)
{
  for (;;) {
    char name[256];
    int n = 0;
    skip_ws(p);
    if (!*p->s) { if (until_brace) p->err = 1; return; }
    if (*p->s == '}') { if (until_brace) { p->s++; return; } p->err = 1; return; }
    while (*p->s && (isalnum((unsigned char)*p->s) || *p->s == '_' || *p->s == '-' || *p->s == '*') && n < 255)
      name[n++] = *p->s++;
    name[n] = 0;
    if (n == 0) { p->err = 1; return; }
    skip_ws(p);
    if (*p->s != ':' && *p->s != '=') { p->err = 1; return; }
    p->s++;
    {
      config_setting_t *v = parse_value(p, name);
      if (!v) { p->err = 1; return; }
      add_child(group, v);
    }
    skip_ws(p);
    if (*p->s == ';' || *p->s == ',') p->s++;
    if (p->err) return;
  }
}

static config_setting_t *parse_value(parser_t *p, const char *name)
{
  skip_ws(p);
  if (*p->s == '{') {
    config_setting_t *g = new_setting(name, CONFIG_TYPE_GROUP);
    p->s++;
    parse_settings(p, g, 1);
    return g;
  }
  if (*p->s == '[' || *p->s == '(') {
    char close = (*p->s == '[') ? ']' : ')';
    config_setting_t *a = new_setting(name, close == ']' ? CONFIG_TYPE_ARRAY : CONFIG_TYPE_LIST);
    p->s++;
    for (;;) {
      skip_ws(p);
      if (*p->s == close) { p->s++; break; }
      if (!*p->s) { p->err = 1; break; }
      {
        config_setting_t *e = parse_value(p, NULL);
        if (!e) { p->err = 1; break; }
        add_child(a, e);
      }
      skip_ws(p);
      if (*p->s == ',') p->s++;
    }
    return a;
  }
  if (*p->s == '"') {
    config_setting_t *s = new_setting(name, CONFIG_TYPE_STRING);
    size_t cap = 64, n = 0;
    s->sval = malloc(cap);
    /* adjacent string literals are concatenated, as in libconfig */
    while (*p->s == '"') {
      p->s++;
      while (*p->s && *p->s != '"') {
        char ch = *p->s++;
        if (ch == '\\' && *p->s) {
          char e = *p->s++;
          ch = (e == 'n') ? '\n' : (e == 't') ? '\t' : (e == 'r') ? '\r' : e;
        }
        if (n + 2 > cap) { cap *= 2; s->sval = realloc(s->sval, cap); }
        s->sval[n++] = ch;
      }
      if (*p->s == '"') p->s++;
      skip_ws(p);
    }
    s->sval[n] = 0;
    return s;
  }
  if (!strncasecmp(p->s, "true", 4) && !isalnum((unsigned char)p->s[4])) {
    config_setting_t *s = new_setting(name, CONFIG_TYPE_BOOL);
    s->ival = 1; p->s += 4; return s;
  }
  if (!strncasecmp(p->s, "false", 5) && !isalnum((unsigned char)p->s[5])) {
    config_setting_t *s = new_setting(name, CONFIG_TYPE_BOOL);
    s->ival = 0; p->s += 5; return s;
  }
  {
    /* number: int unless it contains '.', 'e' or 'E' */
    const char *q = p->s;
    int isfloat = 0;
    char *end;
    if (*q == '+' || *q == '-') q++;
    if (!isdigit((unsigned char)*q) && *q != '.') return NULL;
    if (q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) {
      config_setting_t *s = new_setting(name, CONFIG_TYPE_INT);
      s->ival = strtoll(p->s, &end, 16); s->fval = (double)s->ival; p->s = end; return s;
    }
    while (isdigit((unsigned char)*q) || *q == '.' || *q == 'e' || *q == 'E' ||
           ((*q == '+' || *q == '-') && (q[-1] == 'e' || q[-1] == 'E'))) {
      if (*q == '.' || *q == 'e' || *q == 'E') isfloat = 1;
      q++;
    }
    if (isfloat) {
      config_setting_t *s = new_setting(name, CONFIG_TYPE_FLOAT);
      s->fval = strtod(p->s, &end); s->ival = (long long)s->fval; p->s = end; return s;
    } else {
      config_setting_t *s = new_setting(name, CONFIG_TYPE_INT);
      s->ival = strtoll(p->s, &end, 10); s->fval = (double)s->ival; p->s = end;
      if (*p->s == 'L') { p->s++; if (*p->s == 'L') p->s++; s->type = CONFIG_TYPE_INT64; }
      return s;
    }
  }
}

void config_init(config_t *c) { memset(c, 0, sizeof(*c)); c->root = new_setting(NULL, CONFIG_TYPE_GROUP); }
void config_destroy(config_t *c) { free_setting(c->root); c->root = NULL; }
void config_set_options(config_t *c, int options) { c->options = options; }

int config_read_file(config_t *c, const char *fname)
{
  FILE *f = fopen(fname, "rb");
  long sz;
  char *buf;
  parser_t p;
  if (!f) { c->error_text = "file I/O error"; return CONFIG_FALSE; }
  fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET);
  buf = malloc(sz + 1);
  if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); return CONFIG_FALSE; }
  buf[sz] = 0;
  fclose(f);
  p.s = buf; p.line = 1; p.err = 0;
  parse_settings(&p, c->root, 0);
  free(buf);
  if (p.err) { c->error_text = "syntax error"; c->error_line = p.line; return CONFIG_FALSE; }
  return CONFIG_TRUE;
}

static void write_setting(FILE *f, const config_setting_t *s, int depth, int in_array)
{
  int i;
  if (!in_array) { for (i = 0; i < depth; i++) fprintf(f, "  "); if (s->name) fprintf(f, "%s%s", s->name, s->type == CONFIG_TYPE_GROUP ? " : " : " = "); }
  switch (s->type) {
  case CONFIG_TYPE_GROUP:
    fprintf(f, "\n"); for (i = 0; i < depth; i++) fprintf(f, "  "); fprintf(f, "{\n");
    for (i = 0; i < s->nchild; i++) write_setting(f, s->child[i], depth + 1, 0);
    for (i = 0; i < depth; i++) fprintf(f, "  "); fprintf(f, "}");
    break;
  case CONFIG_TYPE_ARRAY: case CONFIG_TYPE_LIST:
    fprintf(f, s->type == CONFIG_TYPE_ARRAY ? "[ " : "( ");
    for (i = 0; i < s->nchild; i++) { write_setting(f, s->child[i], 0, 1); if (i + 1 < s->nchild) fprintf(f, ", "); }
    fprintf(f, s->type == CONFIG_TYPE_ARRAY ? " ]" : " )");
    break;
  case CONFIG_TYPE_STRING: fprintf(f, "\"%s\"", s->sval); break;
  case CONFIG_TYPE_BOOL: fprintf(f, s->ival ? "true" : "false"); break;
  case CONFIG_TYPE_FLOAT: fprintf(f, "%.10g", s->fval); if (s->fval == (double)(long long)s->fval) fprintf(f, ".0"); break;
  default: fprintf(f, "%lld", s->ival); break;
  }
  if (!in_array) fprintf(f, "%s\n", depth == 0 && s->type == CONFIG_TYPE_GROUP ? ";" : ";");
}

int config_write_file(config_t *c, const char *fname)
{
  FILE *f = fopen(fname, "w");
  int i;
  if (!f) return CONFIG_FALSE;
  for (i = 0; i < c->root->nchild; i++) write_setting(f, c->root->child[i], 0, 0);
  fclose(f);
  return CONFIG_TRUE;
}

config_setting_t *config_lookup(const config_t *c, const char *path)
{
  config_setting_t *cur = c->root;
  const char *s = path;
  while (*s && cur) {
    char name[256];
    int n = 0, i;
    config_setting_t *next = NULL;
    while (*s && *s != '.' && *s != '/' && *s != ':' && n < 255) name[n++] = *s++;
    name[n] = 0;
    if (*s) s++;
    if (n == 0) continue;
    if (cur->type != CONFIG_TYPE_GROUP) return NULL;
    for (i = 0; i < cur->nchild; i++)
      if (cur->child[i]->name && !strcmp(cur->child[i]->name, name)) { next = cur->child[i]; break; }
    cur = next;
  }
  return cur;
}

static int is_number(const config_setting_t *s)
{ return s->type == CONFIG_TYPE_INT || s->type == CONFIG_TYPE_INT64 || s->type == CONFIG_TYPE_FLOAT; }

int config_lookup_int(const config_t *c, const char *path, int *value)
{
  config_setting_t *s = config_lookup(c, path);
  if (!s) return CONFIG_FALSE;
  if (s->type == CONFIG_TYPE_INT) { *value = (int)s->ival; return CONFIG_TRUE; }
  if (s->type == CONFIG_TYPE_FLOAT && (c->options & CONFIG_OPTION_AUTOCONVERT)) { *value = (int)s->fval; return CONFIG_TRUE; }
  return CONFIG_FALSE;
}

int config_lookup_float(const config_t *c, const char *path, double *value)
{
  config_setting_t *s = config_lookup(c, path);
  if (!s) return CONFIG_FALSE;
  if (s->type == CONFIG_TYPE_FLOAT) { *value = s->fval; return CONFIG_TRUE; }
  if (is_number(s) && (c->options & CONFIG_OPTION_AUTOCONVERT)) { *value = (double)s->ival; return CONFIG_TRUE; }
  return CONFIG_FALSE;
}

int config_lookup_bool(const config_t *c, const char *path, int *value)
{
  config_setting_t *s = config_lookup(c, path);
  if (!s || s->type != CONFIG_TYPE_BOOL) return CONFIG_FALSE;
  *value = (int)s->ival;
  return CONFIG_TRUE;
}

int config_lookup_string(const config_t *c, const char *path, const char **value)
{
  config_setting_t *s = config_lookup(c, path);
  if (!s || s->type != CONFIG_TYPE_STRING) return CONFIG_FALSE;
  *value = s->sval;
  return CONFIG_TRUE;
}

int config_setting_length(const config_setting_t *s)
{
  if (s->type == CONFIG_TYPE_GROUP || s->type == CONFIG_TYPE_ARRAY || s->type == CONFIG_TYPE_LIST) return s->nchild;
  return 0;
}

double config_setting_get_float_elem(const config_setting_t *s, int idx)
{
  if (idx < 0 || idx >= s->nchild) return 0;
  /* libconfig only auto-converts when the option is set on the config; CoLoRe always sets it */
  return is_number(s->child[idx]) ? s->child[idx]->fval : 0;
}
