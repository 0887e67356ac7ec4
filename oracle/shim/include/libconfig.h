/* Minimal libconfig-compatible shim (TEST INFRASTRUCTURE ONLY): the subset of the
 * hyperrealm libconfig C API used by io.c:160-426 of the reference. */
#ifndef SHIM_LIBCONFIG_H
#define SHIM_LIBCONFIG_H
#include <stdio.h>
#define CONFIG_TRUE 1
#define CONFIG_FALSE 0
#define CONFIG_OPTION_AUTOCONVERT 0x01
enum { CONFIG_TYPE_NONE = 0, CONFIG_TYPE_GROUP, CONFIG_TYPE_INT, CONFIG_TYPE_INT64, CONFIG_TYPE_FLOAT,
       CONFIG_TYPE_STRING, CONFIG_TYPE_BOOL, CONFIG_TYPE_ARRAY, CONFIG_TYPE_LIST };
typedef struct config_setting_t {
  char *name;
  int type;
  long long ival;
  double fval;
  char *sval;
  int nchild;
  struct config_setting_t **child;
} config_setting_t;
typedef struct config_t {
  config_setting_t *root;
  int options;
  const char *error_text;
  int error_line;
} config_t;
void config_init(config_t *c);
void config_destroy(config_t *c);
void config_set_options(config_t *c, int options);
int config_read_file(config_t *c, const char *fname);
int config_write_file(config_t *c, const char *fname);
config_setting_t *config_lookup(const config_t *c, const char *path);
int config_lookup_int(const config_t *c, const char *path, int *value);
int config_lookup_float(const config_t *c, const char *path, double *value);
int config_lookup_bool(const config_t *c, const char *path, int *value);
int config_lookup_string(const config_t *c, const char *path, const char **value);
int config_setting_length(const config_setting_t *s);
double config_setting_get_float_elem(const config_setting_t *s, int idx);
#endif
