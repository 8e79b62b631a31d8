/*
 * ref_nc.c -- TEST INFRASTRUCTURE (oracle/_ref only): the reference's int16 scale/offset packing of NetCDF grids,
 * compress_2d / decompress_2d (model/nc.c:247-320), compiled from model/nc.c WHERE IT LIES: that file is included
 * below unmodified. libnetcdf is not in this image, so the nc_* calls of its two I/O functions (which this wrapper
 * never runs) resolve to local no-op stand-ins, and its write_nc / read_nc are renamed out of the way of the harness'
 * own write_nc stub.
 */
#include <float.h>
#include <limits.h>
#include <stdlib.h>

#define NC_CLOBBER 0
#define NC_NOWRITE 0
#define NC_FLOAT 5
#define NC_SHORT 3
#define NC_GLOBAL (-1)
static int nc_close() { return 0; }
static int nc_create() { return 0; }
static int nc_def_dim() { return 0; }
static int nc_def_var() { return 0; }
static int nc_enddef() { return 0; }
static int nc_get_att_float() { return 0; }
static int nc_get_att_short() { return 0; }
static int nc_get_var() { return 0; }
static int nc_inq_dimid() { return 0; }
static int nc_inq_dimlen() { return 0; }
static int nc_inq_varid() { return 0; }
static int nc_open() { return 0; }
static int nc_put_att_float() { return 0; }
static int nc_put_att_short() { return 0; }
static int nc_put_att_text() { return 0; }
static int nc_put_var_float() { return 0; }
static int nc_put_vara_short() { return 0; }
#define write_nc ref_nc_unused_write_nc
#define read_nc ref_nc_unused_read_nc
#include "nc.c" /* -I$(REF): /root/reference/model/nc.c */
#undef write_nc
#undef read_nc

static float **rows_of(const float *flat, int nrows, int ncols) {
  float **g = (float **)malloc((size_t)nrows * sizeof(float *));
  for (int i = 0; i < nrows; i++) g[i] = (float *)flat + (size_t)i * ncols;
  return g;
}

/* compress_2d (nc.c:271-320) on a flat [nrows][ncols] grid; out3 = add_offset, scale_factor */
int ref_nc_pack(int nrows, int ncols, const float *grid, double spval, short *packed, float *offset_scale, short *missing) {
  float **g = rows_of(grid, nrows, ncols);
  compress_2d(g, packed, ncols, nrows, spval, &offset_scale[0], &offset_scale[1], missing);
  free(g);
  return 0;
}

/* decompress_2d (nc.c:247-266) */
int ref_nc_unpack(int nrows, int ncols, const short *packed, float add_offset, float scale_factor, short missing,
                  double spval, float *grid) {
  float **g = rows_of(grid, nrows, ncols);
  decompress_2d((short *)packed, g, ncols, nrows, add_offset, scale_factor, missing, spval);
  free(g);
  return 0;
}
