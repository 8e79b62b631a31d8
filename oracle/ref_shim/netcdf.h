/* Empty shim: netCDF is absent in this image; the reference hot path never calls it except write_nc (stubbed in ref_harness.c). */
#define nc_strerror(e) "netcdf absent"
