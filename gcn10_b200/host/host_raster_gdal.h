/* gcn10_b200/host/host_raster_gdal.h -- the optional GDAL input backend (host_raster_gdal.c, -DGCN10_WITH_GDAL). */
#ifndef GCN10_HOST_RASTER_GDAL_H
#define GCN10_HOST_RASTER_GDAL_H
#include <stddef.h>
#include <stdint.h>

typedef struct gh_gdal gh_gdal;

int gh_gdal_open(const char *path, gh_gdal **out, int *w, int *h, double gt[6], char *err, size_t errlen);
int gh_gdal_read_window(gh_gdal *g, int xoff, int yoff, int xcount, int ycount, uint8_t *dst, size_t pitch, char *err,
                        size_t errlen);
void gh_gdal_close(gh_gdal *g);

#endif
