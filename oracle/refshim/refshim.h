/* oracle/refshim/refshim.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Internal glue between fake_gdal.c (the RAM GDAL/OGR stand-in) and ref_api.c
 * (the ctypes-facing entry points of oracle/_ref/libgcn10_ref.so). */
#ifndef REFSHIM_H
#define REFSHIM_H
#include <stddef.h>
#include <stdint.h>

#define REFSHIM_MAX_PLANES 18

typedef struct {
    uint8_t *out;               /* NULL: planes are counted and timed but not kept */
    size_t plane_capacity;      /* bytes available per plane in out */
    int count;                  /* planes delivered so far */
    int w, h;                   /* size of the last plane */
    double gt[6];               /* geotransform set on the last created dataset */
    char paths[REFSHIM_MAX_PLANES][128];
    char last_options[4][64];
    double t_start, t_plane[REFSHIM_MAX_PLANES];
    int overflow;
} refshim_sink;

typedef struct { int xoff, yoff, xsize, ysize; } refshim_lastread;

void refshim_reset(void);
int refshim_add_raster(const char *path, const uint8_t *data, int w, int h, const double t[6]);
int refshim_set_blocks(const char *path, int n, const int *ids, const double *bboxes);
refshim_sink *refshim_get_sink(void);
const refshim_lastread *refshim_get_lastread(void);
void refshim_sink_deliver(refshim_sink *s, const char *path, const void *buf, int w, int h);
double refshim_now(void);

#endif
