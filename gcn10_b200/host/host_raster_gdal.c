/* gcn10_b200/host/host_raster_gdal.c -- optional input backend: open a raster with GDAL itself.
 *
 * Built only with -DGCN10_WITH_GDAL (make GDAL=1: flags from gdal-config).  The program's own readers
 * (host_tiff.c, host_raster.c) cover what the reference's shipped configuration uses -- DEFLATE GeoTIFFs and the VRT
 * mosaic of them -- and hand the COMPRESSED tiles to the GPU.  Everything else GDALOpen() accepts (remote /vsicurl/
 * sources, LZW / ZSTD / JPEG tiles, other formats, warped VRTs) goes through this file: the calls are the ones the
 * reference makes in /root/reference/src/raster.c:106-189 -- GDALOpen, GDALGetGeoTransform, GDALGetRasterBand(1),
 * GDALGetRasterXSize / YSize, one GDALRasterIO(GF_Read, GDT_Byte) per window -- and the decoded bytes enter the
 * library through the raster-in entry point (gcn10_cuda_block_deflate_rows) instead of the compressed-tile one.
 * Output stays with host_tiff.c: the files are the tiled DEFLATE GeoTIFFs of raster.c:191-227 either way; the
 * spatial reference of a GDAL-opened input is not carried over as GeoKeys (raster.c:164-165 copies the WKT), the
 * writer's EPSG:4326 default stands.
 *
 * A GDAL dataset handle must not be used from two threads at once: reads are serialised per handle.
 */
#ifdef GCN10_WITH_GDAL
#include "host_raster_gdal.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <gdal.h>

struct gh_gdal {
    GDALDatasetH ds;
    GDALRasterBandH band;
    int w, h;
    pthread_mutex_t lock;
};

static pthread_once_t g_register_once = PTHREAD_ONCE_INIT;
static void register_drivers(void) { GDALAllRegister(); }       /* raster.c:14-19, called at raster.c:118 */

int gh_gdal_open(const char *path, gh_gdal **out, int *w, int *h, double gt[6], char *err, size_t errlen)
{
    *out = NULL;
    pthread_once(&g_register_once, register_drivers);
    GDALDatasetH ds = GDALOpen(path, GA_ReadOnly);              /* raster.c:119 */
    if (!ds) {
        if (err && errlen)
            snprintf(err, errlen, "gdal open failed: %s", path);        /* raster.c:121 */
        return -1;
    }
    gh_gdal *g = calloc(1, sizeof *g);
    if (!g) {
        GDALClose(ds);
        return -1;
    }
    g->ds = ds;
    /* raster.c:126 ignores the result: a dataset without a geotransform keeps GDAL's identity default */
    if (GDALGetGeoTransform(ds, gt) != CE_None) {
        gt[0] = 0.0, gt[1] = 1.0, gt[2] = 0.0;
        gt[3] = 0.0, gt[4] = 0.0, gt[5] = 1.0;
    }
    g->band = GDALGetRasterBand(ds, 1);                         /* raster.c:177 */
    g->w = GDALGetRasterXSize(ds);
    g->h = GDALGetRasterYSize(ds);
    if (!g->band || g->w <= 0 || g->h <= 0) {
        if (err && errlen)
            snprintf(err, errlen, "gdal open failed: %s (no band 1)", path);
        GDALClose(ds);
        free(g);
        return -1;
    }
    pthread_mutex_init(&g->lock, NULL);
    *w = g->w;
    *h = g->h;
    *out = g;
    return 0;
}

int gh_gdal_read_window(gh_gdal *g, int xoff, int yoff, int xcount, int ycount, uint8_t *dst, size_t pitch, char *err,
                        size_t errlen)
{
    if (xoff < 0 || yoff < 0 || xcount <= 0 || ycount <= 0 || (long long)xoff + xcount > g->w ||
        (long long)yoff + ycount > g->h || pitch < (size_t)xcount || pitch > 0x7FFFFFFFu) {
        if (err && errlen)
            snprintf(err, errlen, "gdalrasterio error 3 (window %d,%d %dx%d outside %dx%d)", xoff, yoff, xcount, ycount,
                     g->w, g->h);
        return -1;
    }
    pthread_mutex_lock(&g->lock);
    /* raster.c:177-179, with the caller's row pitch as the line spacing */
    const CPLErr e = GDALRasterIO(g->band, GF_Read, xoff, yoff, xcount, ycount, dst, xcount, ycount, GDT_Byte, 1, (int)pitch);
    pthread_mutex_unlock(&g->lock);
    if (e != CE_None) {
        if (err && errlen)
            snprintf(err, errlen, "gdalrasterio error %d", (int)e);     /* raster.c:182 */
        return -1;
    }
    return 0;
}

void gh_gdal_close(gh_gdal *g)
{
    if (!g)
        return;
    GDALClose(g->ds);
    pthread_mutex_destroy(&g->lock);
    free(g);
}
#else
typedef int gcn10_host_raster_gdal_not_built;   /* ISO C forbids an empty translation unit */
#endif
