/* oracle/refshim/fake_gdal.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A RAM-backed stand-in for the handful of GDAL / OGR / CPL calls made by the
 * reference's src/raster.c and src/cn.c, so that those two files can be compiled
 * UNMODIFIED from /root/reference/src and executed on synthetic rasters:
 *
 *   - "datasets" opened by path are looked up in a small registry of in-memory
 *     byte rasters (path, pointer, width, height, geotransform);
 *   - GDALRasterIO(GF_Read) copies the requested window out of the registered
 *     raster (this is what raster.c:177-179 asks GDAL to do);
 *   - GDALCreate + GDALRasterIO(GF_Write) (raster.c:210-219) hand each finished
 *     plane to a sink owned by ref_api.c, in creation order, together with the
 *     path and the geotransform the reference set on it;
 *   - the OGR calls (cn.c:155-184) serve bounding boxes from a block table
 *     keyed by the integer in the "\"ID\"=<n>" attribute filter.
 *
 * Nothing here computes curve numbers: every CN byte that reaches the sink was
 * produced by the reference's own object code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "gdal.h"
#include "refshim.h"

/* ---------------------------------------------------------------- registry */

#define MAX_RASTERS 8
#define MAX_BLOCKS  4096

typedef struct {
    char path[256];
    const uint8_t *data;        /* may be NULL: window-math-only probes */
    int w, h;
    double t[6];
    int used;
} mem_raster;

typedef struct {
    int kind;                   /* 0 = opened for read, 1 = created for write */
    mem_raster *src;
    char path[512];
    int w, h;
    double t[6];
    int has_t;
} fake_ds;

typedef struct { int id; double minx, miny, maxx, maxy; } block_row;

static mem_raster g_rasters[MAX_RASTERS];
static block_row g_blocks[MAX_BLOCKS];
static int g_nblocks = 0;
static char g_blocks_path[256];

/* layer state: one data source open at a time is all the reference needs */
static int g_filter_id = -1;    /* -1: no filter */
static int g_cursor = 0;

static refshim_sink g_sink;
static refshim_lastread g_lastread;

void refshim_reset(void)
{
    memset(g_rasters, 0, sizeof(g_rasters));
    g_nblocks = 0;
    g_blocks_path[0] = 0;
    g_filter_id = -1;
    g_cursor = 0;
    memset(&g_sink, 0, sizeof(g_sink));
    memset(&g_lastread, 0, sizeof(g_lastread));
}

int refshim_add_raster(const char *path, const uint8_t *data, int w, int h, const double t[6])
{
    for (int i = 0; i < MAX_RASTERS; i++) {
        if (!g_rasters[i].used) {
            snprintf(g_rasters[i].path, sizeof(g_rasters[i].path), "%s", path);
            g_rasters[i].data = data;
            g_rasters[i].w = w;
            g_rasters[i].h = h;
            memcpy(g_rasters[i].t, t, sizeof(double) * 6);
            g_rasters[i].used = 1;
            return 0;
        }
    }
    return -1;
}

int refshim_set_blocks(const char *path, int n, const int *ids, const double *bboxes /* n x (minx,miny,maxx,maxy) */)
{
    if (n > MAX_BLOCKS)
        return -1;
    snprintf(g_blocks_path, sizeof(g_blocks_path), "%s", path);
    for (int i = 0; i < n; i++) {
        g_blocks[i].id = ids[i];
        g_blocks[i].minx = bboxes[4 * i + 0];
        g_blocks[i].miny = bboxes[4 * i + 1];
        g_blocks[i].maxx = bboxes[4 * i + 2];
        g_blocks[i].maxy = bboxes[4 * i + 3];
    }
    g_nblocks = n;
    return 0;
}

refshim_sink *refshim_get_sink(void) { return &g_sink; }
const refshim_lastread *refshim_get_lastread(void) { return &g_lastread; }

/* ------------------------------------------------------------------- GDAL */

void GDALAllRegister(void) {}
void OGRRegisterAll(void) {}

GDALDatasetH GDALOpen(const char *path, GDALAccess access)
{
    (void)access;
    for (int i = 0; i < MAX_RASTERS; i++) {
        if (g_rasters[i].used && strcmp(g_rasters[i].path, path) == 0) {
            fake_ds *ds = calloc(1, sizeof(*ds));
            ds->kind = 0;
            ds->src = &g_rasters[i];
            ds->w = g_rasters[i].w;
            ds->h = g_rasters[i].h;
            memcpy(ds->t, g_rasters[i].t, sizeof(ds->t));
            ds->has_t = 1;
            return ds;
        }
    }
    return NULL;
}

void GDALClose(GDALDatasetH h) { free(h); }

CPLErr GDALGetGeoTransform(GDALDatasetH h, double *t)
{
    fake_ds *ds = h;
    memcpy(t, ds->t, sizeof(double) * 6);
    return CE_None;
}

CPLErr GDALSetGeoTransform(GDALDatasetH h, double *t)
{
    fake_ds *ds = h;
    memcpy(ds->t, t, sizeof(double) * 6);
    ds->has_t = 1;
    if (ds->kind == 1)
        memcpy(g_sink.gt, t, sizeof(double) * 6);
    return CE_None;
}

int GDALGetRasterXSize(GDALDatasetH h) { return ((fake_ds *)h)->w; }
int GDALGetRasterYSize(GDALDatasetH h) { return ((fake_ds *)h)->h; }
const char *GDALGetProjectionRef(GDALDatasetH h) { (void)h; return "GEOGCS[\"refshim WGS 84\"]"; }
CPLErr GDALSetProjection(GDALDatasetH h, const char *wkt) { (void)h; (void)wkt; return CE_None; }
GDALRasterBandH GDALGetRasterBand(GDALDatasetH h, int band) { (void)band; return h; }

CPLErr GDALRasterIO(GDALRasterBandH band, GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize,
                    void *buf, int bxsize, int bysize, GDALDataType type, int pixel_space, int line_space)
{
    fake_ds *ds = band;
    /* 0 = GDAL's default spacing (what raster.c passes): packed bytes, rows of bxsize */
    if (type != GDT_Byte || bxsize != xsize || bysize != ysize || (pixel_space != 0 && pixel_space != 1) ||
        (line_space != 0 && line_space < xsize))
        return CE_Failure;
    const size_t row_bytes = line_space ? (size_t)line_space : (size_t)xsize;
    if (rw == GF_Read) {
        if (ds->kind != 0)
            return CE_Failure;
        if (xoff < 0 || yoff < 0 || xoff + xsize > ds->w || yoff + ysize > ds->h)
            return CE_Failure;
        g_lastread.xoff = xoff;
        g_lastread.yoff = yoff;
        g_lastread.xsize = xsize;
        g_lastread.ysize = ysize;
        if (ds->src->data) {
            for (int y = 0; y < ysize; y++)
                memcpy((uint8_t *)buf + (size_t)y * row_bytes,
                       ds->src->data + (size_t)(yoff + y) * ds->w + xoff, (size_t)xsize);
        }
        else {
            for (int y = 0; y < ysize; y++)
                memset((uint8_t *)buf + (size_t)y * row_bytes, 0, (size_t)xsize);
        }
        return CE_None;
    }
    /* GF_Write on a created dataset: deliver the plane to the sink */
    if (ds->kind != 1 || xoff != 0 || yoff != 0 || xsize != ds->w || ysize != ds->h)
        return CE_Failure;
    refshim_sink_deliver(&g_sink, ds->path, buf, xsize, ysize);
    return CE_None;
}

GDALDriverH GDALGetDriverByName(const char *name)
{
    static int gtiff_driver;
    return strcmp(name, "GTiff") == 0 ? (GDALDriverH)&gtiff_driver : NULL;
}

GDALDatasetH GDALCreate(GDALDriverH drv, const char *path, int xsize, int ysize, int bands,
                        GDALDataType type, char **opts)
{
    (void)drv; (void)opts;
    if (bands != 1 || type != GDT_Byte)
        return NULL;
    fake_ds *ds = calloc(1, sizeof(*ds));
    ds->kind = 1;
    ds->w = xsize;
    ds->h = ysize;
    snprintf(ds->path, sizeof(ds->path), "%s", path);
    return ds;
}

/* creation options are only recorded (NAME=VALUE strings) so a test can assert
 * that the reference asked for COMPRESS=DEFLATE and TILED=YES (raster.c:206-207) */
char **CSLSetNameValue(char **list, const char *name, const char *value)
{
    int n = 0;
    if (list)
        while (list[n])
            n++;
    list = realloc(list, sizeof(char *) * (n + 2));
    size_t len = strlen(name) + strlen(value) + 2;
    list[n] = malloc(len);
    snprintf(list[n], len, "%s=%s", name, value);
    list[n + 1] = NULL;
    snprintf(g_sink.last_options[n < 4 ? n : 3], sizeof(g_sink.last_options[0]), "%s", list[n]);
    return list;
}

void CSLDestroy(char **list)
{
    if (!list)
        return;
    for (int i = 0; list[i]; i++)
        free(list[i]);
    free(list);
}

void CPLFree(void *p) { free(p); }

OGRSpatialReferenceH OSRNewSpatialReference(const char *wkt)
{
    /* the reference never destroys these handles (raster.c:165); keep them tiny */
    static char handle[] = "GEOGCS[\"refshim WGS 84\"]";
    (void)wkt;
    return handle;
}

OGRErr OSRExportToWkt(OGRSpatialReferenceH srs, char **wkt)
{
    *wkt = strdup((const char *)srs);
    return 0;
}

/* -------------------------------------------------------------------- OGR */

static int ds_token, layer_token;

OGRDataSourceH OGROpen(const char *path, int update, OGRSFDriverH *drv)
{
    (void)update; (void)drv;
    if (!g_blocks_path[0] || strcmp(path, g_blocks_path) != 0)
        return NULL;
    g_filter_id = -1;
    g_cursor = 0;
    return &ds_token;
}

void OGR_DS_Destroy(OGRDataSourceH ds) { (void)ds; }
OGRLayerH OGR_DS_GetLayer(OGRDataSourceH ds, int i) { (void)ds; (void)i; return &layer_token; }

OGRErr OGR_L_SetAttributeFilter(OGRLayerH layer, const char *filter)
{
    int id;
    (void)layer;
    g_cursor = 0;
    if (filter && sscanf(filter, "\"ID\"=%d", &id) == 1)
        g_filter_id = id;
    else
        g_filter_id = -1;
    return 0;
}

OGRFeatureH OGR_L_GetNextFeature(OGRLayerH layer)
{
    (void)layer;
    while (g_cursor < g_nblocks) {
        block_row *b = &g_blocks[g_cursor++];
        if (g_filter_id < 0 || b->id == g_filter_id)
            return b;
    }
    return NULL;
}

void OGR_L_ResetReading(OGRLayerH layer) { (void)layer; g_cursor = 0; }
int OGR_L_GetFeatureCount(OGRLayerH layer, int force) { (void)layer; (void)force; return g_nblocks; }
OGRGeometryH OGR_F_GetGeometryRef(OGRFeatureH feat) { return feat; }

void OGR_G_GetEnvelope(OGRGeometryH geom, OGREnvelope *env)
{
    block_row *b = geom;
    env->MinX = b->minx;
    env->MaxX = b->maxx;
    env->MinY = b->miny;
    env->MaxY = b->maxy;
}

void OGR_F_Destroy(OGRFeatureH feat) { (void)feat; }
int OGR_F_GetFieldIndex(OGRFeatureH feat, const char *name) { (void)feat; return strcmp(name, "ID") == 0 ? 1 : -1; }
int OGR_F_GetFieldAsInteger(OGRFeatureH feat, int field) { return field == 1 ? ((block_row *)feat)->id : 0; }
