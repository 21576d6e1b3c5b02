/* gcn10_b200/host/host_tiff.c -- minimal GeoTIFF reader / writer for the gcn10 host program.
 *
 * Stands in for the two GDAL calls on either side of the hot path, neither of which is available
 * in this image (no libgdal):
 *   load_raster()  /root/reference/src/raster.c:106-189  -> gh_tiff_open + gh_tiff_read_window
 *   save_raster()  /root/reference/src/raster.c:192-227  -> gh_tiffw_* / gh_tiff_write
 *
 * Reader: single band, 8 bit, strips or tiles, classic TIFF or BigTIFF, either byte order,
 * compression none (1), LZW (5), DEFLATE (8 / 32946), predictor 1 or 2.
 * Writer: what GDAL's GTiff driver produces for COMPRESS=DEFLATE, TILED=YES with defaults
 * (raster.c:206-207): 256 x 256 tiles, zlib level 6, no predictor, PixelIsArea, EPSG:4326 keys,
 * no NoData tag.  Tiles are compressed in parallel (pthreads) and appended band by band, so a
 * 36000 x 36000 plane never has to be resident in full.
 */
#define _GNU_SOURCE
#include "gcn10_host.h"
#include "host_tiff.h"

#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <zlib.h>

static void set_err(char *err, size_t errlen, const char *fmt, ...)
{
    if (!err || !errlen)
        return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

/* ------------------------------------------------------------------------------- parallel for */

typedef void (*job_fn)(void *arg, int index);

/* Every calling thread (a GPU worker, its loader, its band reader) owns a small pool of helper threads that lives as
 * long as the caller: the tile sinks run a parallel-for per strip, 18 strips per block, and creating the helpers anew
 * each time cost more than the work they did.  A job is "run fn(arg, i) for i in [0, n)"; indices are handed out under
 * the pool's mutex, the caller takes part, and returns when all indices are done. */
typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t work, idle;
    pthread_t *th;
    int nthreads;               /* helpers */
    job_fn fn;
    void *arg;
    int n, next, active, generation, quit;
    int limit;                  /* helpers allowed to join the current job */
} job_pool;

static pthread_key_t g_pool_key;
static pthread_once_t g_pool_once = PTHREAD_ONCE_INIT;

static void *pool_helper(void *p)
{
    job_pool *jp = p;
    int seen = 0;
    pthread_mutex_lock(&jp->mu);
    for (;;) {
        while (!jp->quit && (jp->generation == seen || jp->next >= jp->n || jp->active >= jp->limit))
            pthread_cond_wait(&jp->work, &jp->mu);
        if (jp->quit)
            break;
        seen = jp->generation;
        jp->active++;
        while (jp->next < jp->n) {
            int i = jp->next++;
            pthread_mutex_unlock(&jp->mu);
            jp->fn(jp->arg, i);
            pthread_mutex_lock(&jp->mu);
        }
        jp->active--;
        if (jp->active == 0)
            pthread_cond_signal(&jp->idle);
    }
    pthread_mutex_unlock(&jp->mu);
    return NULL;
}

static void pool_destroy(void *p)
{
    job_pool *jp = p;
    if (!jp)
        return;
    pthread_mutex_lock(&jp->mu);
    jp->quit = 1;
    pthread_cond_broadcast(&jp->work);
    pthread_mutex_unlock(&jp->mu);
    for (int t = 0; t < jp->nthreads; t++)
        pthread_join(jp->th[t], NULL);
    pthread_mutex_destroy(&jp->mu);
    pthread_cond_destroy(&jp->work);
    pthread_cond_destroy(&jp->idle);
    free(jp->th);
    free(jp);
}

static void pool_key_init(void) { pthread_key_create(&g_pool_key, pool_destroy); }

/* the calling thread's pool, grown to at least `helpers` helper threads */
static job_pool *pool_get(int helpers)
{
    pthread_once(&g_pool_once, pool_key_init);
    job_pool *jp = pthread_getspecific(g_pool_key);
    if (!jp) {
        jp = calloc(1, sizeof *jp);
        if (!jp)
            return NULL;
        pthread_mutex_init(&jp->mu, NULL);
        pthread_cond_init(&jp->work, NULL);
        pthread_cond_init(&jp->idle, NULL);
        pthread_setspecific(g_pool_key, jp);
    }
    if (jp->nthreads < helpers) {
        pthread_t *th = realloc(jp->th, sizeof(pthread_t) * (size_t)helpers);
        if (!th)
            return jp;
        jp->th = th;
        while (jp->nthreads < helpers) {
            if (pthread_create(&jp->th[jp->nthreads], NULL, pool_helper, jp) != 0)
                break;
            jp->nthreads++;
        }
    }
    return jp;
}

static void parallel_for(int n, int threads, job_fn fn, void *arg)
{
    if (threads > n)
        threads = n;
    job_pool *jp = threads > 1 ? pool_get(threads - 1) : NULL;
    if (!jp || jp->nthreads == 0) {
        for (int i = 0; i < n; i++)
            fn(arg, i);
        return;
    }
    pthread_mutex_lock(&jp->mu);
    jp->fn = fn;
    jp->arg = arg;
    jp->n = n;
    jp->next = 0;
    jp->limit = threads - 1;
    jp->generation++;
    pthread_cond_broadcast(&jp->work);
    while (jp->next < jp->n) {                 /* the caller works too */
        int i = jp->next++;
        pthread_mutex_unlock(&jp->mu);
        fn(arg, i);
        pthread_mutex_lock(&jp->mu);
    }
    while (jp->active > 0)
        pthread_cond_wait(&jp->idle, &jp->mu);
    jp->n = 0;                                  /* late wakers find nothing to do */
    pthread_mutex_unlock(&jp->mu);
}

void gh_parallel_for(int n, int threads, void (*fn)(void *arg, int index), void *arg)
{
    parallel_for(n, threads, fn, arg);
}

/* ------------------------------------------------------------------------------------- reader */

struct gh_tiff {
    int fd;
    int big, swap;
    int w, h;
    int compression, predictor;
    int tiled, tw, th;          /* tile (or strip: tw = w, th = rows per strip) geometry */
    int tiles_x, tiles_y;
    uint64_t *offsets, *counts;
    uint64_t nchunks;
    uint64_t file_size;
    int has_gt;
    double gt[6];
    /* GeoTIFF georeferencing tags as they lie in the file (34735 key directory, 34736 double params, 34737 ascii
     * params): handed to the outputs verbatim, as the reference copies the source's projection (raster.c:164-165,
     * 212-214) */
    uint16_t *geo_keys;
    size_t n_geo_keys;
    double *geo_doubles;
    size_t n_geo_doubles;
    char *geo_ascii;
    size_t n_geo_ascii;
};

static uint16_t rd16(const gh_tiff *t, const unsigned char *p)
{
    return t->swap ? (uint16_t)(p[0] << 8 | p[1]) : (uint16_t)(p[1] << 8 | p[0]);
}

static uint32_t rd32(const gh_tiff *t, const unsigned char *p)
{
    return t->swap ? ((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3])
                   : ((uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]);
}

static uint64_t rd64(const gh_tiff *t, const unsigned char *p)
{
    uint64_t v = 0;
    if (t->swap)
        for (int i = 0; i < 8; i++) v = v << 8 | p[i];
    else
        for (int i = 7; i >= 0; i--) v = v << 8 | p[i];
    return v;
}

static int pread_all(int fd, void *buf, size_t n, uint64_t off)
{
    unsigned char *p = buf;
    while (n) {
        ssize_t r = pread(fd, p, n, (off_t)off);
        if (r <= 0)
            return -1;
        p += r;
        n -= (size_t)r;
        off += (uint64_t)r;
    }
    return 0;
}

static size_t type_size(int type)
{
    switch (type) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    }
    return 0;
}

/* Fetches the values of one IFD entry as doubles or uint64 (whichever array is non-NULL). */
static int entry_values(gh_tiff *t, int type, uint64_t count, const unsigned char *valfield, size_t valfield_len,
                        uint64_t *u, double *d, uint64_t max)
{
    size_t ts = type_size(type);
    if (!ts || count == 0)
        return -1;
    const int is_inline = count <= valfield_len / ts;            /* decided by the stored count */
    if (count > max)
        count = max;
    /* a table cannot be longer than the file that holds it (a damaged count must not size an allocation) */
    if (!is_inline && count > t->file_size / ts)
        return -1;
    size_t bytes = ts * (size_t)count;
    unsigned char *buf = malloc(bytes);
    if (!buf)
        return -1;
    if (is_inline) {
        memcpy(buf, valfield, bytes);
    }
    else {
        uint64_t off = t->big ? rd64(t, valfield) : rd32(t, valfield);
        if (off > t->file_size || bytes > t->file_size - off || pread_all(t->fd, buf, bytes, off)) {
            free(buf);
            return -1;
        }
    }
    for (uint64_t i = 0; i < count; i++) {
        const unsigned char *p = buf + ts * i;
        uint64_t uv = 0;
        double dv = 0;
        switch (type) {
        case 1: case 2: case 7: uv = p[0]; dv = (double)uv; break;
        case 3: uv = rd16(t, p); dv = (double)uv; break;
        case 4: case 13: uv = rd32(t, p); dv = (double)uv; break;
        case 16: case 18: uv = rd64(t, p); dv = (double)uv; break;
        case 12: { uint64_t b = rd64(t, p); memcpy(&dv, &b, 8); uv = dv >= 0.0 && dv < 1.8e19 ? (uint64_t)dv : 0; break; }
        case 11: { uint32_t b = rd32(t, p); float f; memcpy(&f, &b, 4); dv = f; uv = dv >= 0.0 && dv < 1.8e19 ? (uint64_t)dv : 0; break; }
        default: free(buf); return -1;
        }
        if (u) u[i] = uv;
        if (d) d[i] = dv;
    }
    free(buf);
    return 0;
}

int gh_tiff_open(const char *path, gh_tiff **out, char *err, size_t errlen)
{
    *out = NULL;
    if (err && errlen)
        err[0] = '\0';
    gh_tiff *t = calloc(1, sizeof *t);
    if (!t)
        return -1;
    t->fd = open(path, O_RDONLY);
    unsigned char hdr[16];
    struct stat st;
    if (t->fd < 0 || fstat(t->fd, &st) || st.st_size < 8 || pread_all(t->fd, hdr, 8, 0)) {
        set_err(err, errlen, "gdal open failed: %s", path);                 /* raster.c:121 */
        goto fail;
    }
    t->file_size = (uint64_t)st.st_size;
    if (hdr[0] == 'I' && hdr[1] == 'I') t->swap = 0;
    else if (hdr[0] == 'M' && hdr[1] == 'M') t->swap = 1;
    else { set_err(err, errlen, "gdal open failed: %s (not a TIFF)", path); goto fail; }
    uint16_t magic = rd16(t, hdr + 2);
    uint64_t ifd;
    if (magic == 42) {
        ifd = rd32(t, hdr + 4);
    }
    else if (magic == 43) {
        t->big = 1;
        if (pread_all(t->fd, hdr, 16, 0)) goto fail;
        ifd = rd64(t, hdr + 8);
    }
    else { set_err(err, errlen, "gdal open failed: %s (bad TIFF magic %u)", path, magic); goto fail; }

    unsigned char nb[8];
    if (pread_all(t->fd, nb, t->big ? 8 : 2, ifd)) goto fail;
    uint64_t nent = t->big ? rd64(t, nb) : rd16(t, nb);
    size_t esz = t->big ? 20 : 12;
    if (nent == 0 || nent > 65535 || ifd > t->file_size || esz * nent > t->file_size - ifd) {
        set_err(err, errlen, "gdal open failed: %s (damaged TIFF directory)", path);
        goto fail;
    }
    unsigned char *ents = malloc(esz * (size_t)nent);
    if (!ents || pread_all(t->fd, ents, esz * (size_t)nent, ifd + (t->big ? 8 : 2))) { free(ents); goto fail; }

    int bits = 1, spp = 1, rows_per_strip = 0;
    uint64_t off_tag_count = 0;
    const unsigned char *off_ent = NULL, *cnt_ent = NULL;
    int off_type = 0, cnt_type = 0;
    double scale[3] = { 0 }, tie[6] = { 0 }, xf[16] = { 0 };
    int have_scale = 0, have_tie = 0, have_xf = 0;
    t->compression = 1;
    t->predictor = 1;
    for (uint64_t i = 0; i < nent; i++) {
        const unsigned char *e = ents + esz * i;
        int tag = rd16(t, e), type = rd16(t, e + 2);
        uint64_t count = t->big ? rd64(t, e + 4) : rd32(t, e + 4);
        const unsigned char *val = e + (t->big ? 12 : 8);
        size_t vlen = t->big ? 8 : 4;
        uint64_t u[1] = { 0 };
        switch (tag) {
        case 256: entry_values(t, type, count, val, vlen, u, NULL, 1); t->w = (int)u[0]; break;
        case 257: entry_values(t, type, count, val, vlen, u, NULL, 1); t->h = (int)u[0]; break;
        case 258: entry_values(t, type, count, val, vlen, u, NULL, 1); bits = (int)u[0]; break;
        case 259: entry_values(t, type, count, val, vlen, u, NULL, 1); t->compression = (int)u[0]; break;
        case 277: entry_values(t, type, count, val, vlen, u, NULL, 1); spp = (int)u[0]; break;
        case 278: entry_values(t, type, count, val, vlen, u, NULL, 1); rows_per_strip = (int)(u[0] > 0x7fffffff ? 0x7fffffff : u[0]); break;
        case 317: entry_values(t, type, count, val, vlen, u, NULL, 1); t->predictor = (int)u[0]; break;
        case 322: entry_values(t, type, count, val, vlen, u, NULL, 1); t->tw = (int)u[0]; t->tiled = 1; break;
        case 323: entry_values(t, type, count, val, vlen, u, NULL, 1); t->th = (int)u[0]; t->tiled = 1; break;
        case 273: case 324: off_ent = e; off_type = type; off_tag_count = count; break;
        case 279: case 325: cnt_ent = e; cnt_type = type; break;
        case 33550: have_scale = entry_values(t, type, count, val, vlen, NULL, scale, 3) == 0 && count >= 2; break;
        case 33922: have_tie = entry_values(t, type, count, val, vlen, NULL, tie, 6) == 0 && count >= 6; break;
        case 34264: have_xf = entry_values(t, type, count, val, vlen, NULL, xf, 16) == 0 && count >= 16; break;
        case 34735:
            if (type == 3 && count >= 4 && count < 4096 && !t->geo_keys) {
                uint64_t *tmp = malloc(sizeof(uint64_t) * (size_t)count);
                t->geo_keys = malloc(sizeof(uint16_t) * (size_t)count);
                if (tmp && t->geo_keys && entry_values(t, type, count, val, vlen, tmp, NULL, count) == 0) {
                    for (uint64_t k = 0; k < count; k++)
                        t->geo_keys[k] = (uint16_t)tmp[k];
                    t->n_geo_keys = (size_t)count;
                }
                free(tmp);
            }
            break;
        case 34736:
            if (type == 12 && count > 0 && count < 4096 && !t->geo_doubles) {
                t->geo_doubles = malloc(sizeof(double) * (size_t)count);
                if (t->geo_doubles && entry_values(t, type, count, val, vlen, NULL, t->geo_doubles, count) == 0)
                    t->n_geo_doubles = (size_t)count;
            }
            break;
        case 34737:
            if (type == 2 && count > 0 && count < 65536 && !t->geo_ascii) {
                uint64_t *tmp = malloc(sizeof(uint64_t) * (size_t)count);
                t->geo_ascii = malloc((size_t)count + 1);
                if (tmp && t->geo_ascii && entry_values(t, type, count, val, vlen, tmp, NULL, count) == 0) {
                    for (uint64_t k = 0; k < count; k++)
                        t->geo_ascii[k] = (char)tmp[k];
                    t->geo_ascii[count] = '\0';
                    t->n_geo_ascii = (size_t)count;
                }
                free(tmp);
            }
            break;
        default: break;
        }
    }
    if (t->w <= 0 || t->h <= 0 || bits != 8 || spp != 1 || !off_ent || !cnt_ent) {
        set_err(err, errlen, "gdal open failed: %s (need a single-band 8-bit raster; got %d bits x %d samples)",
                path, bits, spp);
        free(ents);
        goto fail;
    }
    if (t->compression != 1 && t->compression != 5 && t->compression != 8 && t->compression != 32946) {
        set_err(err, errlen, "gdal open failed: %s (unsupported TIFF compression %d)", path, t->compression);
        free(ents);
        goto fail;
    }
    if (!t->tiled) {
        t->tw = t->w;
        t->th = (rows_per_strip <= 0 || rows_per_strip > t->h) ? t->h : rows_per_strip;
    }
    /* tile geometry out of a damaged directory: no division by zero, no tile that cannot be allocated */
    if (t->tw <= 0 || t->th <= 0 || (uint64_t)t->tw * (uint64_t)t->th > (1ull << 33)) {
        set_err(err, errlen, "gdal open failed: %s (bad tile size %d x %d)", path, t->tw, t->th);
        free(ents);
        goto fail;
    }
    t->tiles_x = (int)(((int64_t)t->w + t->tw - 1) / t->tw);
    t->tiles_y = (int)(((int64_t)t->h + t->th - 1) / t->th);
    t->nchunks = (uint64_t)t->tiles_x * (uint64_t)t->tiles_y;
    if (off_tag_count < t->nchunks || t->nchunks > t->file_size) {
        set_err(err, errlen, "gdal open failed: %s (offset table shorter than the tile grid)", path);
        free(ents);
        goto fail;
    }
    t->offsets = malloc(sizeof(uint64_t) * t->nchunks);
    t->counts = malloc(sizeof(uint64_t) * t->nchunks);
    size_t vlen = t->big ? 8 : 4;
    if (!t->offsets || !t->counts ||
        entry_values(t, off_type, off_tag_count, off_ent + (t->big ? 12 : 8), vlen, t->offsets, NULL, t->nchunks) ||
        entry_values(t, cnt_type, off_tag_count, cnt_ent + (t->big ? 12 : 8), vlen, t->counts, NULL, t->nchunks)) {
        set_err(err, errlen, "gdal open failed: %s (cannot read the tile tables)", path);
        free(ents);
        goto fail;
    }
    free(ents);
    if (have_scale && have_tie) {
        /* PixelIsArea north-up: origin = tiepoint world coords minus tiepoint raster coords * scale */
        t->gt[1] = scale[0];
        t->gt[5] = -scale[1];
        t->gt[0] = tie[3] - tie[0] * scale[0];
        t->gt[3] = tie[4] + tie[1] * scale[1];
        t->has_gt = 1;
        /* GTRasterTypeGeoKey (1025) = RasterPixelIsPoint (2): the tiepoint names a pixel CENTRE; GDAL's geotransform
         * is that of the pixel's corner, half a pixel up and to the left */
        for (size_t k = 4; k + 3 < t->n_geo_keys; k += 4)
            if (t->geo_keys[k] == 1025 && t->geo_keys[k + 1] == 0 && t->geo_keys[k + 3] == 2) {
                t->gt[0] -= 0.5 * t->gt[1];
                t->gt[3] -= 0.5 * t->gt[5];
            }
    }
    else if (have_xf) {
        t->gt[0] = xf[3]; t->gt[1] = xf[0]; t->gt[2] = xf[1];
        t->gt[3] = xf[7]; t->gt[4] = xf[4]; t->gt[5] = xf[5];
        t->has_gt = 1;
    }
    else {
        t->gt[1] = 1.0;         /* GDAL's default for an ungeoreferenced raster */
        t->gt[5] = 1.0;
    }
    *out = t;
    return 0;
fail:
    if (err && !*err)
        set_err(err, errlen, "gdal open failed: %s", path);
    gh_tiff_close(t);
    return -1;
}

int gh_tiff_size(const gh_tiff *t, int *w, int *h) { *w = t->w; *h = t->h; return 0; }
int gh_tiff_geotransform(const gh_tiff *t, double gt[6]) { memcpy(gt, t->gt, sizeof t->gt); return t->has_gt ? 0 : 1; }

void gh_tiff_close(gh_tiff *t)
{
    if (!t)
        return;
    if (t->fd >= 0)
        close(t->fd);
    free(t->offsets);
    free(t->counts);
    free(t->geo_keys);
    free(t->geo_doubles);
    free(t->geo_ascii);
    free(t);
}

int gh_tiff_geokeys(const gh_tiff *t, gh_geokeys *out)
{
    memset(out, 0, sizeof *out);
    if (!t->n_geo_keys)
        return 1;
    out->keys = t->geo_keys;
    out->n_keys = t->n_geo_keys;
    out->doubles = t->geo_doubles;
    out->n_doubles = t->n_geo_doubles;
    out->ascii = t->geo_ascii;
    out->n_ascii = t->n_geo_ascii;
    return 0;
}

/* TIFF LZW: MSB-first variable width codes (9..12 bits), 256 = clear, 257 = end of information,
 * width grows one code early ("early change"). */
static int lzw_decode(const unsigned char *src, size_t n, unsigned char *dst, size_t cap)
{
    enum { CLEAR = 256, EOI = 257, FIRST = 258, MAXC = 4096 };
    uint16_t *prefix = malloc(sizeof(uint16_t) * MAXC);
    unsigned char *suffix = malloc(MAXC), *first = malloc(MAXC);
    uint16_t *length = malloc(sizeof(uint16_t) * MAXC);
    if (!prefix || !suffix || !first || !length) { free(prefix); free(suffix); free(first); free(length); return -1; }
    for (int i = 0; i < 256; i++) { prefix[i] = 0xFFFF; suffix[i] = first[i] = (unsigned char)i; length[i] = 1; }
    size_t out = 0, bitpos = 0, nbits = n * 8;
    int width = 9, next = FIRST, prev = -1, rc = 0;
    while (bitpos + (size_t)width <= nbits) {
        uint32_t code = 0;
        for (int b = 0; b < width; b++) {
            size_t bp = bitpos + (size_t)b;
            code = code << 1 | ((src[bp >> 3] >> (7 - (bp & 7))) & 1u);
        }
        bitpos += (size_t)width;
        if (code == EOI)
            break;
        if (code == CLEAR) {
            width = 9;
            next = FIRST;
            prev = -1;
            continue;
        }
        int cur = (int)code;
        if (prev < 0) {
            if (cur >= 256) { rc = -1; break; }
            if (out < cap) dst[out] = (unsigned char)cur;
            out++;
            prev = cur;
            continue;
        }
        int emit;
        if (cur < next) {
            emit = cur;
        }
        else if (cur == next) {
            emit = -1;          /* KwKwK: prev string + its own first byte */
        }
        else { rc = -1; break; }
        /* add table entry prev + first(emit string) */
        unsigned char fb = emit >= 0 ? first[emit] : first[prev];
        if (next < MAXC) {
            prefix[next] = (uint16_t)prev;
            suffix[next] = fb;
            first[next] = first[prev];
            length[next] = (uint16_t)(length[prev] + 1);
            next++;
        }
        int s = emit >= 0 ? emit : next - 1;
        size_t len = length[s];
        if (out + len <= cap) {
            size_t pos = out + len;
            int c = s;
            while (c != 0xFFFF && pos > out) {
                dst[--pos] = suffix[c];
                c = prefix[c];
            }
        }
        out += len;
        prev = cur;
        if (next + 1 >= (1 << width) && width < 12)
            width++;
    }
    free(prefix); free(suffix); free(first); free(length);
    (void)out;
    return rc;
}

typedef struct {
    gh_tiff *t;
    int xoff, yoff, xcount, ycount;
    uint8_t *dst;
    size_t pitch;
    int tx0, ty0, ntx, nty;
    int failed;
} read_job;

static void read_chunk(void *arg, int index)
{
    read_job *j = arg;
    gh_tiff *t = j->t;
    int tx = j->tx0 + index % j->ntx, ty = j->ty0 + index / j->ntx;
    uint64_t ci = (uint64_t)ty * (uint64_t)t->tiles_x + (uint64_t)tx;
    int rows = t->th, cols = t->tw;
    if (!t->tiled && (ty + 1) * t->th > t->h)
        rows = t->h - ty * t->th;           /* the last strip may be short */
    size_t raw = (size_t)rows * (size_t)cols;
    unsigned char *buf = malloc(raw ? raw : 1);
    unsigned char *comp = NULL;
    if (!buf) { j->failed = 1; return; }
    uint64_t n = t->counts[ci];
    if (n == 0) {
        memset(buf, 0, raw);                /* sparse tile: GDAL serves zeros */
    }
    else if (t->compression == 1) {
        size_t take = n < raw ? (size_t)n : raw;
        if (pread_all(t->fd, buf, take, t->offsets[ci])) j->failed = 1;
        if (take < raw) memset(buf + take, 0, raw - take);
    }
    else if (t->offsets[ci] > t->file_size || n > t->file_size - t->offsets[ci]) {
        j->failed = 1;                      /* the tile table points outside the file */
    }
    else {
        comp = malloc((size_t)n);
        if (!comp || pread_all(t->fd, comp, (size_t)n, t->offsets[ci])) {
            j->failed = 1;
        }
        else if (t->compression == 5) {
            memset(buf, 0, raw);
            if (lzw_decode(comp, (size_t)n, buf, raw)) j->failed = 1;
        }
        else {
            uLongf dl = (uLongf)raw;
            int zr = uncompress(buf, &dl, comp, (uLong)n);
            if (zr != Z_OK && zr != Z_BUF_ERROR) j->failed = 1;
            if (dl < raw) memset(buf + dl, 0, raw - dl);
        }
    }
    free(comp);
    if (t->predictor == 2) {
        for (int r = 0; r < rows; r++) {
            unsigned char *p = buf + (size_t)r * cols;
            for (int c = 1; c < cols; c++)
                p[c] = (unsigned char)(p[c] + p[c - 1]);
        }
    }
    /* copy the intersection with the window */
    int cx0 = tx * t->tw, cy0 = ty * t->th;
    int x_lo = cx0 > j->xoff ? cx0 : j->xoff, x_hi = cx0 + cols < j->xoff + j->xcount ? cx0 + cols : j->xoff + j->xcount;
    int y_lo = cy0 > j->yoff ? cy0 : j->yoff, y_hi = cy0 + rows < j->yoff + j->ycount ? cy0 + rows : j->yoff + j->ycount;
    for (int y = y_lo; y < y_hi; y++)
        memcpy(j->dst + (size_t)(y - j->yoff) * j->pitch + (size_t)(x_lo - j->xoff),
               buf + (size_t)(y - cy0) * cols + (size_t)(x_lo - cx0), (size_t)(x_hi - x_lo));
    free(buf);
}

int gh_tiff_read_window(gh_tiff *t, int xoff, int yoff, int xcount, int ycount, uint8_t *dst, size_t pitch,
                        int threads, char *err, size_t errlen)
{
    if (xoff < 0 || yoff < 0 || xcount <= 0 || ycount <= 0 || xoff + xcount > t->w || yoff + ycount > t->h ||
        pitch < (size_t)xcount) {
        set_err(err, errlen, "gdalrasterio error 3 (window %d,%d %dx%d outside %dx%d)", xoff, yoff, xcount, ycount,
                t->w, t->h);
        return -1;
    }
    read_job j = { t, xoff, yoff, xcount, ycount, dst, pitch, 0, 0, 0, 0, 0 };
    j.tx0 = xoff / t->tw;
    j.ty0 = yoff / t->th;
    j.ntx = (xoff + xcount - 1) / t->tw - j.tx0 + 1;
    j.nty = (yoff + ycount - 1) / t->th - j.ty0 + 1;
    parallel_for(j.ntx * j.nty, threads, read_chunk, &j);
    if (j.failed) {
        set_err(err, errlen, "gdalrasterio error 3 (tile decode failed)");
        return -1;
    }
    return 0;
}

/* ---- compressed tiles of a window, untouched (the GPU inflates them) */

int gh_tiff_window_tiles_plan(const gh_tiff *t, int xoff, int yoff, int xcount, int ycount, gh_tile_plan *plan)
{
    if (!t->tiled || (t->compression != 8 && t->compression != 32946) || t->predictor != 1)
        return 1;
    if (xoff < 0 || yoff < 0 || xcount <= 0 || ycount <= 0 || xoff + xcount > t->w || yoff + ycount > t->h)
        return 1;
    plan->tile_w = t->tw;
    plan->tile_h = t->th;
    plan->tx0 = xoff / t->tw;
    plan->ty0 = yoff / t->th;
    plan->tiles_x = (xoff + xcount - 1) / t->tw - plan->tx0 + 1;
    plan->tiles_y = (yoff + ycount - 1) / t->th - plan->ty0 + 1;
    plan->x_in = xoff - plan->tx0 * t->tw;
    plan->y_in = yoff - plan->ty0 * t->th;
    size_t total = 0;
    for (int ty = 0; ty < plan->tiles_y; ty++)
        for (int tx = 0; tx < plan->tiles_x; tx++) {
            const uint64_t ci = (uint64_t)(plan->ty0 + ty) * (uint64_t)t->tiles_x + (uint64_t)(plan->tx0 + tx);
            uint64_t n = t->counts[ci];
            /* a tile the file cannot hold: leave it to the host path, whose read fails like GDAL's would */
            if (n > 0xFFFFFFFFull || t->offsets[ci] > t->file_size || n > t->file_size - t->offsets[ci])
                return 1;
            total += (size_t)n;
        }
    plan->blob_bytes = total;
    return 0;
}

typedef struct {
    gh_tiff *t;
    const gh_tile_plan *plan;
    uint8_t *blob;
    const uint64_t *offsets;
    int failed;
} tiles_job;

/* one job = one tile row of the window: runs of tiles that are adjacent in the file become one pread */
static void read_tile_row(void *arg, int ty)
{
    tiles_job *j = arg;
    const gh_tile_plan *p = j->plan;
    const gh_tiff *t = j->t;
    int tx = 0;
    while (tx < p->tiles_x) {
        uint64_t ci = (uint64_t)(p->ty0 + ty) * (uint64_t)t->tiles_x + (uint64_t)(p->tx0 + tx);
        uint64_t start = t->offsets[ci], len = t->counts[ci];
        int run = 1;
        while (tx + run < p->tiles_x && t->counts[ci + run] > 0 && t->offsets[ci + run] == start + len) {
            len += t->counts[ci + run];
            run++;
        }
        if (len && pread_all(t->fd, j->blob + j->offsets[(size_t)ty * p->tiles_x + tx], (size_t)len, start))
            j->failed = 1;
        tx += run;
    }
}

int gh_tiff_window_tiles_read(gh_tiff *t, const gh_tile_plan *plan, uint8_t *blob, uint64_t *offsets, uint32_t *sizes,
                              int threads, char *err, size_t errlen)
{
    uint64_t pos = 0;
    for (int ty = 0; ty < plan->tiles_y; ty++)
        for (int tx = 0; tx < plan->tiles_x; tx++) {
            uint64_t n = t->counts[(uint64_t)(plan->ty0 + ty) * (uint64_t)t->tiles_x + (uint64_t)(plan->tx0 + tx)];
            offsets[(size_t)ty * plan->tiles_x + tx] = pos;
            sizes[(size_t)ty * plan->tiles_x + tx] = (uint32_t)n;
            pos += n;
        }
    tiles_job j = { t, plan, blob, offsets, 0 };
    parallel_for(plan->tiles_y, threads, read_tile_row, &j);
    if (j.failed) {
        set_err(err, errlen, "gdalrasterio error 3 (tile read failed)");
        return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------- writer */

enum { TILE = 256 };

struct gh_tiffw {
    FILE *fp;
    char *path, *tmp;           /* written as <path>.part, renamed on a successful close, removed otherwise */
    uint16_t *geo_keys;         /* georeferencing tags copied from the land-cover source, or NULL = EPSG:4326 */
    size_t n_geo_keys;
    double *geo_doubles;
    size_t n_geo_doubles;
    char *geo_ascii;
    size_t n_geo_ascii;
    int w, h;
    int tiles_x, tiles_y;
    double gt[6];
    uint32_t *offsets, *counts;
    uint64_t pos;
    int next_tile_row;
    int level;
    int failed;
};

typedef struct {
    gh_tiffw *tw;
    const uint8_t *data;        /* first row of this band */
    size_t pitch;
    int rows;                   /* rows available in the band */
    int tile_row0;
    unsigned char **comp;
    uLongf *comp_len;
} write_job;

static void compress_tile(void *arg, int index)
{
    write_job *j = arg;
    gh_tiffw *tw = j->tw;
    int tx = index % tw->tiles_x, tr = index / tw->tiles_x;
    unsigned char tile[TILE * TILE];
    int x0 = tx * TILE, y0 = tr * TILE;
    int cols = tw->w - x0 < TILE ? tw->w - x0 : TILE;
    int rows = j->rows - y0 < TILE ? j->rows - y0 : TILE;
    if (cols < TILE || rows < TILE)
        memset(tile, 0, sizeof tile);       /* edge tiles are padded with zeros */
    for (int r = 0; r < rows; r++)
        memcpy(tile + r * TILE, j->data + (size_t)(y0 + r) * j->pitch + x0, (size_t)cols);
    uLongf cap = compressBound(sizeof tile);
    unsigned char *out = malloc(cap);
    if (!out || compress2(out, &cap, tile, sizeof tile, tw->level) != Z_OK) {
        free(out);
        out = NULL;
        cap = 0;
        tw->failed = 1;
    }
    j->comp[index] = out;
    j->comp_len[index] = cap;
}

int gh_tiffw_open(const char *path, int w, int h, const double gt[6], gh_tiffw **out, char *err, size_t errlen)
{
    *out = NULL;
    if (w <= 0 || h <= 0) {
        set_err(err, errlen, "write error 3 on %s (empty raster)", path);
        return -1;
    }
    gh_tiffw *tw = calloc(1, sizeof *tw);
    if (!tw)
        return -1;
    /* A block that fails half way (read error, damaged tile, full disk) must not leave a truncated file under the
     * reference's output name: with overwrite off a re-run would keep it and write beside it (cn.c:320-360). */
    tw->path = strdup(path);
    tw->tmp = malloc(strlen(path) + 6);
    if (tw->path && tw->tmp) {
        sprintf(tw->tmp, "%s.part", path);
        tw->fp = fopen(tw->tmp, "wb");
    }
    if (!tw->fp) {
        set_err(err, errlen, "write error 3 on %s (%s)", path, strerror(errno));     /* raster.c:221 */
        free(tw->path);
        free(tw->tmp);
        free(tw);
        return -1;
    }
    setvbuf(tw->fp, NULL, _IOFBF, 1 << 20);     /* tiles arrive as ~1.5 KB pieces: one write(2) per MB, not per tile */
    tw->w = w;
    tw->h = h;
    tw->tiles_x = (w + TILE - 1) / TILE;
    tw->tiles_y = (h + TILE - 1) / TILE;
    tw->level = 6;              /* GDAL's default ZLEVEL */
    memcpy(tw->gt, gt, sizeof tw->gt);
    size_t nt = (size_t)tw->tiles_x * (size_t)tw->tiles_y;
    tw->offsets = calloc(nt, sizeof(uint32_t));
    tw->counts = calloc(nt, sizeof(uint32_t));
    unsigned char hdr[8] = { 'I', 'I', 42, 0, 0, 0, 0, 0 };      /* IFD offset patched on close */
    if (!tw->offsets || !tw->counts || fwrite(hdr, 1, 8, tw->fp) != 8) {
        set_err(err, errlen, "write error 3 on %s", path);
        gh_tiffw_abort(tw);
        return -1;
    }
    tw->pos = 8;
    *out = tw;
    return 0;
}

int gh_tiffw_set_geokeys(gh_tiffw *tw, const gh_geokeys *gk)
{
    if (!tw || !gk || !gk->keys || gk->n_keys < 4)
        return -1;
    free(tw->geo_keys);
    free(tw->geo_doubles);
    free(tw->geo_ascii);
    tw->geo_keys = malloc(sizeof(uint16_t) * gk->n_keys);
    tw->geo_doubles = gk->n_doubles ? malloc(sizeof(double) * gk->n_doubles) : NULL;
    tw->geo_ascii = gk->n_ascii ? malloc(gk->n_ascii) : NULL;
    if (!tw->geo_keys || (gk->n_doubles && !tw->geo_doubles) || (gk->n_ascii && !tw->geo_ascii)) {
        free(tw->geo_keys);
        free(tw->geo_doubles);
        free(tw->geo_ascii);
        tw->geo_keys = NULL;
        tw->geo_doubles = NULL;
        tw->geo_ascii = NULL;
        tw->n_geo_keys = tw->n_geo_doubles = tw->n_geo_ascii = 0;
        return -1;
    }
    memcpy(tw->geo_keys, gk->keys, sizeof(uint16_t) * gk->n_keys);
    tw->n_geo_keys = gk->n_keys;
    if (gk->n_doubles)
        memcpy(tw->geo_doubles, gk->doubles, sizeof(double) * gk->n_doubles);
    tw->n_geo_doubles = gk->n_doubles;
    if (gk->n_ascii)
        memcpy(tw->geo_ascii, gk->ascii, gk->n_ascii);
    tw->n_geo_ascii = gk->n_ascii;
    /* the outputs' tiepoint is the corner of pixel (0, 0): whatever the source said, they are PixelIsArea */
    for (size_t k = 4; k + 3 < tw->n_geo_keys; k += 4)
        if (tw->geo_keys[k] == 1025 && tw->geo_keys[k + 1] == 0)
            tw->geo_keys[k + 3] = 1;
    return 0;
}

int gh_tiffw_write_rows(gh_tiffw *tw, const uint8_t *data, size_t pitch, int y0, int nrows, int threads)
{
    if (tw->failed || y0 != tw->next_tile_row * TILE || y0 + nrows > tw->h ||
        (nrows % TILE != 0 && y0 + nrows != tw->h))
        return -1;
    int tile_rows = (nrows + TILE - 1) / TILE;
    int n = tile_rows * tw->tiles_x;
    write_job j = { tw, data, pitch, nrows, tw->next_tile_row, NULL, NULL };
    j.comp = calloc((size_t)n, sizeof(unsigned char *));
    j.comp_len = calloc((size_t)n, sizeof(uLongf));
    if (!j.comp || !j.comp_len) {
        free(j.comp);
        free(j.comp_len);
        tw->failed = 1;
        return -1;
    }
    parallel_for(n, threads, compress_tile, &j);
    for (int i = 0; i < n; i++) {
        size_t ti = (size_t)(tw->next_tile_row + i / tw->tiles_x) * (size_t)tw->tiles_x + (size_t)(i % tw->tiles_x);
        if (!tw->failed && j.comp[i]) {
            if (tw->pos + j.comp_len[i] > 0xFFFFFFF0ull || fwrite(j.comp[i], 1, j.comp_len[i], tw->fp) != j.comp_len[i])
                tw->failed = 1;
            tw->offsets[ti] = (uint32_t)tw->pos;
            tw->counts[ti] = (uint32_t)j.comp_len[i];
            tw->pos += j.comp_len[i];
        }
        free(j.comp[i]);
    }
    free(j.comp);
    free(j.comp_len);
    tw->next_tile_row += tile_rows;
    return tw->failed ? -1 : 0;
}

int gh_tiffw_put_tile_row(gh_tiffw *tw, int tile_row, const uint8_t *blob, const uint64_t *offsets,
                          const uint32_t *sizes)
{
    if (tw->failed || tile_row != tw->next_tile_row || tile_row >= tw->tiles_y)
        return -1;
    for (int tx = 0; tx < tw->tiles_x; tx++) {
        size_t ti = (size_t)tile_row * (size_t)tw->tiles_x + (size_t)tx;
        if (tw->pos + sizes[tx] > 0xFFFFFFF0ull || fwrite(blob + offsets[tx], 1, sizes[tx], tw->fp) != sizes[tx]) {
            tw->failed = 1;
            return -1;
        }
        tw->offsets[ti] = (uint32_t)tw->pos;
        tw->counts[ti] = sizes[tx];
        tw->pos += sizes[tx];
    }
    tw->next_tile_row++;
    return 0;
}

int gh_tiffw_put_tile_rows(gh_tiffw *tw, int tile_row0, int nrows, const uint8_t *blob, const uint64_t *offsets,
                           const uint32_t *sizes)
{
    if (tw->failed || tile_row0 != tw->next_tile_row || nrows < 1 || tile_row0 + nrows > tw->tiles_y)
        return -1;
    const size_t n = (size_t)nrows * (size_t)tw->tiles_x;
    /* laid out in table order with small gaps (the library's "ordered" strips)?  Then the whole share is one write;
     * the gaps (< 16 bytes per tile, alignment padding) go into the file as unused bytes, which TIFF allows */
    int ordered = 1;
    for (size_t i = 0; i + 1 < n && ordered; i++)
        ordered = offsets[i + 1] >= offsets[i] + sizes[i] && offsets[i + 1] - (offsets[i] + sizes[i]) < 64;
    if (!ordered) {
        for (int r = 0; r < nrows; r++)
            if (gh_tiffw_put_tile_row(tw, tile_row0 + r, blob, offsets + (size_t)r * tw->tiles_x, sizes + (size_t)r * tw->tiles_x))
                return -1;
        return 0;
    }
    const uint64_t first = offsets[0], bytes = offsets[n - 1] + sizes[n - 1] - first;
    if (tw->pos + bytes > 0xFFFFFFF0ull || fwrite(blob + first, 1, (size_t)bytes, tw->fp) != (size_t)bytes) {
        tw->failed = 1;
        return -1;
    }
    for (size_t i = 0; i < n; i++) {
        const size_t ti = (size_t)tile_row0 * (size_t)tw->tiles_x + i;
        tw->offsets[ti] = (uint32_t)(tw->pos + (offsets[i] - first));
        tw->counts[ti] = sizes[i];
    }
    tw->pos += bytes;
    tw->next_tile_row += nrows;
    return 0;
}

static void put16(unsigned char **p, uint16_t v) { (*p)[0] = (unsigned char)v; (*p)[1] = (unsigned char)(v >> 8); *p += 2; }
static void put32(unsigned char **p, uint32_t v) { for (int i = 0; i < 4; i++) (*p)[i] = (unsigned char)(v >> (8 * i)); *p += 4; }
static void put_f64(unsigned char **p, double d) { uint64_t v; memcpy(&v, &d, 8); for (int i = 0; i < 8; i++) (*p)[i] = (unsigned char)(v >> (8 * i)); *p += 8; }

static void put_entry(unsigned char **p, uint16_t tag, uint16_t type, uint32_t count, uint32_t value)
{
    put16(p, tag);
    put16(p, type);
    put32(p, count);
    if (type == 3 && count == 1) { put16(p, (uint16_t)value); put16(p, 0); }
    else put32(p, value);
}

int gh_tiffw_close(gh_tiffw *tw)
{
    if (!tw)
        return -1;
    int rc = -1;
    size_t nt = (size_t)tw->tiles_x * (size_t)tw->tiles_y;
    if (!tw->failed && tw->next_tile_row == tw->tiles_y) {
        /* trailing data area: tile offsets, tile byte counts, pixel scale, tiepoint, geo keys, ascii */
        static const uint16_t geokeys[] = {
            1, 1, 0, 4,
            1024, 0, 1, 2,          /* GTModelTypeGeoKey      = ModelTypeGeographic */
            1025, 0, 1, 1,          /* GTRasterTypeGeoKey     = RasterPixelIsArea   */
            2048, 0, 1, 4326,       /* GeographicTypeGeoKey   = WGS 84              */
            2054, 0, 1, 9102,       /* GeogAngularUnitsGeoKey = degree              */
        };
        static const char ascii_default[] = "WGS 84|";
        const uint16_t *keys = tw->geo_keys ? tw->geo_keys : geokeys;
        const size_t nkeys = tw->geo_keys ? tw->n_geo_keys : sizeof geokeys / 2;
        const char *ascii = tw->geo_keys ? tw->geo_ascii : ascii_default;
        const size_t nascii = tw->geo_keys ? tw->n_geo_ascii : sizeof ascii_default;
        const size_t ndbl = tw->geo_keys ? tw->n_geo_doubles : 0;
        if (tw->pos & 1) { fputc(0, tw->fp); tw->pos++; }
        uint64_t off_offsets = tw->pos, off_counts = off_offsets + 4 * nt, off_scale = off_counts + 4 * nt;
        uint64_t off_tie = off_scale + 24, off_keys = off_tie + 48, off_dbl = off_keys + 2 * nkeys;
        uint64_t off_ascii = off_dbl + 8 * ndbl;
        uint64_t off_ifd = (off_ascii + nascii + 1) & ~(uint64_t)1;
        const int nent = 14 + 1 + (ndbl ? 1 : 0) + (nascii ? 1 : 0);
        size_t tail = (size_t)(off_ifd - tw->pos) + 2 + (size_t)nent * 12 + 4;
        unsigned char *buf = calloc(1, tail), *p = buf;
        if (buf && off_ifd + 256 < 0xFFFFFFF0ull) {
            for (size_t i = 0; i < nt; i++) put32(&p, tw->offsets[i]);
            for (size_t i = 0; i < nt; i++) put32(&p, tw->counts[i]);
            put_f64(&p, tw->gt[1]); put_f64(&p, -tw->gt[5]); put_f64(&p, 0.0);
            put_f64(&p, 0.0); put_f64(&p, 0.0); put_f64(&p, 0.0);
            put_f64(&p, tw->gt[0]); put_f64(&p, tw->gt[3]); put_f64(&p, 0.0);
            for (size_t i = 0; i < nkeys; i++) put16(&p, keys[i]);
            for (size_t i = 0; i < ndbl; i++) put_f64(&p, tw->geo_doubles[i]);
            if (nascii)
                memcpy(p, ascii, nascii);
            p = buf + (off_ifd - tw->pos);
            put16(&p, (uint16_t)nent);
            put_entry(&p, 256, 4, 1, (uint32_t)tw->w);
            put_entry(&p, 257, 4, 1, (uint32_t)tw->h);
            put_entry(&p, 258, 3, 1, 8);
            put_entry(&p, 259, 3, 1, 8);                /* Adobe DEFLATE, what GDAL writes */
            put_entry(&p, 262, 3, 1, 1);                /* BlackIsZero */
            put_entry(&p, 277, 3, 1, 1);
            put_entry(&p, 284, 3, 1, 1);
            put_entry(&p, 322, 3, 1, TILE);
            put_entry(&p, 323, 3, 1, TILE);
            put_entry(&p, 324, 4, (uint32_t)nt, nt == 1 ? tw->offsets[0] : (uint32_t)off_offsets);
            put_entry(&p, 325, 4, (uint32_t)nt, nt == 1 ? tw->counts[0] : (uint32_t)off_counts);
            put_entry(&p, 339, 3, 1, 1);                /* SampleFormat = unsigned */
            put_entry(&p, 33550, 12, 3, (uint32_t)off_scale);
            put_entry(&p, 33922, 12, 6, (uint32_t)off_tie);
            put_entry(&p, 34735, 3, (uint32_t)nkeys, (uint32_t)off_keys);
            if (ndbl)
                put_entry(&p, 34736, 12, (uint32_t)ndbl, (uint32_t)off_dbl);
            if (nascii)
                put_entry(&p, 34737, 2, (uint32_t)nascii, (uint32_t)off_ascii);
            put32(&p, 0);
            if (fwrite(buf, 1, tail, tw->fp) == tail && fseek(tw->fp, 4, SEEK_SET) == 0) {
                unsigned char o[4], *q = o;
                put32(&q, (uint32_t)off_ifd);
                if (fwrite(o, 1, 4, tw->fp) == 4)
                    rc = 0;
            }
        }
        free(buf);
    }
    if (fclose(tw->fp) != 0)
        rc = -1;
    if (rc == 0 && rename(tw->tmp, tw->path) != 0)
        rc = -1;
    if (rc != 0)
        unlink(tw->tmp);
    free(tw->geo_keys);
    free(tw->geo_doubles);
    free(tw->geo_ascii);
    free(tw->path);
    free(tw->tmp);
    free(tw->offsets);
    free(tw->counts);
    free(tw);
    return rc;
}

void gh_tiffw_abort(gh_tiffw *tw)
{
    if (!tw)
        return;
    if (tw->fp)
        fclose(tw->fp);
    if (tw->tmp)
        unlink(tw->tmp);
    free(tw->geo_keys);
    free(tw->geo_doubles);
    free(tw->geo_ascii);
    free(tw->path);
    free(tw->tmp);
    free(tw->offsets);
    free(tw->counts);
    free(tw);
}

int gh_tiff_write(const char *path, const uint8_t *data, int w, int h, size_t pitch, const double gt[6],
                  int threads, char *err, size_t errlen)
{
    gh_tiffw *tw;
    if (gh_tiffw_open(path, w, h, gt, &tw, err, errlen))
        return -1;
    /* bands of 16 tile rows keep the compressed scratch small */
    for (int y0 = 0; y0 < h; y0 += 16 * TILE) {
        int rows = h - y0 < 16 * TILE ? h - y0 : 16 * TILE;
        if (gh_tiffw_write_rows(tw, data + (size_t)y0 * pitch, pitch, y0, rows, threads)) {
            set_err(err, errlen, "write error 3 on %s", path);
            gh_tiffw_abort(tw);
            return -1;
        }
    }
    if (gh_tiffw_close(tw)) {
        set_err(err, errlen, "write error 3 on %s", path);
        return -1;
    }
    return 0;
}
