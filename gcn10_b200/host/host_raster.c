/* gcn10_b200/host/host_raster.c -- what the reference opens with GDALOpen() (/root/reference/src/raster.c:119):
 * a GeoTIFF, or a GDAL VRT mosaic of GeoTIFFs.
 *
 * The reference's shipped configuration points esa_data_path at landcover/esa_worldcover_2021.vrt: one
 * 4 320 000 x 1 728 000 Byte band assembled from 2651 36000 x 36000 source files, each placed 1:1 by a
 * <ComplexSource> with <SrcRect> / <DstRect> (no scaling, NODATA 0 on a band whose NoDataValue is 0).  GDAL is not
 * in this image, so the subset of the VRT format that file uses is read here: <VRTDataset rasterXSize rasterYSize>,
 * <GeoTransform>, the first <VRTRasterBand>'s <NoDataValue> and its <SimpleSource> / <ComplexSource> elements with
 * <SourceFilename relativeToVRT>, <SourceBand> (1), <SrcRect> and <DstRect> of equal size.  A window read is the
 * fill value plus, source by source in file order, the intersection of the window with the source's DstRect --
 * either decoded on the host (gh_raster_read_window) or handed over as the sources' compressed tiles
 * (gh_raster_window_parts -> gcn10_cuda_block_parts_deflate, the GPU inflates them).
 *
 * Source files are opened on first use and stay open (one descriptor + tile tables each, at most
 * GH_RASTER_MAX_OPEN at a time, least recently used first out): a worker that walks neighbouring blocks re-reads
 * no IFD.  /vsicurl/ and http(s) sources cannot be fetched here; they are looked up by base name in
 * $GCN10_VRT_SOURCE_DIR or next to the .vrt file, and a source that cannot be opened fails the read like GDAL's
 * RasterIO would (raster.c:182-186: the block is skipped).
 *
 * Built with -DGCN10_WITH_GDAL (host_raster_gdal.c), a path these readers do not take -- a /vsi... name, a file that
 * is neither a VRT nor a TIFF they can decode -- is opened with GDAL itself, and GCN10_RASTER_BACKEND=gdal sends
 * every input there.  Such a raster is read window by window on the host (no compressed tiles for the GPU).
 */
#define _GNU_SOURCE
#include "gcn10_host.h"
#ifdef GCN10_WITH_GDAL
#include "host_raster_gdal.h"
#endif

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { GH_RASTER_MAX_OPEN = 32 };

typedef struct {
    char *path;
    int sx, sy;                 /* SrcRect offset */
    int dx, dy, w, h;           /* DstRect (size equals SrcRect's) */
    gh_tiff *ds;                /* open handle or NULL */
    unsigned long stamp;        /* last use */
} vrt_source;

struct gh_raster {
    gh_tiff *single;            /* plain GeoTIFF */
#ifdef GCN10_WITH_GDAL
    gh_gdal *gdal;              /* opened by GDAL: neither single nor src is set */
#endif
    int w, h;
    double gt[6];
    int fill;                   /* VRT: the band's NoDataValue (0 when absent) */
    vrt_source *src;
    int nsrc, nopen;
    unsigned long clock;
};

static void set_err(char *err, size_t errlen, const char *fmt, ...)
{
    if (!err || !errlen)
        return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

/* ---- a very small XML reader: enough for the VRT subset above ------------------------------------------ */

/* value of attribute `name` inside the tag that starts at `tag` ("<Name ...>"); NULL if absent */
static int attr_value(const char *tag, const char *name, char *out, size_t n)
{
    const char *end = strchr(tag, '>');
    if (!end)
        return -1;
    size_t ln = strlen(name);
    for (const char *p = tag; p < end; p++) {
        if ((p == tag || isspace((unsigned char)p[-1])) && !strncmp(p, name, ln)) {
            const char *q = p + ln;
            while (q < end && isspace((unsigned char)*q))
                q++;
            if (q >= end || *q != '=')
                continue;
            q++;
            while (q < end && isspace((unsigned char)*q))
                q++;
            if (q >= end || (*q != '"' && *q != '\''))
                continue;
            const char quote = *q++;
            const char *e = memchr(q, quote, (size_t)(end - q));
            if (!e)
                return -1;
            size_t len = (size_t)(e - q);
            if (len >= n)
                len = n - 1;
            memcpy(out, q, len);
            out[len] = '\0';
            return 0;
        }
    }
    return -1;
}

static int attr_int(const char *tag, const char *name, int *v)
{
    char buf[64];
    if (attr_value(tag, name, buf, sizeof buf))
        return -1;
    char *e;
    double d = strtod(buf, &e);         /* GDAL writes integers, sometimes as "36000.0" */
    if (e == buf)
        return -1;
    *v = (int)d;
    return (double)*v == d ? 0 : -1;
}

/* first "<name" at or after p and before limit that is a whole element name */
static const char *find_tag(const char *p, const char *limit, const char *name)
{
    size_t ln = strlen(name);
    while (p && p < limit) {
        p = memmem(p, (size_t)(limit - p), "<", 1);
        if (!p)
            return NULL;
        if ((size_t)(limit - p) > ln + 1 && !strncmp(p + 1, name, ln) &&
            (isspace((unsigned char)p[1 + ln]) || p[1 + ln] == '>' || p[1 + ln] == '/'))
            return p;
        p++;
    }
    return NULL;
}

/* text content of <name ...>text</name> -> out (entities &amp; &lt; &gt; &quot; &apos; decoded, trimmed) */
static int element_text(const char *tag, const char *limit, char *out, size_t n)
{
    const char *gt = memchr(tag, '>', (size_t)(limit - tag));
    if (!gt || gt[-1] == '/')
        return -1;
    const char *end = memmem(gt, (size_t)(limit - gt), "</", 2);
    if (!end)
        return -1;
    const char *p = gt + 1;
    while (p < end && isspace((unsigned char)*p))
        p++;
    while (end > p && isspace((unsigned char)end[-1]))
        end--;
    size_t k = 0;
    static const struct { const char *ent; char ch; } ents[] = {
        { "&amp;", '&' }, { "&lt;", '<' }, { "&gt;", '>' }, { "&quot;", '"' }, { "&apos;", '\'' } };
    while (p < end && k + 1 < n) {
        int hit = 0;
        if (*p == '&')
            for (size_t i = 0; i < sizeof ents / sizeof ents[0]; i++) {
                size_t le = strlen(ents[i].ent);
                if ((size_t)(end - p) >= le && !strncmp(p, ents[i].ent, le)) {
                    out[k++] = ents[i].ch;
                    p += le;
                    hit = 1;
                    break;
                }
            }
        if (!hit)
            out[k++] = *p++;
    }
    out[k] = '\0';
    return 0;
}

static char *dir_of(const char *path)
{
    const char *slash = strrchr(path, '/');
    if (!slash)
        return strdup(".");
    return strndup(path, (size_t)(slash - path) ? (size_t)(slash - path) : 1);
}

/* where a <SourceFilename> is looked for on this machine */
static char *resolve_source(const char *vrt_dir, const char *name, int relative)
{
    char *out = NULL;
    const int remote = !strncmp(name, "/vsi", 4) || strstr(name, "://") != NULL;
    if (remote) {
        const char *base = strrchr(name, '/');
        base = base ? base + 1 : name;
        const char *dir = getenv("GCN10_VRT_SOURCE_DIR");
        if (asprintf(&out, "%s/%s", dir && *dir ? dir : vrt_dir, base) < 0)
            return NULL;
        return out;
    }
    if (relative && name[0] != '/') {
        if (asprintf(&out, "%s/%s", vrt_dir, name) < 0)
            return NULL;
        return out;
    }
    return strdup(name);
}

static int parse_vrt(const char *path, const char *xml, size_t len, gh_raster *r, char *err, size_t errlen)
{
    const char *limit = xml + len;
    const char *root = find_tag(xml, limit, "VRTDataset");
    if (!root || attr_int(root, "rasterXSize", &r->w) || attr_int(root, "rasterYSize", &r->h) || r->w <= 0 || r->h <= 0) {
        set_err(err, errlen, "gdal open failed: %s (not a VRT dataset)", path);
        return -1;
    }
    /* GDAL's default for a dataset without georeferencing */
    r->gt[0] = 0; r->gt[1] = 1; r->gt[2] = 0; r->gt[3] = 0; r->gt[4] = 0; r->gt[5] = 1;
    char buf[4096];
    const char *g = find_tag(root, limit, "GeoTransform");
    if (g && !element_text(g, limit, buf, sizeof buf)) {
        char *p = buf;
        for (int i = 0; i < 6; i++) {
            char *e;
            r->gt[i] = strtod(p, &e);   /* the same conversion CPLAtof applies to GDAL's %.16e output */
            if (e == p) {
                set_err(err, errlen, "gdal open failed: %s (bad GeoTransform)", path);
                return -1;
            }
            p = e;
            while (*p == ',' || isspace((unsigned char)*p))
                p++;
        }
    }
    const char *band = find_tag(root, limit, "VRTRasterBand");
    if (!band) {
        set_err(err, errlen, "gdal open failed: %s (no raster band)", path);
        return -1;
    }
    char dtype[32] = "Byte";
    attr_value(band, "dataType", dtype, sizeof dtype);
    if (strcmp(dtype, "Byte")) {
        set_err(err, errlen, "gdal open failed: %s (band data type %s; need Byte)", path, dtype);
        return -1;
    }
    const char *band_end = memmem(band, (size_t)(limit - band), "</VRTRasterBand>", 16);
    if (!band_end)
        band_end = limit;
    const char *nd = find_tag(band, band_end, "NoDataValue");
    if (nd && !element_text(nd, band_end, buf, sizeof buf))
        r->fill = (int)strtod(buf, NULL) & 255;

    char *vdir = dir_of(path);
    int cap = 0;
    const char *p = band;
    while (p < band_end) {
        const char *s1 = find_tag(p, band_end, "SimpleSource"), *s2 = find_tag(p, band_end, "ComplexSource");
        const char *s = !s1 ? s2 : !s2 ? s1 : (s1 < s2 ? s1 : s2);
        if (!s)
            break;
        const char *close_name = s == s1 ? "</SimpleSource>" : "</ComplexSource>";
        const char *e = memmem(s, (size_t)(band_end - s), close_name, strlen(close_name));
        if (!e)
            e = band_end;
        const char *fn = find_tag(s, e, "SourceFilename"), *sr = find_tag(s, e, "SrcRect"), *dr = find_tag(s, e, "DstRect");
        const char *sb = find_tag(s, e, "SourceBand");
        vrt_source v;
        memset(&v, 0, sizeof v);
        int sw = 0, sh = 0, bandno = 1;
        char rel[8] = "0";
        if (sb && !element_text(sb, e, buf, sizeof buf))
            bandno = atoi(buf);
        if (!fn || !sr || !dr || element_text(fn, e, buf, sizeof buf) ||
            attr_int(sr, "xOff", &v.sx) || attr_int(sr, "yOff", &v.sy) || attr_int(sr, "xSize", &sw) || attr_int(sr, "ySize", &sh) ||
            attr_int(dr, "xOff", &v.dx) || attr_int(dr, "yOff", &v.dy) || attr_int(dr, "xSize", &v.w) || attr_int(dr, "ySize", &v.h)) {
            set_err(err, errlen, "gdal open failed: %s (source %d: missing SourceFilename / SrcRect / DstRect)", path, r->nsrc);
            free(vdir);
            return -1;
        }
        if (sw != v.w || sh != v.h || v.w <= 0 || v.h <= 0 || v.sx < 0 || v.sy < 0 || bandno != 1) {
            set_err(err, errlen, "gdal open failed: %s (source %d: only band 1 placed 1:1 is supported, got %dx%d -> %dx%d, band %d)",
                    path, r->nsrc, sw, sh, v.w, v.h, bandno);
            free(vdir);
            return -1;
        }
        attr_value(fn, "relativeToVRT", rel, sizeof rel);
        v.path = resolve_source(vdir, buf, rel[0] == '1');
        if (r->nsrc == cap) {
            cap = cap ? 2 * cap : 64;
            vrt_source *ns = realloc(r->src, (size_t)cap * sizeof *ns);
            if (!ns || !v.path) {
                free(v.path);
                free(vdir);
                return -1;
            }
            r->src = ns;
        }
        r->src[r->nsrc++] = v;
        p = e + 1;
    }
    free(vdir);
    return 0;
}

/* ---- the optional GDAL backend ------------------------------------------------------------------------- */

int gh_raster_have_gdal(void)
{
#ifdef GCN10_WITH_GDAL
    return 1;
#else
    return 0;
#endif
}

static int open_with_gdal(const char *path, gh_raster *r, char *err, size_t errlen)
{
#ifdef GCN10_WITH_GDAL
    return gh_gdal_open(path, &r->gdal, &r->w, &r->h, r->gt, err, errlen);
#else
    (void)r;
    set_err(err, errlen, "gdal open failed: %s (this build has no GDAL backend: make GDAL=1)", path);
    return -1;
#endif
}

const char *gh_raster_backend(const gh_raster *r)
{
#ifdef GCN10_WITH_GDAL
    if (r->gdal)
        return "gdal";
#endif
    return r->single ? "geotiff" : "vrt";
}

int gh_raster_open(const char *path, gh_raster **out, char *err, size_t errlen)
{
    *out = NULL;
    if (err && errlen)
        err[0] = '\0';
    gh_raster *r = calloc(1, sizeof *r);
    if (!r)
        return -1;
    const char *want = getenv("GCN10_RASTER_BACKEND");
    if (want && strcmp(want, "gdal") == 0) {
        if (open_with_gdal(path, r, err, errlen)) {
            free(r);
            return -1;
        }
        *out = r;
        return 0;
    }
    /* a VRT is XML: look at the first bytes instead of trusting the extension, like GDAL's driver probing */
    FILE *f = fopen(path, "rb");
    if (!f) {
        /* not a local file (/vsicurl/..., a driver prefix): GDAL's business when it is there */
        if (gh_raster_have_gdal() && open_with_gdal(path, r, err, errlen) == 0) {
            *out = r;
            return 0;
        }
        set_err(err, errlen, "gdal open failed: %s", path);                 /* raster.c:121 */
        free(r);
        return -1;
    }
    char head[256];
    size_t nh = fread(head, 1, sizeof head - 1, f);
    head[nh] = '\0';
    if (strstr(head, "<VRTDataset")) {
        fseek(f, 0, SEEK_END);
        long len = ftell(f);
        fseek(f, 0, SEEK_SET);
        char *xml = len > 0 ? malloc((size_t)len + 1) : NULL;
        if (!xml || fread(xml, 1, (size_t)len, f) != (size_t)len) {
            set_err(err, errlen, "gdal open failed: %s", path);
            free(xml);
            fclose(f);
            free(r);
            return -1;
        }
        xml[len] = '\0';
        fclose(f);
        int rc = parse_vrt(path, xml, (size_t)len, r, err, errlen);
        free(xml);
        if (rc) {
            gh_raster_close(r);
            /* a VRT outside the subset above (warped, scaled, derived bands): GDAL reads it when it is there */
            char gerr[256];
            if (gh_raster_have_gdal() && (r = calloc(1, sizeof *r)) != NULL) {
                if (open_with_gdal(path, r, gerr, sizeof gerr) == 0) {
                    if (err && errlen)
                        err[0] = '\0';
                    *out = r;
                    return 0;
                }
                free(r);
            }
            return -1;
        }
        *out = r;
        return 0;
    }
    fclose(f);
    if (gh_tiff_open(path, &r->single, err, errlen)) {
        /* not a TIFF this reader decodes (another format, LZW / ZSTD tiles, ...) */
        char gerr[256];
        if (gh_raster_have_gdal() && open_with_gdal(path, r, gerr, sizeof gerr) == 0) {
            if (err && errlen)
                err[0] = '\0';
            *out = r;
            return 0;
        }
        free(r);
        return -1;
    }
    gh_tiff_size(r->single, &r->w, &r->h);
    gh_tiff_geotransform(r->single, r->gt);
    *out = r;
    return 0;
}

int gh_raster_size(const gh_raster *r, int *w, int *h) { *w = r->w; *h = r->h; return 0; }
int gh_raster_geotransform(const gh_raster *r, double gt[6]) { memcpy(gt, r->gt, sizeof r->gt); return 0; }
static int is_gdal(const gh_raster *r)
{
#ifdef GCN10_WITH_GDAL
    return r->gdal != NULL;
#else
    (void)r;
    return 0;
#endif
}
int gh_raster_is_mosaic(const gh_raster *r) { return !r->single && !is_gdal(r); }
int gh_raster_source_count(const gh_raster *r) { return gh_raster_is_mosaic(r) ? r->nsrc : 1; }
int gh_raster_fill(const gh_raster *r) { return r->fill; }

int gh_raster_geokeys(const gh_raster *r, gh_geokeys *out)
{
    memset(out, 0, sizeof *out);
#ifdef GCN10_WITH_GDAL
    if (r->gdal)
        return 1;               /* the writer's EPSG:4326 default (host_raster_gdal.c) */
#endif
    if (r->single)
        return gh_tiff_geokeys(r->single, out);
    for (int i = 0; i < r->nsrc; i++)
        if (r->src[i].ds)
            return gh_tiff_geokeys(r->src[i].ds, out);
    return 1;
}

void gh_raster_close(gh_raster *r)
{
    if (!r)
        return;
#ifdef GCN10_WITH_GDAL
    gh_gdal_close(r->gdal);
#endif
    gh_tiff_close(r->single);
    for (int i = 0; i < r->nsrc; i++) {
        gh_tiff_close(r->src[i].ds);
        free(r->src[i].path);
    }
    free(r->src);
    free(r);
}

/* the open handle of source i (opened on demand; the least recently used handle makes room) */
static gh_tiff *source_handle(gh_raster *r, int i, char *err, size_t errlen)
{
    vrt_source *v = &r->src[i];
    v->stamp = ++r->clock;
    if (v->ds)
        return v->ds;
    if (r->nopen >= GH_RASTER_MAX_OPEN) {
        int old = -1;
        for (int k = 0; k < r->nsrc; k++)
            if (r->src[k].ds && k != i && (old < 0 || r->src[k].stamp < r->src[old].stamp))
                old = k;
        if (old >= 0) {
            gh_tiff_close(r->src[old].ds);
            r->src[old].ds = NULL;
            r->nopen--;
        }
    }
    if (gh_tiff_open(v->path, &v->ds, err, errlen))
        return NULL;
    int sw, sh;
    gh_tiff_size(v->ds, &sw, &sh);
    if (v->sx + v->w > sw || v->sy + v->h > sh) {
        set_err(err, errlen, "gdalrasterio error 3 (%s is %dx%d, the VRT reads %dx%d at %d,%d)", v->path, sw, sh, v->w, v->h,
                v->sx, v->sy);
        gh_tiff_close(v->ds);
        v->ds = NULL;
        return NULL;
    }
    r->nopen++;
    return v->ds;
}

/* intersection of window [xoff, xoff+xcount) x [yoff, ...) with source i: 0 = empty */
static int intersect(const vrt_source *v, int xoff, int yoff, int xcount, int ycount, int *ix, int *iy, int *iw, int *ih)
{
    const int x0 = v->dx > xoff ? v->dx : xoff, y0 = v->dy > yoff ? v->dy : yoff;
    const long long x1 = (long long)v->dx + v->w < (long long)xoff + xcount ? (long long)v->dx + v->w : (long long)xoff + xcount;
    const long long y1 = (long long)v->dy + v->h < (long long)yoff + ycount ? (long long)v->dy + v->h : (long long)yoff + ycount;
    if (x1 <= x0 || y1 <= y0)
        return 0;
    *ix = x0;
    *iy = y0;
    *iw = (int)(x1 - x0);
    *ih = (int)(y1 - y0);
    return 1;
}

int gh_raster_read_window(gh_raster *r, int xoff, int yoff, int xcount, int ycount, uint8_t *dst, size_t pitch,
                          int threads, char *err, size_t errlen)
{
#ifdef GCN10_WITH_GDAL
    if (r->gdal)
        return gh_gdal_read_window(r->gdal, xoff, yoff, xcount, ycount, dst, pitch, err, errlen);
#endif
    if (r->single)
        return gh_tiff_read_window(r->single, xoff, yoff, xcount, ycount, dst, pitch, threads, err, errlen);
    if (xoff < 0 || yoff < 0 || xcount <= 0 || ycount <= 0 || (long long)xoff + xcount > r->w ||
        (long long)yoff + ycount > r->h || pitch < (size_t)xcount) {
        set_err(err, errlen, "gdalrasterio error 3 (window %d,%d %dx%d outside %dx%d)", xoff, yoff, xcount, ycount, r->w, r->h);
        return -1;
    }
    for (int y = 0; y < ycount; y++)
        memset(dst + (size_t)y * pitch, r->fill, (size_t)xcount);
    for (int i = 0; i < r->nsrc; i++) {
        int ix, iy, iw, ih;
        if (!intersect(&r->src[i], xoff, yoff, xcount, ycount, &ix, &iy, &iw, &ih))
            continue;
        gh_tiff *ds = source_handle(r, i, err, errlen);
        if (!ds)
            return -1;
        const vrt_source *v = &r->src[i];
        if (gh_tiff_read_window(ds, v->sx + (ix - v->dx), v->sy + (iy - v->dy), iw, ih,
                                dst + (size_t)(iy - yoff) * pitch + (size_t)(ix - xoff), pitch, threads, err, errlen))
            return -1;
    }
    return 0;
}

int gh_raster_window_parts(gh_raster *r, int xoff, int yoff, int xcount, int ycount, gh_raster_part *parts, int max_parts,
                           int *nparts, char *err, size_t errlen)
{
    *nparts = 0;
    if (err && errlen)
        err[0] = '\0';
#ifdef GCN10_WITH_GDAL
    if (r->gdal)
        return 1;                       /* GDAL hands out decoded pixels only */
#endif
    if (r->single) {
        if (max_parts < 1 || gh_tiff_window_tiles_plan(r->single, xoff, yoff, xcount, ycount, &parts[0].plan))
            return 1;
        parts[0].ds = r->single;
        parts[0].dst_x = parts[0].dst_y = 0;
        parts[0].w = xcount;
        parts[0].h = ycount;
        *nparts = 1;
        return 0;
    }
    if (xoff < 0 || yoff < 0 || xcount <= 0 || ycount <= 0 || (long long)xoff + xcount > r->w ||
        (long long)yoff + ycount > r->h)
        return 1;
    int n = 0;
    for (int i = 0; i < r->nsrc; i++) {
        int ix, iy, iw, ih;
        if (!intersect(&r->src[i], xoff, yoff, xcount, ycount, &ix, &iy, &iw, &ih))
            continue;
        if (n == max_parts)
            return 1;                   /* more sources than the device call takes: decode on the host */
        gh_tiff *ds = source_handle(r, i, err, errlen);
        if (!ds)
            return -1;
        const vrt_source *v = &r->src[i];
        if (gh_tiff_window_tiles_plan(ds, v->sx + (ix - v->dx), v->sy + (iy - v->dy), iw, ih, &parts[n].plan))
            return 1;
        /* parts must not overlap (later sources would have to win pixel by pixel): decode on the host then */
        for (int k = 0; k < n; k++)
            if (ix - xoff < parts[k].dst_x + parts[k].w && parts[k].dst_x < ix - xoff + iw &&
                iy - yoff < parts[k].dst_y + parts[k].h && parts[k].dst_y < iy - yoff + ih)
                return 1;
        parts[n].ds = ds;
        parts[n].dst_x = ix - xoff;
        parts[n].dst_y = iy - yoff;
        parts[n].w = iw;
        parts[n].h = ih;
        n++;
    }
    *nparts = n;
    return 0;
}
