/* gcn10_b200/host/host_core.c -- lookup CSVs, window arithmetic, config file, block ids,
 * block extents (.shp/.dbf) and logging for the gcn10 host program.  See gcn10_host.h for the
 * mapping to the reference's files; citations are relative to /root/reference/.
 *
 * Compiled without -march / -ffast-math and with -ffp-contract=off: the window arithmetic feeds
 * geotransforms to the GPU index maps and must be evaluated as the reference's x86-64 build does.
 */
#define _GNU_SOURCE
#include "gcn10_host.h"

#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

static void set_err(char *err, size_t errlen, const char *fmt, ...)
{
    if (!err || !errlen)
        return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

/* ------------------------------------------------------------------------------ lookup tables */

int gh_load_lookup_table(const char *dir, const char *hc, const char *arc, int table[256][5],
                         char *err, size_t errlen)
{
    char path[PATH_MAX];
    char buf[128];              /* the reference reads through a 128-byte line buffer (cn.c:17) */

    if (snprintf(path, sizeof path, "%s/default_lookup_%s_%s.csv", dir, hc, arc) >= (int)sizeof path) {
        set_err(err, errlen, "lookup table path too long: %s", path);      /* cn.c:23 */
        return -1;
    }
    FILE *fp = fopen(path, "r");
    if (!fp) {
        set_err(err, errlen, "cannot open lookup table %s", path);         /* cn.c:30 */
        return -2;
    }
    for (int i = 0; i < 256 * 5; i++)
        (&table[0][0])[i] = 255;                                            /* cn.c:36-40 */

    if (!fgets(buf, sizeof buf, fp)) {                                      /* header, cn.c:43 */
        set_err(err, errlen, "empty lookup table %s", path);               /* cn.c:44 */
        fclose(fp);
        return -3;
    }
    while (fgets(buf, sizeof buf, fp)) {
        /* first comma-separated token, leading commas skipped like strtok does */
        char *key = buf + strspn(buf, ",");
        if (!*key)
            continue;                                                       /* cn.c:52-54 */
        char *rest = key + strcspn(key, ",");
        if (*rest)
            *rest++ = '\0';
        char *bar = strchr(key, '_');
        if (!bar)
            continue;                                                       /* cn.c:57-62: logged, skipped */
        *bar = '\0';
        int lc = atoi(key);                                                 /* cn.c:65 */
        int sg;
        switch (bar[1]) {                                                   /* cn.c:66 */
        case 'A': sg = 1; break;
        case 'B': sg = 2; break;
        case 'C': sg = 3; break;
        default:  sg = 4; break;
        }
        rest += strspn(rest, ",");
        if (!*rest)
            continue;                                                       /* cn.c:68-73: "missing cn" */
        rest[strcspn(rest, ",")] = '\0';
        int cn = atoi(rest);                                                /* cn.c:74 */
        if (lc >= 0 && lc < 256)                                            /* cn.c:75-77 */
            table[lc][sg] = cn;
    }
    fclose(fp);
    return 0;
}

int gh_load_lookup_tables(const char *dir, int tables[9][256][5], char *err, size_t errlen)
{
    static const char *const hcs[3] = { "p", "f", "g" };
    static const char *const arcs[3] = { "i", "ii", "iii" };

    for (int t = 0; t < 9; t++) {
        int rc = gh_load_lookup_table(dir, hcs[t / 3], arcs[t % 3], tables[t], err, errlen);
        if (rc)
            return rc;
    }
    return 0;
}

/* --------------------------------------------------------------------------- window arithmetic */

/* double -> int the way the reference's x86-64 object code does it (cvttsd2si): truncation,
 * INT_MIN for NaN / out of range */
static int trunc_to_int(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int)v;
}

int gh_raster_window(int raster_w, int raster_h, const double t[6], const double bbox[4], gh_window *win)
{
    const double minx = bbox[0], miny = bbox[1], maxx = bbox[2], maxy = bbox[3];
    int x0 = trunc_to_int(floor((minx - t[0]) / t[1]));         /* raster.c:127 */
    int y0 = trunc_to_int(floor((maxy - t[3]) / t[5]));         /* raster.c:128 */
    int nx = trunc_to_int(ceil((maxx - minx) / t[1]));          /* raster.c:129 */
    int ny = trunc_to_int(ceil((miny - maxy) / t[5]));          /* raster.c:130 */

    if (x0 < 0) { nx += x0; x0 = 0; }                           /* raster.c:134-137 */
    if (y0 < 0) { ny += y0; y0 = 0; }                           /* raster.c:138-141 */
    if (x0 >= raster_w || y0 >= raster_h || nx <= 0 || ny <= 0)
        return 1;                                               /* raster.c:142-147 */
    if (x0 + nx > raster_w) nx = raster_w - x0;                 /* raster.c:148-150 */
    if (y0 + ny > raster_h) ny = raster_h - y0;                 /* raster.c:151-153 */

    win->xoff = x0;
    win->yoff = y0;
    win->xcount = nx;
    win->ycount = ny;
    win->gt[0] = t[0] + x0 * t[1];                              /* raster.c:157 */
    win->gt[1] = t[1];
    win->gt[2] = t[2];
    win->gt[3] = t[3] + y0 * t[5];                              /* raster.c:160 */
    win->gt[4] = t[4];
    win->gt[5] = t[5];
    return 0;
}

/* ---------------------------------------------------------------------------------- config file */

static char *strip(char *s)
{
    while (*s && isspace((unsigned char)*s))
        s++;
    size_t n = strlen(s);
    while (n > 0 && isspace((unsigned char)s[n - 1]))
        s[--n] = '\0';
    return s;
}

int gh_config_parse(const char *path, gh_config *cfg, char *err, size_t errlen)
{
    static const char *const keys[5] = { "hysogs_data_path", "esa_data_path", "blocks_shp_path",
                                         "lookup_table_path", "log_dir" };
    char line[512];             /* config.c:47 */
    char **slots[5] = { &cfg->hysogs_data_path, &cfg->esa_data_path, &cfg->blocks_shp_path,
                        &cfg->lookup_table_path, &cfg->log_dir };

    memset(cfg, 0, sizeof *cfg);
    FILE *fp = fopen(path, "r");
    if (!fp) {
        set_err(err, errlen, "cannot open config '%s'", path);             /* config.c:52 */
        return -1;
    }
    while (fgets(line, sizeof line, fp)) {
        char *p = strip(line);
        if (!*p || *p == '#')                                               /* config.c:58-60 */
            continue;
        char *eq = strchr(p, '=');
        if (!eq)                                                            /* config.c:62-64 */
            continue;
        *eq = '\0';
        char *key = strip(p), *val = strip(eq + 1);
        for (int k = 0; k < 5; k++) {
            if (strcmp(key, keys[k]) == 0) {
                free(*slots[k]);                                            /* last assignment wins */
                *slots[k] = strdup(val);
            }
        }
    }
    fclose(fp);
    for (int k = 0; k < 5; k++) {
        if (!*slots[k]) {                                                   /* config.c:107-113 */
            set_err(err, errlen, "missing one of: hysogs_data_path, esa_data_path,\n"
                                 "blocks_shp_path, lookup_table_path, log_dir");
            gh_config_free(cfg);
            return -2;
        }
    }
    return 0;
}

void gh_config_free(gh_config *cfg)
{
    free(cfg->hysogs_data_path);
    free(cfg->esa_data_path);
    free(cfg->blocks_shp_path);
    free(cfg->lookup_table_path);
    free(cfg->log_dir);
    memset(cfg, 0, sizeof *cfg);
}

/* ------------------------------------------------------------------------------------ block ids */

int gh_read_block_list(const char *path, int **ids, int *n)
{
    FILE *fp = fopen(path, "r");
    *ids = NULL;
    *n = 0;
    if (!fp)
        return -1;                                                          /* raster.c:30-34 */
    int cap = 128, cnt = 0, v;
    int *a = malloc(sizeof(int) * (size_t)cap);
    while (a && fscanf(fp, "%d", &v) == 1) {                                /* raster.c:45 */
        if (cnt == cap) {
            cap *= 2;
            int *b = realloc(a, sizeof(int) * (size_t)cap);
            if (!b) {
                free(a);
                a = NULL;
                break;
            }
            a = b;
        }
        a[cnt++] = v;
    }
    fclose(fp);
    if (!a)
        return -2;
    *ids = a;
    *n = cnt;
    return 0;
}

/* ------------------------------------------------------------------------ block extents (.shp) */

struct gh_blocks {
    int n;
    int *ids;
    double *bbox;               /* n x {minx, miny, maxx, maxy} */
};

static uint32_t be32(const unsigned char *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static uint32_t le32(const unsigned char *p) { return (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]; }
static uint16_t le16(const unsigned char *p) { return (uint16_t)(p[1] << 8 | p[0]); }
static double le_f64(const unsigned char *p)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--)
        v = v << 8 | p[i];
    double d;
    memcpy(&d, &v, 8);
    return d;
}

static unsigned char *slurp(const char *path, size_t *len)
{
    FILE *fp = fopen(path, "rb");
    if (!fp)
        return NULL;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    unsigned char *buf = sz >= 0 ? malloc((size_t)sz + 1) : NULL;
    if (buf && fread(buf, 1, (size_t)sz, fp) != (size_t)sz) {
        free(buf);
        buf = NULL;
    }
    fclose(fp);
    if (buf)
        *len = (size_t)sz;
    return buf;
}

int gh_blocks_open(const char *shp_path, gh_blocks **out, char *err, size_t errlen)
{
    *out = NULL;
    size_t shp_len = 0, dbf_len = 0;
    unsigned char *shp = slurp(shp_path, &shp_len);
    if (!shp || shp_len < 100 || be32(shp) != 9994) {
        set_err(err, errlen, "ogr open failed: %s", shp_path);             /* cn.c:157, raster.c:79 */
        free(shp);
        return -1;
    }
    char dbf_path[PATH_MAX];
    size_t plen = strlen(shp_path);
    if (plen < 4 || plen >= sizeof dbf_path) {
        free(shp);
        set_err(err, errlen, "ogr open failed: %s", shp_path);
        return -1;
    }
    memcpy(dbf_path, shp_path, plen + 1);
    memcpy(dbf_path + plen - 3, isupper((unsigned char)shp_path[plen - 1]) ? "DBF" : "dbf", 3);
    unsigned char *dbf = slurp(dbf_path, &dbf_len);
    if (!dbf || dbf_len < 33) {
        set_err(err, errlen, "ogr open failed: %s (no attribute table %s)", shp_path, dbf_path);
        free(shp);
        free(dbf);
        return -1;
    }

    /* dBASE header: record count, header length, record length; 32-byte field descriptors */
    uint32_t nrec = le32(dbf + 4);
    uint16_t hlen = le16(dbf + 8), rlen = le16(dbf + 10);
    int id_off = -1, id_len = 0, off = 1;       /* byte 0 of a record is the deletion flag */
    const size_t desc_end = hlen < dbf_len ? hlen : dbf_len;    /* a damaged header length must not reach past the file */
    for (size_t d = 32; d + 32 <= desc_end && dbf[d] != 0x0D; d += 32) {
        char name[12] = { 0 };
        memcpy(name, dbf + d, 11);
        int flen = dbf[d + 16];
        if (strcasecmp(name, "ID") == 0) {
            id_off = off;
            id_len = flen;
        }
        off += flen;
    }
    if (id_off < 0 || rlen == 0 || id_off + id_len > (int)rlen || (size_t)hlen + (size_t)nrec * rlen > dbf_len + 1) {
        set_err(err, errlen, "ogr open failed: %s (attribute \"ID\" not found)", shp_path);
        free(shp);
        free(dbf);
        return -1;
    }

    gh_blocks *b = calloc(1, sizeof *b);
    if (b) {
        b->ids = malloc(sizeof(int) * (nrec ? (size_t)nrec : 1));
        b->bbox = malloc(sizeof(double) * 4 * (nrec ? (size_t)nrec : 1));
    }
    if (!b || !b->ids || !b->bbox) {
        set_err(err, errlen, "ogr open failed: %s (out of memory)", shp_path);
        gh_blocks_close(b);
        free(shp);
        free(dbf);
        return -1;
    }
    size_t pos = 100;
    uint32_t i = 0;
    while (i < nrec && pos + 8 <= shp_len) {
        const size_t content = (size_t)be32(shp + pos + 4) * 2u;       /* length in 16-bit words */
        const unsigned char *rec = shp + pos + 8;
        if (content > shp_len - pos - 8)
            break;
        uint32_t shape = content >= 4 ? le32(rec) : 0;
        double *bb = b->bbox + 4 * (size_t)i;
        if ((shape == 5 || shape == 15 || shape == 25 || shape == 3 || shape == 13 || shape == 23) && content >= 36) {
            /* Polygon / PolyLine (plain, Z, M): Xmin Ymin Xmax Ymax follow the shape type */
            bb[0] = le_f64(rec + 4);
            bb[1] = le_f64(rec + 12);
            bb[2] = le_f64(rec + 20);
            bb[3] = le_f64(rec + 28);
        }
        else if ((shape == 1 || shape == 11 || shape == 21) && content >= 20) {
            bb[0] = bb[2] = le_f64(rec + 4);
            bb[1] = bb[3] = le_f64(rec + 12);
        }
        else {
            bb[0] = bb[1] = bb[2] = bb[3] = 0.0;        /* null shape: an empty envelope */
        }
        char num[32] = { 0 };
        memcpy(num, dbf + hlen + (size_t)i * rlen + id_off, id_len < 31 ? (size_t)id_len : 31);
        b->ids[i] = atoi(num);
        pos += 8 + content;
        i++;
    }
    b->n = (int)i;
    free(shp);
    free(dbf);
    *out = b;
    return 0;
}

int gh_blocks_count(const gh_blocks *b) { return b ? b->n : 0; }
int gh_blocks_id(const gh_blocks *b, int index) { return (b && index >= 0 && index < b->n) ? b->ids[index] : -1; }

int gh_blocks_bbox(const gh_blocks *b, int id, double bbox[4])
{
    for (int i = 0; b && i < b->n; i++) {
        if (b->ids[i] == id) {                          /* "\"ID\"=%d" + first feature, cn.c:162-171 */
            memcpy(bbox, b->bbox + 4 * (size_t)i, sizeof(double) * 4);
            return 0;
        }
    }
    return 1;                                           /* "block %d not found", cn.c:172-177 */
}

void gh_blocks_close(gh_blocks *b)
{
    if (!b)
        return;
    free(b->ids);
    free(b->bbox);
    free(b);
}

/* --------------------------------------------------------------------------------------- logging */

struct gh_log {
    FILE *fp;
    int worker;
    pthread_mutex_t mu;
};

static pthread_mutex_t g_console_mu = PTHREAD_MUTEX_INITIALIZER;

static void stamp(char *buf, size_t n)
{
    time_t t = time(NULL);
    struct tm tmv;
    localtime_r(&t, &tmv);
    strftime(buf, n, "%Y-%m-%dT%H:%M:%S", &tmv);        /* log.c:91-98 */
}

gh_log *gh_log_open(const char *log_dir, int worker)
{
    gh_log *lg = calloc(1, sizeof *lg);
    if (!lg)
        return NULL;
    lg->worker = worker;
    pthread_mutex_init(&lg->mu, NULL);
    const char *dir = (log_dir && *log_dir) ? log_dir : ".";
    struct stat st;
    if (stat(dir, &st) != 0 && mkdir(dir, 0775) != 0 && errno != EEXIST)        /* log.c:46-64 */
        fprintf(stderr, "log: failed to create directory '%s': %s\n", dir, strerror(errno));
    char path[4096];
    snprintf(path, sizeof path, "%s/rank_%d.log", dir, worker);                  /* log.c:76 */
    lg->fp = fopen(path, "a");
    if (!lg->fp)
        fprintf(stderr, "log: failed to open %s: %s (fallback to stderr only)\n", path, strerror(errno));
    char ts[64];
    stamp(ts, sizeof ts);
    if (lg->fp) {
        fprintf(lg->fp, "[%s] [rank %d] logging started\n", ts, worker);        /* log.c:111-114 */
        fflush(lg->fp);
    }
    pthread_mutex_lock(&g_console_mu);
    fprintf(stderr, "[%s] [rank %d] logging started\n", ts, worker);
    pthread_mutex_unlock(&g_console_mu);
    return lg;
}

void gh_log_message(gh_log *lg, const char *level, const char *msg, int also_console)
{
    char ts[64];
    stamp(ts, sizeof ts);
    int worker = lg ? lg->worker : 0;
    if (lg && lg->fp) {
        pthread_mutex_lock(&lg->mu);
        fprintf(lg->fp, "[%s] [%s] [rank %d] %s\n", ts, level ? level : "INFO", worker, msg ? msg : "");
        fflush(lg->fp);                                                          /* log.c:157-161 */
        pthread_mutex_unlock(&lg->mu);
    }
    if (also_console) {
        pthread_mutex_lock(&g_console_mu);
        fprintf(stderr, "[%s] [%s] [rank %d] %s\n", ts, level ? level : "INFO", worker, msg ? msg : "");
        pthread_mutex_unlock(&g_console_mu);
    }
}

void gh_log_close(gh_log *lg)
{
    if (!lg)
        return;
    char ts[64];
    stamp(ts, sizeof ts);
    if (lg->fp) {
        fprintf(lg->fp, "[%s] [rank %d] logging finished\n", ts, lg->worker);   /* log.c:262-267 */
        fclose(lg->fp);
    }
    pthread_mutex_lock(&g_console_mu);
    fprintf(stderr, "[%s] [rank %d] logging finished\n", ts, lg->worker);
    pthread_mutex_unlock(&g_console_mu);
    pthread_mutex_destroy(&lg->mu);
    free(lg);
}
