/* oracle/refshim/ref_api.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * ctypes-facing entry points of oracle/_ref/libgcn10_ref.so.  The library links
 * the reference's src/cn.c and src/raster.c, compiled UNMODIFIED from
 * /root/reference/src (recipe: oracle/Makefile), against the RAM GDAL/OGR/MPI
 * stand-ins in this directory.  These wrappers only stage inputs, call the
 * reference's own process_block() / load_raster() (global.h:54-58) and collect
 * what the reference hands to save_raster().
 *
 * It also supplies the symbols cn.c/raster.c import from the reference files we
 * do not link (config.c globals, log.c's log_message/report_block_completion).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <setjmp.h>
#include <time.h>
#include <unistd.h>
#include <limits.h>
#include "gdal.h"
#include "refshim.h"

/* reference prototypes (src/global.h:54-58) */
uint8_t *load_raster(const char *, const double *, int *, int *, double *, OGRSpatialReferenceH *);
void process_block(int, bool, int);

/* config.c globals (src/config.c:13-21) */
char *hysogs_data_path = NULL;
char *esa_data_path = NULL;
char *blocks_shp_path = NULL;
char *lookup_table_path = NULL;
char *log_dir = NULL;
bool use_list_mode = false;
char *block_ids_file = NULL;

static char g_log[16384];
static int g_nerrors = 0;
static int g_ncompleted = 0;
static jmp_buf g_abort_jmp;
static int g_abort_armed = 0;

double refshim_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* log.c stand-ins: keep the text so tests can assert on the reference's messages */
void log_message(const char *level, const char *msg, bool echo)
{
    (void)echo;
    if (strcmp(level, "ERROR") == 0)
        g_nerrors++;
    size_t used = strlen(g_log);
    if (used + strlen(level) + strlen(msg) + 8 < sizeof(g_log))
        snprintf(g_log + used, sizeof(g_log) - used, "[%s] %s\n", level, msg);
}

void report_block_completion(int block_id, int total_blocks)
{
    (void)block_id; (void)total_blocks;
    g_ncompleted++;
}

void refshim_abort(int code)
{
    if (g_abort_armed)
        longjmp(g_abort_jmp, code ? code : 1);
    fprintf(stderr, "refshim: MPI_Abort(%d) outside a guarded call\n", code);
    exit(code ? code : 1);
}

void refshim_sink_deliver(refshim_sink *s, const char *path, const void *buf, int w, int h)
{
    if (s->count >= REFSHIM_MAX_PLANES) {
        s->overflow = 1;
        return;
    }
    int k = s->count;
    s->t_plane[k] = refshim_now() - s->t_start;
    snprintf(s->paths[k], sizeof(s->paths[k]), "%s", path);
    s->w = w;
    s->h = h;
    if (s->out) {
        size_t n = (size_t)w * h;
        if (n > s->plane_capacity)
            s->overflow = 1;
        else
            memcpy(s->out + (size_t)k * s->plane_capacity, buf, n);
    }
    s->count = k + 1;
}

const char *refshim_log(void) { return g_log; }
int refshim_error_count(void) { return g_nerrors; }
int refshim_completed_count(void) { return g_ncompleted; }
const char *refshim_plane_path(int k) { return refshim_get_sink()->paths[k]; }
const char *refshim_create_option(int k) { return refshim_get_sink()->last_options[k]; }
double refshim_plane_time(int k) { return refshim_get_sink()->t_plane[k]; }

/* Run the reference's process_block() on one block.
 *
 *   esa/hsg         full in-memory rasters with their dataset geotransforms
 *   bbox            minx, miny, maxx, maxy of the block (what OGR would return)
 *   lookup_dir      directory holding default_lookup_<hc>_<arc>.csv
 *   scratch_dir     CWD for the call (the reference mkdirs cn_rasters_* there)
 *   out             18 planes, plane_capacity bytes apart, in the reference's
 *                   save order (cn.c:236,258-259); NULL = time only
 * Returns the number of planes the reference saved (18 on success), or a
 * negative number: -1 bad arguments, -2 reference called MPI_Abort,
 * -3 plane larger than plane_capacity.
 */
int refshim_run_block(const uint8_t *esa, int ew, int eh, const double esa_t[6],
                      const uint8_t *hsg, int hw, int hh, const double hsg_t[6],
                      const double bbox[4], const char *lookup_dir, const char *scratch_dir,
                      int block_id, int overwrite,
                      uint8_t *out, size_t plane_capacity,
                      int *out_w, int *out_h, double out_gt[6])
{
    char cwd[PATH_MAX];
    int rc;

    if (!esa || !hsg || !lookup_dir || !scratch_dir)
        return -1;
    refshim_reset();
    g_log[0] = 0;
    g_nerrors = 0;
    g_ncompleted = 0;

    refshim_add_raster("mem:esa", esa, ew, eh, esa_t);
    refshim_add_raster("mem:hsg", hsg, hw, hh, hsg_t);
    refshim_set_blocks("mem:blocks", 1, &block_id, bbox);
    esa_data_path = "mem:esa";
    hysogs_data_path = "mem:hsg";
    blocks_shp_path = "mem:blocks";
    lookup_table_path = (char *)lookup_dir;
    log_dir = (char *)scratch_dir;

    refshim_sink *sink = refshim_get_sink();
    sink->out = out;
    sink->plane_capacity = plane_capacity;

    if (!getcwd(cwd, sizeof(cwd)) || chdir(scratch_dir) != 0)
        return -1;

    g_abort_armed = 1;
    if (setjmp(g_abort_jmp) == 0) {
        sink->t_start = refshim_now();
        process_block(block_id, overwrite != 0, 1);
        rc = sink->overflow ? -3 : sink->count;
    }
    else {
        rc = -2;
    }
    g_abort_armed = 0;
    if (chdir(cwd) != 0)
        rc = -1;

    if (out_w) *out_w = sink->w;
    if (out_h) *out_h = sink->h;
    if (out_gt) memcpy(out_gt, sink->gt, sizeof(double) * 6);
    return rc;
}

/* Run only the reference's window arithmetic (raster.c:126-162) for a raster of
 * rw x rh pixels with dataset geotransform t and a block bbox.  Returns 0 and
 * fills xoff/yoff/xsize/ysize/gt, or 1 when the reference rejected the window
 * ("invalid raster bounds", raster.c:142-147). */
int refshim_window(int rw, int rh, const double t[6], const double bbox[4],
                   int *xoff, int *yoff, int *xsize, int *ysize, double gt[6])
{
    OGRSpatialReferenceH srs;
    uint8_t *buf;

    refshim_reset();
    g_log[0] = 0;
    g_nerrors = 0;
    refshim_add_raster("mem:probe", NULL, rw, rh, t);
    g_abort_armed = 1;
    if (setjmp(g_abort_jmp) != 0) {
        g_abort_armed = 0;
        return 2;
    }
    buf = load_raster("mem:probe", bbox, xsize, ysize, gt, &srs);
    g_abort_armed = 0;
    if (!buf)
        return 1;
    free(buf);
    const refshim_lastread *lr = refshim_get_lastread();
    *xoff = lr->xoff;
    *yoff = lr->yoff;
    return 0;
}
