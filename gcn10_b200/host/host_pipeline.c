/* gcn10_b200/host/host_pipeline.c -- the per-block pipeline and the per-GPU block queue.
 *
 * gh_process_block() is this program's process_block() (/root/reference/src/cn.c:134-384): same
 * inputs (block id -> bbox -> two raster windows), same 18 outputs with the same names, same log
 * lines and the same two-tier error convention (recoverable: log ERROR and skip the block; fatal:
 * exit(1) where the reference calls MPI_Abort).  What changes is the middle: instead of five CPU
 * passes per raster (cn.c:218-290) the block is streamed band by band (2048 rows, a multiple of the
 * 256-row GeoTIFF tiles).  Default path: a reader thread decodes land-cover band i+1 while the GPU runs
 * gcn10_cuda_block_deflate_rows() on band i -- Curve Numbers AND the DEFLATE tiles of save_raster()
 * (raster.c:204-219) are produced on the device -- and the worker only appends the compressed tiles to
 * the 18 GeoTIFFs.  With GCN10_HOST_DEFLATE=1 the raw planes come back instead
 * (gcn10_cuda_block_rows) and a zlib thread pool encodes them, overlapped with the next band.
 *
 * gh_run_blocks() replaces the MPI round-robin of main.c:171: one worker thread and one gcn10_ctx
 * per GPU, block ids popped from a shared atomic counter; the join of the workers is the barrier
 * (main.c:187).  No data moves between workers, exactly as no data moves between ranks.
 */
#define _GNU_SOURCE
#include "gcn10_host.h"
#include "host_tiff.h"
#include "../../include/gcn10_cuda.h"

#include <errno.h>
#include <limits.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

enum { BAND_ROWS = 2048, NPLANES = GCN10_NPLANES };

static const char *const k_conds[2] = { "drained", "undrained" };      /* cn.c:145 */
static const char *const k_hcs[3] = { "p", "f", "g" };                 /* cn.c:146 */
static const char *const k_arcs[3] = { "i", "ii", "iii" };             /* cn.c:147 */

typedef struct {
    uint8_t *esa;                   /* pinned: BAND_ROWS x pitch */
    uint8_t *planes[NPLANES];       /* pinned: BAND_ROWS x pitch each */
    int y0, rows;
    int state;                      /* 0 free, 1 filled (waiting for its consumer), -1 read failed */
} band_buf;

typedef struct worker {
    int index;                      /* worker number: the "rank" of the log lines */
    int device;                     /* its GPU */
    const gh_run_options *opt;
    gh_blocks *blocks;
    const int (*tables)[256][5];
    const int *ids;
    int n_ids;
    atomic_int *next;               /* shared queue head */
    atomic_int *done;               /* blocks that produced all 18 rasters */
    gh_log *log;                    /* rank_<index>.log */
    gh_log *log0;                   /* worker 0's log: progress lines go there (log.c:199-207) */
    gcn10_ctx *ctx;
    size_t pitch_cap;
    int have_planes;
    band_buf bands[2];
    /* compressed land-cover tiles of a block (GPU-inflate path) */
    uint8_t *tile_blob;             /* pinned */
    size_t tile_blob_cap;
    uint64_t *tile_off;
    uint32_t *tile_size;
    size_t tile_cap;
    /* encoder hand-off */
    pthread_mutex_t mu;
    pthread_cond_t cv;
    gh_tiffw *writers[NPLANES];
    size_t pitch;
    int encode_failed;
    int encoder_quit;
} worker;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void fatal(worker *wk, const char *msg)
{
    /* the reference's MPI_Abort(MPI_COMM_WORLD, 1) tier */
    gh_log_message(wk->log, "ERROR", msg, 1);
    exit(1);
}

static int ensure_bands(worker *wk, size_t pitch, int need_planes)
{
    if (wk->pitch_cap >= pitch && (!need_planes || wk->have_planes))
        return 0;
    for (int b = 0; b < 2; b++) {
        gcn10_cuda_host_free(wk->bands[b].esa);
        wk->bands[b].esa = gcn10_cuda_host_alloc(pitch * BAND_ROWS);
        for (int k = 0; k < NPLANES; k++) {
            gcn10_cuda_host_free(wk->bands[b].planes[k]);
            wk->bands[b].planes[k] = need_planes ? gcn10_cuda_host_alloc(pitch * BAND_ROWS) : NULL;
            if (need_planes && !wk->bands[b].planes[k])
                return -1;
        }
        if (!wk->bands[b].esa)
            return -1;
        wk->bands[b].state = 0;
    }
    wk->pitch_cap = pitch;
    wk->have_planes = need_planes;
    return 0;
}

/* encoder thread: compresses and appends filled bands in order */
static void *encoder_main(void *arg)
{
    worker *wk = arg;
    int turn = 0;
    for (;;) {
        pthread_mutex_lock(&wk->mu);
        while (wk->bands[turn].state != 1 && !wk->encoder_quit)
            pthread_cond_wait(&wk->cv, &wk->mu);
        if (wk->bands[turn].state != 1 && wk->encoder_quit) {
            pthread_mutex_unlock(&wk->mu);
            return NULL;
        }
        pthread_mutex_unlock(&wk->mu);
        band_buf *bb = &wk->bands[turn];
        for (int k = 0; k < NPLANES; k++)
            if (gh_tiffw_write_rows(wk->writers[k], bb->planes[k], wk->pitch, bb->y0, bb->rows, wk->opt->io_threads))
                wk->encode_failed = 1;
        pthread_mutex_lock(&wk->mu);
        bb->state = 0;
        pthread_cond_broadcast(&wk->cv);
        pthread_mutex_unlock(&wk->mu);
        turn ^= 1;
    }
}

/* cn.c:293-360: "<outdir>/cn_<hc>_<arc>_<id>.tif", or "..._<id>_.tif" when the file exists and
 * overwrite is off */
static void output_path(const worker *wk, int cond, int hi, int ai, int block_id, char *path, size_t n)
{
    const char *root = (wk->opt->out_root && *wk->opt->out_root) ? wk->opt->out_root : ".";
    snprintf(path, n, "%s/cn_rasters_%s/cn_%s_%s_%d.tif", root, k_conds[cond], k_hcs[hi], k_arcs[ai], block_id);
    if (!wk->opt->overwrite) {
        FILE *f = fopen(path, "r");
        if (f) {
            fclose(f);
            snprintf(path, n, "%s/cn_rasters_%s/cn_%s_%s_%d_.tif", root, k_conds[cond], k_hcs[hi], k_arcs[ai],
                     block_id);
        }
    }
}

/* ---- GPU-deflate path: reader thread -> gcn10_cuda_block_deflate_rows -> append compressed tiles ---- */

typedef struct {
    worker *wk;
    gh_tiff *esa_ds;
    const gh_window *we;
    int w, h;
    size_t pitch;
    double t_read;
    char err[GH_ERRLEN];
} reader_job;

static void *reader_main(void *arg)
{
    reader_job *rj = arg;
    worker *wk = rj->wk;
    int turn = 0;
    for (int y0 = 0; y0 < rj->h; y0 += BAND_ROWS, turn ^= 1) {
        band_buf *bb = &wk->bands[turn];
        pthread_mutex_lock(&wk->mu);
        while (bb->state != 0 && !wk->encoder_quit)
            pthread_cond_wait(&wk->cv, &wk->mu);
        int quit = wk->encoder_quit;
        pthread_mutex_unlock(&wk->mu);
        if (quit)
            return NULL;
        int rows = rj->h - y0 < BAND_ROWS ? rj->h - y0 : BAND_ROWS;
        double t0 = now_s();
        int rc = gh_tiff_read_window(rj->esa_ds, rj->we->xoff, rj->we->yoff + y0, rj->w, rows, bb->esa, rj->pitch,
                                     wk->opt->io_threads, rj->err, sizeof rj->err);
        rj->t_read += now_s() - t0;
        bb->y0 = y0;
        bb->rows = rows;
        pthread_mutex_lock(&wk->mu);
        bb->state = rc ? -1 : 1;
        pthread_cond_broadcast(&wk->cv);
        pthread_mutex_unlock(&wk->mu);
        if (rc)
            return NULL;
    }
    return NULL;
}

typedef struct {
    worker *wk;
    const gcn10_tile_strip *st;
} sink_job;

/* one plane of a strip: append its tile rows to that plane's GeoTIFF */
static void sink_plane(void *arg, int k)
{
    sink_job *j = arg;
    const gcn10_tile_strip *st = j->st;
    gh_tiffw *tw = j->wk->writers[st->plane_ids[k]];
    for (int tr = 0; tr < st->n_tile_rows; tr++) {
        size_t base = ((size_t)k * st->n_tile_rows + tr) * (size_t)st->tiles_x;
        if (gh_tiffw_put_tile_row(tw, st->tile_row0 + tr, st->blob, st->offsets + base, st->sizes + base)) {
            j->wk->encode_failed = 1;
            return;
        }
    }
}

/* the 18 files are independent: their appends run on the I/O threads while the GPU works on the next strips */
static int tile_sink(void *user, const gcn10_tile_strip *st)
{
    worker *wk = user;
    sink_job j = { wk, st };
    gh_parallel_for(st->n_planes, wk->opt->io_threads > 0 ? wk->opt->io_threads : 1, sink_plane, &j);
    return wk->encode_failed ? 1 : 0;
}

/* returns 0 ok, 1 = land-cover read failed (recoverable tier) */
static int bands_gpu_deflate(worker *wk, int block_id, gh_tiff *esa_ds, const gh_window *we, const uint8_t *hsg,
                             const gh_window *wh, size_t pitch, double *t_read, double *t_gpu)
{
    char msg[1024];
    const int w = we->xcount, h = we->ycount;
    reader_job rj = { wk, esa_ds, we, w, h, pitch, 0.0, "" };
    pthread_t rd;
    wk->encoder_quit = 0;
    wk->bands[0].state = wk->bands[1].state = 0;
    if (pthread_create(&rd, NULL, reader_main, &rj) != 0)
        fatal(wk, "cannot start the reader thread");
    int turn = 0, failed = 0;
    for (int y0 = 0; y0 < h; y0 += BAND_ROWS, turn ^= 1) {
        band_buf *bb = &wk->bands[turn];
        pthread_mutex_lock(&wk->mu);
        while (bb->state == 0)
            pthread_cond_wait(&wk->cv, &wk->mu);
        int st = bb->state;
        pthread_mutex_unlock(&wk->mu);
        if (st < 0) {
            gh_log_message(wk->log, "ERROR", rj.err, 1);                            /* raster.c:182-186 */
            failed = 1;
            break;
        }
        double t1 = now_s();
        int rc = gcn10_cuda_block_deflate_rows(wk->ctx, bb->esa, w, h, bb->y0, bb->rows, pitch, we->gt, hsg, wh->xcount,
                                               wh->ycount, (size_t)wh->xcount, wh->gt, GCN10_MASK_ALL, tile_sink, wk);
        if (rc && !wk->encode_failed) {
            snprintf(msg, sizeof msg, "cuda failure on block %d: %s", block_id, gcn10_cuda_last_error());
            fatal(wk, msg);                         /* CUDA errors are the fatal tier; there is no CPU path */
        }
        *t_gpu += now_s() - t1;
        pthread_mutex_lock(&wk->mu);
        bb->state = 0;
        pthread_cond_broadcast(&wk->cv);
        pthread_mutex_unlock(&wk->mu);
        if (wk->encode_failed)
            break;
    }
    pthread_mutex_lock(&wk->mu);
    wk->encoder_quit = 1;
    pthread_cond_broadcast(&wk->cv);
    pthread_mutex_unlock(&wk->mu);
    pthread_join(rd, NULL);
    *t_read += rj.t_read;
    return failed;
}

/* ---- GPU-inflate path: the land-cover window goes to the device as the DEFLATE tiles of the file
 * (gcn10_cuda_block_tiles_deflate): no decode on the host at all.  Returns 0 ok, 1 = land-cover read or tile
 * decode failed (recoverable tier, raster.c:182-186) */
static int block_gpu_tiles(worker *wk, int block_id, gh_tiff *esa_ds, const gh_window *we, const gh_tile_plan *plan,
                           const uint8_t *hsg, const gh_window *wh, double *t_read, double *t_gpu)
{
    char msg[1024], err[GH_ERRLEN] = "";
    const size_t ntiles = (size_t)plan->tiles_x * (size_t)plan->tiles_y;
    if (wk->tile_blob_cap < plan->blob_bytes + 16) {
        gcn10_cuda_host_free(wk->tile_blob);
        wk->tile_blob_cap = plan->blob_bytes + plan->blob_bytes / 4 + (1u << 20);
        wk->tile_blob = gcn10_cuda_host_alloc(wk->tile_blob_cap);
        if (!wk->tile_blob) {
            wk->tile_blob_cap = 0;
            snprintf(msg, sizeof msg, "pinned allocation failed for block %d: %s", block_id, gcn10_cuda_last_error());
            fatal(wk, msg);
        }
    }
    if (wk->tile_cap < ntiles) {
        free(wk->tile_off);
        free(wk->tile_size);
        wk->tile_off = malloc(ntiles * sizeof *wk->tile_off);
        wk->tile_size = malloc(ntiles * sizeof *wk->tile_size);
        if (!wk->tile_off || !wk->tile_size)
            fatal(wk, "out of memory for raster");
        wk->tile_cap = ntiles;
    }
    double t0 = now_s();
    if (gh_tiff_window_tiles_read(esa_ds, plan, wk->tile_blob, wk->tile_off, wk->tile_size, wk->opt->io_threads, err,
                                  sizeof err)) {
        gh_log_message(wk->log, "ERROR", err, 1);
        return 1;
    }
    double t1 = now_s();
    *t_read += t1 - t0;
    gcn10_tile_source src;
    src.tile_w = plan->tile_w;
    src.tile_h = plan->tile_h;
    src.tiles_x = plan->tiles_x;
    src.tiles_y = plan->tiles_y;
    src.x_off = plan->x_in;
    src.y_off = plan->y_in;
    src.blob = wk->tile_blob;
    src.blob_bytes = plan->blob_bytes;
    src.offsets = wk->tile_off;
    src.sizes = wk->tile_size;
    int rc = gcn10_cuda_block_tiles_deflate(wk->ctx, &src, we->xcount, we->ycount, we->gt, hsg, wh->xcount, wh->ycount,
                                            (size_t)wh->xcount, wh->gt, GCN10_MASK_ALL, tile_sink, wk);
    *t_gpu += now_s() - t1;
    if (rc == GCN10_EDATA) {
        snprintf(msg, sizeof msg, "gdalrasterio error 3 (%s)", gcn10_cuda_last_error());
        gh_log_message(wk->log, "ERROR", msg, 1);
        return 1;
    }
    if (rc && !wk->encode_failed) {
        snprintf(msg, sizeof msg, "cuda failure on block %d: %s", block_id, gcn10_cuda_last_error());
        fatal(wk, msg);
    }
    return 0;
}

/* ---- host-deflate path: raw planes back, zlib thread pool encodes band i while band i+1 is computed ---- */

static int bands_host_deflate(worker *wk, int block_id, gh_tiff *esa_ds, const gh_window *we, const uint8_t *hsg,
                              const gh_window *wh, size_t pitch, double *t_read, double *t_gpu)
{
    char msg[1024], err[GH_ERRLEN] = "";
    const int w = we->xcount, h = we->ycount;
    wk->encoder_quit = 0;
    wk->bands[0].state = wk->bands[1].state = 0;
    pthread_t enc;
    if (pthread_create(&enc, NULL, encoder_main, wk) != 0)
        fatal(wk, "cannot start the encoder thread");
    int turn = 0, failed = 0;
    for (int y0 = 0; y0 < h && !failed; y0 += BAND_ROWS, turn ^= 1) {
        band_buf *bb = &wk->bands[turn];
        int rows = h - y0 < BAND_ROWS ? h - y0 : BAND_ROWS;
        pthread_mutex_lock(&wk->mu);
        while (bb->state != 0)
            pthread_cond_wait(&wk->cv, &wk->mu);
        pthread_mutex_unlock(&wk->mu);
        double t0 = now_s();
        if (gh_tiff_read_window(esa_ds, we->xoff, we->yoff + y0, w, rows, bb->esa, pitch, wk->opt->io_threads, err,
                                sizeof err)) {
            gh_log_message(wk->log, "ERROR", err, 1);                               /* raster.c:182-186 */
            failed = 1;
            break;
        }
        double t1 = now_s();
        int rc = gcn10_cuda_block_rows(wk->ctx, bb->esa, w, h, y0, rows, pitch, we->gt, hsg, wh->xcount, wh->ycount,
                                       (size_t)wh->xcount, wh->gt, GCN10_MASK_ALL, bb->planes, pitch);
        if (rc) {
            snprintf(msg, sizeof msg, "cuda failure on block %d: %s", block_id, gcn10_cuda_last_error());
            fatal(wk, msg);
        }
        *t_read += t1 - t0;
        *t_gpu += now_s() - t1;
        bb->y0 = y0;
        bb->rows = rows;
        pthread_mutex_lock(&wk->mu);
        bb->state = 1;
        pthread_cond_broadcast(&wk->cv);
        pthread_mutex_unlock(&wk->mu);
    }
    pthread_mutex_lock(&wk->mu);
    wk->encoder_quit = 1;
    pthread_cond_broadcast(&wk->cv);
    pthread_mutex_unlock(&wk->mu);
    pthread_join(enc, NULL);
    return failed;
}

static int gh_process_block(worker *wk, int block_id, int total_blocks)
{
    char msg[8192], err[GH_ERRLEN] = "";
    double bbox[4];
    double t_start = now_s();

    /* block geometry (cn.c:155-184) */
    if (gh_blocks_bbox(wk->blocks, block_id, bbox)) {
        snprintf(msg, sizeof msg, "block %d not found", block_id);                  /* cn.c:173 */
        gh_log_message(wk->log, "ERROR", msg, 1);
        return -1;
    }

    /* land cover window (cn.c:187 -> raster.c:106-189) */
    gh_tiff *esa_ds = NULL, *hsg_ds = NULL;
    gh_window we, wh;
    int rw, rh;
    double t[6];
    if (gh_tiff_open(wk->opt->cfg.esa_data_path, &esa_ds, err, sizeof err)) {
        gh_log_message(wk->log, "ERROR", err, 1);
        goto esa_failed;
    }
    gh_tiff_size(esa_ds, &rw, &rh);
    gh_tiff_geotransform(esa_ds, t);
    if (gh_raster_window(rw, rh, t, bbox, &we)) {
        snprintf(msg, sizeof msg, "invalid raster bounds for %s", wk->opt->cfg.esa_data_path);   /* raster.c:143 */
        gh_log_message(wk->log, "ERROR", msg, 1);
        goto esa_failed;
    }

    /* soil window (cn.c:195-196) */
    if (gh_tiff_open(wk->opt->cfg.hysogs_data_path, &hsg_ds, err, sizeof err)) {
        gh_log_message(wk->log, "ERROR", err, 1);
        goto hsg_failed;
    }
    gh_tiff_size(hsg_ds, &rw, &rh);
    gh_tiff_geotransform(hsg_ds, t);
    if (gh_raster_window(rw, rh, t, bbox, &wh)) {
        snprintf(msg, sizeof msg, "invalid raster bounds for %s", wk->opt->cfg.hysogs_data_path);
        gh_log_message(wk->log, "ERROR", msg, 1);
        goto hsg_failed;
    }
    uint8_t *hsg = malloc((size_t)wh.xcount * (size_t)wh.ycount);
    if (!hsg)
        fatal(wk, "out of memory for raster");                                      /* raster.c:171 */
    if (gh_tiff_read_window(hsg_ds, wh.xoff, wh.yoff, wh.xcount, wh.ycount, hsg, (size_t)wh.xcount,
                            wk->opt->io_threads, err, sizeof err)) {
        gh_log_message(wk->log, "ERROR", err, 1);
        free(hsg);
        goto hsg_failed;
    }
    gh_tiff_close(hsg_ds);
    hsg_ds = NULL;

    const int w = we.xcount, h = we.ycount;
    const size_t pitch = ((size_t)w + 255) / 256 * 256;
    const char *hd = getenv("GCN10_HOST_DEFLATE");
    const int host_deflate = hd && *hd && *hd != '0';
    const char *hi = getenv("GCN10_HOST_INFLATE");
    const int host_inflate = hi && *hi && *hi != '0';
    /* a tiled DEFLATE land-cover file (what GDAL writes for COMPRESS=DEFLATE TILED=YES, and what the ESA WorldCover
     * files are) is handed to the GPU compressed; anything else is decoded here, band by band */
    gh_tile_plan plan;
    const int gpu_inflate = !host_deflate && !host_inflate &&
                            gh_tiff_window_tiles_plan(esa_ds, we.xoff, we.yoff, w, h, &plan) == 0;
    if (!gpu_inflate && ensure_bands(wk, pitch, host_deflate)) {
        snprintf(msg, sizeof msg, "pinned allocation failed for block %d: %s", block_id, gcn10_cuda_last_error());
        fatal(wk, msg);
    }
    wk->pitch = pitch;

    /* output directories and files (cn.c:236-256, 293-360) */
    const char *root = (wk->opt->out_root && *wk->opt->out_root) ? wk->opt->out_root : ".";
    char paths[NPLANES][PATH_MAX];
    for (int c = 0; c < 2; c++) {
        char outdir[PATH_MAX];
        snprintf(outdir, sizeof outdir, "%s/cn_rasters_%s", root, k_conds[c]);
        if (mkdir(outdir, 0755) != 0 && errno != EEXIST) {
            snprintf(msg, sizeof msg, "failed to create output directory %s", outdir);   /* cn.c:250 */
            fatal(wk, msg);
        }
    }
    int opened = 0;
    for (int k = 0; k < NPLANES; k++) {
        output_path(wk, k / 9, (k % 9) / 3, k % 3, block_id, paths[k], sizeof paths[k]);
        if (gh_tiffw_open(paths[k], w, h, we.gt, &wk->writers[k], err, sizeof err)) {
            gh_log_message(wk->log, "ERROR", err, 1);                               /* raster.c:220-223 */
            break;
        }
        opened++;
    }
    if (opened < NPLANES) {
        for (int k = 0; k < opened; k++)
            gh_tiffw_abort(wk->writers[k]);
        free(hsg);
        gh_tiff_close(esa_ds);
        return -1;
    }

    /* band loop */
    wk->encode_failed = 0;
    double t_read = 0, t_gpu = 0;
    int failed = gpu_inflate ? block_gpu_tiles(wk, block_id, esa_ds, &we, &plan, hsg, &wh, &t_read, &t_gpu)
                 : host_deflate ? bands_host_deflate(wk, block_id, esa_ds, &we, hsg, &wh, pitch, &t_read, &t_gpu)
                                : bands_gpu_deflate(wk, block_id, esa_ds, &we, hsg, &wh, pitch, &t_read, &t_gpu);
    free(hsg);
    gh_tiff_close(esa_ds);

    if (failed) {
        for (int k = 0; k < NPLANES; k++)
            gh_tiffw_abort(wk->writers[k]);
        snprintf(msg, sizeof msg, "esa load failed for block %d", block_id);        /* cn.c:189 */
        gh_log_message(wk->log, "ERROR", msg, 1);
        return -1;
    }
    int ok = 1;
    for (int k = 0; k < NPLANES; k++) {
        if (gh_tiffw_close(wk->writers[k]) || wk->encode_failed) {
            snprintf(msg, sizeof msg, "write error 3 on %s", paths[k]);             /* raster.c:221 */
            gh_log_message(wk->log, "ERROR", msg, 1);
            ok = 0;
        }
        /* the reference logs each raster and reports it to rank 0 whether or not the write worked
         * (cn.c:363-373) */
        snprintf(msg, sizeof msg, "completed condition for %d: %s/%s/%s", block_id, k_conds[k / 9],
                 k_hcs[(k % 9) / 3], k_arcs[k % 3]);                                /* cn.c:366-369 */
        gh_log_message(wk->log, "INFO", msg, 0);
        snprintf(msg, sizeof msg, "progress: completed block %d / total %d", block_id, total_blocks);
        gh_log_message(wk->log0, "INFO", msg, 0);                                   /* log.c:199-207 */
    }
    double dt = now_s() - t_start;
    snprintf(msg, sizeof msg,
             "block %d: %d x %d px, 18 rasters in %.2f s (%.1f Mpx/s; decode %.2f s, gpu+copies %.2f s; %s deflate)%s",
             block_id, w, h, dt, (double)w * h / dt / 1e6, t_read, t_gpu, host_deflate ? "host" : "gpu",
             gpu_inflate ? " [land cover inflated on the gpu]" : "");
    gh_log_message(wk->log, "INFO", msg, 0);
    return ok ? 0 : -1;

hsg_failed:
    gh_tiff_close(hsg_ds);
    gh_tiff_close(esa_ds);
    snprintf(msg, sizeof msg, "hysogs load failed for block %d", block_id);         /* cn.c:198-199 */
    gh_log_message(wk->log, "ERROR", msg, 1);
    return -1;
esa_failed:
    gh_tiff_close(esa_ds);
    snprintf(msg, sizeof msg, "esa load failed for block %d", block_id);            /* cn.c:189 */
    gh_log_message(wk->log, "ERROR", msg, 1);
    return -1;
}

static void *worker_main(void *arg)
{
    worker *wk = arg;
    char msg[256];
    /* keep this worker (and the reader / encoder threads it spawns, which inherit the mask) on the
     * NUMA node of its GPU so that the pinned band buffers are node-local */
    int node = gcn10_cuda_bind_host_thread(wk->device);
    snprintf(msg, sizeof msg, "worker %d on gpu %d, numa node %d", wk->index, wk->device, node);
    gh_log_message(wk->log, "INFO", msg, 0);
    if (gcn10_cuda_create(wk->device, &wk->ctx) || gcn10_cuda_set_luts(wk->ctx, wk->tables)) {
        snprintf(msg, sizeof msg, "cannot initialise GPU %d: %s", wk->device, gcn10_cuda_last_error());
        fatal(wk, msg);
    }
    pthread_mutex_init(&wk->mu, NULL);
    pthread_cond_init(&wk->cv, NULL);
    for (;;) {
        int i = atomic_fetch_add(wk->next, 1);                  /* replaces i = rank; i += size (main.c:171) */
        if (i >= wk->n_ids)
            break;
        snprintf(msg, sizeof msg, "processing block %d", wk->ids[i]);              /* main.c:172-173 */
        gh_log_message(wk->log, "INFO", msg, 1);
        if (gh_process_block(wk, wk->ids[i], wk->n_ids) == 0)
            atomic_fetch_add(wk->done, 1);
    }
    for (int b = 0; b < 2; b++) {
        gcn10_cuda_host_free(wk->bands[b].esa);
        for (int k = 0; k < NPLANES; k++)
            gcn10_cuda_host_free(wk->bands[b].planes[k]);
    }
    gcn10_cuda_host_free(wk->tile_blob);
    free(wk->tile_off);
    free(wk->tile_size);
    gcn10_cuda_destroy(wk->ctx);
    pthread_cond_destroy(&wk->cv);
    pthread_mutex_destroy(&wk->mu);
    return NULL;
}

int gh_run_blocks(const gh_run_options *opt, const int *block_ids, int n_blocks)
{
    char err[GH_ERRLEN] = "", msg[1024];
    static int tables[9][256][5];

    int ngpu = gcn10_cuda_device_count();
    if (ngpu <= 0) {
        fprintf(stderr, "gcn10: no usable CUDA device (%s); there is no CPU fallback\n", gcn10_cuda_last_error());
        exit(1);
    }
    /* several workers per GPU keep more than one block in flight on it: the kernels of one block (inflate, then
     * the fused Curve Number + DEFLATE kernel) are latency bound and overlap with the other block's copies */
    const int gpus = opt->n_gpus > 0 && opt->n_gpus < ngpu ? opt->n_gpus : ngpu;
    int nworkers = gpus * (opt->workers_per_gpu > 0 ? opt->workers_per_gpu : 1);
    if (nworkers > n_blocks)
        nworkers = n_blocks > 0 ? n_blocks : 1;

    gh_log *log0 = gh_log_open(opt->cfg.log_dir, 0);
    /* lookup tables once per run (the reference re-reads the CSV 18 times per block, cn.c:261) */
    if (gh_load_lookup_tables(opt->cfg.lookup_table_path, tables, err, sizeof err)) {
        gh_log_message(log0, "ERROR", err, 1);                  /* fatal in the reference: cn.c:25,32,47 */
        exit(1);
    }
    gh_blocks *blocks = NULL;
    if (gh_blocks_open(opt->cfg.blocks_shp_path, &blocks, err, sizeof err)) {
        gh_log_message(log0, "ERROR", err, 1);
        exit(1);
    }
    snprintf(msg, sizeof msg, "processing %d blocks on %d gpu workers (%d gpus)", n_blocks, nworkers,
             gpus < nworkers ? gpus : nworkers);
    gh_log_message(log0, "INFO", msg, 1);

    atomic_int next = 0, done = 0;
    worker *wks = calloc((size_t)nworkers, sizeof *wks);
    pthread_t *th = calloc((size_t)nworkers, sizeof *th);
    for (int i = 0; i < nworkers; i++) {
        wks[i].index = i;
        wks[i].device = i % gpus;
        wks[i].opt = opt;
        wks[i].blocks = blocks;
        wks[i].tables = tables;
        wks[i].ids = block_ids;
        wks[i].n_ids = n_blocks;
        wks[i].next = &next;
        wks[i].done = &done;
        wks[i].log0 = log0;
        wks[i].log = i == 0 ? log0 : gh_log_open(opt->cfg.log_dir, i);
        pthread_create(&th[i], NULL, worker_main, &wks[i]);
    }
    for (int i = 0; i < nworkers; i++)
        pthread_join(th[i], NULL);              /* the barrier of main.c:187 */
    snprintf(msg, sizeof msg, "processed %d blocks on %d ranks", n_blocks, nworkers);      /* main.c:191-193 */
    gh_log_message(log0, "INFO", msg, 1);
    for (int i = nworkers - 1; i >= 0; i--)
        gh_log_close(wks[i].log);
    gh_blocks_close(blocks);
    free(wks);
    free(th);
    return atomic_load(&done);
}
