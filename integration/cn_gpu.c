/* integration/cn_gpu.c -- the reference-side binding, as code.
 *
 * A replacement for the reference's src/cn.c that a maintainer of clawrim/gcn10 would add to put the per-block
 * hot path on a B200: it compiles against the reference's own src/global.h, keeps the reference's
 * process_block() seam (global.h:58, called from main.c:175), reads its windows with the reference's unmodified
 * load_raster() (raster.c:106-189), writes its rasters with the reference's unmodified save_raster()
 * (raster.c:192-227), logs through the reference's log_message() / report_block_completion() -- and replaces
 * the five CPU passes per raster of cn.c:218-290 with ONE call into libgcn10cuda (include/gcn10_cuda.h).
 *
 *   build in the reference tree:   cc -std=c99 -I<this repo>/include -c cn_gpu.c   (instead of cn.c)
 *                                  link with -lgcn10cuda next to -lgdal and MPI
 *   build here (no MPI / GDAL):    make -C integration      -> oracle/_ref/libgcn10_gpu_ref.so, against the
 *                                  RAM GDAL/OGR/MPI stand-ins of oracle/refshim, exactly like oracle/_ref/
 *                                  libgcn10_ref.so is built from the reference's cn.c.  tests/test_gpu_integration.py
 *                                  runs both libraries on the same rasters and compares the 18 buffers that reach
 *                                  save_raster(), their order, sizes, geotransforms and file names.
 *
 * One MPI rank drives one GPU (rank % device count), so `mpirun -n 8 gcn10 ...` (src/test/run_test.py:66-69)
 * maps onto an 8 x B200 box unchanged.  Observable behaviour kept from cn.c: messages and error tiers
 * (recoverable -> log + return, fatal -> MPI_Abort), output names and the no-overwrite underscore rule
 * (cn.c:293-360), one "completed condition" line and one completion report per raster (cn.c:366-373), loop order
 * cond -> hc -> arc (cn.c:236,258-259).  Differences: the nine lookup CSVs are parsed once per process instead of
 * 18 times per block, and the raster of every (cond, hc, arc) is computed before the first one is saved.
 *
 * The asynchronous form is used: gcn10_cuda_block_async() queues the block and returns, the rank creates the output
 * directories and builds its 18 output paths meanwhile, then waits.  A rank that pipelines its own I/O would
 * instead read block i+1 (load_raster) before gcn10_cuda_wait(block i).
 */
#include "global.h"             /* the reference's own header: MPI, GDAL, prototypes, config globals */

#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "gcn10_cuda.h"

static const char *const k_cond[2] = { "drained", "undrained" };       /* cn.c:145 */
static const char *const k_hc[3] = { "p", "f", "g" };                  /* cn.c:146 */
static const char *const k_arc[3] = { "i", "ii", "iii" };              /* cn.c:147 */

static gcn10_ctx *g_ctx;        /* one context per rank, created with the first block */

static void die(const char *what, int block_id)
{
    char msg[1024];
    snprintf(msg, sizeof msg, "%s (block %d): %s", what, block_id, gcn10_cuda_last_error());
    log_message("ERROR", msg, true);
    MPI_Abort(MPI_COMM_WORLD, 1);          /* there is no CPU fallback */
}

/* default_lookup_<hc>_<arc>.csv -> table[256][5]: same file name, same accepted syntax and the same messages as the
 * reference's static load_lookup_table() (cn.c:13-85): header line skipped, rows "<lc>_<A|B|C|D>,<cn>", any other
 * letter means D, rows without '_' or without a value are reported and skipped, lc outside 0..255 ignored. */
static void read_lookup(const char *hc, const char *arc, int table[256][5])
{
    char path[PATH_MAX], row[128], msg[8192];
    if (snprintf(path, sizeof path, "%s/default_lookup_%s_%s.csv", lookup_table_path, hc, arc) >= (int)sizeof path) {
        snprintf(msg, sizeof msg, "lookup table path too long: %s", path);
        log_message("ERROR", msg, true);
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    FILE *fp = fopen(path, "r");
    if (!fp) {
        snprintf(msg, sizeof msg, "cannot open lookup table %s", path);
        log_message("ERROR", msg, true);
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    for (int i = 0; i < 256 * 5; i++)
        (&table[0][0])[i] = GCN10_NODATA;
    if (!fgets(row, sizeof row, fp)) {
        snprintf(msg, sizeof msg, "empty lookup table %s", path);
        log_message("ERROR", msg, true);
        fclose(fp);
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    while (fgets(row, sizeof row, fp)) {
        char *key = row + strspn(row, ",");                     /* first comma-separated field */
        if (!*key)
            continue;
        char *val = key + strcspn(key, ",");
        if (*val)
            *val++ = '\0';
        char *bar = strchr(key, '_');
        if (!bar) {
            snprintf(msg, sizeof msg, "invalid grid_code %s in %s", key, path);
            log_message("ERROR", msg, true);
            continue;
        }
        *bar = '\0';
        const int lc = atoi(key);
        const int sg = bar[1] == 'A' ? 1 : bar[1] == 'B' ? 2 : bar[1] == 'C' ? 3 : 4;
        val += strspn(val, ",");
        if (!*val) {
            snprintf(msg, sizeof msg, "missing cn value in %s", path);
            log_message("ERROR", msg, true);
            continue;
        }
        val[strcspn(val, ",")] = '\0';
        if (lc >= 0 && lc < 256)
            table[lc][sg] = atoi(val);
    }
    fclose(fp);
}

static void gpu_start(int block_id)
{
    static int tables[GCN10_NVARIANTS][256][5];
    int rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const int ndev = gcn10_cuda_device_count();
    if (ndev <= 0 || gcn10_cuda_create(rank % ndev, &g_ctx) != GCN10_OK)
        die("cannot initialise the GPU", block_id);
    for (int h = 0; h < 3; h++)             /* the order of cn.c:258-261, hoisted out of the block loop */
        for (int a = 0; a < 3; a++)
            read_lookup(k_hc[h], k_arc[a], tables[h * 3 + a]);
    if (gcn10_cuda_set_luts(g_ctx, tables) != GCN10_OK)
        die("cannot install the lookup tables", block_id);
}

/* bbox of the block by its "ID" attribute (cn.c:155-184); 0 ok, 1 = skip the block */
static int block_bbox(int block_id, double bbox[4])
{
    char msg[8192], filter[64];
    OGRDataSourceH ds = OGROpen(blocks_shp_path, FALSE, NULL);
    if (!ds) {
        snprintf(msg, sizeof msg, "ogr open failed: %s", blocks_shp_path);
        log_message("ERROR", msg, true);
        return 1;
    }
    OGRLayerH layer = OGR_DS_GetLayer(ds, 0);
    if (snprintf(filter, sizeof filter, "\"ID\"=%d", block_id) >= (int)sizeof filter) {
        snprintf(msg, sizeof msg, "filter string too long for block %d", block_id);
        log_message("ERROR", msg, true);
        OGR_DS_Destroy(ds);
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
    OGR_L_SetAttributeFilter(layer, filter);
    OGRFeatureH feat = OGR_L_GetNextFeature(layer);
    if (!feat) {
        snprintf(msg, sizeof msg, "block %d not found", block_id);
        log_message("ERROR", msg, true);
        OGR_DS_Destroy(ds);
        return 1;
    }
    OGREnvelope env;
    OGR_G_GetEnvelope(OGR_F_GetGeometryRef(feat), &env);
    bbox[0] = env.MinX;
    bbox[1] = env.MinY;
    bbox[2] = env.MaxX;
    bbox[3] = env.MaxY;
    OGR_F_Destroy(feat);
    OGR_DS_Destroy(ds);
    return 0;
}

void process_block(int block_id, bool overwrite, int total_blocks)
{
    char msg[8192];
    double bbox[4], gt[6], soil_gt[6];
    int w, h, hsx, hsy;
    OGRSpatialReferenceH srs, soil_srs;

    if (block_bbox(block_id, bbox))
        return;

    /* the two windows, read by the reference's own load_raster() (cn.c:187-204) */
    uint8_t *esa = load_raster(esa_data_path, bbox, &w, &h, gt, &srs);
    if (!esa) {
        snprintf(msg, sizeof msg, "esa load failed for block %d", block_id);
        log_message("ERROR", msg, true);
        return;
    }
    uint8_t *hsg = load_raster(hysogs_data_path, bbox, &hsx, &hsy, soil_gt, &soil_srs);
    if (!hsg) {
        snprintf(msg, sizeof msg, "hysogs load failed for block %d", block_id);
        log_message("ERROR", msg, true);
        free(esa);
        return;
    }

    /* cn.c:208-290 for all 18 rasters: one asynchronous call */
    if (!g_ctx)
        gpu_start(block_id);
    const size_t npix = (size_t)w * (size_t)h;
    uint8_t *planes[GCN10_NPLANES];
    for (int k = 0; k < GCN10_NPLANES; k++) {
        planes[k] = malloc(npix);
        if (!planes[k]) {
            snprintf(msg, sizeof msg, "malloc failed for cn raster, block %d", block_id);
            log_message("ERROR", msg, true);
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
    }
    gcn10_event *done = NULL;
    if (gcn10_cuda_block_async(g_ctx, esa, w, h, (size_t)w, gt, hsg, hsx, hsy, (size_t)hsx, soil_gt, GCN10_MASK_ALL,
                               planes, (size_t)w, &done) != GCN10_OK)
        die("gcn10_cuda_block_async failed", block_id);

    /* meanwhile: output directories and the 18 file names (cn.c:236-256, 293-360) */
    char outpath[GCN10_NPLANES][PATH_MAX];
    for (int c = 0; c < 2; c++) {
        char outdir[PATH_MAX];
        if (snprintf(outdir, sizeof outdir, "cn_rasters_%s", k_cond[c]) >= (int)sizeof outdir) {
            snprintf(msg, sizeof msg, "output directory path too long for %s", k_cond[c]);
            log_message("ERROR", msg, true);
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
        if (mkdir(outdir, 0755) != 0 && errno != EEXIST) {
            snprintf(msg, sizeof msg, "failed to create output directory %s", outdir);
            log_message("ERROR", msg, true);
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
    }

    if (gcn10_cuda_wait(done) != GCN10_OK)
        die("curve number kernel failed", block_id);
    free(hsg);
    free(esa);

    /* save in the reference's order; the name of a raster is decided when its turn comes, as in cn.c:320-360 */
    for (int c = 0; c < 2; c++)
        for (int hi = 0; hi < 3; hi++)
            for (int ai = 0; ai < 3; ai++) {
                const int k = c * 9 + hi * 3 + ai;
                char *path = outpath[k];
                if (snprintf(path, PATH_MAX, "cn_rasters_%s/cn_%s_%s_%d.tif", k_cond[c], k_hc[hi], k_arc[ai], block_id) >=
                    PATH_MAX) {
                    snprintf(msg, sizeof msg, "output path too long for block %d", block_id);
                    log_message("ERROR", msg, true);
                    MPI_Abort(MPI_COMM_WORLD, 1);
                }
                if (!overwrite) {
                    FILE *f = fopen(path, "r");
                    if (f) {
                        fclose(f);
                        snprintf(path, PATH_MAX, "cn_rasters_%s/cn_%s_%s_%d_.tif", k_cond[c], k_hc[hi], k_arc[ai],
                                 block_id);
                    }
                }
                save_raster(planes[k], w, h, gt, srs, path);                        /* cn.c:363 */
                snprintf(msg, sizeof msg, "completed condition for %d: %s/%s/%s", block_id, k_cond[c], k_hc[hi],
                         k_arc[ai]);
                log_message("INFO", msg, false);                                    /* cn.c:366-369 */
                report_block_completion(block_id, total_blocks);                    /* cn.c:373 */
                free(planes[k]);
            }
}
