/* oracle/cn_oracle.c -- TEST INFRASTRUCTURE ONLY (see cn_oracle.h for the rules
 * on who may load it and for the parity-pinning statement).
 *
 * Plain-C restatement of the Curve Number hot path of clawrim/gcn10, written
 * from the behaviour of /root/reference/src/cn.c and src/raster.c.  It is NOT the
 * product: the product is the CUDA path behind include/gcn10_cuda.h.
 *
 * Build: oracle/Makefile, -O2 -ffp-contract=off and no -march, so that the fp64
 * index arithmetic is evaluated as separate IEEE multiply / add / subtract /
 * divide operations exactly like the reference's -O3 x86-64 build (which has no
 * FMA instructions available: src/CMakeLists.txt:68 passes no -march).
 */
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cn_oracle.h"

/* (int) of a double as the reference's x86-64 build performs it: cvttsd2si
 * truncates toward zero and yields INT_MIN ("integer indefinite") for NaN and for
 * anything outside int range.  In-range behaviour is ordinary C. */
static int int_from_double_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int)v;
}

static int clamp_index(int i, int n)
{
    /* cn.c:228-229 */
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

int cn_oracle_parse_lookup(const char *csv_path, int table[256][5])
{
    /* cn.c:13-85.  Lines are consumed through a 128-byte fgets buffer (cn.c:17,50),
     * the first line is discarded unseen (cn.c:43), rows are "<lc>_<letter>,<cn>". */
    char line[128];
    FILE *fp = fopen(csv_path, "r");

    if (!fp)
        return -1;                                  /* cn.c:28-33 (fatal there) */
    for (int lc = 0; lc < 256; lc++)
        for (int sg = 0; sg < 5; sg++)
            table[lc][sg] = CN_ORACLE_NODATA;       /* cn.c:36-40 */
    if (!fgets(line, sizeof line, fp)) {            /* cn.c:43-48 */
        fclose(fp);
        return -2;
    }
    while (fgets(line, sizeof line, fp)) {
        char *code = strtok(line, ",");             /* cn.c:51 */
        if (!code)
            continue;
        char *sep = strchr(code, '_');              /* cn.c:56 */
        if (!sep)
            continue;                               /* logged + skipped, cn.c:57-62 */
        *sep = '\0';
        int lc = atoi(code);                        /* cn.c:65 */
        char letter = sep[1];
        int sg = letter == 'A' ? 1 : letter == 'B' ? 2 : letter == 'C' ? 3 : 4;   /* cn.c:66 */
        char *val = strtok(NULL, ",");              /* cn.c:67 */
        if (!val)
            continue;                               /* cn.c:68-73 */
        int cn = atoi(val);                         /* cn.c:74 */
        if (lc >= 0 && lc < 256)                    /* cn.c:75 (sg is always 1..4) */
            table[lc][sg] = cn;
    }
    fclose(fp);
    return 0;
}

int cn_oracle_load_tables(const char *lookup_dir, int tables[9][256][5])
{
    static const char *hc[3] = { "p", "f", "g" };           /* cn.c:146 */
    static const char *arc[3] = { "i", "ii", "iii" };       /* cn.c:147 */
    char path[4096];

    for (int h = 0; h < 3; h++) {
        for (int a = 0; a < 3; a++) {
            /* cn.c:21-22 */
            snprintf(path, sizeof path, "%s/default_lookup_%s_%s.csv", lookup_dir, hc[h], arc[a]);
            int rc = cn_oracle_parse_lookup(path, tables[h * 3 + a]);
            if (rc)
                return rc;
        }
    }
    return 0;
}

int cn_oracle_window(int raster_w, int raster_h, const double t[6], const double bbox[4],
                     int *xoff, int *yoff, int *xcount, int *ycount, double gt[6])
{
    /* raster.c:127-130; bbox = {minx, miny, maxx, maxy} (cn.c:179-182) */
    int xo = int_from_double_x86(floor((bbox[0] - t[0]) / t[1]));
    int yo = int_from_double_x86(floor((bbox[3] - t[3]) / t[5]));
    int xc = int_from_double_x86(ceil((bbox[2] - bbox[0]) / t[1]));
    int yc = int_from_double_x86(ceil((bbox[1] - bbox[3]) / t[5]));

    if (xo < 0) {                                   /* raster.c:134-137 */
        xc += xo;
        xo = 0;
    }
    if (yo < 0) {                                   /* raster.c:138-141 */
        yc += yo;
        yo = 0;
    }
    if (xo >= raster_w || yo >= raster_h || xc <= 0 || yc <= 0)
        return 1;                                   /* raster.c:142-147 */
    if (xo + xc > raster_w)                         /* raster.c:148-153 */
        xc = raster_w - xo;
    if (yo + yc > raster_h)
        yc = raster_h - yo;

    *xoff = xo;
    *yoff = yo;
    *xcount = xc;
    *ycount = yc;
    gt[0] = t[0] + xo * t[1];                       /* raster.c:157-162 */
    gt[1] = t[1];
    gt[2] = t[2];
    gt[3] = t[3] + yo * t[5];
    gt[4] = t[4];
    gt[5] = t[5];
    return 0;
}

void cn_oracle_col_index(int w, const double gt[6], const double soil_gt[6], int hsx, int32_t *ci)
{
    for (int x = 0; x < w; x++) {
        double px = gt[0] + (x + 0.5) * gt[1];              /* cn.c:222 */
        double dc = (px - soil_gt[0]) / soil_gt[1];         /* cn.c:223 */
        ci[x] = clamp_index(int_from_double_x86(round(dc)), hsx);   /* cn.c:225,228 */
    }
}

void cn_oracle_row_index(int h, const double gt[6], const double soil_gt[6], int hsy, int32_t *cj)
{
    for (int y = 0; y < h; y++) {
        double py = gt[3] + (y + 0.5) * gt[5];              /* cn.c:219 */
        double dr = (soil_gt[3] - py) / fabs(soil_gt[5]);   /* cn.c:224 */
        cj[y] = clamp_index(int_from_double_x86(round(dr)), hsy);   /* cn.c:226,229 */
    }
}

void cn_oracle_resample_rows(const uint8_t *coarse, int hsx, int hsy, const double soil_gt[6],
                             int w, int h, const double gt[6], int y0, int y1, uint8_t *dst)
{
    int32_t *ci = malloc(sizeof(int32_t) * (size_t)(w > 0 ? w : 1));
    int32_t *cj = malloc(sizeof(int32_t) * (size_t)(h > 0 ? h : 1));

    cn_oracle_col_index(w, gt, soil_gt, hsx, ci);
    cn_oracle_row_index(h, gt, soil_gt, hsy, cj);
    for (int y = y0; y < y1; y++) {
        const uint8_t *src_row = coarse + (size_t)cj[y] * hsx;
        uint8_t *dst_row = dst + (size_t)(y - y0) * w;
        for (int x = 0; x < w; x++)
            dst_row[x] = src_row[ci[x]];                    /* cn.c:230 */
    }
    free(ci);
    free(cj);
}

void cn_oracle_remap_hsg(uint8_t *hsg, size_t n, int drained)
{
    if (drained) {                                          /* cn.c:92-98 */
        for (size_t i = 0; i < n; i++)
            if (hsg[i] >= 11 && hsg[i] <= 14)
                hsg[i] = 4;
    }
    else {                                                  /* cn.c:99-110 */
        for (size_t i = 0; i < n; i++) {
            uint8_t v = hsg[i];
            if (v >= 11 && v <= 14)
                hsg[i] = (uint8_t)(v - 10);
        }
    }
}

void cn_oracle_apply_table(const uint8_t *esa, const uint8_t *hsg, size_t n,
                           const int table[256][5], uint8_t *out)
{
    memset(out, CN_ORACLE_NODATA, n);                       /* cn.c:289 */
    for (size_t i = 0; i < n; i++) {
        int sg = hsg[i];
        if (sg < 5) {                                       /* cn.c:123-124 */
            int cn = table[esa[i]][sg];
            if (cn < 255)                                   /* cn.c:126 */
                out[i] = (uint8_t)cn;                       /* cn.c:127 */
        }
    }
}

int cn_oracle_block_rows(const uint8_t *esa_rows, int w, int h, const double gt[6],
                         const uint8_t *coarse, int hsx, int hsy, const double soil_gt[6],
                         const int tables[9][256][5], int y0, int y1,
                         uint8_t *out, size_t plane_stride)
{
    size_t n = (size_t)(y1 - y0) * (size_t)w;
    uint8_t *resampled = malloc(n ? n : 1);
    uint8_t *adjusted = malloc(n ? n : 1);

    if (!resampled || !adjusted) {
        free(resampled);
        free(adjusted);
        return -1;
    }
    cn_oracle_resample_rows(coarse, hsx, hsy, soil_gt, w, h, gt, y0, y1, resampled);
    for (int cond = 0; cond < 2; cond++) {                  /* cn.c:145,236: drained, undrained */
        for (int t = 0; t < 9; t++) {                       /* cn.c:258-259 */
            memcpy(adjusted, resampled, n);                 /* cn.c:274 */
            cn_oracle_remap_hsg(adjusted, n, cond == 0);    /* cn.c:275 */
            cn_oracle_apply_table(esa_rows, adjusted, n, tables[t],
                                  out + (size_t)(cond * 9 + t) * plane_stride);
        }
    }
    free(resampled);
    free(adjusted);
    return 0;
}
