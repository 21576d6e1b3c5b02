/* oracle/cn_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of GCN10's per-block Curve Number path, used as the parity
 * checker for the CUDA product path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; nothing
 * under gcn10_b200/ may.  Parity status: PINNED -- tests/test_oracle_pinned.py
 * checks every function here against the reference's own object code
 * (oracle/_ref/libgcn10_ref.so = /root/reference/src/{cn.c,raster.c} compiled
 * unmodified) and against the committed fixtures in tests/golden/ that were
 * generated from it (tests/golden/make_golden.py).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef CN_ORACLE_H
#define CN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define CN_ORACLE_NODATA 255
#define CN_ORACLE_NPLANES 18       /* 2 drainage conditions x 3 hydrologic conditions x 3 ARCs */

/* src/cn.c:13-85 -- one lookup CSV -> table[256][5]; 0 ok, -1 cannot open, -2 empty file */
int cn_oracle_parse_lookup(const char *csv_path, int table[256][5]);

/* src/cn.c:145-147,258-261 -- the nine tables of a lookup directory in the
 * reference's loop order (p,f,g) x (i,ii,iii); 0 ok or the first parse error */
int cn_oracle_load_tables(const char *lookup_dir, int tables[9][256][5]);

/* src/raster.c:126-162 -- bbox -> pixel window and clipped geotransform.
 * 0 ok, 1 = "invalid raster bounds" (raster.c:142-147) */
int cn_oracle_window(int raster_w, int raster_h, const double t[6], const double bbox[4],
                     int *xoff, int *yoff, int *xcount, int *ycount, double gt[6]);

/* src/cn.c:219-229 -- the separable pixel -> HSG cell index maps */
void cn_oracle_col_index(int w, const double gt[6], const double soil_gt[6], int hsx, int32_t *ci);
void cn_oracle_row_index(int h, const double gt[6], const double soil_gt[6], int hsy, int32_t *cj);

/* src/cn.c:218-232 -- nearest-neighbour upsample of rows [y0,y1) */
void cn_oracle_resample_rows(const uint8_t *coarse, int hsx, int hsy, const double soil_gt[6],
                             int w, int h, const double gt[6], int y0, int y1, uint8_t *dst);

/* src/cn.c:88-111 -- dual-HSG remap in place; drained != 0 selects the "drained" branch */
void cn_oracle_remap_hsg(uint8_t *hsg, size_t n, int drained);

/* src/cn.c:289 + 114-131 -- prefill 255 then table pass */
void cn_oracle_apply_table(const uint8_t *esa, const uint8_t *hsg, size_t n,
                           const int table[256][5], uint8_t *out);

/* src/cn.c:218-290 for rows [y0,y1) of a w x h block: writes the 18 planes in the
 * reference's save order (cond-major: drained p_i..g_iii, undrained p_i..g_iii),
 * plane k at out + k*plane_stride, each (y1-y0)*w bytes, row-major.
 * esa_rows points at row y0 of the land-cover window (so a caller that only holds
 * a few rows of a large block can still check them).  0 ok, -1 allocation failure */
int cn_oracle_block_rows(const uint8_t *esa_rows, int w, int h, const double gt[6],
                         const uint8_t *coarse, int hsx, int hsy, const double soil_gt[6],
                         const int tables[9][256][5], int y0, int y1,
                         uint8_t *out, size_t plane_stride);

#endif
