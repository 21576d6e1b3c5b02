/* gcn10_b200/host/host_tiff.h -- incremental GeoTIFF writer used by the block pipeline
 * (the one-shot gh_tiff_write and the reader are declared in gcn10_host.h). */
#ifndef GCN10_HOST_TIFF_H
#define GCN10_HOST_TIFF_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gh_tiffw gh_tiffw;

/* Creates path and reserves the header; tiles are appended with gh_tiffw_write_rows. */
int gh_tiffw_open(const char *path, int w, int h, const double gt[6], gh_tiffw **out, char *err, size_t errlen);
/* Georeferencing tags to write instead of the default EPSG:4326 keys (copied; GTRasterTypeGeoKey is forced to
 * PixelIsArea because the tiepoint written is a pixel corner).  Call before gh_tiffw_close. */
#include "gcn10_host.h"
int gh_tiffw_set_geokeys(gh_tiffw *tw, const gh_geokeys *gk);
/* Appends rows [y0, y0+nrows): y0 must continue where the previous call stopped and be a multiple
 * of 256; nrows must be a multiple of 256 except for the last band of the raster. */
int gh_tiffw_write_rows(gh_tiffw *tw, const uint8_t *data, size_t pitch, int y0, int nrows, int threads);
/* Appends tiles that are already compressed (complete zlib streams of 256 x 256 bytes, e.g. from
 * gcn10_cuda_block_deflate): tile row `tile_row` must be the next one not yet written; sizes / offsets
 * address `blob` for the tiles_x tiles of that row. */
int gh_tiffw_put_tile_row(gh_tiffw *tw, int tile_row, const uint8_t *blob, const uint64_t *offsets,
                          const uint32_t *sizes);
/* Several consecutive tile rows at once ([nrows][tiles_x] offsets / sizes).  When the tiles lie in `blob` in table
 * order with at most alignment gaps between them (gcn10_cuda_set_option "ordered") they are written with one write. */
int gh_tiffw_put_tile_rows(gh_tiffw *tw, int tile_row0, int nrows, const uint8_t *blob, const uint64_t *offsets,
                           const uint32_t *sizes);
/* Runs fn(arg, 0..n-1) on up to `threads` threads (the calling thread included). */
void gh_parallel_for(int n, int threads, void (*fn)(void *arg, int index), void *arg);
/* Writes the tile tables, georeferencing and IFD; 0 ok. */
int gh_tiffw_close(gh_tiffw *tw);
void gh_tiffw_abort(gh_tiffw *tw);

#ifdef __cplusplus
}
#endif
#endif
