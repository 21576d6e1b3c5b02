/* oracle/refshim/include/gdal.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Declarations of the GDAL / OGR / CPL entry points that the reference's
 * src/cn.c (OGR bbox fetch, cn.c:155-184) and src/raster.c (raster.c:17-18,
 * 77-101, 119-180, 204-226) call.  The implementations live in
 * oracle/refshim/fake_gdal.c and serve in-memory rasters and an in-memory block
 * table; no real GDAL is present in this image.  Field order of OGREnvelope
 * follows GDAL's public ogr_core.h (MinX, MaxX, MinY, MaxY).
 */
#ifndef REFSHIM_GDAL_H
#define REFSHIM_GDAL_H

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

typedef void *GDALDatasetH;
typedef void *GDALDriverH;
typedef void *GDALRasterBandH;
typedef void *OGRDataSourceH;
typedef void *OGRLayerH;
typedef void *OGRFeatureH;
typedef void *OGRGeometryH;
typedef void *OGRSpatialReferenceH;
typedef void *OGRSFDriverH;

typedef struct { double MinX, MaxX, MinY, MaxY; } OGREnvelope;

typedef enum { CE_None = 0, CE_Debug = 1, CE_Warning = 2, CE_Failure = 3, CE_Fatal = 4 } CPLErr;
typedef enum { GA_ReadOnly = 0, GA_Update = 1 } GDALAccess;
typedef enum { GF_Read = 0, GF_Write = 1 } GDALRWFlag;
typedef enum { GDT_Unknown = 0, GDT_Byte = 1 } GDALDataType;
typedef int OGRErr;

void GDALAllRegister(void);
void OGRRegisterAll(void);

GDALDatasetH GDALOpen(const char *path, GDALAccess access);
void GDALClose(GDALDatasetH ds);
CPLErr GDALGetGeoTransform(GDALDatasetH ds, double *t);
CPLErr GDALSetGeoTransform(GDALDatasetH ds, double *t);
int GDALGetRasterXSize(GDALDatasetH ds);
int GDALGetRasterYSize(GDALDatasetH ds);
const char *GDALGetProjectionRef(GDALDatasetH ds);
CPLErr GDALSetProjection(GDALDatasetH ds, const char *wkt);
GDALRasterBandH GDALGetRasterBand(GDALDatasetH ds, int band);
CPLErr GDALRasterIO(GDALRasterBandH band, GDALRWFlag rw, int xoff, int yoff, int xsize, int ysize,
                    void *buf, int bxsize, int bysize, GDALDataType type, int pixel_space, int line_space);
GDALDriverH GDALGetDriverByName(const char *name);
GDALDatasetH GDALCreate(GDALDriverH drv, const char *path, int xsize, int ysize, int bands,
                        GDALDataType type, char **opts);

char **CSLSetNameValue(char **list, const char *name, const char *value);
void CSLDestroy(char **list);
void CPLFree(void *p);

OGRSpatialReferenceH OSRNewSpatialReference(const char *wkt);
OGRErr OSRExportToWkt(OGRSpatialReferenceH srs, char **wkt);

OGRDataSourceH OGROpen(const char *path, int update, OGRSFDriverH *drv);
void OGR_DS_Destroy(OGRDataSourceH ds);
OGRLayerH OGR_DS_GetLayer(OGRDataSourceH ds, int i);
OGRErr OGR_L_SetAttributeFilter(OGRLayerH layer, const char *filter);
OGRFeatureH OGR_L_GetNextFeature(OGRLayerH layer);
void OGR_L_ResetReading(OGRLayerH layer);
int OGR_L_GetFeatureCount(OGRLayerH layer, int force);
OGRGeometryH OGR_F_GetGeometryRef(OGRFeatureH feat);
void OGR_G_GetEnvelope(OGRGeometryH geom, OGREnvelope *env);
void OGR_F_Destroy(OGRFeatureH feat);
int OGR_F_GetFieldIndex(OGRFeatureH feat, const char *name);
int OGR_F_GetFieldAsInteger(OGRFeatureH feat, int field);

#endif
