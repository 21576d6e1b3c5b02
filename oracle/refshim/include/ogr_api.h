/* oracle/refshim/include/ogr_api.h -- TEST INFRASTRUCTURE ONLY: forwards to the shim gdal.h. */
#include "gdal.h"
