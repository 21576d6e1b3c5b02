/* oracle/refshim/include/gdal_alg.h -- TEST INFRASTRUCTURE ONLY: forwards to the shim gdal.h. */
#include "gdal.h"
