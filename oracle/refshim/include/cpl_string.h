/* oracle/refshim/include/cpl_string.h -- TEST INFRASTRUCTURE ONLY: forwards to the shim gdal.h. */
#include "gdal.h"
