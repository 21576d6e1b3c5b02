/* oracle/refshim/include/cpl_conv.h -- TEST INFRASTRUCTURE ONLY: forwards to the shim gdal.h. */
#include "gdal.h"
