/* Stand-in for Intel IPP's <ippdefs.h>; everything lives in ipp.h. */
#include "ipp.h"
