/* Stand-in for Intel IPP's <ippi.h>; everything lives in ipp.h. */
#include "ipp.h"
