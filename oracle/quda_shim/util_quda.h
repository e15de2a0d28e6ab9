// util_quda.h — see quda_shim_core.h (oracle/ test infrastructure: QUDA is not installed in this image)
#pragma once
#include "quda_shim_core.h"
