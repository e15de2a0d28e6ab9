// enum_mugiq.h — forwards to mugiq_api.h (see there for the reference lines each definition follows).
#include "mugiq_api.h"
