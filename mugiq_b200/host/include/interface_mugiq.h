// interface_mugiq.h — the reference's file of this name only forward-declares QUDA's checkGauge
// (/root/reference/include/interface_mugiq.h:9); the loop path's checks live in Displace.
#include "mugiq_api.h"
namespace quda {
bool checkGauge(const QudaGaugeParam *param);  // extents positive and even in x, precision single or double
}
