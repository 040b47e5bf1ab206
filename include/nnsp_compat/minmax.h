/* forwarding header: lets unmodified ns-nnsp callers and def_nn*.c tables build against nnsp-b200 */
#include "nnsp_legacy_api.h"
