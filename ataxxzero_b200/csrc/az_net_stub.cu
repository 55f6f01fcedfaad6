#include "az_common.h"
void az_net_release(az_context *) {}
