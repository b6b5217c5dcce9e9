// build_sah.cu -- placeholder, replaced by the binned-SAH builder
#include <stdio.h>
#include "rtk_device.h"
namespace rtk {
int build_sah_tree(const PrimRef*, uint32_t*, uint32_t, Node2*, const float*, cudaStream_t, char* errmsg) {
  snprintf(errmsg, 256, "SAH builder not built");
  return -1;
}
}
