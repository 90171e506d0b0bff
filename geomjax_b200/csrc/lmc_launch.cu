#include "launch.h"
namespace gb {
int GB_LPC_NAME(launch_lmc)(const TransArgs&, const gb200_target_desc&, LayoutChoice, int, cudaStream_t) {
  set_error("lmc: not built yet");
  return GB200_ERR_UNSUPPORTED;
}
}  // namespace gb
