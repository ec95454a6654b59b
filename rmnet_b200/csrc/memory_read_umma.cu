// placeholder until the tcgen05 kernel lands (keeps the library linkable)
#include "common.cuh"
namespace rmnet {
bool umma_supported(int) { return false; }
int launch_memory_read_umma(const BankView &, const float *, long long, const int *, int, int, int, int, int, int,
                            const ReadWorkspace &, cudaStream_t) {
  set_error("tcgen05 memory-read kernel not built");
  return RMNET_E_UNSUPPORTED;
}
}  // namespace rmnet
