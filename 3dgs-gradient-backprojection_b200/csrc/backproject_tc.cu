// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
namespace gwbp {
size_t fpack_bytes(int, int, int) { return 0; }
bool tc_supported(int) { return false; }
int launch_backproject_tc(const TileCtx &, const float *, int64_t, int64_t, int64_t, int, float *, float *, void *,
                          long long *, cudaStream_t) {
    set_error("tcgen05 path not built");
    return -1;
}
}  // namespace gwbp
