// placeholder until the tcgen05 kernel lands (next commit): reports "unsupported" for every shape.
#include "common.cuh"
namespace gnnb200 {
int gemm_tf32_supported(const float*, int64_t, int, const float*, int64_t, int, const float*, int64_t, int64_t,
                        int64_t, int64_t) { return 0; }
int gemm_tf32(const float*, int64_t, int, const float*, int64_t, int, float*, int64_t, int64_t, int64_t, int64_t,
              const float*, int, void*, size_t*, cudaStream_t) { return GNNB200_EUNSUPPORTED; }
}  // namespace gnnb200
