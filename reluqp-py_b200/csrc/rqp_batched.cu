// Batched solve (placeholder until the batched kernels land): reports UNSUPPORTED loudly.
#include "rqp_common.cuh"
#include "rqp_host.h"

namespace rqp {
int batch_workspace_size(const rqp_problem*, const rqp_settings*, int32_t, const rqp_caps&, size_t*) {
    return RQP_ERR_UNSUPPORTED;
}
int launch_batched(const rqp_problem*, const rqp_settings*, rqp_batch*, void*, size_t, int32_t*, const rqp_caps&,
                   cudaStream_t) {
    return RQP_ERR_UNSUPPORTED;
}
}  // namespace rqp
