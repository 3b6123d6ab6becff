// Host-side declarations shared by the translation units of librqp.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/rqp.h"

namespace rqp {

struct SinglePlan {
    int grid, block, cpt, rpc, rows_smem, rmode, ring;
    int cpl;         // row-per-warp mode: 16-byte pieces of its row a lane keeps (template value; 0 = off)
    int cl, cl_cps;  // cluster (2-D) mode: cluster size (0 = off) and vector columns of v per CTA of a cluster
    int check_tpw;   // > 0: the residual-check matrix rows of each warp stay in shared memory (tasks per warp)
    size_t smem_bytes;
    size_t vcells_bytes, pcells_bytes, ws_bytes;
};

// rqp_single.cu
int plan_single(const rqp_problem* prob, const rqp_settings* stng, const rqp_caps& caps, SinglePlan* plan);
int launch_single(const rqp_problem* prob, const rqp_settings* stng, rqp_state* state, rqp_result* result_dev,
                  double* trace_dev, int32_t trace_cap, void* ws, size_t ws_bytes, const rqp_caps& caps,
                  cudaStream_t stream);
// rqp_struct.cu
int struct_workspace_size(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng,
                          const rqp_caps& caps, size_t* bytes);
int launch_struct(const rqp_problem* prob, const rqp_structured* sp, const rqp_settings* stng, rqp_state* state,
                  rqp_result* result_dev, double* trace_dev, int32_t trace_cap, void* ws, size_t ws_bytes,
                  const rqp_caps& caps, cudaStream_t stream);
// rqp_batched.cu
int batch_workspace_size(const rqp_problem* prob, const rqp_settings* stng, int32_t B, const rqp_caps& caps,
                         size_t* bytes);
int launch_batched(const rqp_problem* prob, const rqp_settings* stng, rqp_batch* batch, void* ws, size_t ws_bytes,
                   int32_t* sweeps_host, const rqp_caps& caps, cudaStream_t stream);

}  // namespace rqp
