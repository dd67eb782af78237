// placeholder until the sequence kernels land (same round)
#include "common.cuh"
extern "C" {
size_t slb_seq_step_workspace_bytes(const slb_seq_step_args*) { return 0; }
int slb_seq_train_step(const slb_seq_step_args*, slb_stream_t) { slb_set_error("seq_train_step: not built"); return SLB_EINVAL; }
int slb_seq_representation(const slb_seq_step_args*, float*, slb_stream_t) { slb_set_error("seq_representation: not built"); return SLB_EINVAL; }
}
