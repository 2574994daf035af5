// Instantiates the min-sum edge-parallel kernels (bp_edge.cuh).
#include "bp_edge.cuh"
namespace bpb {
EdgeKernel pick_edge_ms(bool llr, bool msg_global) { return pick_edge_kernel<kMinimumSum>(llr, msg_global); }
}  // namespace bpb
