// Instantiates the product-sum edge-parallel kernels (bp_edge.cuh).
#include "bp_edge.cuh"
namespace bpb {
EdgeKernel pick_edge_ps(bool llr, bool msg_global) { return pick_edge_kernel<kProductSum>(llr, msg_global); }
}  // namespace bpb
