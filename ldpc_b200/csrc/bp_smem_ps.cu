// Instantiates the product-sum on-chip kernels (bp_smem.cuh); 512 threads per CTA (more registers per thread).
#include "bp_smem.cuh"
namespace bpb {
SmemKernel pick_smem_ps(int dc, int dv, bool llr) { return pick_smem_bucket<kProductSum>(dc, dv, llr); }
}  // namespace bpb
