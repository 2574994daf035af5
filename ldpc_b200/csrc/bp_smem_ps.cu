// Instantiates the product-sum on-chip kernels (bp_smem.cuh) for every degree bucket.
#include "bp_smem.cuh"
namespace bpb {
SmemKernel pick_smem_ps(int dc, int dv, bool regular, bool llr) {
    return pick_smem_bucket<kProductSum>(dc, dv, regular, llr);
}
}  // namespace bpb
