// Instantiates the min-sum on-chip kernels (bp_smem.cuh) for every degree bucket.
#include "bp_smem.cuh"
namespace bpb {
SmemKernel pick_smem_ms(int dc, int dv, bool regular, bool llr) {
    return pick_smem_bucket<kMinimumSum>(dc, dv, regular, llr);
}
}  // namespace bpb
