// dist.cu — placeholder until the NCCL layer lands (kept so ctx.cu links).
#include "common.cuh"
namespace lorb {
void dist_destroy(lorb_ctx*) {}
}  // namespace lorb
