#include "step3p_impl.cuh"
namespace phet {
template int run_step3_prune<float>(phet_ctx *c, const Step3Args &a);
}  // namespace phet
