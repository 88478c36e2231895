#include "step3_impl.cuh"
namespace phet {
int run_step3_f32(phet_ctx *c, const Step3Args &a) { return run_step3_T<float>(c, a); }
}  // namespace phet
