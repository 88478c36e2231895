#include "step3h_impl.cuh"
namespace phet {
int run_step3_f32(phet_ctx *c, const Step3Args &a) { return run_step3_any<float>(c, a); }
}  // namespace phet
