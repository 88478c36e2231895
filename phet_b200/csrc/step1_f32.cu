#include "step1h_impl.cuh"
namespace phet {
int run_step1_f32(phet_ctx *c, const Step1Args &a) { return run_step1_any<float>(c, a); }
}  // namespace phet
