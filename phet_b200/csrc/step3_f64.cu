#include "step3h_impl.cuh"
namespace phet {
int run_step3_f64(phet_ctx *c, const Step3Args &a) { return run_step3_any<double>(c, a); }
}  // namespace phet
