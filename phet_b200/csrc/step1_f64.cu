#include "step1h_impl.cuh"
namespace phet {
int run_step1_f64(phet_ctx *c, const Step1Args &a) { return run_step1_any<double>(c, a); }
}  // namespace phet
