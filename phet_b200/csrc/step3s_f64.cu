#include "step3s_impl.cuh"
namespace phet {
template int run_step3_scatter<double>(phet_ctx *c, const Step3Args &a);
bool step3_scatter_applies_f64(const phet_ctx *c, const Step3Args &a) { return step3_scatter_applies<double>(c, a); }
}  // namespace phet
