#include "step3s_impl.cuh"
namespace phet {
template int run_step3_scatter<float>(phet_ctx *c, const Step3Args &a);
bool step3_scatter_applies_f32(const phet_ctx *c, const Step3Args &a) { return step3_scatter_applies<float>(c, a); }
}  // namespace phet
