// curve_p192.cu — instantiates every batch kernel for p192 (the primeorder template on a third limb count,
// L = 6; p192/src/arithmetic.rs, p192/src/arithmetic/field.rs, p192/src/ecdsa.rs).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_p192() { return Launch<CurveP192>::table(); }
}  // namespace ecb
