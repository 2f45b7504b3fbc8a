// curve_p224.cu — instantiates every batch kernel for p224 (the primeorder template on an odd limb count,
// L = 7: column-wise multiplier, row-wise Montgomery reduction, Tonelli-Shanks square root; p224/src/arithmetic.rs,
// p224/src/arithmetic/field.rs, p224/src/ecdsa.rs).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_p224() { return Launch<CurveP224>::table(); }
}  // namespace ecb
