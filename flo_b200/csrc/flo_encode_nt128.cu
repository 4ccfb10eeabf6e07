// flo_encode_nt128.cu -- the frame-encode kernel built for 128 threads per CTA, 4 CTA(s) per SM.
#include <cstdio>
#include <type_traits>

#include "flo_internal.h"

#define FLO_VARIANT_NT 128
#define FLO_VARIANT_CTAS 4
#define FLO_VARIANT_CTAS_FIXED FLO_VARIANT_CTAS

namespace flo {
namespace nt128 {

typedef unsigned long long u64;
typedef long long i64;
typedef uint32_t u32;
typedef int32_t i32;

#include "encode_v3_body.cuh"

}  // namespace nt128

extern const EncodeVariant g_variant_nt128 = {128, FLO_VARIANT_CTAS, FLO_VARIANT_CTAS_FIXED, nt128::encode_static_smem, nt128::variant_configure,
                                              nt128::variant_launch, nt128::variant_occupancy};

}  // namespace flo
