// flo_encode_nt256.cu -- the frame-encode kernel built for 256 threads per CTA, 2 CTA(s) per SM.
#include <cstdio>
#include <type_traits>

#include "flo_internal.h"

#define FLO_VARIANT_NT 256
#ifndef FLO_VARIANT_CTAS
#define FLO_VARIANT_CTAS 2
#endif
// the instantiation without LPC (levels 0-3) fits 85 registers: 3 CTAs per SM (level 0: 1.57 -> 1.42 ms per hour of CD audio)
#ifndef FLO_VARIANT_CTAS_FIXED
#define FLO_VARIANT_CTAS_FIXED 3
#endif

namespace flo {
namespace nt256 {

typedef unsigned long long u64;
typedef long long i64;
typedef uint32_t u32;
typedef int32_t i32;

#include "encode_v3_body.cuh"

}  // namespace nt256

extern const EncodeVariant g_variant_nt256 = {256, FLO_VARIANT_CTAS, FLO_VARIANT_CTAS_FIXED, nt256::encode_static_smem, nt256::variant_configure,
                                              nt256::variant_launch, nt256::variant_occupancy};

}  // namespace flo
