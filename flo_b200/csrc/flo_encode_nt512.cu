// flo_encode_nt512.cu -- the frame-encode kernel built for 512 threads per CTA, 1 CTA(s) per SM.
#include <cstdio>
#include <type_traits>

#include "flo_internal.h"

#define FLO_VARIANT_NT 512
#define FLO_VARIANT_CTAS 1
#define FLO_VARIANT_CTAS_FIXED FLO_VARIANT_CTAS

namespace flo {
namespace nt512 {

typedef unsigned long long u64;
typedef long long i64;
typedef uint32_t u32;
typedef int32_t i32;

#include "encode_v3_body.cuh"

}  // namespace nt512

extern const EncodeVariant g_variant_nt512 = {512, FLO_VARIANT_CTAS, FLO_VARIANT_CTAS_FIXED, nt512::encode_static_smem, nt512::variant_configure,
                                              nt512::variant_launch, nt512::variant_occupancy};

}  // namespace flo
