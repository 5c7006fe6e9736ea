// ll_exchange.cuh -- flagged ("LL") words for fence-free exchanges between the CTAs of a cooperative kernel
// (LU panel pivot exchange, QR panel reductions).
#pragma once
#include "la_common.cuh"

namespace la {
// Flagged ("LL") exchange words: every 4-byte half of a value travels next to a 4-byte tag in the same naturally
// atomic 8-byte unit, so a reader that sees the expected tag has the data -- no fence, no atomic, no grid barrier.
// All accesses are RELAXED at gpu scope (served by L2, free to overlap): volatile ones would be kept in program
// order by the hardware, which serialises a poll of n words into n L2 round trips.
__device__ __forceinline__ void st_relaxed_2x64(void* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_2x64(const void* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ll_pack(unsigned v, unsigned tag) {
  return ((unsigned long long)tag << 32) | v;
}
template <typename T>
struct LL;
template <>
struct LL<double> {
  typedef uint4 word;
  static __device__ __forceinline__ void store(word* p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    st_relaxed_2x64(p, ll_pack((unsigned)b, tag), ll_pack((unsigned)(b >> 32), tag));
  }
  static __device__ __forceinline__ bool load(const word* p, unsigned tag, double& v) {
    unsigned long long a, b;
    ld_relaxed_2x64(p, a, b);
    v = __longlong_as_double((long long)((b << 32) | (a & 0xffffffffull)));
    return (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag;
  }
};
template <>
struct LL<float> {
  typedef uint2 word;
  static __device__ __forceinline__ void store(word* p, float v, unsigned tag) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(ll_pack(__float_as_uint(v), tag)) : "memory");
  }
  static __device__ __forceinline__ bool load(const word* p, unsigned tag, float& v) {
    unsigned long long a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    v = __uint_as_float((unsigned)a);
    return (unsigned)(a >> 32) == tag;
  }
};

}  // namespace la
