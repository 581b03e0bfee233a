// movers.cuh -- device helpers shared by the rows->columns kernels: W-byte element movers (row image -> registers ->
// column), the 32x32 bit-matrix transpose of the validity bits, and 32-bit shared-space load/store wrappers.
#pragma once
#include "common.cuh"

namespace srj {

// ---- element movers: load W bytes from the row image into registers, store them to the column -------
template <int W>
struct Reg {
  uint32_t v[(W + 3) / 4];
};

template <int W, bool SAFE>
__device__ __forceinline__ Reg<W> ld_elem(const uint8_t* src)
{
  Reg<W> r;
  if constexpr (SAFE) {
#pragma unroll
    for (int i = 0; i < (W + 3) / 4; ++i) r.v[i] = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) r.v[i / 4] |= static_cast<uint32_t>(src[i]) << (8 * (i & 3));
  } else {
    if constexpr (W == 1) {
      r.v[0] = *src;
    } else if constexpr (W == 2) {
      r.v[0] = *reinterpret_cast<const uint16_t*>(src);
    } else if constexpr (W == 4) {
      r.v[0] = *reinterpret_cast<const uint32_t*>(src);
    } else if constexpr (W == 8) {
      const uint2 a = *reinterpret_cast<const uint2*>(src);
      r.v[0]        = a.x;
      r.v[1]        = a.y;
    } else {
      // rows are only 8-byte aligned (JCUDF_ROW_ALIGNMENT, RC:63): two 8-byte reads
      const uint2 a = *reinterpret_cast<const uint2*>(src);
      const uint2 b = *reinterpret_cast<const uint2*>(src + 8);
      r.v[0]        = a.x;
      r.v[1]        = a.y;
      r.v[2]        = b.x;
      r.v[3]        = b.y;
    }
  }
  return r;
}

// explicit st.global: the column pointers come out of a shared-memory table, so the compiler cannot
// infer the address space on its own (it would emit generic ST)
template <int W>
__device__ __forceinline__ void st_elem(uint8_t* dst, const Reg<W>& r)
{
  if constexpr (W == 1) {
    asm volatile("st.global.u8 [%0], %1;" ::"l"(dst), "r"(r.v[0]));
  } else if constexpr (W == 2) {
    asm volatile("st.global.u16 [%0], %1;" ::"l"(dst), "h"(static_cast<uint16_t>(r.v[0])));
  } else if constexpr (W == 4) {
    asm volatile("st.global.u32 [%0], %1;" ::"l"(dst), "r"(r.v[0]));
  } else if constexpr (W == 8) {
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(r.v[0]), "r"(r.v[1]));
  } else {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
                 "r"(r.v[3]));
  }
}

// ---- validity: 32x32 bit-matrix transpose across the warp ----------------------------------------------
// lane = row holds 32 validity bits (32 columns) of its row; five shuffle/xor butterfly steps leave
// lane = column holding the 32-row mask word of that column (RC:1062-1071 builds the same word with
// one __ballot_sync per column; this does 32 columns in ~30 instructions).
__device__ __forceinline__ uint32_t transpose32(uint32_t r, int lane)
{
  uint32_t m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint32_t other = __shfl_xor_sync(0xffffffffu, r, j);
    if ((lane & j) == 0) {
      const uint32_t t = ((r >> j) ^ other) & m;
      r ^= t << j;
    } else {
      const uint32_t t = ((other >> j) ^ r) & m;
      r ^= t;
    }
    m ^= m << (j >> 1);
  }
  return r;
}


// ---- explicit shared-space accesses through 32-bit addresses (the generic forms compile to LD/ST + 64-bit math) ---
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// W bytes of a row image at shared address `a` (8-byte aligned rows: 16-byte fields are read as two 8-byte halves)
template <int W>
__device__ __forceinline__ Reg<W> lds_elem(uint32_t a)
{
  Reg<W> r;
  if constexpr (W == 1) {
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r.v[0]) : "r"(a));
  } else if constexpr (W == 2) {
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r.v[0]) : "r"(a));
  } else if constexpr (W == 4) {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.v[0]) : "r"(a));
  } else if constexpr (W == 8) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "r"(a));
  } else {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1]) : "r"(a));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a + 8));
  }
  return r;
}

}  // namespace srj
