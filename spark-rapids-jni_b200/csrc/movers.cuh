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

// lane's string = L bytes (<= 32) at SHARED address srcs -> staging byte ds.  Nine aligned 32-bit source words are read
// unconditionally (the stage has slack on both sides), funnel-shifted to the staging alignment and stored as whole
// words where the string covers a whole staging word; the <= 3 bytes of a partial first / last staging word are taken
// from the shifted words (the last one is rebuilt from two more loads: its index is not a compile-time constant).
__device__ __forceinline__ void copy_shared_to_staging(uint32_t srcs, uint32_t ds, int L)
{
  const uint32_t dsh = ds & 3u;
  const uint32_t ssh = srcs & 3u;
  const bool back    = ssh < dsh;                            // the word stream starts one word before the string
  const uint32_t sp  = (srcs - ssh) - (back ? 4u : 0u);      // aligned; staging word k <- stream words k, k + 1
  const uint32_t sh  = ((ssh - dsh) & 3u) * 8u;
  const uint32_t w0s = ds - dsh;
  const int end      = static_cast<int>(dsh) + L;            // one past the last staging byte, relative to word 0
  const int kfull1   = end >> 2;                             // whole words: [dsh ? 1 : 0, kfull1)
  // The kernel is bound by shared-memory wavefronts (lanes hit random banks), not by issue slots: every load is
  // predicated on the lane needing that word, which thins the active lanes of the later words and with them the
  // bank conflicts (average string 13 bytes, longest of a warp ~32).
  const int need = L > 0 ? (end + 3) >> 2 : -1;  // staging words 0 .. need-1 are touched; word k needs stream words k, k+1
  uint32_t w[9];  // end <= 35: whole words are k <= 7, built from stream words 0..8
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w[k] = 0;
    if (k <= need) w[k] = lds_u32(sp + 4 * k);
  }
  const int nt      = (kfull1 > 0 || dsh == 0) ? (end & 3) : 0;
  const uint32_t ta = sp + 4u * static_cast<uint32_t>(kfull1);
  uint32_t t0 = 0, t1 = 0;
  if (nt > 0) {
    t0 = lds_u32(ta);
    t1 = lds_u32(ta + 4);
  }
  const uint32_t y0 = __funnelshift_r(w[0], w[1], sh);
  if (dsh == 0 && kfull1 > 0) sts_u32(w0s, y0);
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const uint32_t y = __funnelshift_r(w[k], w[k + 1], sh);
    if (k < kfull1) sts_u32(w0s + 4 * k, y);
  }
  // partial first word: bytes [dsh, min(4, end)) when dsh > 0
  const int hend = dsh ? tmin(end, 4) : 0;
#pragma unroll
  for (int b = 1; b < 4; ++b)
    if (b >= static_cast<int>(dsh) && b < hend) sts_u8(w0s + b, y0 >> (8 * b));
  // partial last word: bytes [0, end & 3) of word kfull1, unless that is the first word again
  const uint32_t tw = __funnelshift_r(t0, t1, sh);
  const uint32_t tb = w0s + 4u * static_cast<uint32_t>(kfull1);
#pragma unroll
  for (int b = 0; b < 3; ++b)
    if (b < nt) sts_u8(tb + b, tw >> (8 * b));
}

}  // namespace srj
