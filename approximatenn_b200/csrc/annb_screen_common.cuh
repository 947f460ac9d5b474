// annb_screen_common.cuh — pieces shared by the fp16 screens of S3 (annb_leaf_screen.cuh) and S5
// (annb_supercharge_screen.cuh): the operand layout of a screened row and the tensor-core product.
#pragma once

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// D/4 coordinates of an fp16 row as NR = D/8 registers of two: lane t of a quad takes the 16
// bytes at 64 v + 16 t of the row for v = 0 .. NR/4 - 1 (the quad reads whole 32-byte sectors
// with every instruction).  Register pair (2ks, 2ks+1) goes to k-step ks as logical
// k = (2t, 2t+1) and (2t+8, 2t+9): queries and candidates use the same map, so the dot
// product covers every coordinate exactly once.
template <int D>
struct ScreenRow {
  static constexpr int NR = D / 8;
  u32 r[NR];
  __device__ __forceinline__ void load(const unsigned short *__restrict__ row16, int t) {
    if (NR >= 4) {
#pragma unroll
      for (int v = 0; v < NR / 4; v++) {
        uint4 q = *reinterpret_cast<const uint4 *>(row16 + 32 * v + 8 * t);
        r[4 * v] = q.x; r[4 * v + 1] = q.y; r[4 * v + 2] = q.z; r[4 * v + 3] = q.w;
      }
    } else {
      uint2 q = *reinterpret_cast<const uint2 *>(row16 + 4 * t);
      r[0] = q.x; r[1] = q.y;
    }
  }
};

