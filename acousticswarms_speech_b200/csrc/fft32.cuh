// fft32.cuh -- 32-point register-resident DFT used by the warp-per-FFT kernels (stft_cc_warp.cu, gcc.cu).
// A 1024-point complex FFT is two of these around one shared-memory transpose: lane t holds z[32 q + t],
// DFT over q, twiddle by W_1024^{t k1}, transpose, DFT over t; lane k1 then owns Z[k1 + 32 k2].
#pragma once
#include "common.cuh"

namespace asw {

__device__ __forceinline__ constexpr int bitrev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// cos(2 pi j / 32), sin(2 pi j / 32), j = 0..15
__device__ constexpr float kC32[16] = {1.0f,           0.98078528040f, 0.92387953251f, 0.83146961230f,
                                       0.70710678119f, 0.55557023302f, 0.38268343237f, 0.19509032202f,
                                       0.0f,           -0.19509032202f, -0.38268343237f, -0.55557023302f,
                                       -0.70710678119f, -0.83146961230f, -0.92387953251f, -0.98078528040f};
__device__ constexpr float kS32[16] = {0.0f,           0.19509032202f, 0.38268343237f, 0.55557023302f,
                                       0.70710678119f, 0.83146961230f, 0.92387953251f, 0.98078528040f,
                                       1.0f,           0.98078528040f, 0.92387953251f, 0.83146961230f,
                                       0.70710678119f, 0.55557023302f, 0.38268343237f, 0.19509032202f};

// Forward DFT of 32 register-resident points, radix-2 decimation in frequency, fully unrolled.
// Output X[k] is left at v[bitrev5(k)].
__device__ __forceinline__ void dft32(float2 (&v)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1;
        const int tstep = 32 / len;  // W_len^j = W_32^{j * tstep}
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 a = v[blk + j], b = v[blk + j + half];
                v[blk + j] = make_float2(a.x + b.x, a.y + b.y);
                const float dx = a.x - b.x, dy = a.y - b.y;
                const int tw = j * tstep;  // compile-time after unrolling
                if (tw == 0) {
                    v[blk + j + half] = make_float2(dx, dy);
                } else if (tw == 8) {  // multiply by -i
                    v[blk + j + half] = make_float2(dy, -dx);
                } else {
                    const float c = kC32[tw], s = kS32[tw];  // exp(-i theta) = c - i s
                    v[blk + j + half] = make_float2(fmaf(dx, c, dy * s), fmaf(dy, c, -dx * s));
                }
            }
        }
    }
}

// ---- packed fp32x2 arithmetic (sm_100a) ----------------------------------------------------------------
// Blackwell issues FADD2 / FMUL2 / FFMA2 on a 64-bit register pair: one instruction, two IEEE fp32 results.
// ptxas folds half swaps, per-half negation and scalar broadcast of an operand into the instruction
// (`R10.F32x2.LO_HI.NP`, `R5.F32`), so a complex butterfly with a twiddle is 4 instructions instead of 8 and a
// complex multiply 2 instead of 4.  A complex number is one 64-bit value (re = low half).  Every routine below
// rounds exactly like its scalar twin above/below (same products, same fused adds), so results are bit-identical.
typedef unsigned long long c64;

__device__ __forceinline__ c64 pk(float lo, float hi) {
    c64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk(c64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ c64 add2(c64 a, c64 b) {
    c64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ c64 sub2(c64 a, c64 b) {
    c64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ c64 mul2(c64 a, c64 b) {
    c64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ c64 fma2(c64 a, c64 b, c64 c) {
    c64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// a * b, rounding as cmul(): (fma(a.x, b.x, -(a.y b.y)), fma(a.x, b.y, a.y b.x))
__device__ __forceinline__ c64 cmul2(c64 a, c64 b) {
    const float2 af = upk(a), bf = upk(b);
    return fma2(pk(af.x, af.x), b, mul2(pk(af.y, af.y), pk(-bf.y, bf.x)));
}

// dft32 on packed complex values; bit-identical to the float2 version.
__device__ __forceinline__ void dft32(c64 (&v)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1;
        const int tstep = 32 / len;
#pragma unroll
        for (int blk = 0; blk < 32; blk += len) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const c64 a = v[blk + j], b = v[blk + j + half];
                v[blk + j] = add2(a, b);
                const c64 d = sub2(a, b);
                const int tw = j * tstep;
                if (tw == 0) {
                    v[blk + j + half] = d;
                } else {
                    const float2 df = upk(d);
                    const c64 rot = pk(df.y, -df.x);              // -i d
                    if (tw == 8) {
                        v[blk + j + half] = rot;
                    } else {
                        const float c = kC32[tw], s = kS32[tw];   // (dx c + dy s, dy c - dx s)
                        v[blk + j + half] = fma2(d, pk(c, c), mul2(rot, pk(s, s)));
                    }
                }
            }
        }
    }
}

// Pass-1 epilogue shared by both users: twiddle Y[k1] (held at v[bitrev5(k1)]) by W_1024^{lane * k1} (seeded
// exactly every 8 steps from the table) and store transposed into the per-warp tile, a c64[32][33] array:
// 64-bit accesses are served per half-warp, the row stride of 33 * 8 bytes puts the 16 lanes of a half-warp on 16
// distinct bank pairs in both directions, so the write (lane-contiguous) and the read (lane-strided) are conflict-free.
__device__ __forceinline__ void twiddle_and_transpose(const c64 (&v)[32], float* tile_f, int lane, float2 w1f, float2 w8,
                                                      float2 w16, float2 w24) {
    c64* tile = reinterpret_cast<c64*>(tile_f);
    c64 tw = pk(1.f, 0.f);
    const c64 w1 = pk(w1f.x, w1f.y);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
        if (k1 == 8) tw = pk(w8.x, w8.y);
        if (k1 == 16) tw = pk(w16.x, w16.y);
        if (k1 == 24) tw = pk(w24.x, w24.y);
        tile[k1 * 33 + lane] = (k1 == 0) ? v[0] : cmul2(v[bitrev5(k1)], tw);
        if ((k1 & 7) != 7) tw = cmul2(tw, w1);
    }
}

__device__ __forceinline__ void load_transposed(c64 (&v)[32], const float* tile_f, int lane) {
    const c64* tile = reinterpret_cast<const c64*>(tile_f);
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = tile[lane * 33 + t];
}

}  // namespace asw
