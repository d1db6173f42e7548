// lsq_device.cuh -- sm_100a device code for the LSQ+ fake-quantize hot path.
//
// One kernel family serves per-tensor, per-channel and multi-tensor launches: a tensor is a
// contiguous (outer, C, inner) box; a CTA owns one TILE = (channel c, split j) = a contiguous
// slice of channel c's element space, so scale / shift / zero-point are CTA-uniform registers
// and the backward's grad_scale / grad_shift partial sums stay in registers until one
// warp-shuffle -> shared-memory -> (last-arriver, fixed-order) reduction per tile.
//
// Arithmetic contract (file:line under /root/reference/torchlsq/csrc/ops/):
//   kernels/lsq_kernel.h:12-13    zp, forward value            -> fq_forward()
//   kernels/lsq_kernel.h:33-36    clamped un-rounded xq, mask   -> fq_backward()
//   kernels/lsq_kernel.h:56-61    xfq, learned-init grad, dS
//   kernels/lsq_kernel.h:85-86    dB
//   kernels/lsq_kernel.h:157-159  per-channel eps clamp + 1/s   -> make_chan()
//   cuda/lsq_cuda.cu:52-55        per-tensor eps clamp + 1/s (host side in the reference)
//   global_scope.h:39,51-52       nearbyint, fminf/fmaxf
// nvcc fuses x*inv_s+zp and (r-zp)*s-x into FFMAs in the reference CUDA build (checked in its
// sm_100a SASS); this file spells every operation with explicit-rounding intrinsics so that
// exactly those two are fused and nothing else is.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace lsqb200 {

enum : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2, DT_F64 = 3 };   // DT_F64: lsq_f64.cuh
// MODE: how the affine value v = x/s + zp is formed
enum : int { M_FP32 = 0,        // fp32 internal math (fp32 tensors; fp16/bf16 tensors up-cast)
             M_HALF_EXACT = 1,  // fp16 x with fp16 params: round to half after every operator (c10::Half)
             M_F64 = 2,         // float64 x with float64 params (lsq_f64.cuh)
             M_FP32_RELU = 3,      // M_FP32 with the ReLU prologue fused in: fake-quant of relu(x) in one pass
             M_FP32_ADD_RELU = 4,  // ... of relu(x + x2): the residual join of a ResNet block
             M_FP32_ADD = 5 };     // ... of x + x2
// Prologue fusion (not in the reference: there `relu` is a separate ATen pass that writes relu(x) to HBM and the
// fake-quant reads it back).  lsq(relu(x)): forward x' = relu(x) feeds fq_forward; backward the upstream gradient
// goes through the fake-quant's mask on x' and then through relu's own backward (threshold_backward: x <= 0 -> exact 0).
// With an ADD prologue the fake-quant's input is the residual sum x + x2 ROUNDED TO THE TENSOR TYPE (what ATen's add
// writes and the next op reads back), so results stay bit-identical to the three-pass sequence; the single grad_x the
// backward writes is the gradient of both addends.
__host__ __device__ constexpr bool mode_relu(int m) { return m == M_FP32_RELU || m == M_FP32_ADD_RELU; }
__host__ __device__ constexpr bool mode_add(int m) { return m == M_FP32_ADD_RELU || m == M_FP32_ADD; }
enum : int { B_NORMAL = 0, B_INIT = 1, B_EVAL = 2, B_EVAL_INIT = 3 };
__host__ __device__ constexpr bool bmode_passthrough(int b) { return b == B_INIT || b == B_EVAL_INIT; }
__host__ __device__ constexpr bool bmode_reduces(int b) { return b == B_NORMAL || b == B_INIT; }

// ---------------------------------------------------------------------------------------------
// Segment descriptor (one fake-quant site).  Passed by value for single launches, read from a
// device table for multi-tensor plans.
// ---------------------------------------------------------------------------------------------
struct Seg {
    const void* x;
    const void* x2;        // second addend (ADD prologues), same layout as x; else unused
    void* y;
    const void* g;
    void* gx;
    const void* scale;
    const void* shift;
    void* gscale;
    void* gshift;
    double* partials;      // [tiles_of_segment * 2] when splits > 1
    unsigned* counters;    // [C] when splits > 1 (self-resetting tickets)
    float* stats_out;      // weight-init: float[C]
    long long C;
    long long vpr;         // units per row (regime 1); huge for regime 0
    long long row_stride;  // C * vpr, in units (regime 1)
    long long chan_stride; // units between channel starts: vpr (regime 1); unused (regime 0)
    long long inner;       // elements per row
    long long chan_units;  // units in one channel's space (body only for regime 0)
    long long units_per_split;
    long long tile_begin;  // first global tile id of this segment (plans)
    long long chan_elems;  // outer * inner (elements per channel), for stats
    double gs;             // grad scale factor (already includes grad_scaler)
    float qmin, qmax, tmin, tmax;
    float stats_denom;     // 2^bitness
    int splits;
    int regime;            // 0: channel space contiguous (outer == 1), 1: strided rows
    int per_channel;
    int pdt;
    int sym;
    int vec;               // elements per unit actually used
    int interleave;        // 1: splits of a channel are interleaved at group granularity
    int group;             // threads per tile-owning group (32 or THREADS)
    // observer step (lsq_observe_kernel): x -> running min/max (gscale/gshift slots hold float* state) -> scale/shift
    float obs_c, obs_eps;  // averaging constant, eps
    int obs_flags;         // bit 0: symmetric qscheme, bit 1: moving average (else running extrema)
    int obs_zp_sym;        // zero point of the symmetric scheme (0, 128 or (qmin+qmax)//2)
    // integer export (lsq_export.cuh): y holds uint8 / int8 codes
    int code_signed;       // 1: int8 codes (qint8), 0: uint8 (quint8)
    int flags;             // L2-prefetch depth: this many of a thread's first units are prefetched before the dependency wait (Tuning::l2_prefetch, 0 = off)
};

// ---------------------------------------------------------------------------------------------
// memory access policies
// ---------------------------------------------------------------------------------------------
enum : int { LD_DEFAULT = 0, LD_NC_NOALLOC = 1, LD_EVICT_FIRST = 2 };
enum : int { ST_DEFAULT = 0, ST_CS = 1, ST_NOALLOC = 2 };

// A unit is NW 32-bit words moved by ONE instruction: NW = 8 -> LDG.E.256 / STG.E.256 (new on
// sm_100), 4 -> 128-bit, 2 -> 64-bit.  NW = 0 denotes the scalar (element-at-a-time) path.
template <int NW> struct Raw { uint32_t w[NW == 0 ? 1 : NW]; };

template <int LD, int NW>
__device__ __forceinline__ Raw<NW> ld_unit(const void* p) {
    Raw<NW> r;
    if (NW == 8) {
        if (LD == LD_EVICT_FIRST)
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4 % NW]), "=r"(r.w[5 % NW]), "=r"(r.w[6 % NW]), "=r"(r.w[7 % NW]) : "l"(p));
        else if (LD == LD_NC_NOALLOC)
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4 % NW]), "=r"(r.w[5 % NW]), "=r"(r.w[6 % NW]), "=r"(r.w[7 % NW]) : "l"(p));
        else
            asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4 % NW]), "=r"(r.w[5 % NW]), "=r"(r.w[6 % NW]), "=r"(r.w[7 % NW]) : "l"(p));
    } else if (NW == 4) {
        if (LD == LD_DEFAULT)
            asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1 % NW]), "=r"(r.w[2 % NW]), "=r"(r.w[3 % NW]) : "l"(p));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1 % NW]), "=r"(r.w[2 % NW]), "=r"(r.w[3 % NW]) : "l"(p));
    } else {
        if (LD == LD_DEFAULT)
            asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1 % NW]) : "l"(p));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1 % NW]) : "l"(p));
    }
    return r;
}
template <int ST, int NW>
__device__ __forceinline__ void st_unit(void* p, const Raw<NW>& v) {
    if (NW == 8) {
        if (ST == ST_CS)
            asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4 % NW]), "r"(v.w[5 % NW]), "r"(v.w[6 % NW]), "r"(v.w[7 % NW]) : "memory");
        else if (ST == ST_NOALLOC)
            asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4 % NW]), "r"(v.w[5 % NW]), "r"(v.w[6 % NW]), "r"(v.w[7 % NW]) : "memory");
        else
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4 % NW]), "r"(v.w[5 % NW]), "r"(v.w[6 % NW]), "r"(v.w[7 % NW]) : "memory");
    } else if (NW == 4) {
        if (ST == ST_CS)
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1 % NW]), "r"(v.w[2 % NW]), "r"(v.w[3 % NW]) : "memory");
        else if (ST == ST_NOALLOC)
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1 % NW]), "r"(v.w[2 % NW]), "r"(v.w[3 % NW]) : "memory");
        else
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1 % NW]), "r"(v.w[2 % NW]), "r"(v.w[3 % NW]) : "memory");
    } else {
        asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1 % NW]) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// element <-> float conversion; T is the storage type (float, __half, __nv_bfloat16)
// ---------------------------------------------------------------------------------------------
template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
    static constexpr int PER_WORD = 1;
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
    static __device__ __forceinline__ void unpack_word(uint32_t w, float* f) { f[0] = __uint_as_float(w); }
    static __device__ __forceinline__ uint32_t pack_word(const float* f) { return __float_as_uint(f[0]); }
};
template <> struct ElemTraits<__half> {
    static constexpr int PER_WORD = 2;
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
    static __device__ __forceinline__ void unpack_word(uint32_t w, float* f) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
        f[0] = t.x; f[1] = t.y;
    }
    static __device__ __forceinline__ uint32_t pack_word(const float* f) {
        const __half2 h = __floats2half2_rn(f[0], f[1]);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
};
template <> struct ElemTraits<__nv_bfloat16> {
    static constexpr int PER_WORD = 2;
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void unpack_word(uint32_t w, float* f) {
        f[0] = __uint_as_float(w << 16); f[1] = __uint_as_float(w & 0xffff0000u);
    }
    static __device__ __forceinline__ uint32_t pack_word(const float* f) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
};
// elements per unit
template <typename T, int NW> struct UnitOf { static constexpr int VEC = NW == 0 ? 1 : NW * ElemTraits<T>::PER_WORD; };
template <typename T, int NW>
__device__ __forceinline__ void unpack_unit(const Raw<NW>& r, float* f) {
#pragma unroll
    for (int i = 0; i < NW; i++) ElemTraits<T>::unpack_word(r.w[i], f + i * ElemTraits<T>::PER_WORD);
}
template <typename T, int NW>
__device__ __forceinline__ Raw<NW> pack_unit(const float* f) {
    Raw<NW> r;
#pragma unroll
    for (int i = 0; i < NW; i++) r.w[i] = ElemTraits<T>::pack_word(f + i * ElemTraits<T>::PER_WORD);
    return r;
}

// residual sum as ATen's add kernel forms it: fp32 add, one rounding to the tensor type
template <typename T>
__device__ __forceinline__ float pre_add(float a, float b) {
    return ElemTraits<T>::to_f(ElemTraits<T>::from_f(__fadd_rn(a, b)));
}

__device__ __forceinline__ float load_param(const void* p, long long i, int pdt) {
    if (pdt == DT_F32) return reinterpret_cast<const float*>(p)[i];
    if (pdt == DT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void store_param(void* p, long long i, int pdt, double v) {
    if (pdt == DT_F32) reinterpret_cast<float*>(p)[i] = (float)v;
    else if (pdt == DT_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn((float)v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn((float)v);
}

// ---------------------------------------------------------------------------------------------
// scalar math
// ---------------------------------------------------------------------------------------------
struct Chan {
    float s, inv_s, zp;
    float c_lo, c_hi;   // qmin - zp, qmax - zp (border multipliers of dS)
    float qmin, qmax;
};

__device__ __forceinline__ float hround(float f) { return __half2float(__float2half_rn(f)); }

// torch.relu == clamp_min(x, 0): NaN stays NaN, -0 and every negative become +0 (max.NaN orders -0 < +0)
template <int MODE>
__device__ __forceinline__ float pre_op(float x) {
    if (mode_relu(MODE)) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(0.0f)); return r; }
    return x;
}

template <int MODE>
__device__ __forceinline__ Chan make_chan(float scale, float shift, const Seg& sg) {
    Chan c;
    const float a = fabsf(scale);
    if (MODE == M_HALF_EXACT) {
        const float eps = 0.0009765625f;                           // numeric_limits<c10::Half>::epsilon
        c.s = sg.per_channel ? fmaxf(eps, a) : ((a < eps) ? eps : a);
        c.inv_s = hround(__fdiv_rn(1.0f, c.s));
        const float t = hround(__fmul_rn(-shift, c.inv_s));
        c.zp = hround(rintf(fminf(sg.tmax, fmaxf(sg.tmin, t))));
    } else {
        const float eps = 1.1920928955078125e-07f;                 // numeric_limits<float>::epsilon
        // per-tensor: std::max(|s|, eps) (NaN stays NaN); per-channel: fmaxf(eps, |s|)
        c.s = sg.per_channel ? fmaxf(eps, a) : ((a < eps) ? eps : a);
        c.inv_s = __fdiv_rn(1.0f, c.s);
        const float t = __fmul_rn(-shift, c.inv_s);
        c.zp = rintf(fminf(sg.tmax, fmaxf(sg.tmin, t)));
    }
    c.qmin = sg.qmin; c.qmax = sg.qmax;
    c.c_lo = __fsub_rn(sg.qmin, c.zp);
    c.c_hi = __fsub_rn(sg.qmax, c.zp);
    return c;
}

// The raw scale / shift loads are issued at kernel entry but the derived constants are formed
// only after the first data loads of the tile are in flight, so a tile's two dependent DRAM
// round trips (parameters, then data) overlap instead of adding up.
template <int MODE>
struct LazyChan {
    float sraw, braw;
    Chan c;
    bool ready;
    __device__ __forceinline__ void issue(const Seg& sg, long long pidx) {
        sraw = load_param(sg.scale, pidx, sg.pdt);
        braw = load_param(sg.shift, pidx, sg.pdt);
        ready = false;
    }
    __device__ __forceinline__ const Chan& get(const Seg& sg) {
        if (!ready) { c = make_chan<MODE>(sraw, braw, sg); ready = true; }
        return c;
    }
};

template <int MODE>
__device__ __forceinline__ float affine_v(float x, const Chan& c) {
    if (MODE == M_HALF_EXACT) return hround(__fadd_rn(hround(__fmul_rn(x, c.inv_s)), c.zp));
    return __fmaf_rn(x, c.inv_s, c.zp);
}

template <int MODE>
__device__ __forceinline__ float fq_forward(float x_in, const Chan& c) {
    const float x = pre_op<MODE>(x_in);
    const float v = affine_v<MODE>(x, c);
    const float r = rintf(fminf(c.qmax, fmaxf(c.qmin, v)));       // max first: NaN -> qmin
    return __fmul_rn(__fsub_rn(r, c.zp), c.s);
}

// returns dX (as float); accumulates the two parameter-gradient terms (not yet multiplied by gs).
//
// Same values as lsq_kernel.h:33-61,85-86 with fewer instructions (the bf16 / fp16 backward is
// issue-bound, not memory-bound, at 6 bytes per element):
//   * mask = (qmin < xq) && (xq < qmax) is tested on the un-clamped v: identical for
//     qmin < qmax (checked on the host), including NaN (both false);
//   * the border factor (xq <= qmin ? qmin - zp : qmax - zp) IS r - zp, because outside the
//     open range the clamped, rounded r equals qmin or qmax exactly (NaN -> qmax, as in the
//     reference where fmin drops the NaN first);
//   * in the streaming kernels the interior term g'*(xfq - x)*inv_s is formed as
//     g' * ((xfq - x)*inv_s) and fused into the accumulation (<= 2 ulp per term from the
//     reference's left-to-right product; only sums are observable and they are held to 1e-6);
//   * dB = (!mask) * g' is exact, so fma(g', 1 - m, acc) adds exactly the reference's term.
// EXACT selects the reference's fp32 terms bit for bit (kernels that own short channels, where
// few terms are summed and nothing averages the 2-ulp difference out); ACC is the accumulator
// type (float partials are promoted to double after <= 32-64 terms).
template <int MODE, int BMODE, bool EXACT, typename ACC>
__device__ __forceinline__ float fq_backward(float g, float x_in, const Chan& c, ACC& accS, ACC& accB) {
    const float x = pre_op<MODE>(x_in);                                // fused prologue: the fake-quant sees relu(x)
    const float v = affine_v<MODE>(x, c);
    const bool mask = (v > c.qmin) && (v < c.qmax);
    const float m = mask ? 1.0f : 0.0f;
    float dX = bmode_passthrough(BMODE) ? g : __fmul_rn(g, m);         // g*mask keeps -0 / NaN like the reference
    if (mode_relu(MODE)) dX = (x <= 0.0f) ? 0.0f : dX;                 // relu backward (threshold_backward): exact 0, NaN x passes
    if (bmode_reduces(BMODE)) {
        const float xq = fmaxf(fminf(v, c.qmax), c.qmin);          // min first: NaN -> qmax
        const float t = __fsub_rn(rintf(xq), c.zp);
        const float d = __fmaf_rn(t, c.s, -x);                     // xfq - x, fused as in the reference build
        const float gg = (BMODE == B_INIT) ? __fmul_rn(2.0f, d) : g;
        const float nm = __fsub_rn(1.0f, m);
        if (!EXACT) {
            // streaming kernels: one select + two fused accumulations
            const float w = mask ? __fmul_rn(d, c.inv_s) : t;
            accS = __fmaf_rn(gg, w, accS);
            accB = __fmaf_rn(gg, nm, accB);
        } else {
            // warp-group kernels (short channels): the reference's fp32 terms bit for bit,
            // (g'*d)*inv_s or g'*(r - zp), summed in double
            const float dS = mask ? __fmul_rn(__fmul_rn(gg, d), c.inv_s) : __fmul_rn(gg, t);
            accS += (ACC)dS;
            accB += (ACC)__fmul_rn(gg, nm);
        }
    }
    return dX;
}

// ---------------------------------------------------------------------------------------------
// tile bookkeeping.  A tile is owned by a GROUP of G threads: the whole CTA (G == THREADS) for
// long channel slices, or one warp (G == 32) for short ones (e.g. conv-weight rows), so that
// eight independent rows share a 256-thread CTA and reductions never leave the warp.
// ---------------------------------------------------------------------------------------------
template <int G, int THREADS>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp(); else __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// fixed-order group reduction of two doubles; result valid in group thread 0
template <int G, int THREADS>
__device__ __forceinline__ void group_sum2(double& a, double& b, double* sm /* [64] per CTA */) {
    a = warp_sum(a); b = warp_sum(b);
    if (G == 32) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < NW ? sm[lane] : 0.0;
        b = lane < NW ? sm[32 + lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
    }
    __syncthreads();   // sm[] may be reused by the caller
}

// Programmatic dependent launch (sm_90+): let the NEXT kernel of the stream start launching while
// this grid drains, and - when this grid itself was launched as a dependent - wait for the
// previous grid to complete before touching any global memory.  No-ops for ordinary launches.
// The wait is placed as LATE as possible: descriptor staging (kernel parameters / plan tables, which no
// kernel ever writes), tile geometry and walker set-up run while the predecessor is still draining;
// only the first access to tensor or parameter memory has to sit behind pdl_wait().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }
// diagnostic build (-DLSQ_FLAT_TRACE): per-CTA globaltimer stamps of the lean per-tensor kernels into a caller-provided buffer that
// the host passes in Seg::stats_out (unused by these kernels); tools/flattrace.py
#ifdef LSQ_FLAT_TRACE
#define LSQ_FTRACE(i) do { if (threadIdx.x == 0 && sg.stats_out) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    unsigned long long* tb_ = reinterpret_cast<unsigned long long*>(sg.stats_out); tb_[(blockIdx.x + 1) * 4 + (i)] = t_; \
    if ((i) == 0 && blockIdx.x == 0) { tb_[0] = (unsigned long long)sg.inner; tb_[1] = gridDim.x; tb_[2] = sg.g ? 2 : 1; } } } while (0)
#else
#define LSQ_FTRACE(i) do {} while (0)
#endif
// Non-binding L2 prefetch of the line a thread will load first, issued BEFORE griddepcontrol.wait: while the predecessor's tail
// drains (DRAM otherwise idle) the kernel's first round of loads is already on its way into L2, so the ramp after the wait starts
// from L2 latency instead of DRAM latency.  Harmless if the predecessor is still producing the data: L2 is the point of coherence,
// a later write updates the prefetched line in place and the loads after the wait see it.
// gpu-scope acquire-release fence for the publish -> ticket -> last-arriver chains: MEMBAR.ALL.GPU, where __threadfence() is the
// sequentially consistent MEMBAR.SC.GPU (the chains need release / acquire ordering only)
#ifndef LSQ_LIGHT_FENCE
#define LSQ_LIGHT_FENCE 1
#endif
__device__ __forceinline__ void fence_acq_rel_gpu() {
#if LSQ_LIGHT_FENCE
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
    __threadfence();
#endif
}
__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// A group may walk several CONSECUTIVE tiles (rows of the same weight tensor, as a rule), so
// the descriptor look-up and staging are paid once per few rows, not once per row.
__host__ __device__ constexpr int tiles_per_group(int group) { return (void)group, 1; }   // r1: 4 rows per warp measured slower (fewer CTAs in flight); kept as a knob

// Copy the tile's descriptor into shared memory (uniform broadcast reads afterwards) unless the
// group already holds it (`staged` = index of the staged segment, -1 = the single descriptor).
template <int G, int THREADS>
__device__ __forceinline__ void stage_segment(const Seg& single, const Seg* table, const int* tile_seg, int nseg,
                                              long long gtile, Seg* dst_seg, int& staged) {
    int want = -1;
    if (table != nullptr) {
        if (tile_seg != nullptr) {
            want = tile_seg[gtile];               // plan-time lookup: one load instead of a dependent search
        } else {
            int lo = 0, hi = nseg - 1;            // last segment with tile_begin <= gtile
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (table[mid].tile_begin <= gtile) lo = mid; else hi = mid - 1;
            }
            want = lo;
        }
    }
    if (want == staged) return;                   // uniform across the group
    const int tg = threadIdx.x % G;
    uint32_t* dst = reinterpret_cast<uint32_t*>(dst_seg);
    constexpr int NW = sizeof(Seg) / 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(want < 0 ? &single : &table[want]);
    if (staged != -2) group_sync<G, THREADS>();   // readers of the previous descriptor are done
    for (int i = tg; i < NW; i += G) dst[i] = src[i];
    group_sync<G, THREADS>();
    staged = want;
}

// 64-bit division costs ~100 instructions; operands almost always fit 32 bits (uniform branch)
__device__ __forceinline__ long long fast_div(long long a, long long b) {
    if ((((unsigned long long)a | (unsigned long long)b) >> 32) == 0) return (long long)((unsigned)a / (unsigned)b);
    return a / b;
}

// Walks the units [u0, u1) of one channel slice, G threads interleaved, yielding unit addresses.
// Units are addressed as (row n, column col) with address = base + n*row_stride + col; regime 0
// (contiguous channel) uses artificial rows of vpr = row_stride = 2^30 units, so all per-thread
// state is 32-bit and one code path serves both regimes.
struct Walker {
    long long base, row_stride;
    int n, col, dn, dcol, vpr;
    unsigned left;            // units this thread still has to visit, in steps of G (tile-local)
    // contiguous tiles: this group walks [u0, u1) in steps of G.  Interleaved tiles
    // (sg.interleave): split j of a channel starts at unit j*G and steps by splits*G, so at any
    // moment all CTAs of a launch stream through one narrow window of memory.
    template <int G>
    __device__ __forceinline__ void init(const Seg& sg, long long u0, long long u1, long long base_unit, int tg) {
        vpr = (int)sg.vpr; row_stride = sg.row_stride; base = base_unit;
        const long long u = u0 + tg;
        const long long step = sg.interleave ? (long long)sg.splits * G : (long long)G;
        left = (u < u1) ? (unsigned)fast_div(u1 - u + step - 1, step) : 0u;
        if (sg.regime == 0) {                       // artificial rows of 2^30 units: shifts, no division
            n = (int)(u >> 30); col = (int)(u & ((1LL << 30) - 1));
            dn = (int)(step >> 30); dcol = (int)(step & ((1LL << 30) - 1));
        } else {
            const long long nn = fast_div(u, vpr);
            n = (int)nn; col = (int)(u - nn * vpr);
            dn = (int)fast_div(step, vpr); dcol = (int)(step - (long long)dn * vpr);
        }
    }
    __device__ __forceinline__ bool more() const { return left != 0u; }
    __device__ __forceinline__ bool next(long long& addr) {
        const bool ok = left != 0u;
        addr = base + (long long)n * row_stride + col;
        left -= ok ? 1u : 0u;
        col += dcol; n += dn;
        if (col >= vpr) { col -= vpr; n++; }
        return ok;
    }
};

// Tile geometry shared by all kernels.
struct TileCtx {
    long long c, ltile, pidx, u0, u1, base_unit;
    long long peel_begin0, peel_n0, peel_begin1, peel_n1;   // scalar head / tail (regime 0, VEC > 1, split 0)
    int j;
};
template <int VEC>
__device__ __forceinline__ TileCtx make_tile(const Seg& sg, long long gtile) {
    TileCtx t;
    t.ltile = gtile - sg.tile_begin;
    if (sg.splits == 1) { t.c = t.ltile; t.j = 0; }
    else { t.c = fast_div(t.ltile, sg.splits); t.j = (int)(t.ltile - t.c * sg.splits); }
    t.pidx = sg.per_channel ? t.c : 0;
    long long chan_units = sg.chan_units;
    t.peel_n0 = t.peel_n1 = 0; t.peel_begin0 = t.peel_begin1 = 0;
    if (sg.regime == 0) {
        // channel c is the contiguous element range [c*inner, (c+1)*inner); with VEC > 1 only the
        // 16-byte aligned body is vectorised, head / tail elements are peeled by split 0
        const long long rb = t.c * sg.inner, re = rb + sg.inner;
        long long bb = rb, be = re;
        if constexpr (VEC > 1) {
            bb = (rb + VEC - 1) / VEC * VEC;
            be = re / VEC * VEC;
            if (bb > be) { bb = re; be = re; }
            if (t.j == 0) { t.peel_begin0 = rb; t.peel_n0 = (bb < re ? bb : re) - rb; t.peel_begin1 = be; t.peel_n1 = re - be; }
        }
        chan_units = (be - bb) / VEC;
        t.base_unit = bb / VEC;
    } else {
        t.base_unit = t.c * sg.vpr;   // channel c starts vpr units after channel c-1 inside a row group
    }
    if (sg.interleave) {
        t.u0 = (long long)t.j * sg.group;          // first unit of this split; Walker strides by splits * G
        t.u1 = chan_units;
    } else {
        t.u0 = (long long)t.j * sg.units_per_split;
        const long long e = t.u0 + sg.units_per_split;
        t.u1 = e < chan_units ? e : chan_units;
    }
    if (t.u0 > t.u1) t.u0 = t.u1;
    return t;
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int NW, bool INIT, int G, int THREADS, int UNROLL, int LD, int ST, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_fwd_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                long long total_tiles) {
    using Tr = ElemTraits<T>;
    constexpr int VEC = UnitOf<T, NW>::VEC;
    constexpr int UB = NW * 4;   // unit bytes
    constexpr int GROUPS = THREADS / G;
    constexpr bool RAWCOPY = INIT && !mode_relu(MODE) && !mode_add(MODE);   // learned init: y is x's bits (lsq_kernel.h:13) unless a prologue is fused in
    constexpr bool ADD = mode_add(MODE);
    __shared__ Seg smem_seg[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    constexpr int ROWS = tiles_per_group(G);
    pdl_trigger();
    int staged = -2;
    for (int row_i = 0; row_i < ROWS; row_i++) {
    const long long gtile = ((long long)blockIdx.x * GROUPS + grp) * ROWS + row_i;
    if (gtile >= total_tiles) break;              // whole group leaves together
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    const T* __restrict__ xp = reinterpret_cast<const T*>(sg.x);
    const T* __restrict__ x2p = reinterpret_cast<const T*>(sg.x2);
    T* __restrict__ yp = reinterpret_cast<T*>(sg.y);
    Walker w{};
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    if constexpr (VEC > 1) {
        if (sg.flags) {
            Walker pw = w;                        // a copy walks ahead; the real walker is untouched
            for (int d = 0; d < sg.flags && pw.more(); d++) {
                long long a0;
                pw.next(a0);
                l2_prefetch(reinterpret_cast<const char*>(xp) + a0 * UB);
                if constexpr (ADD) l2_prefetch(reinterpret_cast<const char*>(x2p) + a0 * UB);
            }
        }
    }
    pdl_wait();                                   // first touch of tensor / parameter memory is below
    LazyChan<MODE> lch;
    lch.issue(sg, tl.pidx);

    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            const float xin = ADD ? pre_add<T>(Tr::to_f(xp[e]), Tr::to_f(x2p[e])) : Tr::to_f(xp[e]);
            if constexpr (INIT) yp[e] = RAWCOPY ? xp[e] : Tr::from_f(pre_op<MODE>(xin));
            else yp[e] = Tr::from_f(fq_forward<MODE>(xin, lch.get(sg)));
        }
    }
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        if constexpr (VEC > 1) {
            Raw<NW> xr[UNROLL], x2r[ADD ? UNROLL : 1];
#pragma unroll
            for (int k = 0; k < UNROLL; k++) { // tail lanes re-read unit 0 (always valid) so loads stay unconditional
                xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + (ok[k] ? addr[k] : addr[0]) * UB);
                if constexpr (ADD) x2r[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(x2p) + (ok[k] ? addr[k] : addr[0]) * UB);
            }
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                if (RAWCOPY) { st_unit<ST, NW>(reinterpret_cast<char*>(yp) + addr[k] * UB, xr[k]); continue; }
                float f[VEC];
                unpack_unit<T, NW>(xr[k], f);
                if constexpr (ADD) {
                    float f2[VEC];
                    unpack_unit<T, NW>(x2r[k], f2);
#pragma unroll
                    for (int e = 0; e < VEC; e++) f[e] = pre_add<T>(f[e], f2[e]);
                }
                if constexpr (INIT) {                   // learned init behind a fused prologue: y = relu(x)
#pragma unroll
                    for (int e = 0; e < VEC; e++) f[e] = pre_op<MODE>(f[e]);
                } else {
                    const Chan& ch = lch.get(sg);
#pragma unroll
                    for (int e = 0; e < VEC; e++) f[e] = fq_forward<MODE>(f[e], ch);
                }
                st_unit<ST, NW>(reinterpret_cast<char*>(yp) + addr[k] * UB, pack_unit<T, NW>(f));
            }
        } else {
            T xr[UNROLL], x2r[ADD ? UNROLL : 1];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                if (ok[k]) { xr[k] = xp[addr[k]]; if constexpr (ADD) x2r[k] = x2p[addr[k]]; }
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                if (ok[k]) {
                    const float xin = ADD ? pre_add<T>(Tr::to_f(xr[k]), Tr::to_f(x2r[ADD ? k : 0])) : Tr::to_f(xr[k]);
                    if constexpr (INIT) yp[addr[k]] = RAWCOPY ? xr[k] : Tr::from_f(pre_op<MODE>(xin));
                    else yp[addr[k]] = Tr::from_f(fq_forward<MODE>(xin, lch.get(sg)));
                }
        }
    }
    }   // tiles of this group
}

// ---------------------------------------------------------------------------------------------
// Cross-tile finalisation shared by backward and stats: publish this tile's two partial sums,
// take a ticket, and let the LAST arriver of the channel add all partials of the channel in a
// fixed order (bit-reproducible whatever the arrival order).  Returns true in every thread of
// the finishing group, with the channel totals valid in group thread 0.
// ---------------------------------------------------------------------------------------------
template <int G, int THREADS>
__device__ __forceinline__ bool channel_finish(const Seg& sg, const TileCtx& tl, double& a, double& b,
                                               double* red, int* last_flag, int tg) {
    group_sum2<G, THREADS>(a, b, red);
    if (sg.splits == 1) return true;
    if (tg == 0) {
        reinterpret_cast<double2*>(sg.partials)[tl.ltile] = make_double2(a, b);
        fence_acq_rel_gpu();
        const unsigned prev = atomicAdd(&sg.counters[tl.c], 1u);
        *last_flag = (prev == (unsigned)sg.splits - 1u);
    }
    group_sync<G, THREADS>();
    const bool last = *last_flag != 0;
    if (!last) return false;
    fence_acq_rel_gpu();
    a = 0.0; b = 0.0;
    // every thread first ISSUES all of its partial-pair loads (one 128-bit L2 read each, four per round: a single-wave
    // launch has <= 4 * G splits), then adds them in index order: one L2 round trip on the launch's critical tail instead
    // of one per pair, same fixed summation order
    const double2* p = reinterpret_cast<const double2*>(sg.partials) + tl.c * sg.splits;   // 16-byte aligned: workspace layouts keep it so
    for (int base = tg; base < sg.splits; base += 4 * G) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = base + k * G;
            v[k] = i < sg.splits ? __ldcg(p + i) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { a += v[k].x; b += v[k].y; }
    }
    group_sum2<G, THREADS>(a, b, red);
    if (tg == 0) sg.counters[tl.c] = 0u;          // leave the workspace zeroed for the next call
    return true;
}

// ---------------------------------------------------------------------------------------------
// backward kernel: gx write + grad_scale / grad_shift reduction; x and grad are read exactly once
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int NW, int BMODE, int G, int THREADS, int UNROLL, int LD, int ST, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_bwd_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                long long total_tiles) {
    using Tr = ElemTraits<T>;
    constexpr int VEC = UnitOf<T, NW>::VEC;
    constexpr int UB = NW * 4;   // unit bytes
    constexpr int GROUPS = THREADS / G;
    constexpr bool RAWPASS = bmode_passthrough(BMODE) && !mode_relu(MODE);   // learned init: gx is grad's bits (lsq_kernel.h:35)
    constexpr bool ADD = mode_add(MODE);
    __shared__ Seg smem_seg[GROUPS];
    __shared__ double red[64];
    __shared__ int last_flag[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    constexpr int ROWS = tiles_per_group(G);
    pdl_trigger();
    int staged = -2;
    for (int row_i = 0; row_i < ROWS; row_i++) {
    const long long gtile = ((long long)blockIdx.x * GROUPS + grp) * ROWS + row_i;
    if (gtile >= total_tiles) break;              // whole group leaves together
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    const T* __restrict__ xp = reinterpret_cast<const T*>(sg.x);
    const T* __restrict__ x2p = reinterpret_cast<const T*>(sg.x2);
    const T* __restrict__ gp = reinterpret_cast<const T*>(sg.g);
    T* __restrict__ gxp = reinterpret_cast<T*>(sg.gx);
    const bool write_gx = gxp != nullptr;
    Walker w{};
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    if constexpr (VEC > 1) {
        if (sg.flags) {
            Walker pw = w;                        // a copy walks ahead; the real walker is untouched
            for (int d = 0; d < sg.flags && pw.more(); d++) {
                long long a0;
                pw.next(a0);
                l2_prefetch(reinterpret_cast<const char*>(xp) + a0 * UB);
                l2_prefetch(reinterpret_cast<const char*>(gp) + a0 * UB);
                if constexpr (ADD) l2_prefetch(reinterpret_cast<const char*>(x2p) + a0 * UB);
            }
        }
    }
    pdl_wait();                                   // first touch of tensor / parameter memory is below
    LazyChan<MODE> lch;
    lch.issue(sg, tl.pidx);

    double accS = 0.0, accB = 0.0;
    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            const float xin = ADD ? pre_add<T>(Tr::to_f(xp[e]), Tr::to_f(x2p[e])) : Tr::to_f(xp[e]);
            const float dx = fq_backward<MODE, BMODE, true>(Tr::to_f(gp[e]), xin, lch.get(sg), accS, accB);
            if (write_gx) gxp[e] = RAWPASS ? gp[e] : Tr::from_f(dx);
        }
    }
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        // streaming (CTA-group) kernels: fp32 partial over <= UNROLL*VEC terms, then promoted to
        // double; warp-group kernels (short channels, cancellation-prone sums): double per term
        using Acc = typename std::conditional<G == 32, double, float>::type;
        Acc ls = 0, lb = 0;
        if constexpr (VEC > 1) {
            Raw<NW> xr[UNROLL], gr[UNROLL], x2r[ADD ? UNROLL : 1];
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {   // tail lanes re-read unit 0 (always valid) so loads stay unconditional
                const long long a = ok[k] ? addr[k] : addr[0];
                xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + a * UB);
                if constexpr (ADD) x2r[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(x2p) + a * UB);
                gr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(gp) + a * UB);
            }
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                const Chan& ch = lch.get(sg);
                float fx[VEC], fg[VEC];
                unpack_unit<T, NW>(xr[k], fx);
                if constexpr (ADD) {
                    float f2[VEC];
                    unpack_unit<T, NW>(x2r[k], f2);
#pragma unroll
                    for (int e = 0; e < VEC; e++) fx[e] = pre_add<T>(fx[e], f2[e]);
                }
                unpack_unit<T, NW>(gr[k], fg);
#pragma unroll
                for (int e = 0; e < VEC; e++) fg[e] = fq_backward<MODE, BMODE, G == 32>(fg[e], fx[e], ch, ls, lb);
                if (write_gx) {
                    if (RAWPASS) st_unit<ST, NW>(reinterpret_cast<char*>(gxp) + addr[k] * UB, gr[k]);
                    else st_unit<ST, NW>(reinterpret_cast<char*>(gxp) + addr[k] * UB, pack_unit<T, NW>(fg));
                }
            }
        } else {
            T xr[UNROLL], gr[UNROLL], x2r[ADD ? UNROLL : 1];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                if (ok[k]) { xr[k] = xp[addr[k]]; gr[k] = gp[addr[k]]; if constexpr (ADD) x2r[k] = x2p[addr[k]]; }
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                const float xin = ADD ? pre_add<T>(Tr::to_f(xr[k]), Tr::to_f(x2r[ADD ? k : 0])) : Tr::to_f(xr[k]);
                const float dx = fq_backward<MODE, BMODE, G == 32>(Tr::to_f(gr[k]), xin, lch.get(sg), ls, lb);
                if (write_gx) gxp[addr[k]] = RAWPASS ? gr[k] : Tr::from_f(dx);
            }
        }
        accS += (double)ls; accB += (double)lb;
    }

    if (!bmode_reduces(BMODE)) {   // eval: parameters get exact zeros (lsq_kernel.h:143-144)
        if (tl.j == 0 && tg == 0) {
            store_param(sg.gscale, tl.pidx, sg.pdt, 0.0);
            store_param(sg.gshift, tl.pidx, sg.pdt, 0.0);
        }
        continue;
    }
    if (!channel_finish<G, THREADS>(sg, tl, accS, accB, red, &last_flag[grp], tg)) continue;
    if (tg == 0) {
        store_param(sg.gscale, tl.pidx, sg.pdt, accS * sg.gs);
        store_param(sg.gshift, tl.pidx, sg.pdt, sg.sym ? 0.0 : accB * sg.gs);
    }
    }   // tiles of this group
}

// ---------------------------------------------------------------------------------------------
// mu +- 3 sigma weight-init statistics (observers.py:329-337): ONE read of w.
// Shifted sums in double: S1 = sum(w - p), S2 = sum((w - p)^2), p = first element of the channel.
// ---------------------------------------------------------------------------------------------
// scale = max(|mu - 3 sd|, |mu + 3 sd|) / 2^bitness from the shifted sums S1 = sum(w - p), S2 = sum((w - p)^2) of K elements
__device__ __forceinline__ float stats_scale(double s1, double s2, float pivot, double K, float denom) {
    const double mean = (double)pivot + s1 / K;
    double var = (s2 - s1 * s1 / K) / (K - 1.0);                   // unbiased (torch.std default); K == 1 -> NaN
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, sd = (float)sqrt(var);
    const float lo = fabsf(__fsub_rn(mu, __fmul_rn(3.0f, sd)));
    const float hi = fabsf(__fadd_rn(mu, __fmul_rn(3.0f, sd)));
    return __fdiv_rn(fmaxf(lo, hi), denom);
}

template <typename T, int NW, int G, int THREADS, int UNROLL, int LD, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_stats_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                long long total_tiles) {
    using Tr = ElemTraits<T>;
    constexpr int VEC = UnitOf<T, NW>::VEC;
    constexpr int UB = NW * 4;   // unit bytes
    constexpr int GROUPS = THREADS / G;
    __shared__ Seg smem_seg[GROUPS];
    __shared__ double red[64];
    __shared__ int last_flag[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    constexpr int ROWS = tiles_per_group(G);
    pdl_trigger();
    int staged = -2;
    for (int row_i = 0; row_i < ROWS; row_i++) {
    const long long gtile = ((long long)blockIdx.x * GROUPS + grp) * ROWS + row_i;
    if (gtile >= total_tiles) break;              // whole group leaves together
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    const T* __restrict__ xp = reinterpret_cast<const T*>(sg.x);
    pdl_wait();
    const float pivot = Tr::to_f(xp[tl.c * sg.inner]);      // element (0, c, 0): shift that keeps the sums small

    // per unit: d = w - pivot, sum d and sum d^2 over the unit's <= 16 elements in fp32 (FADD / FFMA),
    // then one promotion to fp64 per unit - the fp64 pipe sees 2 adds per unit instead of 4 ops per element
    double s1 = 0.0, s2 = 0.0;
    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            const float d = __fsub_rn(Tr::to_f(xp[e]), pivot);
            s1 += (double)d; s2 += (double)__fmul_rn(d, d);
        }
    }
    Walker w{};
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        if constexpr (VEC > 1) {
            Raw<NW> xr[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + (ok[k] ? addr[k] : addr[0]) * UB);
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                float f[VEC];
                unpack_unit<T, NW>(xr[k], f);
                float u1 = 0.f, u2 = 0.f;
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    const float d = __fsub_rn(f[e], pivot);
                    u1 = __fadd_rn(u1, d); u2 = __fmaf_rn(d, d, u2);
                }
                s1 += (double)u1; s2 += (double)u2;
            }
        } else {
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                const float d = __fsub_rn(Tr::to_f(xp[addr[k]]), pivot);
                s1 += (double)d; s2 += (double)__fmul_rn(d, d);
            }
        }
    }
    if (!channel_finish<G, THREADS>(sg, tl, s1, s2, red, &last_flag[grp], tg)) continue;
    if (tg == 0) sg.stats_out[tl.c] = stats_scale(s1, s2, pivot, (double)sg.chan_elems, sg.stats_denom);
    }   // tiles of this group
}

// Row c of a contiguous (C, inner) tensor whose base is 32-byte aligned: scalar head up to the first 32-byte boundary, whole
// 256-bit units, scalar tail.  Rows that are 32-byte multiples (all but a first conv layer's) have neither head nor tail.
template <typename T, int VEC>
struct RowGeom {
    long long e0;          // element index of the row's first element
    int head, units, tail; // elements, units, elements
    __device__ __forceinline__ void init(long long c, long long inner) {
        e0 = c * inner;
        const int mis = (int)((e0 * (long long)sizeof(T)) & 31);
        long long h = mis ? (32 - mis) / (int)sizeof(T) : 0;
        if (h > inner) h = inner;
        head = (int)h;
        const long long body = inner - h;
        units = (int)(body / VEC);
        tail = (int)(body - (long long)units * VEC);
    }
    // i-th scalar element (0 <= i < head + tail) -> element index
    __device__ __forceinline__ long long scalar_elem(int i) const {
        return i < head ? e0 + i : e0 + head + (long long)units * VEC + (i - head);
    }
};

// ---------------------------------------------------------------------------------------------
// Weight rows, the statistics kernel's everyday case (every conv / linear weight with axis 0: contiguous rows of a few
// hundred to a few thousand elements, 32-byte aligned): ONE WARP PER ROW with nothing but the row walk in the loop.
// The general kernel above spends ~600 warp instructions per row on descriptor staging, tile geometry, peel logic and the
// unit walker - more than the arithmetic of a 4 KB row (ncu, profiles/); here the descriptor fields are read straight from
// the plan table (plan-time constants in L2), the row is base + lane*32 B + k*1 KB, and UNROLL 256-bit loads per lane
// are in flight (two at 6 CTAs/SM measured best: rows are short, more resident warps beat deeper prefetch).  Same arithmetic as lsq_stats_kernel (fp32 within a unit, fp64 across units, shifted by the row's first
// element), so the two kernels agree to the last bit of the fp64 sums' rounding.
// ---------------------------------------------------------------------------------------------
template <typename T, int THREADS, int UNROLL, int LD, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_rowstats_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                    long long total_tiles) {
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    const long long gtile = (long long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    if (gtile >= total_tiles) return;
    const Seg* sg = &single;
    if (table != nullptr) {                       // plan tables are written at plan creation only: safe to read before the wait
        int want;
        if (tile_seg != nullptr) want = tile_seg[gtile];
        else {
            int lo = 0, hi = nseg - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (table[mid].tile_begin <= gtile) lo = mid; else hi = mid - 1; }
            want = lo;
        }
        sg = &table[want];
    }
    const long long c = gtile - sg->tile_begin, inner = sg->inner;
    RowGeom<T, VEC> rg;
    rg.init(c, inner);
    const int units = rg.units;                   // <= Tuning::warp_units
    const T* xp = reinterpret_cast<const T*>(sg->x);
    float* out = sg->stats_out;
    const float denom = sg->stats_denom;
    pdl_wait();
    const float pivot = ElemTraits<T>::to_f(xp[rg.e0]);
    const char* p = reinterpret_cast<const char*>(xp + rg.e0 + rg.head) + lane * UB;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < rg.head + rg.tail; i += 32) {
        const float d = __fsub_rn(ElemTraits<T>::to_f(xp[rg.scalar_elem(i)]), pivot);
        s1 += (double)d; s2 += (double)__fmul_rn(d, d);
    }
    for (int u = lane; u < units; u += 32 * UNROLL) {
        Raw<NW> r[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++)          // lanes past the row end re-read their first unit: loads stay unconditional
            r[k] = ld_unit<LD, NW>(p + (u + 32 * k < units ? k : 0) * (32 * UB));
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (u + 32 * k >= units) continue;
            float f[VEC];
            unpack_unit<T, NW>(r[k], f);
            float u1 = 0.f, u2 = 0.f;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const float d = __fsub_rn(f[e], pivot);
                u1 = __fadd_rn(u1, d); u2 = __fmaf_rn(d, d, u2);
            }
            s1 += (double)u1; s2 += (double)u2;
        }
        p += UNROLL * 32 * UB;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) out[c] = stats_scale(s1, s2, pivot, (double)inner, denom);
}


__device__ __forceinline__ const Seg* find_segment(const Seg& single, const Seg* table, const int* tile_seg, int nseg, long long gtile) {
    if (table == nullptr) return &single;
    if (tile_seg != nullptr) return &table[tile_seg[gtile]];
    int lo = 0, hi = nseg - 1;                    // last segment with tile_begin <= gtile
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (table[mid].tile_begin <= gtile) lo = mid; else hi = mid - 1; }
    return &table[lo];
}

// ---------------------------------------------------------------------------------------------
// lsq_rowstats_kernel with a SHORTER DEPENDENT CHAIN (round 2).  With 27 560 rows of ~4 KB (all conv / linear weights of
// ResNet-50) the one-warp-per-row kernel is bound by the life of a warp, not by bandwidth: the same launch over bf16 weights
// (half the bytes) takes 21.4 us against 26.6 us for fp32.  A warp's life is a chain of dependent round trips - tile map ->
// descriptor -> pivot element -> row data -> reduction; here the plan holds one 32-byte ROW ENTRY per row (row pointer, output
// slot, length, denominator: one broadcast load instead of two dependent ones) and the pivot comes out of the loaded data by
// shuffle (lane 0's first unit starts with the row's first element), so the chain is entry -> data -> reduction.
// Same arithmetic and summation order as lsq_rowstats_kernel: bit-identical scales.  (A persistent variant that also kept the
// next row's entry and first units in flight measured slower, profiles/r2_rowstats.md.)
// `tile_seg` carries the row-entry table; `single.stats_out` the base of the class's output slots.
// ---------------------------------------------------------------------------------------------
struct RowEntry {
    const void* row;      // first element of the row
    long long out_rel;    // output slot relative to the class's first segment
    int inner;            // elements in the row
    float denom;          // 2^bitness
    int pad[2];
};
static_assert(sizeof(RowEntry) == 32, "one 256-bit load per row entry");

template <typename T, int THREADS, int UNROLL, int LD, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_rowstats3_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                     long long total_tiles) {
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32, WARPS = THREADS / 32;
    // the CTA's rows finish together: (s1, s2, pivot, length, denominator, slot) of every warp's row, then ONE warp runs the
    // fp64 division / square-root chains of all of them side by side - they are ~45 % of the warp instructions of a launch
    // when every warp runs them for its single row (ncu: 360 warp instructions per 4 KB row, 160 of them here)
    __shared__ double fin_s1[WARPS], fin_s2[WARPS];
    __shared__ float fin_pivot[WARPS], fin_denom[WARPS];
    __shared__ int fin_inner[WARPS];
    __shared__ long long fin_out[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_trigger();
    const long long gtile = (long long)blockIdx.x * WARPS + warp;
    const bool live = gtile < total_tiles;
    double s1 = 0.0, s2 = 0.0;
    float pivot = 0.f, denom = 1.f;
    int inner = 0;
    long long out_rel = 0;
    if (live) {
        // plan tables are written at plan creation only: safe to read before the wait
        const Raw<8> er = ld_unit<LD_DEFAULT, 8>(reinterpret_cast<const RowEntry*>(tile_seg) + gtile);
        const T* rowp = reinterpret_cast<const T*>(((unsigned long long)er.w[1] << 32) | er.w[0]);
        out_rel = (long long)(((unsigned long long)er.w[3] << 32) | er.w[2]);
        inner = (int)er.w[4];
        denom = __uint_as_float(er.w[5]);
        const int mis = (int)(reinterpret_cast<uintptr_t>(rowp) & 31);
        int head = mis ? (32 - mis) / (int)sizeof(T) : 0;
        if (head > inner) head = inner;
        const int units = (inner - head) / VEC;
        const int tail = inner - head - units * VEC;
        const char* p = reinterpret_cast<const char*>(rowp + head) + lane * UB;
        if (single.flags && lane < units) l2_prefetch(p);
        pdl_wait();
        Raw<NW> r[UNROLL];
        if (units > 0) {
#pragma unroll
            for (int k = 0; k < UNROLL; k++)          // lanes past the row end re-read a valid unit: loads stay unconditional
                r[k] = ld_unit<LD, NW>(lane + 32 * k < units ? p + k * (32 * UB) : (lane < units ? p : reinterpret_cast<const char*>(rowp + head)));
        }
        if (head == 0 && units > 0) {                 // lane 0's first unit starts with the row's first element
            float f0[VEC];
            unpack_unit<T, NW>(r[0], f0);
            pivot = __shfl_sync(0xffffffffu, f0[0], 0);
        } else {
            pivot = ElemTraits<T>::to_f(rowp[0]);
        }
        for (int i = lane; i < head + tail; i += 32) {
            const int e = i < head ? i : head + units * VEC + (i - head);
            const float d = __fsub_rn(ElemTraits<T>::to_f(rowp[e]), pivot);
            s1 += (double)d; s2 += (double)__fmul_rn(d, d);
        }
        for (int u = lane; u < units; u += 32 * UNROLL) {
            if (u != lane) {
#pragma unroll
                for (int k = 0; k < UNROLL; k++) r[k] = ld_unit<LD, NW>(p + (u + 32 * k < units ? k : 0) * (32 * UB));
            }
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (u + 32 * k >= units) continue;
                float f[VEC];
                unpack_unit<T, NW>(r[k], f);
                float u1 = 0.f, u2 = 0.f;
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    const float d = __fsub_rn(f[e], pivot);
                    u1 = __fadd_rn(u1, d); u2 = __fmaf_rn(d, d, u2);
                }
                s1 += (double)u1; s2 += (double)u2;
            }
            p += UNROLL * 32 * UB;
        }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
    }
    if (lane == 0) {
        fin_s1[warp] = s1; fin_s2[warp] = s2; fin_pivot[warp] = pivot; fin_denom[warp] = denom;
        fin_inner[warp] = live ? inner : 0; fin_out[warp] = out_rel;
    }
    __syncthreads();
    if (warp == 0 && lane < WARPS && fin_inner[lane] > 0)
        single.stats_out[fin_out[lane]] = stats_scale(fin_s1[lane], fin_s2[lane], fin_pivot[lane], (double)fin_inner[lane], fin_denom[lane]);
}

// ---------------------------------------------------------------------------------------------
// Weight-row statistics through a BULK-COPY RING (north_star: "TMA staging only where the per-channel layout - weight axis 0 -
// makes tiles pay off").  The register-staged row kernels above keep 2 units per lane in flight for about a third of a warp's
// short life (entry load - data - reduction): ~26 KB in flight per SM, 4.1 TB/s on the 27 560 rows of ResNet-50.  Here a warp
// is persistent (rows r, r + W, ...; gridDim = SM count), prefetches the entries of its next 32 rows with ONE coalesced gather,
// and its lane 0 keeps a private ring of S stages x 2 KB filled by cp.async.bulk (1-D bulk copies: rows are contiguous, no
// tensor map needed) S chunks ahead of the lanes that consume them from shared memory: bytes in flight are set by the ring
// (128 KB per SM), not by registers x occupancy x duty cycle, and nobody waits for a descriptor.
// Only rows that are whole 32-byte units in 32-byte aligned tensors take this path (every conv / linear weight but a first
// layer with K = 147); chunk = 64 units, lane l takes units l and l + 32 of each chunk, i.e. exactly the unit order of
// lsq_rowstats_kernel: same fp32-per-unit / fp64-across-units arithmetic, bit-identical scales.
// mbarrier / bulk-copy helpers are those of lsq_column.cuh (declared there; repeated locally to keep this header standalone).
// ---------------------------------------------------------------------------------------------
namespace rowring {
constexpr int kStages = 4, kStageBytes = 2048, kWarps = 8, kCtasPerSm = 3;
constexpr int kSmemBytes = kWarps * kStages * kStageBytes;      // 64 KB dynamic shared memory per CTA, three CTAs per SM
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ Raw<8> lds_unit32(uint32_t addr) {
    Raw<8> r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "r"(addr + 16));
    return r;
}
}  // namespace rowring

template <typename T>
__global__ void __launch_bounds__(rowring::kWarps * 32, rowring::kCtasPerSm)
lsq_rowstats_ring_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                         long long total_tiles) {
    using namespace rowring;
    constexpr int VEC = UnitOf<T, 8>::VEC;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ __align__(8) unsigned long long bars[kWarps * kStages];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_trigger();
    const long long W = (long long)gridDim.x * kWarps;
    const long long gw = (long long)blockIdx.x * kWarps + warp;
    const uint32_t ring0 = s32(ring_smem) + (uint32_t)warp * (kStages * kStageBytes);
    const uint32_t bar0 = s32(&bars[warp * kStages]);
    if (lane == 0) {
        for (int st = 0; st < kStages; st++) bar_init(bar0 + 8 * st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (gw >= total_tiles) return;
    const long long my_rows = (total_tiles - gw + W - 1) / W;           // rows gw, gw + W, ...
    const RowEntry* entries = reinterpret_cast<const RowEntry*>(tile_seg);
    pdl_wait();
    uint32_t issued = 0, consumed = 0;          // chunk counters of this warp (uniform across lanes)
    for (long long base = 0; base < my_rows; base += 32) {
        const int nb = my_rows - base < 32 ? (int)(my_rows - base) : 32;
        // lane i: entry of the warp's (base + i)-th row - one gather, one latency for up to 32 rows
        unsigned long long e_row = 0; long long e_out = 0; int e_bytes = 0; float e_den = 1.f; int e_inner = 0;
        if (lane < nb) {
            const Raw<8> er = ld_unit<LD_DEFAULT, 8>(entries + (gw + (base + lane) * W));
            e_row = ((unsigned long long)er.w[1] << 32) | er.w[0];
            e_out = (long long)(((unsigned long long)er.w[3] << 32) | er.w[2]);
            e_inner = (int)er.w[4];
            e_den = __uint_as_float(er.w[5]);
            e_bytes = e_inner * (int)sizeof(T);
        }
        int prow = 0, poff = 0;                 // producer cursor: next chunk to issue (row index in this batch, byte offset)
        int pbytes = __shfl_sync(0xffffffffu, e_bytes, 0);
        unsigned long long pptr = __shfl_sync(0xffffffffu, e_row, 0);
        auto issue_ahead = [&]() {              // uniform control flow; only lane 0 touches the barrier / copy engine
            while (prow < nb && issued - consumed < (uint32_t)kStages) {
                const int n = pbytes - poff < kStageBytes ? pbytes - poff : kStageBytes;
                const uint32_t st = issued % kStages;
                if (lane == 0) {
                    bar_expect_tx(bar0 + 8 * st, (uint32_t)n);
                    bulk_load(ring0 + st * kStageBytes, reinterpret_cast<const char*>(pptr) + poff, (uint32_t)n, bar0 + 8 * st);
                }
                issued++;
                poff += n;
                if (poff >= pbytes) {
                    prow++; poff = 0;
                    if (prow < nb) {
                        pbytes = __shfl_sync(0xffffffffu, e_bytes, prow);
                        pptr = __shfl_sync(0xffffffffu, e_row, prow);
                    }
                }
            }
        };
        issue_ahead();
        double k1 = 0.0, k2 = 0.0;              // lane i keeps the sums and the pivot of the batch's i-th row
        float kp = 0.f;
        for (int row = 0; row < nb; row++) {
            const int rbytes = __shfl_sync(0xffffffffu, e_bytes, row);
            float pivot = 0.f;
            double s1 = 0.0, s2 = 0.0;
            for (int off = 0; off < rbytes; off += kStageBytes) {
                const int n = rbytes - off < kStageBytes ? rbytes - off : kStageBytes;
                const uint32_t st = consumed % kStages;
                bar_wait(bar0 + 8 * st, (consumed / kStages) & 1u);
                const uint32_t sbase = ring0 + st * kStageBytes;
                if (off == 0) {                 // the row's first element: the shift that keeps the sums small
                    uint32_t w0;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(sbase));
                    float f0[ElemTraits<T>::PER_WORD];
                    ElemTraits<T>::unpack_word(w0, f0);
                    pivot = f0[0];
                }
                const int units = n >> 5;       // whole 32-byte units (rows on this path are multiples of 32 bytes)
#pragma unroll
                for (int k = 0; k < kStageBytes / 32 / 32; k++) {
                    const int u = lane + 32 * k;
                    if (u < units) {
                        const Raw<8> r = lds_unit32(sbase + (uint32_t)u * 32u);
                        float f[VEC];
                        unpack_unit<T, 8>(r, f);
                        float u1 = 0.f, u2 = 0.f;
#pragma unroll
                        for (int e = 0; e < VEC; e++) {
                            const float d = __fsub_rn(f[e], pivot);
                            u1 = __fadd_rn(u1, d); u2 = __fmaf_rn(d, d, u2);
                        }
                        s1 += (double)u1; s2 += (double)u2;
                    }
                }
                __syncwarp();                   // every lane is done with stage st: it may be refilled
                consumed++;
                issue_ahead();
            }
            s1 = warp_sum(s1); s2 = warp_sum(s2);       // butterfly: every lane holds the totals
            if (lane == row) { k1 = s1; k2 = s2; kp = pivot; }
        }
        // the fp64 division / square root chains of up to 32 rows run side by side, one row per lane
        if (lane < nb) single.stats_out[e_out] = stats_scale(k1, k2, kp, (double)e_inner, e_den);
    }
}

// ---------------------------------------------------------------------------------------------
// Weight rows, forward and backward: the same lean warp-per-row shape as lsq_rowstats_kernel for the rows the plans of a
// model consist of (every conv / linear weight with axis 0, 32-byte aligned rows of <= Tuning::warp_units units).  The
// arithmetic is fq_forward / fq_backward<EXACT> exactly as in the warp-group instantiations of the general kernels (the
// reference's fp32 terms bit for bit, summed in fp64); what is gone is the per-row descriptor staging, tile geometry, peel
// logic and unit walker (~600 warp instructions per row, more than the work of a 4 KB row).
// ---------------------------------------------------------------------------------------------

template <typename T, int MODE, bool INIT, int THREADS, int UNROLL, int LD, int ST, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_rowfwd_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                  long long total_tiles) {
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    const long long gtile = (long long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    if (gtile >= total_tiles) return;
    const Seg* sg = find_segment(single, table, tile_seg, nseg, gtile);     // plan-time constants: readable before the wait
    const long long c = gtile - sg->tile_begin, inner = sg->inner;
    RowGeom<T, VEC> rg;
    rg.init(c, inner);
    const int units = rg.units;
    const T* xp = reinterpret_cast<const T*>(sg->x);
    T* yp = reinterpret_cast<T*>(sg->y);
    const char* px = reinterpret_cast<const char*>(xp + rg.e0 + rg.head) + lane * UB;
    char* py = reinterpret_cast<char*>(yp + rg.e0 + rg.head) + lane * UB;
    const long long pidx = sg->per_channel ? c : 0;
    if (sg->flags && lane < units) l2_prefetch(px);
    pdl_wait();
    Chan ch;
    float sraw = 0.f, braw = 0.f;
    if (!INIT) { sraw = load_param(sg->scale, pidx, sg->pdt); braw = load_param(sg->shift, pidx, sg->pdt); }
    bool have = false;
    if (rg.head + rg.tail > 0) {                  // rows that do not start / end on a 32-byte boundary (uniform per warp)
        if (!INIT) { ch = make_chan<MODE>(sraw, braw, *sg); have = true; }
        for (int i = lane; i < rg.head + rg.tail; i += 32) {
            const long long e = rg.scalar_elem(i);
            yp[e] = INIT ? xp[e] : ElemTraits<T>::from_f(fq_forward<MODE>(ElemTraits<T>::to_f(xp[e]), ch));
        }
    }
    for (int u = lane; u < units; u += 32 * UNROLL) {
        Raw<NW> r[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) r[k] = ld_unit<LD, NW>(px + (u + 32 * k < units ? k : 0) * (32 * UB));
        if (!INIT && !have) { ch = make_chan<MODE>(sraw, braw, *sg); have = true; }    // after the first data loads are in flight
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (u + 32 * k >= units) continue;
            if (INIT) { st_unit<ST, NW>(py + k * (32 * UB), r[k]); continue; }
            float f[VEC];
            unpack_unit<T, NW>(r[k], f);
#pragma unroll
            for (int e = 0; e < VEC; e++) f[e] = fq_forward<MODE>(f[e], ch);
            st_unit<ST, NW>(py + k * (32 * UB), pack_unit<T, NW>(f));
        }
        px += UNROLL * 32 * UB; py += UNROLL * 32 * UB;
    }
}

template <typename T, int MODE, int BMODE, int THREADS, int UNROLL, int LD, int ST, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_rowbwd_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                  long long total_tiles) {
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    const long long gtile = (long long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    if (gtile >= total_tiles) return;
    const Seg* sg = find_segment(single, table, tile_seg, nseg, gtile);
    const long long c = gtile - sg->tile_begin, inner = sg->inner;
    RowGeom<T, VEC> rg;
    rg.init(c, inner);
    const int units = rg.units;
    const T* xp = reinterpret_cast<const T*>(sg->x);
    const T* gp = reinterpret_cast<const T*>(sg->g);
    T* gxp = reinterpret_cast<T*>(sg->gx);
    const long long row_off = (rg.e0 + rg.head) * (long long)sizeof(T) + lane * UB;
    const char* px = reinterpret_cast<const char*>(sg->x) + row_off;
    const char* pg = reinterpret_cast<const char*>(sg->g) + row_off;
    char* pgx = sg->gx ? reinterpret_cast<char*>(sg->gx) + row_off : nullptr;
    const long long pidx = sg->per_channel ? c : 0;
    if (sg->flags && lane < units) { l2_prefetch(px); l2_prefetch(pg); }
    pdl_wait();
    const float sraw = load_param(sg->scale, pidx, sg->pdt), braw = load_param(sg->shift, pidx, sg->pdt);
    Chan ch;
    bool have = false;
    double accS = 0.0, accB = 0.0;
    if (rg.head + rg.tail > 0) {                  // rows that do not start / end on a 32-byte boundary (uniform per warp)
        ch = make_chan<MODE>(sraw, braw, *sg); have = true;
        for (int i = lane; i < rg.head + rg.tail; i += 32) {
            const long long e = rg.scalar_elem(i);
            const float dx = fq_backward<MODE, BMODE, true>(ElemTraits<T>::to_f(gp[e]), ElemTraits<T>::to_f(xp[e]), ch, accS, accB);
            if (gxp) gxp[e] = bmode_passthrough(BMODE) ? gp[e] : ElemTraits<T>::from_f(dx);
        }
    }
    for (int u = lane; u < units; u += 32 * UNROLL) {
        Raw<NW> xr[UNROLL], gr[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const int o = (u + 32 * k < units ? k : 0) * (32 * UB);
            xr[k] = ld_unit<LD, NW>(px + o);
            gr[k] = ld_unit<LD, NW>(pg + o);
        }
        if (!have) { ch = make_chan<MODE>(sraw, braw, *sg); have = true; }
        double ls = 0.0, lb = 0.0;
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (u + 32 * k >= units) continue;
            float fx[VEC], fg[VEC];
            unpack_unit<T, NW>(xr[k], fx);
            unpack_unit<T, NW>(gr[k], fg);
#pragma unroll
            for (int e = 0; e < VEC; e++) fg[e] = fq_backward<MODE, BMODE, true>(fg[e], fx[e], ch, ls, lb);
            if (pgx) {
                if (bmode_passthrough(BMODE)) st_unit<ST, NW>(pgx + k * (32 * UB), gr[k]);
                else st_unit<ST, NW>(pgx + k * (32 * UB), pack_unit<T, NW>(fg));
            }
        }
        accS += ls; accB += lb;
        px += UNROLL * 32 * UB; pg += UNROLL * 32 * UB;
        if (pgx) pgx += UNROLL * 32 * UB;
    }
    if (!bmode_reduces(BMODE)) {                  // eval: exact zeros (lsq_kernel.h:143-144)
        if (lane == 0) { store_param(sg->gscale, pidx, sg->pdt, 0.0); store_param(sg->gshift, pidx, sg->pdt, 0.0); }
        return;
    }
    accS = warp_sum(accS); accB = warp_sum(accB);
    if (lane == 0) {
        store_param(sg->gscale, pidx, sg->pdt, accS * sg->gs);
        store_param(sg->gshift, pidx, sg->pdt, sg->sym ? 0.0 : accB * sg->gs);
    }
}

// ---------------------------------------------------------------------------------------------
// Per-tensor sites launched one at a time (the activation sites of a training step): the lean form of the general kernels
// for ONE contiguous channel in 32-byte aligned buffers.  Same tiles, same interleaved unit <-> thread mapping, same
// arithmetic and the same fixed-order reduction as lsq_fwd_kernel / lsq_bwd_kernel - results are bit-identical - but the
// descriptor stays in the constant bank (no shared-memory staging and barriers), there is no tile geometry, peel or walker
// state to set up (~600 warp instructions per CTA before the first load, 1-1.5 us of every launch's ramp), and the loop
// advances one pointer instead of a (row, column) pair.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, bool INIT, int THREADS, int LD, int ST, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_flatfwd_kernel(const __grid_constant__ Seg sg) {
    using Tr = ElemTraits<T>;
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32;
    pdl_trigger();
    LSQ_FTRACE(0);
    const long long units = sg.inner / VEC;
    const long long stride = (long long)gridDim.x * THREADS;
    long long u = (long long)blockIdx.x * THREADS + threadIdx.x;
    const char* px = reinterpret_cast<const char*>(sg.x) + u * UB;
    char* py = reinterpret_cast<char*>(sg.y) + u * UB;
    const long long sb = stride * UB;
    if (sg.flags && u < units) l2_prefetch(px);     // one unit: the next one is a whole grid stride away (depth 1 / 2 / 3: 6207 / 6197 / 6153 GB/s on the step)
    pdl_wait();                                   // first touch of tensor / parameter memory is below
    LSQ_FTRACE(1);
    LazyChan<MODE> lch;
    lch.issue(sg, 0);
    if (blockIdx.x == 0) {                        // the < VEC elements behind the last whole unit
        const T* xp = reinterpret_cast<const T*>(sg.x);
        T* yp = reinterpret_cast<T*>(sg.y);
        for (long long e = units * VEC + threadIdx.x; e < sg.inner; e += THREADS)
            yp[e] = INIT ? xp[e] : Tr::from_f(fq_forward<MODE>(Tr::to_f(xp[e]), lch.get(sg)));
    }
    for (; u < units; u += stride) {
        const Raw<NW> xr = ld_unit<LD, NW>(px);
        if (INIT) st_unit<ST, NW>(py, xr);
        else {
            const Chan& ch = lch.get(sg);
            float f[VEC];
            unpack_unit<T, NW>(xr, f);
#pragma unroll
            for (int e = 0; e < VEC; e++) f[e] = fq_forward<MODE>(f[e], ch);
            st_unit<ST, NW>(py, pack_unit<T, NW>(f));
        }
        px += sb; py += sb;
    }
    LSQ_FTRACE(2);
}

template <typename T, int MODE, int BMODE, int THREADS, int LD, int ST, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_flatbwd_kernel(const __grid_constant__ Seg sg) {
    using Tr = ElemTraits<T>;
    constexpr int NW = 8, VEC = UnitOf<T, NW>::VEC, UB = 32;
    __shared__ double red[64];
    __shared__ int last_flag;
    pdl_trigger();
    LSQ_FTRACE(0);
    const long long units = sg.inner / VEC;
    const long long stride = (long long)gridDim.x * THREADS;
    long long u = (long long)blockIdx.x * THREADS + threadIdx.x;
    const char* px = reinterpret_cast<const char*>(sg.x) + u * UB;
    const char* pg = reinterpret_cast<const char*>(sg.g) + u * UB;
    char* pgx = sg.gx ? reinterpret_cast<char*>(sg.gx) + u * UB : nullptr;
    const long long sb = stride * UB;
    if (sg.flags && u < units) { l2_prefetch(px); l2_prefetch(pg); }   // one unit per operand, see lsq_flatfwd_kernel
    pdl_wait();                                   // first touch of tensor / parameter memory is below
    LSQ_FTRACE(1);
    LazyChan<MODE> lch;
    lch.issue(sg, 0);
    double accS = 0.0, accB = 0.0;
    if (blockIdx.x == 0) {                        // the < VEC elements behind the last whole unit (exact terms, as the general kernel's peel)
        const T* xp = reinterpret_cast<const T*>(sg.x);
        const T* gp = reinterpret_cast<const T*>(sg.g);
        T* gxp = reinterpret_cast<T*>(sg.gx);
        for (long long e = units * VEC + threadIdx.x; e < sg.inner; e += THREADS) {
            const float dx = fq_backward<MODE, BMODE, true>(Tr::to_f(gp[e]), Tr::to_f(xp[e]), lch.get(sg), accS, accB);
            if (gxp) gxp[e] = bmode_passthrough(BMODE) ? gp[e] : Tr::from_f(dx);
        }
    }
    for (; u < units; u += stride) {
        const Raw<NW> xr = ld_unit<LD, NW>(px);
        const Raw<NW> gr = ld_unit<LD, NW>(pg);
        const Chan& ch = lch.get(sg);
        float fx[VEC], fg[VEC];
        unpack_unit<T, NW>(xr, fx);
        unpack_unit<T, NW>(gr, fg);
        float ls = 0.f, lb = 0.f;                 // fp32 partial over one unit's <= 16 terms, then promoted (as lsq_bwd_kernel)
#pragma unroll
        for (int e = 0; e < VEC; e++) fg[e] = fq_backward<MODE, BMODE, false>(fg[e], fx[e], ch, ls, lb);
        if (pgx) {
            if (bmode_passthrough(BMODE)) st_unit<ST, NW>(pgx, gr);
            else st_unit<ST, NW>(pgx, pack_unit<T, NW>(fg));
            pgx += sb;
        }
        accS += (double)ls; accB += (double)lb;
        px += sb; pg += sb;
    }
    LSQ_FTRACE(2);
    if (!bmode_reduces(BMODE)) {                  // eval: parameters get exact zeros (lsq_kernel.h:143-144)
        if (blockIdx.x == 0 && threadIdx.x == 0) { store_param(sg.gscale, 0, sg.pdt, 0.0); store_param(sg.gshift, 0, sg.pdt, 0.0); }
        return;
    }
    TileCtx tl;
    tl.c = 0; tl.ltile = blockIdx.x; tl.j = (int)blockIdx.x; tl.pidx = 0;
    const bool fin_ = channel_finish<THREADS, THREADS>(sg, tl, accS, accB, red, &last_flag, threadIdx.x);
    LSQ_FTRACE(3);
    if (!fin_) return;
    if (threadIdx.x == 0) {
        store_param(sg.gscale, 0, sg.pdt, accS * sg.gs);
        store_param(sg.gshift, 0, sg.pdt, sg.sym ? 0.0 : accB * sg.gs);
    }
}

// ---------------------------------------------------------------------------------------------
// Observer step (init_mode='observer', observers.py:446-449): ONE read of x gives the per-tensor /
// per-channel min and max; the last tile of a channel then does, on the device and with the same
// fp32 operations in the same order, what torch's observers and the module do on the host:
//   MinMaxObserver / MovingAverage(PerChannel)MinMaxObserver.forward   (running extrema or EMA)
//   UniformQuantizationObserverBase._calculate_qparams                 (scale, zero_point)
//   LSQFakeQuantizer._set_weights                                      (scale, shift = -zp * scale)
// NaN propagates like torch.aminmax (min.NaN / max.NaN).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float nan_min(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float nan_max(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <typename T, int NW, int G, int THREADS, int UNROLL, int LD, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_observe_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                long long total_tiles) {
    using Tr = ElemTraits<T>;
    constexpr int VEC = UnitOf<T, NW>::VEC;
    constexpr int UB = NW * 4;   // unit bytes
    constexpr int GROUPS = THREADS / G;
    __shared__ Seg smem_seg[GROUPS];
    __shared__ float red[64];
    __shared__ int last_flag[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    pdl_trigger();
    int staged = -2;
    const long long gtile = (long long)blockIdx.x * GROUPS + grp;
    if (gtile >= total_tiles) return;
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);   // descriptors are launch constants
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    const T* __restrict__ xp = reinterpret_cast<const T*>(sg.x);
    if constexpr (VEC > 1) {
        if (sg.flags) {                           // first units on their way into L2 while the predecessor drains
            Walker pw{};
            pw.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
            for (int d = 0; d < sg.flags && pw.more(); d++) {
                long long a0;
                pw.next(a0);
                l2_prefetch(reinterpret_cast<const char*>(xp) + a0 * UB);
            }
        }
    }
    pdl_wait();
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            const float v = Tr::to_f(xp[e]);
            mn = nan_min(mn, v); mx = nan_max(mx, v);
        }
    }
    Walker w{};
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        if constexpr (VEC > 1) {
            Raw<NW> xr[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + (ok[k] ? addr[k] : addr[0]) * UB);
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                float f[VEC];
                unpack_unit<T, NW>(xr[k], f);          // re-read unit 0 on tail lanes is harmless for min / max
#pragma unroll
                for (int e = 0; e < VEC; e++) { mn = nan_min(mn, f[e]); mx = nan_max(mx, f[e]); }
            }
        } else {
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                const float v = Tr::to_f(xp[addr[k]]);
                mn = nan_min(mn, v); mx = nan_max(mx, v);
            }
        }
    }
    // group reduction (min / max are order independent: deterministic by construction)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (G != 32) {
        const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
        constexpr int NWARP = THREADS / 32;
        if (lane == 0) { red[wi] = mn; red[32 + wi] = mx; }
        __syncthreads();
        if (wi == 0) {
            mn = lane < NWARP ? red[lane] : __int_as_float(0x7f800000);
            mx = lane < NWARP ? red[32 + lane] : __int_as_float(0xff800000);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
        }
        __syncthreads();
    }
    if (sg.splits > 1) {
        if (tg == 0) {
            sg.partials[2 * tl.ltile] = (double)mn;
            sg.partials[2 * tl.ltile + 1] = (double)mx;
            fence_acq_rel_gpu();
            const unsigned prev = atomicAdd(&sg.counters[tl.c], 1u);
            last_flag[grp] = (prev == (unsigned)sg.splits - 1u);
        }
        group_sync<G, THREADS>();
        if (!last_flag[grp]) return;
        fence_acq_rel_gpu();
        mn = __int_as_float(0x7f800000); mx = __int_as_float(0xff800000);
        const double* p = sg.partials + 2 * (tl.c * sg.splits);
        for (int i = tg; i < sg.splits; i += G) { mn = nan_min(mn, (float)__ldcg(p + 2 * i)); mx = nan_max(mx, (float)__ldcg(p + 2 * i + 1)); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (G != 32) {
            const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
            constexpr int NWARP = THREADS / 32;
            if (lane == 0) { red[wi] = mn; red[32 + wi] = mx; }
            __syncthreads();
            if (wi == 0) {
                mn = lane < NWARP ? red[lane] : __int_as_float(0x7f800000);
                mx = lane < NWARP ? red[32 + lane] : __int_as_float(0xff800000);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
            }
        }
        if (tg == 0) sg.counters[tl.c] = 0u;
    }
    if (tg != 0) return;
    // ---- observer update (torch/ao/quantization/observer.py: MinMaxObserver.forward, MovingAverageMinMaxObserver.forward)
    float* min_state = reinterpret_cast<float*>(sg.gscale);
    float* max_state = reinterpret_cast<float*>(sg.gshift);
    float smin = min_state[tl.pidx], smax = max_state[tl.pidx];
    const bool empty = (smin == __int_as_float(0x7f800000)) && (smax == __int_as_float(0xff800000));
    if (sg.obs_flags & 2) {
        if (empty) { smin = mn; smax = mx; }
        else {
            smin = __fadd_rn(smin, __fmul_rn(sg.obs_c, __fsub_rn(mn, smin)));
            smax = __fadd_rn(smax, __fmul_rn(sg.obs_c, __fsub_rn(mx, smax)));
        }
    } else {
        smin = nan_min(mn, smin); smax = nan_max(mx, smax);
    }
    min_state[tl.pidx] = smin; max_state[tl.pidx] = smax;
    // ---- qparams (UniformQuantizationObserverBase._calculate_qparams) and LSQ parameters (observers.py:346-373)
    float* scale_out = reinterpret_cast<float*>(sg.y);
    float* shift_out = reinterpret_cast<float*>(sg.gx);
    if (scale_out == nullptr) return;
    const float min_neg = nan_min(smin, 0.0f);
    float max_pos = nan_max(smax, 0.0f);
    // torch divides a CUDA tensor by a Python scalar as a multiplication by the fp32 reciprocal
    // (ATen BinaryDivTrueKernel.cu: inv_b = 1 / b; a * inv_b); tensor / tensor is a true division
    const float range = __fsub_rn(sg.qmax, sg.qmin);            // float(quant_max - quant_min), exact for 8-bit ranges
    float scale, zp;
    if (sg.obs_flags & 1) {
        max_pos = nan_max(-min_neg, max_pos);
        scale = nan_max(__fmul_rn(max_pos, __fdiv_rn(1.0f, __fmul_rn(range, 0.5f))), sg.obs_eps);
        zp = (float)sg.obs_zp_sym;
    } else {
        scale = nan_max(__fmul_rn(__fsub_rn(max_pos, min_neg), __fdiv_rn(1.0f, range)), sg.obs_eps);
        zp = __fsub_rn(sg.qmin, rintf(__fdiv_rn(min_neg, scale)));
        zp = fminf(fmaxf(zp, sg.qmin), sg.qmax);
    }
    scale_out[tl.pidx] = scale;
    if (shift_out != nullptr) shift_out[tl.pidx] = __fmul_rn(-zp, scale);
}

}  // namespace lsqb200
