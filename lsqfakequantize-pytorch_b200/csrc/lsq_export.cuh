// lsq_export.cuh -- the step after QAT: turn the learned LSQ+ parameters into
// real integer tensors on the device.
//
//   lsq_export_kernel<DIR = 0>   x (f32 / f16 / bf16)  ->  codes (uint8 / int8): R x + W 1 byte per element
//   lsq_export_kernel<DIR = 1>   codes -> y = (code - zp) * s                   : R 1 byte + W y
//   lsq_qparams_kernel           (scale, shift) -> (max(scale, eps), int64 zero_point), on the device
//
// Two code semantics (file:line under /root/reference/torchlsq/):
//   SEM_LSQ    the integer the training forward already forms, csrc/ops/kernels/lsq_kernel.h:12-13:
//              code = rint(clamp(fma(x, 1/s, zp), quant_min, quant_max)),  s = max(|scale|, eps),
//              zp = rint(clamp(-shift/s, type_min, type_max)).  dequantize(code) is bit-identical to lsq(x).
//   SEM_TORCH  what `torch.quantization.convert` does with `LSQFakeQuantizer.calculate_qparams()`
//              (quantized/modules/observers.py:378-422): scale = max(scale, eps) (no abs),
//              zero_point = clamp(round(-shift / scale), type range), and then torch's CUDA
//              quantize_per_tensor / quantize_per_channel:
//              code = clamp(int64(nearbyint(double(x) / double(scale)) + zero_point), type range).
//              (SEM_TORCH_CPU: torch's CPU quantizers instead, clamp(nearbyint(x * (1.0f/scale)) + zero_point).)
//              The double-precision quotient is reproduced with fp32 operations only: q = RN32(x / s) lies on
//              the same side of every half-integer as the exact quotient unless q IS a half-integer, and then
//              the sign of the exact remainder fma(-q, s, x) decides (0 = a true tie: half-even, as in double).
#pragma once
#include "lsq_device.cuh"

namespace lsqb200 {

enum : int { SEM_LSQ = 0, SEM_TORCH = 1 /* torch CUDA ops */, SEM_TORCH_CPU = 2 /* torch CPU (fbgemm) ops */ };
enum : int { DIR_QUANT = 0, DIR_DEQUANT = 1 };

template <int MODE, int SEM>
__device__ __forceinline__ Chan make_chan_export(float scale, float shift, const Seg& sg) {
    if (SEM == SEM_LSQ) return make_chan<MODE>(scale, shift, sg);
    Chan c;
    c.s = nan_max(scale, 1.1920928955078125e-07f);             // torch.max(scale, eps), observers.py:416
    c.inv_s = __fdiv_rn(1.0f, c.s);                             // torch's CPU quantizers multiply by the fp32 reciprocal
    const float z = rintf(__fdiv_rn(-shift, c.s));              // (-shift / scale).round_()   observers.py:399-400
    c.zp = fminf(fmaxf(z, sg.tmin), sg.tmax);                   // .clamp_(tmin, tmax)
    c.qmin = sg.tmin; c.qmax = sg.tmax;                         // torch clamps codes to the integer type's range
    c.c_lo = c.c_hi = 0.0f;
    return c;
}

template <int MODE, int SEM>
__device__ __forceinline__ int quant_code(float x, const Chan& c, bool perch) {
    if (SEM == SEM_LSQ) {
        const float v = affine_v<MODE>(x, c);
        return __float2int_rn(fminf(c.qmax, fmaxf(c.qmin, v)));          // max first: NaN -> quant_min, as the forward
    }
    if (SEM == SEM_TORCH_CPU) {
        // torch 2.11 CPU quantizers on x * (1.0f / scale).  Finite x: clamp(zero_point + nearbyint(p), type range).
        // Non-finite / huge x follow what the two torch code paths do (pinned by tests/golden/ref_export.npz):
        const float p = __fmul_rn(x, c.inv_s);
        if (perch) {
            // quantize_per_channel: scalar quantize_val, clamp in float; NaN ends as integer 0 (cast of NaN)
            const float t = __fadd_rn(c.zp, rintf(p));
            return (t != t) ? 0 : __float2int_rn(fminf(fmaxf(t, c.qmin), c.qmax));
        }
        // quantize_per_tensor: vector path - min_ps(p, 2^31 - 128) (NaN -> the bound), cvtps_epi32, int32 wrap-around add
        const float pc = (p != p) ? 2147483520.0f : fminf(p, 2147483520.0f);
        const int q = (int)((unsigned)__float2int_rn(pc) + (unsigned)(int)c.zp);
        return min(max(q, (int)c.qmin), (int)c.qmax);
    }
    const float f = __fdiv_rn(x, c.s);
    float r = rintf(f);
    if (fabsf(__fsub_rn(f, r)) == 0.5f) {                                // fp32 quotient landed on a half-integer
        const float rem = __fmaf_rn(-f, c.s, x);                         // exact: x - f*s
        if (rem > 0.0f) r = ceilf(f); else if (rem < 0.0f) r = floorf(f);
    }
    r = fminf(fmaxf(r, -1048576.0f), 1048576.0f);                        // int64 saturation is far outside the clamp
    const float q = (f != f) ? 0.0f : __fadd_rn(r, c.zp);               // cvt of NaN gives 0 on the device
    return __float2int_rn(fminf(fmaxf(q, c.qmin), c.qmax));
}

__device__ __forceinline__ float dequant_code(int code, const Chan& c) {
    return __fmul_rn(__fsub_rn((float)code, c.zp), c.s);                 // lsq_kernel.h:13 / torch dequantize
}

// VEC codes (bytes) moved by one instruction
template <int VEC> struct Codes { uint32_t w[(VEC + 3) / 4]; };
template <int VEC>
__device__ __forceinline__ Codes<VEC> ld_codes(const uint8_t* p) {
    Codes<VEC> c;
    if constexpr (VEC == 16) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w; }
    else if constexpr (VEC == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); c.w[0] = v.x; c.w[1] = v.y; }
    else if constexpr (VEC == 4) { c.w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    else if constexpr (VEC == 2) { c.w[0] = __ldg(reinterpret_cast<const uint16_t*>(p)); }
    else { c.w[0] = __ldg(p); }
    return c;
}
template <int VEC>
__device__ __forceinline__ void st_codes(uint8_t* p, const Codes<VEC>& c) {
    if constexpr (VEC == 16) *reinterpret_cast<uint4*>(p) = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
    else if constexpr (VEC == 8) *reinterpret_cast<uint2*>(p) = make_uint2(c.w[0], c.w[1]);
    else if constexpr (VEC == 4) *reinterpret_cast<uint32_t*>(p) = c.w[0];
    else if constexpr (VEC == 2) *reinterpret_cast<uint16_t*>(p) = (uint16_t)c.w[0];
    else *p = (uint8_t)c.w[0];
}
template <int VEC>
__device__ __forceinline__ int code_at(const Codes<VEC>& c, int i, bool is_signed) {
    const uint32_t b = (c.w[i >> 2] >> (8 * (i & 3))) & 0xffu;
    return is_signed ? (int)(int8_t)b : (int)b;
}

// x is always the floating tensor and y the code tensor in the Seg (whatever the direction); tiles, walkers
// and units are those of the forward kernel, a unit being NW words of the FLOATING tensor.
template <typename T, int MODE, int NW, int SEM, int DIR, int G, int THREADS, int UNROLL, int LD, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_export_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                  long long total_tiles) {
    using Tr = ElemTraits<T>;
    constexpr int VEC = UnitOf<T, NW>::VEC;
    constexpr int UB = NW * 4;
    constexpr int GROUPS = THREADS / G;
    __shared__ Seg smem_seg[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    pdl_trigger();
    int staged = -2;
    const long long gtile = (long long)blockIdx.x * GROUPS + grp;
    if (gtile >= total_tiles) return;
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    T* __restrict__ fp = reinterpret_cast<T*>(const_cast<void*>(sg.x));
    uint8_t* __restrict__ cp = reinterpret_cast<uint8_t*>(sg.y);
    const bool is_signed = sg.code_signed != 0;
    const bool perch = sg.per_channel != 0;
    Walker w;
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    pdl_wait();
    const Chan ch = make_chan_export<MODE, SEM>(load_param(sg.scale, tl.pidx, sg.pdt), load_param(sg.shift, tl.pidx, sg.pdt), sg);

    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            if (DIR == DIR_QUANT) cp[e] = (uint8_t)quant_code<MODE, SEM>(Tr::to_f(fp[e]), ch, perch);
            else fp[e] = Tr::from_f(dequant_code(is_signed ? (int)(int8_t)cp[e] : (int)cp[e], ch));
        }
    }
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        if constexpr (VEC > 1) {
            if (DIR == DIR_QUANT) {
                Raw<NW> xr[UNROLL];
#pragma unroll
                for (int k = 0; k < UNROLL; k++)
                    xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(fp) + (ok[k] ? addr[k] : addr[0]) * UB);
#pragma unroll
                for (int k = 0; k < UNROLL; k++) {
                    if (!ok[k]) continue;
                    float f[VEC];
                    unpack_unit<T, NW>(xr[k], f);
                    Codes<VEC> c;
#pragma unroll
                    for (int i = 0; i < (VEC + 3) / 4; i++) c.w[i] = 0u;
#pragma unroll
                    for (int e = 0; e < VEC; e++)
                        c.w[e >> 2] |= ((uint32_t)quant_code<MODE, SEM>(f[e], ch, perch) & 0xffu) << (8 * (e & 3));
                    st_codes<VEC>(cp + addr[k] * VEC, c);
                }
            } else {
                Codes<VEC> cr[UNROLL];
#pragma unroll
                for (int k = 0; k < UNROLL; k++) cr[k] = ld_codes<VEC>(cp + (ok[k] ? addr[k] : addr[0]) * VEC);
#pragma unroll
                for (int k = 0; k < UNROLL; k++) {
                    if (!ok[k]) continue;
                    float f[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; e++) f[e] = dequant_code(code_at<VEC>(cr[k], e, is_signed), ch);
                    st_unit<ST_DEFAULT, NW>(reinterpret_cast<char*>(fp) + addr[k] * UB, pack_unit<T, NW>(f));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                if (!ok[k]) continue;
                if (DIR == DIR_QUANT) cp[addr[k]] = (uint8_t)quant_code<MODE, SEM>(Tr::to_f(fp[addr[k]]), ch, perch);
                else fp[addr[k]] = Tr::from_f(dequant_code(is_signed ? (int)(int8_t)cp[addr[k]] : (int)cp[addr[k]], ch));
            }
        }
    }
}

// calculate_qparams on the device (observers.py:403-422 + :378-401): no .cpu() round trip, no host sync.
static __global__ void lsq_qparams_kernel(const void* __restrict__ scale, const void* __restrict__ shift, float* __restrict__ scale_out,
                                   long long* __restrict__ zp_out, long long n, int pdt, float tmin, float tmax) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = nan_max(load_param(scale, i, pdt), 1.1920928955078125e-07f);
    const float z = fminf(fmaxf(rintf(__fdiv_rn(-load_param(shift, i, pdt), s)), tmin), tmax);
    scale_out[i] = s;
    if (zp_out != nullptr) zp_out[i] = (long long)z;
}

}  // namespace lsqb200
