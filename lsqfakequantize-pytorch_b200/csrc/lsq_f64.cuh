// lsq_f64.cuh -- float64 tensors (the reference dispatches double on CPU and CUDA:
// AT_DISPATCH_FLOATING_TYPES_AND_HALF, csrc/ops/cuda/lsq_cuda.cu:45,113,186,266).
//
// Arithmetic contract = the reference CUDA build's double instantiation of
// csrc/ops/kernels/lsq_kernel.h, read from its sm_100a SASS (the scalar templates compiled as they lie):
//   * FMIN / FMAX are ::fminf / ::fmaxf for EVERY dtype (csrc/ops/global_scope.h:51-52), so each
//     clamp operand is first rounded to float (F2F.F32.F64): v = fma(x, 1/s, zp) is a double, the
//     clamped value is a float; the zero point is rounded from a float; the per-channel scale
//     fmaxf(eps, |s|) (lsq_kernel.h:157) is the scale ROUNDED TO FLOAT;
//   * x*inv_s + zp and (r - zp)*s - x are single DFMAs; 2*(xfq - x) is d + d;
//   * dS = (g'*d)*inv_s (two DMULs) or g'*(q - zp); dB = (double)(!mask) * g'.
// Sums are plain double sums of those terms (the reference sums term*grad_scaler with at::sum;
// order unpinned), multiplied by gs once.
//
// Same tiling as lsq_device.cuh: (channel, split) tiles owned by CTA or warp groups, 256-bit
// units (4 doubles), fixed-order last-arriver reduction.
#pragma once
#include "lsq_device.cuh"

namespace lsqb200 {

struct ChanD {
    double s, inv_s, zp;
    double c_lo, c_hi;   // qmin - zp, qmax - zp
    float qmin, qmax;
};

__device__ __forceinline__ ChanD make_chan_d(double scale, double shift, const Seg& sg) {
    ChanD c;
    const double a = fabs(scale);
    const double eps = 2.220446049250313e-16;                       // numeric_limits<double>::epsilon = 2^-52
    if (sg.per_channel) c.s = (double)fmaxf((float)eps, __double2float_rn(a));   // ::fmaxf on doubles: float-rounded scale
    else c.s = (a < eps) ? eps : a;                                 // std::max(|s|, eps) on the host (lsq_cuda.cu:53-54)
    c.inv_s = __ddiv_rn(1.0, c.s);
    const float t = __double2float_rn(__dmul_rn(-shift, c.inv_s));
    c.zp = (double)rintf(fminf(sg.tmax, fmaxf(sg.tmin, t)));
    c.qmin = sg.qmin; c.qmax = sg.qmax;
    c.c_lo = __dsub_rn((double)sg.qmin, c.zp);
    c.c_hi = __dsub_rn((double)sg.qmax, c.zp);
    return c;
}

__device__ __forceinline__ double fq_forward_d(double x, const ChanD& c) {
    const float vf = __double2float_rn(__fma_rn(x, c.inv_s, c.zp));
    const float r = rintf(fminf(c.qmax, fmaxf(c.qmin, vf)));       // max first: NaN -> qmin
    return __dmul_rn(__dsub_rn((double)r, c.zp), c.s);
}

template <int BMODE>
__device__ __forceinline__ double fq_backward_d(double g, double x, const ChanD& c, double& accS, double& accB) {
    const float vf = __double2float_rn(__fma_rn(x, c.inv_s, c.zp));
    const float xq = fmaxf(fminf(vf, c.qmax), c.qmin);             // min first: NaN -> qmax
    const bool mask = (c.qmin < xq) && (xq < c.qmax);
    const double m = mask ? 1.0 : 0.0;
    const double dX = bmode_passthrough(BMODE) ? g : __dmul_rn(g, m);
    if (bmode_reduces(BMODE)) {
        const double t = __dsub_rn((double)rintf(xq), c.zp);
        const double d = __fma_rn(t, c.s, -x);                     // xfq - x, one DFMA as in the reference build
        const double gg = (BMODE == B_INIT) ? __dadd_rn(d, d) : g;
        const double dS = mask ? __dmul_rn(__dmul_rn(gg, d), c.inv_s) : __dmul_rn(gg, (xq <= c.qmin) ? c.c_lo : c.c_hi);
        accS = __dadd_rn(accS, dS);
        accB = __dadd_rn(accB, __dmul_rn(gg, __dsub_rn(1.0, m)));
    }
    return dX;
}

template <int NW>
__device__ __forceinline__ void unpack_d(const Raw<NW>& r, double* d) {
#pragma unroll
    for (int i = 0; i < NW / 2; i++) d[i] = __hiloint2double((int)r.w[2 * i + 1], (int)r.w[2 * i]);
}
template <int NW>
__device__ __forceinline__ Raw<NW> pack_d(const double* d) {
    Raw<NW> r;
#pragma unroll
    for (int i = 0; i < NW / 2; i++) { r.w[2 * i] = (uint32_t)__double2loint(d[i]); r.w[2 * i + 1] = (uint32_t)__double2hiint(d[i]); }
    return r;
}

// float64 tensors carry float64 parameters (the reference requires scale.dtype == x.dtype, lsq_cuda.cu:34-35)
struct LazyChanD {
    double sraw, braw;
    ChanD c;
    bool ready;
    __device__ __forceinline__ void issue(const Seg& sg, long long pidx) {
        sraw = reinterpret_cast<const double*>(sg.scale)[pidx];
        braw = reinterpret_cast<const double*>(sg.shift)[pidx];
        ready = false;
    }
    __device__ __forceinline__ const ChanD& get(const Seg& sg) {
        if (!ready) { c = make_chan_d(sraw, braw, sg); ready = true; }
        return c;
    }
};

// NW in {8, 4, 2}: 4 / 2 / 1 doubles per unit (a double is always 8-byte aligned, so there is no scalar path)
template <int NW, bool INIT, int G, int THREADS, int UNROLL, int LD, int ST, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_fwd_f64_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                   long long total_tiles) {
    constexpr int VEC = NW / 2;
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
    const double* __restrict__ xp = reinterpret_cast<const double*>(sg.x);
    double* __restrict__ yp = reinterpret_cast<double*>(sg.y);
    Walker w;
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    pdl_wait();
    LazyChanD lch;
    lch.issue(sg, tl.pidx);
    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            yp[e] = INIT ? xp[e] : fq_forward_d(xp[e], lch.get(sg));
        }
    }
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        Raw<NW> xr[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++)
            xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + (ok[k] ? addr[k] : addr[0]) * UB);
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (!ok[k]) continue;
            if (INIT) { st_unit<ST, NW>(reinterpret_cast<char*>(yp) + addr[k] * UB, xr[k]); continue; }
            const ChanD& ch = lch.get(sg);
            double f[VEC];
            unpack_d<NW>(xr[k], f);
#pragma unroll
            for (int e = 0; e < VEC; e++) f[e] = fq_forward_d(f[e], ch);
            st_unit<ST, NW>(reinterpret_cast<char*>(yp) + addr[k] * UB, pack_d<NW>(f));
        }
    }
}

template <int NW, int BMODE, int G, int THREADS, int UNROLL, int LD, int ST, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
lsq_bwd_f64_kernel(const __grid_constant__ Seg single, const Seg* __restrict__ table, const int* __restrict__ tile_seg, int nseg,
                   long long total_tiles) {
    constexpr int VEC = NW / 2;
    constexpr int UB = NW * 4;
    constexpr int GROUPS = THREADS / G;
    __shared__ Seg smem_seg[GROUPS];
    __shared__ double red[64];
    __shared__ int last_flag[GROUPS];
    const int grp = threadIdx.x / G, tg = threadIdx.x % G;
    pdl_trigger();
    int staged = -2;
    const long long gtile = (long long)blockIdx.x * GROUPS + grp;
    if (gtile >= total_tiles) return;
    stage_segment<G, THREADS>(single, table, tile_seg, nseg, gtile, &smem_seg[grp], staged);
    const Seg& sg = smem_seg[grp];
    const TileCtx tl = make_tile<VEC>(sg, gtile);
    const double* __restrict__ xp = reinterpret_cast<const double*>(sg.x);
    const double* __restrict__ gp = reinterpret_cast<const double*>(sg.g);
    double* __restrict__ gxp = reinterpret_cast<double*>(sg.gx);
    const bool write_gx = gxp != nullptr;
    Walker w;
    w.init<G>(sg, tl.u0, tl.u1, tl.base_unit, tg);
    pdl_wait();
    LazyChanD lch;
    lch.issue(sg, tl.pidx);

    double accS = 0.0, accB = 0.0;
    if constexpr (VEC > 1) {
        for (long long i = tg; i < tl.peel_n0 + tl.peel_n1; i += G) {
            const long long e = i < tl.peel_n0 ? tl.peel_begin0 + i : tl.peel_begin1 + (i - tl.peel_n0);
            const double dx = fq_backward_d<BMODE>(gp[e], xp[e], lch.get(sg), accS, accB);
            if (write_gx) gxp[e] = dx;
        }
    }
    while (w.more()) {
        long long addr[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) ok[k] = w.next(addr[k]);
        Raw<NW> xr[UNROLL], gr[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const long long a = ok[k] ? addr[k] : addr[0];
            xr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(xp) + a * UB);
            gr[k] = ld_unit<LD, NW>(reinterpret_cast<const char*>(gp) + a * UB);
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (!ok[k]) continue;
            const ChanD& ch = lch.get(sg);
            double fx[VEC], fg[VEC];
            unpack_d<NW>(xr[k], fx);
            unpack_d<NW>(gr[k], fg);
#pragma unroll
            for (int e = 0; e < VEC; e++) fg[e] = fq_backward_d<BMODE>(fg[e], fx[e], ch, accS, accB);
            if (write_gx) {
                if (bmode_passthrough(BMODE)) st_unit<ST, NW>(reinterpret_cast<char*>(gxp) + addr[k] * UB, gr[k]);
                else st_unit<ST, NW>(reinterpret_cast<char*>(gxp) + addr[k] * UB, pack_d<NW>(fg));
            }
        }
    }
    if (!bmode_reduces(BMODE)) {   // eval: parameters get exact zeros (lsq_kernel.h:143-144)
        if (tl.j == 0 && tg == 0) {
            reinterpret_cast<double*>(sg.gscale)[tl.pidx] = 0.0;
            reinterpret_cast<double*>(sg.gshift)[tl.pidx] = 0.0;
        }
        return;
    }
    if (!channel_finish<G, THREADS>(sg, tl, accS, accB, red, &last_flag[grp], tg)) return;
    if (tg == 0) {
        reinterpret_cast<double*>(sg.gscale)[tl.pidx] = accS * sg.gs;
        reinterpret_cast<double*>(sg.gshift)[tl.pidx] = sg.sym ? 0.0 : accB * sg.gs;
    }
}

}  // namespace lsqb200
