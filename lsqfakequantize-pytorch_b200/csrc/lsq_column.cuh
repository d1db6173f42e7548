// lsq_column.cuh -- per-channel kernels for SHORT channel rows: channels-last (inner == 1), 7x7 /
// 14x14 NCHW maps, Linear activations.  There a 16-byte unit spans several channels, so the
// row-tiled kernels of lsq_device.cuh would fall back to element-at-a-time accesses.
//
// View the tensor as a matrix: `outer` rows of L = C*inner contiguous elements.  A thread owns
// ONE 16-byte column unit (VEC elements = fixed channels for the whole kernel) and walks down
// the rows, so
//   * every access is a coalesced 128-bit load / store (consecutive threads, consecutive units),
//   * the per-channel constants of its VEC element slots live in registers (no per-element
//     channel arithmetic, no parameter reloads),
//   * grad_scale / grad_shift partial sums stay in per-slot fp32 registers, are promoted every 32
//     rows into the thread's private fp64 cells in shared memory, and leave the CTA once, as one
//     fp64 atomic pair per channel run into a [C][2] accumulator; the last CTA (ticket) turns
//     the accumulator into the outputs and leaves it zeroed.  CTAs are sized to whole waves.
// Arithmetic is the same fq_forward / fq_backward as everywhere else.
#pragma once
#include "lsq_device.cuh"

namespace lsqb200 {

struct ColSeg {
    const void* x;
    void* y;
    const void* g;
    void* gx;
    const void* scale;
    const void* shift;
    void* gscale;
    void* gshift;
    double* acc;          // [C][2] zero on entry, zero on exit
    unsigned* counter;    // one ticket for the whole launch
    long long outer, C, inner;
    long long units_per_row;   // L / VEC
    long long rows_per_split;
    double gs;
    float qmin, qmax, tmin, tmax;
    int tx, ty;                // logical CTA shape: tx column units x ty rows, tx * ty == THREADS
    int pdt, sym, per_channel_unused;
    unsigned total_ctas;
};

constexpr int kColThreads = 256;

// CNW: 32-bit words per column unit (4 = 128-bit, 2 = 64-bit accesses); fewer slots per thread
// mean fewer registers (more resident CTAs) at the price of narrower accesses
template <typename T, int CNW>
struct ColVec { static constexpr int NW = CNW; static constexpr int VEC = UnitOf<T, CNW>::VEC; };

// per-slot channel constants
template <typename T, int MODE, int CNW>
struct SlotParams {
    static constexpr int VEC = ColVec<T, CNW>::VEC;
    float s[VEC], inv_s[VEC], zp[VEC];
    int ch[VEC];
    __device__ __forceinline__ void load(const ColSeg& cs, long long unit_col) {
        Seg fake;                       // make_chan only reads these fields
        fake.per_channel = 1; fake.tmin = cs.tmin; fake.tmax = cs.tmax; fake.qmin = cs.qmin; fake.qmax = cs.qmax;
        // one 32-bit division for the first slot, then walk: L = C*inner < 2^31 is checked on the host
        const unsigned j0 = (unsigned)unit_col * VEC, inner = (unsigned)cs.inner;
        unsigned c = j0 / inner, r = j0 - c * inner;
        float sraw[VEC], braw[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) {          // issue every parameter load before the first use
            ch[k] = (int)c;
            sraw[k] = load_param(cs.scale, c, cs.pdt);
            braw[k] = load_param(cs.shift, c, cs.pdt);
            if (++r == inner) { r = 0; ++c; }
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const Chan cc = make_chan<MODE>(sraw[k], braw[k], fake);
            s[k] = cc.s; inv_s[k] = cc.inv_s; zp[k] = cc.zp;
        }
    }
    __device__ __forceinline__ Chan chan(int k, const ColSeg& cs) const {
        Chan c;
        c.s = s[k]; c.inv_s = inv_s[k]; c.zp = zp[k]; c.qmin = cs.qmin; c.qmax = cs.qmax;
        c.c_lo = 0.f; c.c_hi = 0.f;
        return c;
    }
};

template <typename T, int MODE, bool INIT, int CNW, int kColUnroll, int MINB, int LD, int ST>
__global__ void __launch_bounds__(kColThreads, MINB)
lsq_col_fwd_kernel(const __grid_constant__ ColSeg cs) {
    constexpr int NW = CNW, VEC = ColVec<T, CNW>::VEC, UB = CNW * 4;
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tx = threadIdx.x % cs.tx, ty = threadIdx.x / cs.tx;
    const long long uc = (long long)blockIdx.x * cs.tx + tx;
    if (uc >= cs.units_per_row) return;
    const char* __restrict__ xp = reinterpret_cast<const char*>(cs.x);
    char* __restrict__ yp = reinterpret_cast<char*>(cs.y);
    SlotParams<T, MODE, CNW> sp;
    if (!INIT) sp.load(cs, uc);
    long long n = (long long)blockIdx.y * cs.rows_per_split + ty;
    long long n_end = ((long long)blockIdx.y + 1) * cs.rows_per_split;
    if (n_end > cs.outer) n_end = cs.outer;
    for (; n < n_end; n += (long long)cs.ty * kColUnroll) {
        Raw<NW> xr[kColUnroll];
        long long a[kColUnroll];
#pragma unroll
        for (int r = 0; r < kColUnroll; r++) {
            const long long nn = n + (long long)r * cs.ty;
            a[r] = (nn < n_end ? nn : n) * cs.units_per_row + uc;
            xr[r] = ld_unit<LD, NW>(xp + a[r] * UB);
        }
#pragma unroll
        for (int r = 0; r < kColUnroll; r++) {
            if (n + (long long)r * cs.ty >= n_end) continue;
            if (INIT) { st_unit<ST, NW>(yp + a[r] * UB, xr[r]); continue; }
            float f[VEC];
            unpack_unit<T, NW>(xr[r], f);
#pragma unroll
            for (int k = 0; k < VEC; k++) f[k] = fq_forward<MODE>(f[k], sp.chan(k, cs));
            st_unit<ST, NW>(yp + a[r] * UB, pack_unit<T, NW>(f));
        }
    }
}

template <typename T, int MODE, int BMODE, int CNW, int kColUnroll, int MINB, int LD, int ST>
__global__ void __launch_bounds__(kColThreads, MINB)
lsq_col_bwd_kernel(const __grid_constant__ ColSeg cs) {
    constexpr int NW = CNW, VEC = ColVec<T, CNW>::VEC, UB = CNW * 4;
    constexpr int FLUSH_ITERS = 32 / kColUnroll;             // promote fp32 partials to fp64 every 32 rows
    // private fp64 accumulators of every thread's element slots, [S|B][slot][thread]: no atomics, no conflicts
    __shared__ double sacc[2][VEC][kColThreads];
    __shared__ int last_flag;
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tx = threadIdx.x % cs.tx, ty = threadIdx.x / cs.tx;
    const long long uc = (long long)blockIdx.x * cs.tx + tx;
    const bool active = uc < cs.units_per_row;
    const char* __restrict__ xp = reinterpret_cast<const char*>(cs.x);
    const char* __restrict__ gp = reinterpret_cast<const char*>(cs.g);
    char* __restrict__ gxp = reinterpret_cast<char*>(cs.gx);
    const bool write_gx = gxp != nullptr;
    float accS[VEC], accB[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) { accS[k] = 0.f; accB[k] = 0.f; sacc[0][k][threadIdx.x] = 0.0; sacc[1][k][threadIdx.x] = 0.0; }
    SlotParams<T, MODE, CNW> sp;
    if (active) {
        sp.load(cs, uc);
        long long n = (long long)blockIdx.y * cs.rows_per_split + ty;
        long long n_end = ((long long)blockIdx.y + 1) * cs.rows_per_split;
        if (n_end > cs.outer) n_end = cs.outer;
        int since_flush = 0;
        for (; n < n_end; n += (long long)cs.ty * kColUnroll) {
            Raw<NW> xr[kColUnroll], gr[kColUnroll];
            long long a[kColUnroll];
#pragma unroll
            for (int r = 0; r < kColUnroll; r++) {
                const long long nn = n + (long long)r * cs.ty;
                a[r] = (nn < n_end ? nn : n) * cs.units_per_row + uc;
                xr[r] = ld_unit<LD, NW>(xp + a[r] * UB);
                gr[r] = ld_unit<LD, NW>(gp + a[r] * UB);
            }
#pragma unroll
            for (int r = 0; r < kColUnroll; r++) {
                if (n + (long long)r * cs.ty >= n_end) continue;
                float fx[VEC], fg[VEC];
                unpack_unit<T, NW>(xr[r], fx);
                unpack_unit<T, NW>(gr[r], fg);
#pragma unroll
                for (int k = 0; k < VEC; k++)
                    fg[k] = fq_backward<MODE, BMODE, true>(fg[k], fx[k], sp.chan(k, cs), accS[k], accB[k]);
                if (write_gx) {
                    if (bmode_passthrough(BMODE)) st_unit<ST, NW>(gxp + a[r] * UB, gr[r]);
                    else st_unit<ST, NW>(gxp + a[r] * UB, pack_unit<T, NW>(fg));
                }
            }
            if (bmode_reduces(BMODE) && ++since_flush == FLUSH_ITERS) {
                since_flush = 0;
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    sacc[0][k][threadIdx.x] += (double)accS[k]; accS[k] = 0.f;
                    sacc[1][k][threadIdx.x] += (double)accB[k]; accB[k] = 0.f;
                }
            }
        }
    }
    if constexpr (!bmode_reduces(BMODE)) {      // eval: exact zeros, written by the first row-split
        if (active && blockIdx.y == 0 && ty == 0) {
#pragma unroll
            for (int k = 0; k < VEC; k++)
                if (k == 0 || sp.ch[k] != sp.ch[k > 0 ? k - 1 : 0]) {
                    store_param(cs.gscale, sp.ch[k], cs.pdt, 0.0);
                    store_param(cs.gshift, sp.ch[k], cs.pdt, 0.0);
                }
        }
    } else {
#pragma unroll
    for (int k = 0; k < VEC; k++) {
        sacc[0][k][threadIdx.x] += (double)accS[k];
        sacc[1][k][threadIdx.x] += (double)accB[k];
    }
    __syncthreads();
    // thread row 0 adds the CTA's thread rows in a fixed order, merges neighbouring slots of the same
    // channel and issues one fp64 atomic pair per channel run
    if (active && ty == 0) {
        double rs = 0.0, rb = 0.0;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            for (int r = 0; r < cs.ty; r++) { rs += sacc[0][k][r * cs.tx + tx]; rb += sacc[1][k][r * cs.tx + tx]; }
            if (k == VEC - 1 || sp.ch[k + 1 < VEC ? k + 1 : k] != sp.ch[k]) {
                atomicAdd(cs.acc + 2 * (long long)sp.ch[k], rs);
                atomicAdd(cs.acc + 2 * (long long)sp.ch[k] + 1, rb);
                rs = 0.0; rb = 0.0;
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(cs.counter, 1u);
        last_flag = (prev == cs.total_ctas - 1u);
    }
    __syncthreads();
    if (!last_flag) return;
    __threadfence();
    for (long long c = threadIdx.x; c < cs.C; c += kColThreads) {
        const double a = __ldcg(cs.acc + 2 * c), b = __ldcg(cs.acc + 2 * c + 1);
        store_param(cs.gscale, c, cs.pdt, a * cs.gs);
        store_param(cs.gshift, c, cs.pdt, cs.sym ? 0.0 : b * cs.gs);
        cs.acc[2 * c] = 0.0; cs.acc[2 * c + 1] = 0.0;      // leave the workspace zeroed
    }
    if (threadIdx.x == 0) *cs.counter = 0u;
    }
}

using ColKernelFn = void (*)(const ColSeg);
ColKernelFn get_col_fwd_kernel(int xdtype, int mode, bool init, int variant);
ColKernelFn get_col_bwd_kernel(int xdtype, int mode, int bmode, int variant);
// variant -> (unit words, rows in flight, min CTAs/SM); index with Tuning::col_variant
constexpr int kColVariants = 4;
constexpr int kColVariantNW[kColVariants] = {4, 4, 2, 2};

}  // namespace lsqb200
