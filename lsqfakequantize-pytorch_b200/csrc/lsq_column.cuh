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
    const void* x2;       // second addend (ADD prologues), same layout as x
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
    int l2_prefetch;           // rows per thread L2-prefetched before the dependency wait (0 = off)
};

constexpr int kColThreads = 256;
#ifndef LSQ_COL_EARLY_LOADS
#define LSQ_COL_EARLY_LOADS 1
#endif

// diagnostic build (-DLSQ_COL_TRACE): per-CTA globaltimer stamps + SM id into the partials region of the workspace (tools/coltrace.py)
#ifdef LSQ_COL_TRACE
#define LSQ_TRACE(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(cs.counter) + 16384)[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (i)] = t_; } } while (0)
#define LSQ_TRACE_SM() do { if (threadIdx.x == 0) { unsigned s_; asm volatile("mov.u32 %0, %%smid;" : "=r"(s_)); \
    reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(cs.counter) + 16384)[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = s_; } } while (0)
#else
#define LSQ_TRACE(i) do {} while (0)
#define LSQ_TRACE_SM() do {} while (0)
#endif

// CNW: 32-bit words per column unit (4 = 128-bit, 2 = 64-bit accesses); fewer slots per thread
// mean fewer registers (more resident CTAs) at the price of narrower accesses
template <typename T, int CNW>
struct ColVec { static constexpr int NW = CNW; static constexpr int VEC = UnitOf<T, CNW>::VEC; };

// per-slot channel constants
template <typename T, int MODE, int CNW>
struct SlotParams {
    static constexpr int VEC = ColVec<T, CNW>::VEC;
    float s[VEC], inv_s[VEC], zp[VEC];
    // The raw parameter loads are ISSUED first (issue), the caller then puts its first data loads in flight, and only then are
    // the derived constants formed (finish): the thread's two dependent DRAM round trips (parameters, data) overlap instead of
    // adding up, and make_chan (an IEEE division) runs once per DISTINCT channel of the unit, not once per slot (a 16-byte
    // unit of a 7x7 / 14x14 map spans at most two channels).
    float sraw[VEC], braw[VEC];
    __device__ __forceinline__ void issue(const ColSeg& cs, long long unit_col) {
        // one 32-bit division for the first slot, then walk: L = C*inner < 2^31 is checked on the host
        const unsigned j0 = (unsigned)unit_col * VEC, inner = (unsigned)cs.inner;
        unsigned c = j0 / inner, r = j0 - c * inner;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            sraw[k] = load_param(cs.scale, c, cs.pdt);
            braw[k] = load_param(cs.shift, c, cs.pdt);
            if (++r == inner) { r = 0; ++c; }
        }
        r0 = j0 - (j0 / inner) * inner;
    }
    unsigned r0;     // position of slot 0 inside its channel
    __device__ __forceinline__ void finish(const ColSeg& cs) {
        Seg fake;                       // make_chan only reads these fields
        fake.per_channel = 1; fake.tmin = cs.tmin; fake.tmax = cs.tmax; fake.qmin = cs.qmin; fake.qmax = cs.qmax;
        const unsigned inner = (unsigned)cs.inner;
        unsigned r = r0;
        Chan cc;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            if (k == 0 || r == 0) cc = make_chan<MODE>(sraw[k], braw[k], fake);     // slot k starts a new channel
            s[k] = cc.s; inv_s[k] = cc.inv_s; zp[k] = cc.zp;
            if (++r == inner) r = 0;
        }
    }
    __device__ __forceinline__ void load(const ColSeg& cs, long long unit_col) { issue(cs, unit_col); finish(cs); }
    __device__ __forceinline__ Chan chan(int k, const ColSeg& cs) const {
        Chan c;
        c.s = s[k]; c.inv_s = inv_s[k]; c.zp = zp[k]; c.qmin = cs.qmin; c.qmax = cs.qmax;
        c.c_lo = 0.f; c.c_hi = 0.f;
        return c;
    }
};

// The scale / shift entries of a thread's first channel on their way into L2 before the dependency wait: a prefetch binds no value,
// so whatever the predecessor still writes is what the loads behind the wait see, and the parameter request at the front of the
// kernel is an L2 hit instead of a DRAM round trip shared by every CTA of the launch.
__device__ __forceinline__ void prefetch_params(const ColSeg& cs, long long elem_col) {
    const unsigned c = (unsigned)elem_col / (unsigned)cs.inner;
    const unsigned ps = cs.pdt == DT_F32 ? 4u : 2u;
    l2_prefetch(reinterpret_cast<const char*>(cs.scale) + (size_t)c * ps);
    l2_prefetch(reinterpret_cast<const char*>(cs.shift) + (size_t)c * ps);
}

// rows this thread visits inside its row split, and the byte offset of the first one (ty is a power of two)
struct ColWalk {
    long long off, stride;   // bytes
    int cnt;
    __device__ __forceinline__ void init(const ColSeg& cs, long long uc, int ty, int ub) {
        const long long n0 = (long long)blockIdx.y * cs.rows_per_split + ty;
        long long n_end = ((long long)blockIdx.y + 1) * cs.rows_per_split;
        if (n_end > cs.outer) n_end = cs.outer;
        const int sh = __ffs(cs.ty) - 1;
        cnt = n0 < n_end ? (int)((unsigned)(n_end - n0 + cs.ty - 1) >> sh) : 0;    // rows_per_split < 2^31
        off = (n0 * cs.units_per_row + uc) * ub;
        stride = ((long long)cs.units_per_row << sh) * ub;
    }
};

// The launch's last CTA turns the [C][2] fp64 accumulator into grad_scale / grad_shift and leaves it zeroed.  Eight channel
// pairs per thread are loaded before the first store, so the tail of the launch costs one L2 round trip per 8 * THREADS
// channels instead of one per THREADS (the stores may alias the loads as far as the compiler knows).
template <int THREADS>
__device__ __forceinline__ void col_finalise(const ColSeg& cs) {
    constexpr int B = 8;
    for (long long c0 = threadIdx.x; c0 < cs.C; c0 += (long long)B * THREADS) {
        double2 v[B];
#pragma unroll
        for (int i = 0; i < B; i++) {
            const long long c = c0 + (long long)i * THREADS;
            v[i] = c < cs.C ? __ldcg(reinterpret_cast<const double2*>(cs.acc) + c) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int i = 0; i < B; i++) {
            const long long c = c0 + (long long)i * THREADS;
            if (c >= cs.C) continue;
            store_param(cs.gscale, c, cs.pdt, v[i].x * cs.gs);
            store_param(cs.gshift, c, cs.pdt, cs.sym ? 0.0 : v[i].y * cs.gs);
            reinterpret_cast<double2*>(cs.acc)[c] = make_double2(0.0, 0.0);      // leave the workspace zeroed
        }
    }
    if (threadIdx.x == 0) *cs.counter = 0u;
}

template <typename T, int MODE, bool INIT, int CNW, int kColUnroll, int MINB, int LD, int ST>
__global__ void __launch_bounds__(kColThreads, MINB)
lsq_col_fwd_kernel(const __grid_constant__ ColSeg cs) {
    constexpr int NW = CNW, VEC = ColVec<T, CNW>::VEC, UB = CNW * 4;
    constexpr bool RAWCOPY = INIT && !mode_relu(MODE) && !mode_add(MODE);
    constexpr bool ADD = mode_add(MODE);
    asm volatile("griddepcontrol.launch_dependents;");
    const int tx = threadIdx.x % cs.tx, ty = threadIdx.x / cs.tx;
    const long long uc = (long long)blockIdx.x * cs.tx + tx;
    if (uc >= cs.units_per_row) return;
    ColWalk w;
    w.init(cs, uc, ty, UB);
    const char* __restrict__ px = reinterpret_cast<const char*>(cs.x) + w.off;
    const char* __restrict__ px2 = ADD ? reinterpret_cast<const char*>(cs.x2) + w.off : nullptr;
    char* __restrict__ py = reinterpret_cast<char*>(cs.y) + w.off;
    for (int r = 0; r < cs.l2_prefetch && r < w.cnt; r++) {      // first rows on their way into L2 while the predecessor drains
        l2_prefetch(px + r * w.stride);
        if constexpr (ADD) l2_prefetch(px2 + r * w.stride);
    }
    if (!INIT && cs.l2_prefetch) prefetch_params(cs, uc * VEC);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int cnt = w.cnt;
    Raw<NW> xr[kColUnroll], x2r[ADD ? kColUnroll : 1];
    auto load_group = [&]() {
#pragma unroll
        for (int r = 0; r < kColUnroll; r++) {
            xr[r] = ld_unit<LD, NW>(px + r * w.stride);
            if constexpr (ADD) x2r[r] = ld_unit<LD, NW>(px2 + r * w.stride);
        }
    };
    // the first group's loads are in flight before the parameters are requested: one memory round trip at the front, not two
    // (16-bit tensors: +4 % on 7x7 maps; fp32 units leave too few registers at 6 CTAs/SM and lose 3-5 %, so they keep the plain order)
    const bool pre = LSQ_COL_EARLY_LOADS && sizeof(T) == 2 && !INIT && cnt >= kColUnroll;
    if (pre) load_group();
    SlotParams<T, MODE, CNW> sp;
    if (!INIT) sp.load(cs, uc);
    // whole groups of kColUnroll rows: unconditional loads and stores
    for (bool first = pre; cnt >= kColUnroll; cnt -= kColUnroll, first = false) {
        if (!first) load_group();
#pragma unroll
        for (int r = 0; r < kColUnroll; r++) {
            if (RAWCOPY) { st_unit<ST, NW>(py + r * w.stride, xr[r]); continue; }
            float f[VEC];
            unpack_unit<T, NW>(xr[r], f);
            if constexpr (ADD) {
                float f2[VEC];
                unpack_unit<T, NW>(x2r[r], f2);
#pragma unroll
                for (int k = 0; k < VEC; k++) f[k] = pre_add<T>(f[k], f2[k]);
            }
#pragma unroll
            for (int k = 0; k < VEC; k++) f[k] = INIT ? pre_op<MODE>(f[k]) : fq_forward<MODE>(f[k], sp.chan(k, cs));
            st_unit<ST, NW>(py + r * w.stride, pack_unit<T, NW>(f));
        }
        px += kColUnroll * w.stride; py += kColUnroll * w.stride;
        if constexpr (ADD) px2 += kColUnroll * w.stride;
    }
    for (; cnt > 0; cnt--) {
        const Raw<NW> xs = ld_unit<LD, NW>(px);
        if (RAWCOPY) st_unit<ST, NW>(py, xs);
        else {
            float f[VEC];
            unpack_unit<T, NW>(xs, f);
            if constexpr (ADD) {
                float f2[VEC];
                unpack_unit<T, NW>(ld_unit<LD, NW>(px2), f2);
#pragma unroll
                for (int k = 0; k < VEC; k++) f[k] = pre_add<T>(f[k], f2[k]);
                px2 += w.stride;
            }
#pragma unroll
            for (int k = 0; k < VEC; k++) f[k] = INIT ? pre_op<MODE>(f[k]) : fq_forward<MODE>(f[k], sp.chan(k, cs));
            st_unit<ST, NW>(py, pack_unit<T, NW>(f));
        }
        px += w.stride; py += w.stride;
    }
}

// A channel here sums outer*inner >= thousands of terms, so the column backward uses the streaming form of
// fq_backward (fused accumulation, lsq_device.cuh) like the row-tiled kernels do for long channels.
template <typename T, int MODE, int BMODE, int CNW, int kColUnroll, int MINB, int LD, int ST>
__global__ void __launch_bounds__(kColThreads, MINB)
lsq_col_bwd_kernel(const __grid_constant__ ColSeg cs) {
    constexpr int NW = CNW, VEC = ColVec<T, CNW>::VEC, UB = CNW * 4;
    constexpr int FLUSH_ROWS = 32;                            // promote fp32 partials to fp64 every 32 rows
    constexpr bool ADD = mode_add(MODE);
    // private fp64 accumulators of every thread's element slots, [S|B][slot][thread]: no atomics, no conflicts
    __shared__ double sacc[2][VEC][kColThreads];
    __shared__ int last_flag;
    asm volatile("griddepcontrol.launch_dependents;");
    LSQ_TRACE(0); LSQ_TRACE_SM();
    const int tx = threadIdx.x % cs.tx, ty = threadIdx.x / cs.tx;
    const long long uc = (long long)blockIdx.x * cs.tx + tx;
    const bool active = uc < cs.units_per_row;
    float accS[VEC], accB[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) { accS[k] = 0.f; accB[k] = 0.f; sacc[0][k][threadIdx.x] = 0.0; sacc[1][k][threadIdx.x] = 0.0; }
    if (active && cs.l2_prefetch) {               // first rows on their way into L2 while the predecessor drains
        ColWalk pw;
        pw.init(cs, uc, ty, UB);
        for (int r = 0; r < cs.l2_prefetch && r < pw.cnt; r++) {
            l2_prefetch(reinterpret_cast<const char*>(cs.x) + pw.off + r * pw.stride);
            l2_prefetch(reinterpret_cast<const char*>(cs.g) + pw.off + r * pw.stride);
            if constexpr (ADD) l2_prefetch(reinterpret_cast<const char*>(cs.x2) + pw.off + r * pw.stride);
        }
        prefetch_params(cs, uc * VEC);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    LSQ_TRACE(1);
    if (active) {
        ColWalk w;
        w.init(cs, uc, ty, UB);
        const char* __restrict__ px = reinterpret_cast<const char*>(cs.x) + w.off;
        const char* __restrict__ pg = reinterpret_cast<const char*>(cs.g) + w.off;
        const char* __restrict__ px2 = ADD ? reinterpret_cast<const char*>(cs.x2) + w.off : nullptr;
        char* __restrict__ pgx = cs.gx ? reinterpret_cast<char*>(cs.gx) + w.off : nullptr;
        int cnt = w.cnt, since = 0;
        Raw<NW> xr[kColUnroll], gr[kColUnroll], x2r[ADD ? kColUnroll : 1];
        auto load_group = [&]() {
#pragma unroll
            for (int r = 0; r < kColUnroll; r++) {
                xr[r] = ld_unit<LD, NW>(px + r * w.stride);
                if constexpr (ADD) x2r[r] = ld_unit<LD, NW>(px2 + r * w.stride);
                gr[r] = ld_unit<LD, NW>(pg + r * w.stride);
            }
        };
        // (loads ahead of the parameter request, as the forward does for 16-bit tensors, cost the backward 4 %: 128 registers are full)
        SlotParams<T, MODE, CNW> sp;
        sp.load(cs, uc);
        LSQ_TRACE(2);
        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                sacc[0][k][threadIdx.x] += (double)accS[k]; accS[k] = 0.f;
                sacc[1][k][threadIdx.x] += (double)accB[k]; accB[k] = 0.f;
            }
        };
        auto row = [&](const Raw<NW>& xr, const Raw<NW>& x2r, const Raw<NW>& gr, char* dst) {
            float fx[VEC], fg[VEC];
            unpack_unit<T, NW>(xr, fx);
            if constexpr (ADD) {
                float f2[VEC];
                unpack_unit<T, NW>(x2r, f2);
#pragma unroll
                for (int k = 0; k < VEC; k++) fx[k] = pre_add<T>(fx[k], f2[k]);
            }
            unpack_unit<T, NW>(gr, fg);
#pragma unroll
            for (int k = 0; k < VEC; k++)
                fg[k] = fq_backward<MODE, BMODE, false>(fg[k], fx[k], sp.chan(k, cs), accS[k], accB[k]);
            if (dst) {
                if (bmode_passthrough(BMODE) && !mode_relu(MODE)) st_unit<ST, NW>(dst, gr);
                else st_unit<ST, NW>(dst, pack_unit<T, NW>(fg));
            }
        };
        for (; cnt >= kColUnroll; cnt -= kColUnroll) {
            load_group();
#pragma unroll
            for (int r = 0; r < kColUnroll; r++) row(xr[r], x2r[ADD ? r : 0], gr[r], pgx ? pgx + r * w.stride : nullptr);
            px += kColUnroll * w.stride; pg += kColUnroll * w.stride;
            if constexpr (ADD) px2 += kColUnroll * w.stride;
            if (pgx) pgx += kColUnroll * w.stride;
            if (bmode_reduces(BMODE) && (since += kColUnroll) >= FLUSH_ROWS) { since = 0; flush(); }
        }
        for (; cnt > 0; cnt--) {
            const Raw<NW> xs = ld_unit<LD, NW>(px), gsr = ld_unit<LD, NW>(pg);
            Raw<NW> x2s;
            if constexpr (ADD) { x2s = ld_unit<LD, NW>(px2); px2 += w.stride; } else x2s = xs;
            row(xs, x2s, gsr, pgx);
            px += w.stride; pg += w.stride;
            if (pgx) pgx += w.stride;
        }
        if (bmode_reduces(BMODE)) flush();
    }
    // channel of element slot k of this thread's column unit (L = C*inner < 2^31, checked on the host)
    auto slot_channel = [&](int k) { return (int)(((unsigned)uc * VEC + (unsigned)k) / (unsigned)cs.inner); };
    if constexpr (!bmode_reduces(BMODE)) {      // eval: exact zeros, written by the first row-split
        if (active && blockIdx.y == 0 && ty == 0) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const int c = slot_channel(k);
                if (k == 0 || c != slot_channel(k > 0 ? k - 1 : 0)) {
                    store_param(cs.gscale, c, cs.pdt, 0.0);
                    store_param(cs.gshift, c, cs.pdt, 0.0);
                }
            }
        }
    } else {
    LSQ_TRACE(3);
    __syncthreads();
    // Thread row 0 adds the CTA's thread rows in a fixed order and merges neighbouring slots of the same channel into runs.
    // inner >= VEC (7x7, 14x14 maps ...): a unit spans at most two channels, a channel several threads - the runs are parked in the
    // thread's own shared cells and ONE thread per channel of the CTA adds them in column order (fixed order, no shared atomics:
    // fp64 atomics on shared memory are CAS loops), so a single fp64 atomic pair per (CTA, channel) reaches L2 instead of one per
    // (thread, run) - 12x fewer on a 7x7 map, 25x on 14x14.  inner < VEC (channels-last): every run goes straight to L2.
    const unsigned inner = (unsigned)cs.inner;
    const bool merge = inner >= (unsigned)VEC;
    if (active && ty == 0) {
        double rs = 0.0, rb = 0.0, hs = 0.0, hb = 0.0;
        bool tail = false;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            for (int r = 0; r < cs.ty; r++) { rs += sacc[0][k][r * cs.tx + tx]; rb += sacc[1][k][r * cs.tx + tx]; }
            const int c = slot_channel(k);
            if (k == VEC - 1 || slot_channel(k + 1 < VEC ? k + 1 : k) != c) {
                if (merge) {
                    if (!tail) { hs = rs; hb = rb; tail = true; }       // head run; what follows (if anything) is the one tail run
                } else {
                    atomicAdd(cs.acc + 2 * (long long)c, rs);
                    atomicAdd(cs.acc + 2 * (long long)c + 1, rb);
                }
                if (k < VEC - 1) { rs = 0.0; rb = 0.0; }
            }
        }
        if (merge) {
            const bool two = slot_channel(VEC - 1) != slot_channel(0);
            sacc[0][0][tx] = hs; sacc[1][0][tx] = hb;                   // this thread's own cells (thread row 0 of column tx)
            sacc[0][1][tx] = two ? rs : 0.0; sacc[1][1][tx] = two ? rb : 0.0;
        }
    }
    if (merge) {
        __syncthreads();
        const unsigned u_first = blockIdx.x * (unsigned)cs.tx;
        unsigned u_last = u_first + (unsigned)cs.tx;
        if (u_last > (unsigned)cs.units_per_row) u_last = (unsigned)cs.units_per_row;
        const unsigned c_first = u_first * VEC / inner, c_last = (u_last * VEC - 1u) / inner;
        for (unsigned c = c_first + threadIdx.x; c <= c_last; c += kColThreads) {
            const unsigned e0 = c * inner, e1 = e0 + inner - 1u;         // element columns of channel c
            unsigned u = e0 / VEC, ue = e1 / VEC;
            if (u < u_first) u = u_first;
            if (ue > u_last - 1u) ue = u_last - 1u;
            double s0 = 0.0, b0 = 0.0;
            for (; u <= ue; u++) {
                const int j = u * VEC >= e0 ? 0 : 1;                      // the unit starts inside c: head run; before c: tail run
                s0 += sacc[0][j][u - u_first]; b0 += sacc[1][j][u - u_first];
            }
            atomicAdd(cs.acc + 2 * (long long)c, s0);
            atomicAdd(cs.acc + 2 * (long long)c + 1, b0);
        }
    }
    LSQ_TRACE(4);
    // release: the barrier orders every thread's atomics before thread 0's fence, the fence before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_acq_rel_gpu();
        const unsigned prev = atomicAdd(cs.counter, 1u);
        last_flag = (prev == cs.total_ctas - 1u);
    }
    __syncthreads();
    LSQ_TRACE(5);
    if (!last_flag) return;
    fence_acq_rel_gpu();
    col_finalise<kColThreads>(cs);
    LSQ_TRACE(6);
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged column backward (north_star: "TMA staging only where the per-channel layout makes tiles pay
// off").  Same thread <-> column-unit mapping and arithmetic as lsq_col_bwd_kernel, but the R x 4 KB tile of
// x and of grad (R rows of this CTA's 256 column units) is brought into shared memory by the bulk-copy
// engine (cp.async.bulk, one 1-D copy per row and operand, no tensor map needed because every row piece is
// contiguous), S stages deep, by a dedicated producer warp; the 8 consumer warps read their 16-byte units
// from shared memory.  No registers hold loads in flight, so bytes in flight per SM are set by the ring
// (S*R*8 KB per CTA) instead of by occupancy x unroll.  grad_x leaves through ordinary coalesced 128-bit stores.
// Measured against the register-staged kernel in profiles/ (see DESIGN.md section 4).
// ---------------------------------------------------------------------------------------------
constexpr int kTmaConsumers = 256;                 // one 16-byte column unit each
constexpr int kTmaThreads = kTmaConsumers + 32;    // + the producer warp
constexpr int kTmaRowBytes = kTmaConsumers * 16;   // 4 KB of every row per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ Raw<4> lds_unit(uint32_t addr) {
    Raw<4> r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "r"(addr));
    return r;
}

template <typename T, int MODE, int BMODE, int R, int S, int ST>
__global__ void __launch_bounds__(kTmaThreads, 2)
lsq_col_bwd_tma_kernel(const __grid_constant__ ColSeg cs) {
    constexpr int NW = 4, VEC = ColVec<T, 4>::VEC, UB = 16;
    constexpr int FLUSH_ROWS = 32;
    constexpr uint32_t STAGE_BYTES = 2u * R * kTmaRowBytes;          // x rows then grad rows
    extern __shared__ __align__(128) unsigned char ring[];           // [S][2][R][4 KB]
    __shared__ __align__(8) unsigned long long bars[2 * S];          // full[S], empty[S]
    __shared__ int last_flag;
    asm volatile("griddepcontrol.launch_dependents;");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long u0 = (long long)blockIdx.x * kTmaConsumers;      // first column unit of this CTA
    const long long left = cs.units_per_row - u0;
    const int ncol = left < kTmaConsumers ? (int)left : kTmaConsumers;
    const uint32_t row_bytes = (uint32_t)ncol * UB;
    const long long n0 = (long long)blockIdx.y * cs.rows_per_split;
    long long n_end = n0 + cs.rows_per_split;
    if (n_end > cs.outer) n_end = cs.outer;
    const int nrows = n_end > n0 ? (int)(n_end - n0) : 0;
    const int niter = (nrows + R - 1) / R;
    const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, kTmaConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long row_pitch = cs.units_per_row * UB;               // bytes between rows
    if (warp == kTmaConsumers / 32) {
        // ---- producer warp: one lane feeds the ring
        if (lane == 0) {
            const char* src_x = reinterpret_cast<const char*>(cs.x) + (n0 * cs.units_per_row + u0) * UB;
            const char* src_g = reinterpret_cast<const char*>(cs.g) + (n0 * cs.units_per_row + u0) * UB;
            for (int it = 0; it < niter; it++) {
                const int s = it % S;
                const uint32_t k = (uint32_t)(it / S);
                mbar_wait(empty0 + 8 * s, (k & 1u) ^ 1u);             // first pass over the ring falls through
                const int rows = nrows - it * R < R ? nrows - it * R : R;
                mbar_expect_tx(full0 + 8 * s, 2u * (uint32_t)rows * row_bytes);
                const uint32_t dst = ring0 + (uint32_t)s * STAGE_BYTES;
                for (int r = 0; r < rows; r++) {
                    tma_load_1d(dst + (uint32_t)r * kTmaRowBytes, src_x, row_bytes, full0 + 8 * s);
                    tma_load_1d(dst + (uint32_t)(R + r) * kTmaRowBytes, src_g, row_bytes, full0 + 8 * s);
                    src_x += row_pitch; src_g += row_pitch;
                }
            }
        }
    } else {
        // ---- consumer warps: thread t owns column unit u0 + t
        const int t = threadIdx.x;
        const bool active = t < ncol;
        const long long uc = u0 + t;
        SlotParams<T, MODE, 4> sp;
        float accS[VEC], accB[VEC];
        double sumS[VEC], sumB[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) { accS[k] = 0.f; accB[k] = 0.f; sumS[k] = 0.0; sumB[k] = 0.0; }
        if (active) sp.load(cs, uc);
        char* pgx = cs.gx ? reinterpret_cast<char*>(cs.gx) + (n0 * cs.units_per_row + uc) * UB : nullptr;
        auto row = [&](uint32_t ax, uint32_t ag) {
            const Raw<NW> xr = lds_unit(ax), gr = lds_unit(ag);
            float fx[VEC], fg[VEC];
            unpack_unit<T, NW>(xr, fx);
            unpack_unit<T, NW>(gr, fg);
#pragma unroll
            for (int k = 0; k < VEC; k++)
                fg[k] = fq_backward<MODE, BMODE, false>(fg[k], fx[k], sp.chan(k, cs), accS[k], accB[k]);
            if (pgx) {
                if (bmode_passthrough(BMODE) && !mode_relu(MODE)) st_unit<ST, NW>(pgx, gr);
                else st_unit<ST, NW>(pgx, pack_unit<T, NW>(fg));
                pgx += row_pitch;
            }
        };
        int since = 0;
        for (int it = 0; it < niter; it++) {
            const int s = it % S;
            const uint32_t k = (uint32_t)(it / S);
            mbar_wait(full0 + 8 * s, k & 1u);
            const int rows = nrows - it * R < R ? nrows - it * R : R;
            if (active) {
                const uint32_t ax = ring0 + (uint32_t)s * STAGE_BYTES + (uint32_t)t * UB;
                if (rows == R) {
#pragma unroll
                    for (int r = 0; r < R; r++) row(ax + (uint32_t)r * kTmaRowBytes, ax + (uint32_t)(R + r) * kTmaRowBytes);
                } else {
                    for (int r = 0; r < rows; r++) row(ax + (uint32_t)r * kTmaRowBytes, ax + (uint32_t)(R + r) * kTmaRowBytes);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);               // this warp is done reading stage s
            if (bmode_reduces(BMODE) && (since += rows) >= FLUSH_ROWS) {
                since = 0;
#pragma unroll
                for (int q = 0; q < VEC; q++) { sumS[q] += (double)accS[q]; accS[q] = 0.f; sumB[q] += (double)accB[q]; accB[q] = 0.f; }
            }
        }
        auto slot_channel = [&](int q) { return (int)(((unsigned)uc * VEC + (unsigned)q) / (unsigned)cs.inner); };
        if (active) {
            if constexpr (!bmode_reduces(BMODE)) {
                if (blockIdx.y == 0) {
#pragma unroll
                    for (int q = 0; q < VEC; q++) {
                        const int c = slot_channel(q);
                        if (q == 0 || c != slot_channel(q > 0 ? q - 1 : 0)) {
                            store_param(cs.gscale, c, cs.pdt, 0.0);
                            store_param(cs.gshift, c, cs.pdt, 0.0);
                        }
                    }
                }
            } else {
                double rs = 0.0, rb = 0.0;
#pragma unroll
                for (int q = 0; q < VEC; q++) {
                    rs += sumS[q] + (double)accS[q]; rb += sumB[q] + (double)accB[q];
                    const int c = slot_channel(q);
                    if (q == VEC - 1 || slot_channel(q + 1 < VEC ? q + 1 : q) != c) {
                        atomicAdd(cs.acc + 2 * (long long)c, rs);
                        atomicAdd(cs.acc + 2 * (long long)c + 1, rb);
                        rs = 0.0; rb = 0.0;
                    }
                }
            }
        }
    }
    if constexpr (bmode_reduces(BMODE)) {
        fence_acq_rel_gpu();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(cs.counter, 1u);
            last_flag = (prev == cs.total_ctas - 1u);
        }
        __syncthreads();
        if (!last_flag) return;
        fence_acq_rel_gpu();
        for (long long c = threadIdx.x; c < cs.C; c += kTmaThreads) {
            const double a = __ldcg(cs.acc + 2 * c), b = __ldcg(cs.acc + 2 * c + 1);
            store_param(cs.gscale, c, cs.pdt, a * cs.gs);
            store_param(cs.gshift, c, cs.pdt, cs.sym ? 0.0 : b * cs.gs);
            cs.acc[2 * c] = 0.0; cs.acc[2 * c + 1] = 0.0;
        }
        if (threadIdx.x == 0) *cs.counter = 0u;
    }
}

using ColKernelFn = void (*)(const ColSeg);
ColKernelFn get_col_fwd_kernel(int xdtype, int mode, bool init, int variant);
ColKernelFn get_col_bwd_kernel(int xdtype, int mode, int bmode, int variant);
// TMA-staged backward: tma_variant 1 = (R 4 rows, S 3 stages, 96 KB ring), 2 = (2, 4, 64 KB), 3 = (8, 3, 192 KB); sets *smem_bytes
ColKernelFn get_col_bwd_tma_kernel(int xdtype, int mode, int bmode, int tma_variant, int* smem_bytes);
// kern_pre_*.cu: fused prologues, default variant only (the variant knob is an experiment switch of the plain kernels)
#define LSQ_DECL_PRE_COL(P)                                      \
    ColKernelFn get_col_fwd_kernel_pre_##P(int xdtype, bool init); \
    ColKernelFn get_col_bwd_kernel_pre_##P##_f32(int bmode);       \
    ColKernelFn get_col_bwd_kernel_pre_##P##_f16(int bmode);       \
    ColKernelFn get_col_bwd_kernel_pre_##P##_bf16(int bmode);
LSQ_DECL_PRE_COL(relu)
LSQ_DECL_PRE_COL(addrelu)
LSQ_DECL_PRE_COL(add)
#undef LSQ_DECL_PRE_COL
constexpr int kColVariantRelu = 1;
// variant -> (unit words, rows in flight, min CTAs/SM); index with Tuning::col_variant
constexpr int kColVariants = 6;
constexpr int kColVariantNW[kColVariants] = {4, 4, 2, 2, 4, 4};

}  // namespace lsqb200
