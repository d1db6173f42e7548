// lsq_host.h -- host-side launch planning shared by the C-ABI library and the tuning tool.
// Pure host code: turns a contiguous (outer, C, inner) tensor into the Seg descriptor the
// kernels in lsq_device.cuh consume.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>

#include "../../include/lsq_b200.h"
#include "lsq_device.cuh"

namespace lsqb200 {

// Compile-time launch shape of the product kernels (tools/tune.cu instantiates alternatives).
constexpr int kThreads = 256;
// Launch shapes picked from the round-1 sweep on B200 (profiles/r1_tune_sweep.md): 256-thread
// CTAs, ONE 256-bit unit in flight per thread and operand (two 128-bit units on the narrower
// path), 8 resident CTAs/SM forward and 4 backward.
constexpr int kUnrollFwd = 2;
constexpr int kUnrollBwd = 2;
constexpr int kUnrollStats = 4;
// units in flight per thread and operand: keep ~64 bytes per operand whatever the unit width
// warp groups own short rows and have little else to overlap their latency with: twice the units in flight
// Warp groups (one warp per short channel row, e.g. every ResNet-50 weight row): same units in flight and the same
// resident CTAs as the CTA groups.  Session-7 A/B on the 54-weight plan (tools/ab_variants.sh): 1 unit in flight x 4 / 6
// CTAs per SM 4835 GB/s, 2 units x 3 CTAs 4655, 2 units x 2 CTAs (the earlier default, 116 registers) 4186 - the rows'
// set-up and reduce phases overlap better with more warps than with deeper per-warp prefetch.
#ifndef LSQ_WG_UMUL
#define LSQ_WG_UMUL 1          // warp groups: units in flight relative to CTA groups (experiment knob)
#endif
#ifndef LSQ_WG_MINB_NUM
#define LSQ_WG_MINB_NUM 1      // warp groups: min CTAs/SM = base * NUM / DEN
#define LSQ_WG_MINB_DEN 1
#endif
constexpr int unroll_for(int base, int nw, int group = 256) {
    return (nw == 8 ? (base / 2 > 0 ? base / 2 : 1) : base) * (group == 32 ? LSQ_WG_UMUL : 1);
}
// the read-only statistics kernel keeps two units in flight per lane for warp groups (2490 vs 2264 GB/s on the 54 weights)
constexpr int unroll_for_stats(int base, int nw, int group) { return (nw == 8 ? (base / 2 > 0 ? base / 2 : 1) : base) * (group == 32 ? 2 : 1); }
constexpr int kMinBlocksStats = 2;
constexpr int kResidentStats = 4;   // CTAs/SM the statistics / observer kernels really get (<= 64 registers): whole-wave rounding uses this
constexpr int minb_for(int base, int group) { return group == 32 ? (base * LSQ_WG_MINB_NUM / LSQ_WG_MINB_DEN > 0 ? base * LSQ_WG_MINB_NUM / LSQ_WG_MINB_DEN : 1) : base; }
constexpr int kMinBlocksFwd = 6;   // __launch_bounds__ min CTAs/SM -> register cap 40
#ifndef LSQ_MINB_BWD
#define LSQ_MINB_BWD 4
#endif
constexpr int kMinBlocksBwd = LSQ_MINB_BWD;   // -> register cap 64
constexpr int kLd = LD_NC_NOALLOC;   // streaming loads: read-only path, no L1 allocation
constexpr int kSt = ST_DEFAULT;

using KernelFn = void (*)(const Seg, const Seg*, const int*, int, long long);
// defined in kern_fwd.cu / kern_bwd_*.cu / kern_stats.cu (one translation unit per family so they build in parallel)
KernelFn get_fwd_kernel(int xdtype, int mode, int nw, bool init, int group);
// kern_f64.cu: float64 tensors (nullptr for nw == 0: doubles are 8-byte aligned or the call is rejected)
KernelFn get_fwd_kernel_f64(int nw, bool init, int group);
KernelFn get_bwd_kernel_f64(int nw, int bmode, int group);
KernelFn get_bwd_kernel_f32(int mode, int nw, int bmode, int group);
KernelFn get_bwd_kernel_f16_mixed(int mode, int nw, int bmode, int group);
KernelFn get_bwd_kernel_f16_exact(int mode, int nw, int bmode, int group);
KernelFn get_bwd_kernel_bf16(int mode, int nw, int bmode, int group);
// kern_pre_*.cu: fused prologues (MODE = M_FP32_RELU / M_FP32_ADD_RELU / M_FP32_ADD: fake-quant of relu(x), relu(x + x2),
// x + x2 in one pass); fp32 / fp16 / bf16 tensors, fp32-internal arithmetic
#define LSQ_DECL_PRE(P)                                                   \
    KernelFn get_fwd_kernel_pre_##P(int xdtype, int nw, bool init, int group); \
    KernelFn get_bwd_kernel_pre_##P##_f32(int nw, int bmode, int group);   \
    KernelFn get_bwd_kernel_pre_##P##_f16(int nw, int bmode, int group);   \
    KernelFn get_bwd_kernel_pre_##P##_bf16(int nw, int bmode, int group);
LSQ_DECL_PRE(relu)
LSQ_DECL_PRE(addrelu)
LSQ_DECL_PRE(add)
#undef LSQ_DECL_PRE
inline KernelFn get_fwd_kernel_pre(int mode, int xdtype, int nw, bool init, int group) {
    if (mode == M_FP32_RELU) return get_fwd_kernel_pre_relu(xdtype, nw, init, group);
    if (mode == M_FP32_ADD_RELU) return get_fwd_kernel_pre_addrelu(xdtype, nw, init, group);
    return get_fwd_kernel_pre_add(xdtype, nw, init, group);
}
inline KernelFn get_bwd_kernel_pre(int mode, int xdtype, int nw, int bmode, int group) {
#define LSQ_PICK_PRE(P) (xdtype == DT_F32 ? get_bwd_kernel_pre_##P##_f32(nw, bmode, group) \
                         : (xdtype == DT_F16 ? get_bwd_kernel_pre_##P##_f16(nw, bmode, group) : get_bwd_kernel_pre_##P##_bf16(nw, bmode, group)))
    if (mode == M_FP32_RELU) return LSQ_PICK_PRE(relu);
    if (mode == M_FP32_ADD_RELU) return LSQ_PICK_PRE(addrelu);
    return LSQ_PICK_PRE(add);
#undef LSQ_PICK_PRE
}
inline KernelFn get_bwd_kernel(int xdtype, int mode, int nw, int bmode, int group) {
    if (mode_relu(mode) || mode_add(mode)) return get_bwd_kernel_pre(mode, xdtype, nw, bmode, group);
    if (xdtype == DT_F64) return get_bwd_kernel_f64(nw, bmode, group);
    if (xdtype == DT_F32) return get_bwd_kernel_f32(mode, nw, bmode, group);
    if (xdtype == DT_F16) return mode == M_HALF_EXACT ? get_bwd_kernel_f16_exact(mode, nw, bmode, group)
                                                      : get_bwd_kernel_f16_mixed(mode, nw, bmode, group);
    return get_bwd_kernel_bf16(mode, nw, bmode, group);
}
KernelFn get_stats_kernel(int xdtype, int nw, int group);
// kern_flat.cu: lean per-tensor kernels for single launches (M_FP32 arithmetic only); the descriptor is their one parameter
using FlatKernelFn = void (*)(const Seg);
FlatKernelFn get_flatfwd_kernel(int xdtype, bool init);
FlatKernelFn get_flatbwd_kernel(int xdtype, int bmode);
// kern_rows.cu: lean warp-per-row forward / backward for aligned weight rows (M_FP32 arithmetic only)
KernelFn get_rowfwd_kernel(int xdtype, bool init);
KernelFn get_rowbwd_kernel(int xdtype, int bmode);
constexpr int kRowUnrollFwd = 2, kRowMinBlocksFwd = 6;   // units in flight per lane, CTAs/SM
constexpr int kRowUnrollBwd = 1, kRowMinBlocksBwd = 4;
// weight rows (contiguous, 32-byte aligned, one warp each): see lsq_rowstats_kernel
KernelFn get_rowstats_kernel(int xdtype, int variant);
KernelFn get_rowstats_ring_kernel(int xdtype);
constexpr int kRowStatsUnroll = 4;       // variant 1: 256-bit loads in flight per lane
constexpr int kRowStatsMinBlocks = 4;    // variant 1: CTAs/SM (register cap 64: four 8-word units in flight need 32 registers alone)
KernelFn get_observe_kernel(int xdtype, int nw, int group);
// kern_export.cu: dir 0 = quantize (x -> uint8 / int8 codes), 1 = dequantize; sem 0 = LSQ forward's integer, 1 = torch.quantize_per_*
KernelFn get_export_kernel(int xdtype, int mode, int nw, int sem, int dir, int group);
// kern_optim.cu: fused SGD / Adam step over flat fp32 buffers
int launch_flat_optim(float* p, const float* g, float* s1, float* s2, long long n, const ::lsqb200_optim_args* o, bool pdl, cudaStream_t st);
int launch_flat_optim_sites(float* p, const float* g, float* s1, float* s2, int* steps, const unsigned char* active, long long n,
                            const ::lsqb200_optim_args* o, bool pdl, cudaStream_t st);
int launch_qparams(const void* scale, const void* shift, float* scale_out, long long* zp_out, long long n, int pdt,
                   float tmin, float tmax, bool pdl, cudaStream_t st);

// Fixed workspace layout (see lsqb200_workspace_bytes): tickets first, partials after.
constexpr long long kMaxCounters = 4096;     // channels that may be split across tiles
constexpr long long kMaxSplitTiles = 16384;  // tiles of a split launch (2 doubles each)
// third region: [C][2] fp64 accumulators of the column-layout backward; unlike the partials (scratch,
// always overwritten before being read) this region is ZERO between calls
constexpr long long kMaxColumnChannels = 16384;
constexpr size_t kColAccOffset = kMaxCounters * 4 + kMaxSplitTiles * 16;
constexpr size_t kWorkspaceBytes = kColAccOffset + kMaxColumnChannels * 16;

struct Tuning {
    int sm_count = 148;
    // Tile = the slice of one channel a thread group owns.  Small tiles keep all SMs busy to the
    // end of the launch (SMs do not progress at the same speed; a single wave of big equal
    // slices costs ~10 %); the backward pays one block reduction + ticket per tile, so its
    // tiles are larger.  Sizes are bytes of ONE operand.
    // In-step sweep on the ResNet-50 step (tools/gpu_tilesweep.sh, two alternating rounds per value, +-5 GB/s run to run): forward 128 /
    // 64 / 48 / 32 / 24 / 16 KB -> 6149 / 6222 / 6257 / 6270-6287 / 6283 / 6256-6312 GB/s; backward 1024 / 512 / 384 / 256 / 192 / 128 / 64 KB
    // -> 6104 / 6222 / 6253 / 6251-6287 / 6267 / 6246 / 6181 (profiles/r2_tile_sweep.md).  Round 1 sized the tiles on isolated 1 GiB
    // launches (64 / 512 KB); inside the step smaller tiles shorten each launch's tail of straggling CTAs, which under programmatic
    // dependent launch is idle time the successor cannot use.
    int fwd_tile_kb = 32;
    int bwd_tile_kb = 256;
    int stats_tile_kb = 128;
    // small tensors: shrink tiles until the machine is full - ONE wave of resident CTAs (6 / 4 per SM), not more:
    // in-stream sweep over the ResNet-50 site sizes (tools/site_sweep.py, profiles/r1_site_sweep.md): a second,
    // partly filled wave costs 2-3.5 us per backward launch on 6-50 M-element sites
    int fwd_min_tiles_per_sm = 6;
    int bwd_min_tiles_per_sm = 4;
    int warp_units = 1024;      // channels with <= this many units are owned by warp groups (all ResNet-50 weight rows)
    int min_iters = 1;          // never split below min_iters full group iterations
    int interleave = 1;         // 1: interleave the splits of a channel (grid-stride style), 0: contiguous slices
    int column_path = 1;        // short channel rows (channels-last, 7x7 / 14x14 maps) use the column-layout kernels
    int col_variant = 1;        // column BACKWARD: 0 = 128-bit units x2 rows, 1 = x4 rows (default, r1 + r2 sweeps), 2 = 64-bit units x4 rows, 3 = x8 rows
    int col_variant_fwd = 2;    // column FORWARD: 64-bit units x4 rows at 6 CTAs/SM (r2 sweep: +4-5 % over variant 1 on 7x7 / 14x14 / channels-last; the backward loses with it)
    int col_waves = 2;          // column forward: CTAs <= col_waves * sm_count * (resident CTAs/SM of the kernel), rounded DOWN to whole waves
    int col_tma = 0;            // column backward staged by the bulk-copy engine (lsq_col_bwd_tma_kernel): 0 off, 1 = 4 rows x 3 stages, 2 = 2 x 4, 3 = 8 x 3
    int col_waves_bwd = 1;      // column backward: one wave (per-thread set-up and the per-CTA atomics are paid once per row split; s6 sweep)
    int column_max_row_bytes = 512;   // rows shorter than this (or not 16 B multiples) take the column path
    int whole_waves = 1;        // round big-tensor tile counts to whole waves of resident CTAs
    int pdl = 1;                // launch with programmatic stream serialization (prologue overlaps predecessor's tail)
    int l2_prefetch = 2;        // units per thread L2-prefetched before the dependency wait (row-tiled / flat / weight-row forward and backward); B200 step: 0 -> 6158, 1 -> 6231, 2 -> 6234, 4 -> 6238, 8 -> 6232 GB/s
    int max_unit_bytes = 32;    // 32 -> LDG.E.256 / STG.E.256 (sm_100), 16 -> 128-bit accesses
    // single per-tensor launches in 32-byte aligned buffers: the lean kernels (bit-identical results).  1 = forward only,
    // 2 = forward and backward (default since round 2), 0 = off.  B200, bf16 sites: lean forward is faster alone (ncu: 120.7 vs
    // 123.5 us on the largest site, 8.5 vs 10.0 on the smallest) and in a stream (123.6 vs 127.0); the lean backward is faster alone
    // (186.4 vs 190.7) but was SLOWER back to back under programmatic dependent launch in round 1 (193.6 vs 188.0 per launch: the
    // general kernel's long set-up runs before its griddepcontrol.wait and overlaps the predecessor's reduction tail).  With the L2
    // prefetch of the first units in front of the wait (l2_prefetch) the lean backward has useful work for that window too:
    // 71 + 54-site step 6193 GB/s (1), 6213 (2), 6203 (0); bf16 backward in the step 6115 / 6147 / 6105 GB/s.
    int flatkernels = 2;
    int rowkernels = 1;         // forward / backward over aligned weight rows: the lean warp-per-row kernels (0: the general warp-group kernels)
    int rowstats = 4;           // mu +- 3 sigma over aligned weight rows (kern_stats.cu): 1 / 2 / 3 descriptor form (x4 / x2 / x1 units in flight), 4 / 5 / 6 row-entry form (plans; default 4: x2, 6 CTAs/SM), 7 bulk-copy ring; 0: the general kernel
                                // 54 ResNet-50 weights under ncu on B200: general kernel 34.5 us, variant 1 28.7, 2 25.2, 3 26.3
    // resident CTAs/SM the kernel family really gets (its __launch_bounds__): whole-wave rounding uses these.  The two-operand
    // ADD prologues run with 4 / 3 (tuning_for_mode below)
    int fwd_resident = kMinBlocksFwd;
    int bwd_resident = kMinBlocksBwd;
};
constexpr int kMinBlocksFwdAdd = 4, kMinBlocksBwdAdd = 3;   // LSQ_PRE_MINB of kern_pre_*_add*.cu
inline Tuning tuning_for_mode(const Tuning& tn, int mode) {
    if (mode == M_HALF_EXACT) {      // reference-exact fp16 runs the general kernels (longer per-tile set-up): round-1 tile sizes, 0.917 vs 0.884 on 205 MB
        Tuning t = tn;
        t.fwd_tile_kb = tn.fwd_tile_kb * 2; t.bwd_tile_kb = tn.bwd_tile_kb * 2;
        return t;
    }
    if (!mode_add(mode)) return tn;
    Tuning t = tn;
    t.fwd_resident = kMinBlocksFwdAdd; t.bwd_resident = kMinBlocksBwdAdd;
    if (t.fwd_min_tiles_per_sm > kMinBlocksFwdAdd) t.fwd_min_tiles_per_sm = kMinBlocksFwdAdd;
    if (t.bwd_min_tiles_per_sm > kMinBlocksBwdAdd) t.bwd_min_tiles_per_sm = kMinBlocksBwdAdd;
    return t;
}

enum : int { K_FWD = 0, K_BWD = 1, K_STATS = 2 };

struct Geometry {
    int regime, vec, nw, group, splits, interleave;   // nw: 32-bit words per unit (8 / 4 / 2), 0 = scalar path
    long long outer, C, inner;   // after collapsing C == 1
    long long vpr, row_stride, chan_units, units_per_split, tiles, grid;
};

inline int elem_size(int dt) { return dt == DT_F64 ? 8 : (dt == DT_F32 ? 4 : 2); }
// one contiguous channel streamed by CTA groups with interleaved (or single) splits, 32-byte aligned buffers: the lean per-tensor kernels apply
inline bool flat_eligible(const Geometry& g, int xdtype) {
    return g.regime == 0 && g.C == 1 && g.nw == 8 && g.group != 32 && (g.splits == 1 || g.interleave) && xdtype != DT_F64 && g.inner > 0;
}
// contiguous channel rows that one warp owns, in tensors whose base pointers are 32-byte aligned: the lean row kernels apply
inline bool row_kernels_eligible(const Geometry& g, int xdtype) {
    return g.regime == 0 && g.group == 32 && g.nw == 8 && g.splits == 1 && xdtype != DT_F64 && g.inner > 0;   // rows need not be 32-byte multiples: scalar head / tail
}

inline Geometry plan_geometry(long long outer, long long C, long long inner, int xdtype, int kind,
                              int align_bytes, const Tuning& tn, int threads = kThreads, int unroll_override = 0) {
    Geometry g{};
    if (C == 1) { inner *= outer; outer = 1; }      // per-tensor: one contiguous channel
    g.outer = outer; g.C = C; g.inner = inner;
    const int es = elem_size(xdtype);
    g.regime = (outer == 1) ? 0 : 1;
    // widest unit (32 / 16 / 8 bytes) the base pointers - and, for strided rows, every row start -
    // are aligned to; otherwise the scalar path
    g.nw = 0;
    for (int ub : {32, 16, 8}) {
        if (ub > tn.max_unit_bytes || align_bytes % ub != 0) continue;
        if (g.regime == 1 && (inner * es) % ub != 0) continue;
        g.nw = ub / 4;
        break;
    }
    g.vec = g.nw ? g.nw * 4 / es : 1;
    if (g.regime == 0) {
        g.vpr = 1LL << 30; g.row_stride = 1LL << 30;   // artificial rows: contiguous, 32-bit walker state
        g.chan_units = (g.vec == 1) ? inner : (inner + g.vec - 1) / g.vec;   // upper bound of the aligned body
    } else {
        g.vpr = inner / g.vec; g.row_stride = C * g.vpr;
        g.chan_units = outer * g.vpr;
    }
    const int unroll = unroll_override ? unroll_override : (kind == K_FWD ? kUnrollFwd : (kind == K_BWD ? kUnrollBwd : kUnrollStats));
    // how many tiles: by size, but at least enough to fill the machine
    const long long unit_bytes = g.nw ? g.nw * 4 : es;
    const int tile_kb = kind == K_FWD ? tn.fwd_tile_kb : (kind == K_BWD ? tn.bwd_tile_kb : tn.stats_tile_kb);
    long long tile_units = (long long)tile_kb * 1024 / unit_bytes;
    if (tile_units < 1) tile_units = 1;
    const long long min_total = (long long)tn.sm_count * (kind == K_FWD ? tn.fwd_min_tiles_per_sm : tn.bwd_min_tiles_per_sm);
    long long splits = (g.chan_units + tile_units - 1) / tile_units;
    const long long by_fill = (min_total + C - 1) / C;
    if (splits < by_fill) splits = by_fill;
    const long long min_units = (long long)threads * unroll * tn.min_iters;
    long long max_splits = g.chan_units / min_units;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    // whole waves: when a channel is cut into more tiles than fit on the machine at once, round
    // the count up to a multiple of the resident CTA slots so the last wave is full
    if (C == 1) {
        const long long resident = (long long)tn.sm_count * (kind == K_BWD ? tn.bwd_resident : (kind == K_FWD ? tn.fwd_resident : kResidentStats));
        if (tn.whole_waves && splits > resident) {
            const long long r = (splits + resident - 1) / resident * resident;
            if (r <= max_splits) splits = r;
        }
    }
    if (kind == K_FWD) {
        // no reduction: nothing limits the split count but launch granularity
    } else {
        if (C > kMaxCounters) splits = 1;
        while (splits > 1 && C * splits > kMaxSplitTiles) splits--;
    }
    if (splits < 1) splits = 1;
    long long ups = (g.chan_units + splits - 1) / splits;
    if (ups < 1) ups = 1;
    // whole group iterations per split keep every split's access pattern identical
    // warp groups own whole short channels (conv-weight rows, small-map activations); anything
    // longer is streamed by CTA groups
    g.group = (g.chan_units <= tn.warp_units) ? 32 : threads;
    if (g.group == 32) ups = g.chan_units > 0 ? g.chan_units : 1;
    const long long q = (long long)g.group;
    ups = (ups + q - 1) / q * q;
    splits = (g.chan_units + ups - 1) / ups;
    if (splits < 1) splits = 1;
    g.units_per_split = ups;
    g.interleave = (tn.interleave && splits > 1) ? 1 : 0;
    g.splits = (int)splits;
    g.tiles = C * splits;
    const long long gpc = (long long)(threads / g.group) * tiles_per_group(g.group);
    g.grid = (g.tiles + gpc - 1) / gpc;
    return g;
}

struct SegArgs {
    const void* x2;        // second addend of the ADD prologues (else nullptr)
    const void* x; void* y; const void* g; void* gx;
    const void* scale; const void* shift; void* gscale; void* gshift;
    float* stats_out;
    long long outer, C, inner;
    int xdtype, pdtype, per_channel;
    long long qmin, qmax, tmin, tmax;
    double grad_scaler;
    int use_grad_scaling, sym;
};

// largest power of two (<= 32) dividing every non-null pointer
inline int common_alignment(std::initializer_list<const void*> ps) {
    uintptr_t bits = 32;
    for (const void* p : ps) if (p) bits |= reinterpret_cast<uintptr_t>(p);
    return (int)(bits & (~bits + 1));
}

inline Seg make_seg(const SegArgs& a, const Geometry& g, double* partials, unsigned* counters, long long tile_begin) {
    Seg s;
    std::memset(&s, 0, sizeof(s));
    s.x = a.x; s.x2 = a.x2; s.y = a.y; s.g = a.g; s.gx = a.gx;
    s.scale = a.scale; s.shift = a.shift; s.gscale = a.gscale; s.gshift = a.gshift;
    s.partials = partials; s.counters = counters; s.stats_out = a.stats_out;
    s.C = g.C; s.vpr = g.vpr; s.row_stride = g.row_stride; s.chan_stride = g.vpr;
    s.inner = g.inner; s.chan_units = g.chan_units; s.units_per_split = g.units_per_split;
    s.tile_begin = tile_begin; s.chan_elems = g.outer * g.inner;
    const double numel = (double)a.outer * (double)a.C * (double)a.inner;
    s.gs = a.use_grad_scaling ? a.grad_scaler / std::sqrt(numel * (double)a.qmax) : a.grad_scaler;
    s.qmin = (float)a.qmin; s.qmax = (float)a.qmax; s.tmin = (float)a.tmin; s.tmax = (float)a.tmax;
    // bitness = ceil(log(qmax - qmin) / log 2) - 1   (observers.py:333)
    const int bitness = (int)std::ceil(std::log((double)(a.qmax - a.qmin)) / std::log(2.0)) - 1;
    s.stats_denom = (float)std::ldexp(1.0, bitness);
    s.splits = g.splits; s.regime = g.regime; s.per_channel = a.per_channel; s.pdt = a.pdtype;
    s.sym = a.sym; s.vec = g.vec; s.interleave = g.interleave; s.group = g.group;
    return s;
}

}  // namespace lsqb200
