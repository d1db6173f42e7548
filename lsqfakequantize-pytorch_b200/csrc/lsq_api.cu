// lsq_api.cu -- extern "C" entry points of libtorchlsq_b200.so (see include/lsq_b200.h).
// Host code only: argument checks, launch geometry, kernel selection, launches.  No allocation
// and no synchronisation on the per-call paths; plans allocate once at creation.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/lsq_b200.h"
#include "lsq_host.h"
#include "lsq_column.cuh"
#include "lsq_export.cuh"

using namespace lsqb200;

namespace {

thread_local std::string g_last_error;
Tuning g_tuning;
std::once_flag g_sm_once;

int fail(int code, const char* msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}

const Tuning& tuning() {
    std::call_once(g_sm_once, [] {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            g_tuning.sm_count = sms;
        if (const char* env = std::getenv("LSQB200_TUNE")) lsqb200_set_tuning(env);
    });
    return g_tuning;
}

// (x dtype, param dtype) -> arithmetic mode; returns -1 when the pair is not supported
int pick_mode_plain(int xdt, int pdt) {
    if (xdt == DT_F32 && pdt == DT_F32) return M_FP32;
    if (xdt == DT_F16 && pdt == DT_F32) return M_FP32;
    if (xdt == DT_F16 && pdt == DT_F16) return M_HALF_EXACT;   // reference-exact c10::Half semantics
    if (xdt == DT_BF16 && (pdt == DT_F32 || pdt == DT_BF16)) return M_FP32;
    if (xdt == DT_F64 && pdt == DT_F64) return M_F64;           // reference rule: scale.dtype == x.dtype (lsq_cuda.cu:34-35)
    return -1;
}
// prologue fusion (LSQB200_PRE_RELU) exists for the fp32-internal arithmetic only: fp32 / fp16 / bf16 tensors with
// fp32 (or, for bf16, bf16) parameters; the c10::Half-exact and float64 contracts mirror reference inputs and stay plain
int pick_mode(int xdt, int pdt, int prologue = LSQB200_PRE_NONE) {
    const int m = pick_mode_plain(xdt, pdt);
    if (prologue == LSQB200_PRE_NONE || m < 0) return m;
    if (m != M_FP32) return -1;
    if (prologue == LSQB200_PRE_RELU) return M_FP32_RELU;
    if (prologue == LSQB200_PRE_ADD_RELU) return M_FP32_ADD_RELU;
    if (prologue == LSQB200_PRE_ADD) return M_FP32_ADD;
    return -1;
}
bool prologue_known(int p) { return p == LSQB200_PRE_NONE || p == LSQB200_PRE_RELU || p == LSQB200_PRE_ADD_RELU || p == LSQB200_PRE_ADD; }
bool prologue_adds(int p) { return p == LSQB200_PRE_ADD_RELU || p == LSQB200_PRE_ADD; }
int check_prologue(int prologue, const void* x2, bool nonempty) {
    if (!prologue_known(prologue)) return fail(LSQB200_ERR_ARG, "unknown prologue");
    if (prologue_adds(prologue) && nonempty && !x2) return fail(LSQB200_ERR_ARG, "this prologue needs the second addend x2");
    return 0;
}

int check_q(const lsqb200_qargs* q) {
    if (!q) return fail(LSQB200_ERR_ARG, "qargs is NULL");
    if (q->quant_min >= q->quant_max) return fail(LSQB200_ERR_ARG, "quant_min must be smaller than quant_max");
    return 0;
}

int bmode_of(const lsqb200_qargs* q) {
    if (q->eval_mode) return q->init_mode ? B_EVAL_INIT : B_EVAL;
    return q->init_mode ? B_INIT : B_NORMAL;
}

SegArgs seg_args(const void* x, void* y, const void* g, void* gx, const void* scale, const void* shift,
                 void* gscale, void* gshift, int64_t outer, int64_t C, int64_t inner, int xdt, int pdt,
                 int per_channel, const lsqb200_qargs* q) {
    SegArgs a{};
    a.x = x; a.y = y; a.g = g; a.gx = gx; a.scale = scale; a.shift = shift; a.gscale = gscale; a.gshift = gshift;
    a.stats_out = nullptr;
    a.outer = outer; a.C = C; a.inner = inner; a.xdtype = xdt; a.pdtype = pdt; a.per_channel = per_channel;
    if (q) {
        a.qmin = q->quant_min; a.qmax = q->quant_max; a.tmin = q->type_min; a.tmax = q->type_max;
        a.grad_scaler = q->grad_scaler; a.use_grad_scaling = q->use_grad_scaling; a.sym = q->sym;
    }
    return a;
}

int launch(KernelFn k, const Seg& seg, const Seg* table, const int* tile_seg, int nseg, long long tiles, long long grid,
           cudaStream_t st, int smem_bytes = 0) {
    if (grid <= 0) return 0;
    if (grid > 2147483647LL) return fail(LSQB200_ERR_ARG, "tensor too large for one launch");
    if (smem_bytes > 48 * 1024) {     // opt in to large dynamic shared memory once per kernel
        static std::mutex mu;
        static std::map<const void*, int> done;
        std::lock_guard<std::mutex> lk(mu);
        if (done[(const void*)k] < smem_bytes) {
            cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(max dynamic shared memory)");
            done[(const void*)k] = smem_bytes;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap our prologue with the previous kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, seg, table, tile_seg, nseg, tiles);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return 0;
}

#ifdef LSQ_FLAT_TRACE
// diagnostic build only: launch i: header (elements, CTAs, 1 fwd / 2 bwd) + four stamps per CTA at buf + i * 8192 * 4 words (tools/flattrace.py); not part of include/lsq_b200.h
unsigned long long* g_trace_buf = nullptr;
long long g_trace_launches = 0, g_trace_next = 0;
}  // namespace
extern "C" __attribute__((visibility("default"))) long long lsqb200_trace_arm(void* buf, long long launches) {
    const long long used = g_trace_next;
    g_trace_buf = reinterpret_cast<unsigned long long*>(buf); g_trace_launches = launches; g_trace_next = 0;
    return used;
}
namespace {
#endif
int launch_flat(FlatKernelFn k, const Seg& seg_in, long long grid, cudaStream_t st) {
    if (grid <= 0) return 0;
#ifdef LSQ_FLAT_TRACE
    Seg seg = seg_in;
    seg.stats_out = nullptr;
    if (g_trace_buf && g_trace_next < g_trace_launches && grid < 8192) seg.stats_out = reinterpret_cast<float*>(g_trace_buf + (g_trace_next++) * 8192 * 4);
#else
    const Seg& seg = seg_in;
#endif
    if (grid > 2147483647LL) return fail(LSQB200_ERR_ARG, "tensor too large for one launch");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, seg);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return 0;
}

size_t param_size(int pdt) { return pdt == DT_F64 ? 8 : (pdt == DT_F32 ? 4 : 2); }

// ---- column-layout path (short channel rows: channels-last, 7x7 / 14x14 maps) ------------------
struct ColGeom { bool ok; int tx, ty; long long units_per_row, rows_per_split, col_blocks, row_splits; };

// resident CTAs per SM of a kernel (cached per function)
int occupancy_of(const void* fn, int threads) {
    static std::mutex mu;
    static std::map<const void*, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(fn);
    if (it != cache.end()) return it->second;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, 0) != cudaSuccess || occ < 1) occ = 2;
    cache[fn] = occ;
    return occ;
}

ColGeom plan_column(int64_t outer, int64_t C, int64_t inner, int xdt, int align_bytes, const Tuning& tn, int occ, bool backward,
                    bool relu = false) {
    ColGeom g{};
    const int ub = kColVariantNW[relu ? kColVariantRelu : (backward ? tn.col_variant : tn.col_variant_fwd)] * 4;
    const int es = elem_size(xdt), vec = ub / es;
    const long long L = C * inner;
    g.ok = tn.column_path && outer > 1 && C > 1 && C <= kMaxColumnChannels && L < (1LL << 31) && align_bytes % 16 == 0 && (L * es) % 16 == 0 &&
           ((inner * es) % 16 != 0 || inner * es < tn.column_max_row_bytes);
    if (!g.ok) return g;
    g.units_per_row = L / vec;
    int tx = 1;
    while (tx < kColThreads && tx < g.units_per_row) tx <<= 1;
    g.tx = tx; g.ty = kColThreads / tx;
    g.col_blocks = (g.units_per_row + tx - 1) / tx;
    // whole waves of CTAs (col_waves per resident slot); the per-thread set-up (2*VEC parameter loads,
    // VEC divisions) is amortised over every row a thread visits
    const long long slots = (long long)tn.sm_count * occ;
    // round DOWN: col_blocks * splits must not spill into a partly filled extra wave (686 CTAs on 2 x 296 slots ran 3 waves)
    long long splits = ((long long)(backward ? tn.col_waves_bwd : tn.col_waves) * slots) / g.col_blocks;
    if (splits < 1) splits = 1;
    long long rows_pt = (outer + splits * g.ty - 1) / (splits * g.ty);
    if (rows_pt < 2) rows_pt = 2;
    g.rows_per_split = rows_pt * g.ty;
    g.row_splits = (outer + g.rows_per_split - 1) / g.rows_per_split;
    if (g.row_splits > 65535) { g.rows_per_split = (outer + 65534) / 65535; g.row_splits = (outer + g.rows_per_split - 1) / g.rows_per_split; }
    return g;
}

ColSeg make_colseg(const ColGeom& g, const void* x, const void* x2, void* y, const void* grad, void* gx, const void* scale, const void* shift,
                   void* gscale, void* gshift, int64_t outer, int64_t C, int64_t inner, int pdt, const lsqb200_qargs* q,
                   void* workspace) {
    ColSeg cs{};
    cs.x = x; cs.x2 = x2; cs.y = y; cs.g = grad; cs.gx = gx; cs.scale = scale; cs.shift = shift; cs.gscale = gscale; cs.gshift = gshift;
    if (workspace) {
        cs.counter = reinterpret_cast<unsigned*>(workspace);
        cs.acc = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + kColAccOffset);
    }
    cs.outer = outer; cs.C = C; cs.inner = inner; cs.units_per_row = g.units_per_row; cs.rows_per_split = g.rows_per_split;
    const double numel = (double)outer * (double)C * (double)inner;
    cs.gs = q->use_grad_scaling ? q->grad_scaler / std::sqrt(numel * (double)q->quant_max) : q->grad_scaler;
    cs.qmin = (float)q->quant_min; cs.qmax = (float)q->quant_max; cs.tmin = (float)q->type_min; cs.tmax = (float)q->type_max;
    cs.tx = g.tx; cs.ty = g.ty; cs.pdt = pdt; cs.sym = q->sym;
    cs.total_ctas = (unsigned)(g.col_blocks * g.row_splits);
    cs.l2_prefetch = tuning().l2_prefetch;
    return cs;
}

int launch_col(ColKernelFn k, const ColSeg& cs, const ColGeom& g, cudaStream_t st, int threads = kColThreads, int smem_bytes = 0) {
    if (smem_bytes > 48 * 1024) {     // opt in to large dynamic shared memory once per kernel
        static std::mutex mu;
        static std::map<const void*, int> done;
        std::lock_guard<std::mutex> lk(mu);
        if (done[(const void*)k] < smem_bytes) {
            cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(max dynamic shared memory)");
            done[(const void*)k] = smem_bytes;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)g.col_blocks, (unsigned)g.row_splits);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tuning().pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, cs);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return 0;
}

int forward_common(const void* x, const void* x2, void* y, const void* scale, const void* shift, int64_t outer, int64_t C,
                   int64_t inner, int xdt, int pdt, int per_channel, const lsqb200_qargs* q, int prologue, void* stream) {
    if (int r = check_q(q)) return r;
    if (int r = check_prologue(prologue, x2, outer > 0 && C > 0 && inner > 0)) return r;
    if (!prologue_adds(prologue)) x2 = nullptr;
    if (outer < 0 || C < 0 || inner < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (xdt < 0 || xdt > 3 || pdt < 0 || pdt > 3) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    const int mode = pick_mode(xdt, pdt, prologue);
    if (mode < 0) return fail(LSQB200_ERR_DTYPE, prologue ? "a fused prologue needs float32 / float16 / bfloat16 tensors with float32 scale / shift"
                                                          : "unsupported (x dtype, scale/shift dtype) pair");
    if (outer * C * inner == 0) return 0;
    if (!x || !y || !scale || !shift) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    if (xdt == DT_F64 && common_alignment({x, y, scale, shift}) < 8) return fail(LSQB200_ERR_ARG, "float64 tensors must be 8-byte aligned");
    if (per_channel && xdt != DT_F64) {
        ColKernelFn ck = get_col_fwd_kernel(xdt, mode, q->init_mode != 0, tuning().col_variant_fwd);
        const ColGeom cg = plan_column(outer, C, inner, xdt, common_alignment({x, x2, y}), tuning(), occupancy_of((const void*)ck, kColThreads), false, mode_relu(mode) || mode_add(mode));
        if (cg.ok) {
            const ColSeg cs = make_colseg(cg, x, x2, y, nullptr, nullptr, scale, shift, nullptr, nullptr, outer, C, inner, pdt, q, nullptr);
            return launch_col(ck, cs, cg, (cudaStream_t)stream);
        }
    }
    const Geometry g = plan_geometry(outer, C, inner, xdt, K_FWD, common_alignment({x, x2, y}), tuning_for_mode(tuning(), mode));
    SegArgs a = seg_args(x, y, nullptr, nullptr, scale, shift, nullptr, nullptr, outer, C, inner, xdt, pdt, per_channel, q);
    a.x2 = x2;
    Seg seg = make_seg(a, g, nullptr, nullptr, 0);
    seg.flags = tuning().l2_prefetch;
    if (mode == M_FP32 && tuning().flatkernels && flat_eligible(g, xdt))
        return launch_flat(get_flatfwd_kernel(xdt, q->init_mode != 0), seg, g.grid, (cudaStream_t)stream);
    KernelFn k = (mode == M_FP32 && tuning().rowkernels && row_kernels_eligible(g, xdt)) ? get_rowfwd_kernel(xdt, q->init_mode != 0)
                                                                                      : get_fwd_kernel(xdt, mode, g.nw, q->init_mode != 0, g.group);
    return launch(k, seg, nullptr, nullptr, 0, g.tiles, g.grid, (cudaStream_t)stream);
}

int backward_common(const void* grad, const void* x, const void* x2, void* gx, const void* scale, const void* shift, void* gscale,
                    void* gshift, int64_t outer, int64_t C, int64_t inner, int xdt, int pdt, int per_channel,
                    const lsqb200_qargs* q, int prologue, void* workspace, size_t wbytes, void* stream) {
    if (int r = check_q(q)) return r;
    if (int r = check_prologue(prologue, x2, outer > 0 && C > 0 && inner > 0)) return r;
    if (!prologue_adds(prologue)) x2 = nullptr;
    if (outer < 0 || C < 0 || inner < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (xdt < 0 || xdt > 3 || pdt < 0 || pdt > 3) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    const int mode = pick_mode(xdt, pdt, prologue);
    if (mode < 0) return fail(LSQB200_ERR_DTYPE, prologue ? "a fused prologue needs float32 / float16 / bfloat16 tensors with float32 scale / shift"
                                                          : "unsupported (x dtype, scale/shift dtype) pair");
    if (!gscale || !gshift) return fail(LSQB200_ERR_ARG, "NULL grad_scale / grad_shift pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nslot = per_channel ? C : 1;
    if (outer * C * inner == 0) {   // empty input: defined, all-zero parameter grads (reference returns its inputs, D13)
        if (nslot > 0) {
            cudaError_t e = cudaMemsetAsync(gscale, 0, nslot * param_size(pdt), st);
            if (e == cudaSuccess) e = cudaMemsetAsync(gshift, 0, nslot * param_size(pdt), st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        }
        return 0;
    }
    if (!grad || !x || !scale || !shift) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    if (xdt == DT_F64 && common_alignment({x, grad, gx, scale, shift, gscale, gshift}) < 8)
        return fail(LSQB200_ERR_ARG, "float64 tensors must be 8-byte aligned");
    if (per_channel && xdt != DT_F64) {
        ColKernelFn ck = get_col_bwd_kernel(xdt, mode, bmode_of(q), tuning().col_variant);
        const ColGeom cg = plan_column(outer, C, inner, xdt, common_alignment({x, x2, grad, gx}), tuning(), occupancy_of((const void*)ck, kColThreads), true, mode_relu(mode) || mode_add(mode));
        if (cg.ok) {
            if (!workspace || wbytes < kWorkspaceBytes || (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0)
                return fail(LSQB200_ERR_WORKSPACE, "workspace missing, misaligned or smaller than lsqb200_workspace_bytes()");
            const long long upr16 = C * inner * elem_size(xdt) / 16;
            if (tuning().col_tma > 0 && upr16 >= kTmaConsumers && !mode_relu(mode) && !mode_add(mode)) {
                // TMA-staged variant: a CTA owns 256 column units (4 KB of every row) and a contiguous run of rows
                int smem = 0;
                ColKernelFn tk = get_col_bwd_tma_kernel(xdt, mode, bmode_of(q), tuning().col_tma, &smem);
                ColGeom tg = cg;
                tg.units_per_row = upr16; tg.tx = kTmaConsumers; tg.ty = 1;
                tg.col_blocks = (upr16 + kTmaConsumers - 1) / kTmaConsumers;
                const int occ = smem <= 110 * 1024 ? 2 : 1;
                long long splits = ((long long)tuning().col_waves_bwd * tuning().sm_count * occ) / tg.col_blocks;
                if (splits < 1) splits = 1;
                long long rows = (outer + splits - 1) / splits;
                if (rows < 2) rows = 2;
                tg.rows_per_split = rows;
                tg.row_splits = (outer + rows - 1) / rows;
                if (tg.row_splits > 65535) { tg.rows_per_split = (outer + 65534) / 65535; tg.row_splits = (outer + tg.rows_per_split - 1) / tg.rows_per_split; }
                const ColSeg cs = make_colseg(tg, x, nullptr, nullptr, grad, gx, scale, shift, gscale, gshift, outer, C, inner, pdt, q, workspace);
                return launch_col(tk, cs, tg, st, kTmaThreads, smem);
            }
            const ColSeg cs = make_colseg(cg, x, x2, nullptr, grad, gx, scale, shift, gscale, gshift, outer, C, inner, pdt, q, workspace);
            return launch_col(ck, cs, cg, st);
        }
    }
    const Geometry g = plan_geometry(outer, C, inner, xdt, K_BWD, common_alignment({x, x2, grad, gx}), tuning_for_mode(tuning(), mode));
    double* partials = nullptr;
    unsigned* counters = nullptr;
    if (g.splits > 1) {
        if (!workspace || wbytes < kWorkspaceBytes || (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0)
            return fail(LSQB200_ERR_WORKSPACE, "workspace missing, misaligned or smaller than lsqb200_workspace_bytes()");
        counters = reinterpret_cast<unsigned*>(workspace);
        partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + kMaxCounters * 4);
    }
    SegArgs a = seg_args(x, nullptr, grad, gx, scale, shift, gscale, gshift, outer, C, inner, xdt, pdt, per_channel, q);
    a.x2 = x2;
    Seg seg = make_seg(a, g, partials, counters, 0);
    seg.flags = tuning().l2_prefetch;
    if (mode == M_FP32 && tuning().flatkernels >= 2 && flat_eligible(g, xdt))
        return launch_flat(get_flatbwd_kernel(xdt, bmode_of(q)), seg, g.grid, st);
    KernelFn k = (mode == M_FP32 && tuning().rowkernels && row_kernels_eligible(g, xdt)) ? get_rowbwd_kernel(xdt, bmode_of(q))
                                                                                      : get_bwd_kernel(xdt, mode, g.nw, bmode_of(q), g.group);
    return launch(k, seg, nullptr, nullptr, 0, g.tiles, g.grid, st);
}

int export_common(int dir, const void* fl, void* codes, const void* scale, const void* shift, int64_t outer, int64_t C,
                  int64_t inner, int xdt, int pdt, int per_channel, const lsqb200_qargs* q, int codes_signed, int sem,
                  void* stream) {
    if (int r = check_q(q)) return r;
    if (outer < 0 || C < 0 || inner < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (xdt < 0 || xdt > 2 || pdt < 0 || pdt > 2) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    if (sem != SEM_LSQ && sem != SEM_TORCH && sem != SEM_TORCH_CPU) return fail(LSQB200_ERR_ARG, "unknown code semantics");
    const int mode = pick_mode(xdt, pdt);
    if (mode < 0) return fail(LSQB200_ERR_DTYPE, "unsupported (x dtype, scale/shift dtype) pair");
    if (sem != SEM_LSQ && pdt != DT_F32) return fail(LSQB200_ERR_DTYPE, "torch semantics need float32 scale / shift");
    if (q->type_min < -128 || q->type_max > 255 || q->quant_min < -128 || q->quant_max > 255 ||
        (codes_signed ? (q->quant_max > 127 || (sem != SEM_LSQ && q->type_max > 127))
                      : (q->quant_min < 0 || (sem != SEM_LSQ && q->type_min < 0))))
        return fail(LSQB200_ERR_ARG, "range does not fit the 8-bit code type");
    if (outer * C * inner == 0) return 0;
    if (!fl || !codes || !scale || !shift) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    // a unit of the floating tensor (<= 32 B) covers unit/elem_size code bytes: both sides must be aligned to their unit
    const int es = elem_size(xdt);
    int al = common_alignment({fl});
    const int alc = common_alignment({codes}) * es;
    if (alc < al) al = alc;
    const Geometry g = plan_geometry(outer, C, inner, xdt, K_FWD, al, tuning());
    SegArgs a = seg_args(fl, codes, nullptr, nullptr, scale, shift, nullptr, nullptr, outer, C, inner, xdt, pdt, per_channel, q);
    Seg seg = make_seg(a, g, nullptr, nullptr, 0);
    seg.code_signed = codes_signed ? 1 : 0;
    KernelFn k = get_export_kernel(xdt, mode, g.nw, sem, dir, g.group);
    return launch(k, seg, nullptr, nullptr, 0, g.tiles, g.grid, (cudaStream_t)stream);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
struct lsqb200_plan {
    struct Class {
        KernelFn kernel = nullptr;
        std::vector<Seg> host;
        Seg* dev = nullptr;
        int* dev_tile_seg = nullptr;          // tile -> index into `dev`
        long long tiles = 0, grid = 0;
        int group = kThreads;
        bool rowtable = false;                // dev_tile_seg holds one RowEntry per tile instead of a tile -> segment map (lsq_rowstats3_kernel)
        int smem = 0;                         // dynamic shared memory of the launch (lsq_rowstats_ring_kernel)
        bool resident = false;                // one CTA per SM walks the tiles (lsq_rowstats_ring_kernel)
    };
    std::vector<Class> fwd, bwd, stats;
    std::vector<lsqb200_segment> segs;
    std::vector<long long> stats_offset;   // running sum of C (or 1) per segment
    void* workspace = nullptr;             // counters + partials for every split segment
    float** stats_slot = nullptr;
    const float* stats_uploaded_for = nullptr;   // output buffer the device-side statistics tables currently point at
    int device = 0;
};

namespace {

using ClassKey = std::tuple<int, int, int, int, int>;   // xdtype, mode, vec?, variant (init / bmode), group

int build_classes(lsqb200_plan* p, int kind, std::vector<lsqb200_plan::Class>& out, char* ws_base, size_t& ws_used,
                  bool assign_ws) {
    std::map<ClassKey, size_t> index;
    const Tuning& tn = tuning();
    for (size_t i = 0; i < p->segs.size(); i++) {
        const lsqb200_segment& s = p->segs[i];
        const int mode = pick_mode(s.xdtype, s.pdtype, s.prologue);
        if (mode < 0) return fail(LSQB200_ERR_DTYPE, "unsupported (x dtype, scale/shift dtype, prologue) combination in plan");
        if (s.outer * s.C * s.inner == 0) continue;
        int al = common_alignment({s.x});
        const void* sx2 = prologue_adds(s.prologue) ? s.x2 : nullptr;
        if (kind == K_FWD) al = common_alignment({s.x, sx2, s.y});
        if (kind == K_BWD) al = common_alignment({s.x, sx2, s.grad, s.gx});
        Geometry g = plan_geometry(s.outer, s.C, s.inner, s.xdtype, kind, al, tuning_for_mode(tn, mode));
        int variant = 0;
        KernelFn k = nullptr;
        const bool lean = mode == M_FP32 && tn.rowkernels && row_kernels_eligible(g, s.xdtype);   // aligned weight rows: warp-per-row kernels
        if (kind == K_FWD) {
            variant = s.q.init_mode != 0;
            k = lean ? get_rowfwd_kernel(s.xdtype, variant != 0) : get_fwd_kernel(s.xdtype, mode, g.nw, variant, g.group);
            if (lean) variant += 16;
        } else if (kind == K_BWD) {
            variant = bmode_of(&s.q);
            k = lean ? get_rowbwd_kernel(s.xdtype, variant) : get_bwd_kernel(s.xdtype, mode, g.nw, variant, g.group);
            if (lean) variant += 16;
        }
        else {
            if (s.xdtype == DT_F64) continue;   // no float64 statistics (the module cannot hold float64 weights): slot left untouched
            if (row_kernels_eligible(g, s.xdtype) && tn.rowstats) {
                // rows that are whole 32-byte units take the bulk-copy ring (rowstats = 7); others the register-staged row kernels
                const bool ring = tn.rowstats == 7 && (s.inner * (long long)elem_size(s.xdtype)) % 32 == 0;
                variant = ring ? 3 : (tn.rowstats >= 4 ? 2 : 1);
                k = ring ? get_rowstats_ring_kernel(s.xdtype) : get_rowstats_kernel(s.xdtype, tn.rowstats == 7 ? 4 : tn.rowstats);
            }
            else k = get_stats_kernel(s.xdtype, g.nw, g.group);
        }
        if (!k) return fail(LSQB200_ERR_ARG, "float64 tensors must be 8-byte aligned");
        const ClassKey key{s.xdtype, kind == K_STATS ? 0 : mode, g.nw, variant, g.group};
        auto it = index.find(key);
        if (it == index.end()) {
            it = index.emplace(key, out.size()).first;
            out.emplace_back();
            out.back().kernel = k;
            out.back().group = g.group;
            out.back().rowtable = kind == K_STATS && variant >= 2;
            if (kind == K_STATS && variant == 3) { out.back().smem = rowring::kSmemBytes; out.back().resident = true; }
        }
        lsqb200_plan::Class& c = out[it->second];
        double* partials = nullptr;
        unsigned* counters = nullptr;
        if (kind != K_FWD && g.splits > 1) {
            const size_t need = (size_t)g.C * 4 + 16 + (size_t)g.tiles * 16;
            if (assign_ws) {
                counters = reinterpret_cast<unsigned*>(ws_base + ws_used);
                partials = reinterpret_cast<double*>(ws_base + ws_used + (((size_t)g.C * 4 + 15) / 16) * 16);
            }
            ws_used += ((need + 15) / 16) * 16;
        }
        SegArgs a = seg_args(s.x, s.y, s.grad, s.gx, s.scale, s.shift, s.gscale, s.gshift, s.outer, s.C, s.inner,
                             s.xdtype, s.pdtype, s.per_channel, &s.q);
        a.x2 = sx2;
        Seg seg = make_seg(a, g, partials, counters, c.tiles);
        seg.flags = tn.l2_prefetch;
        seg.stats_out = nullptr;   // patched per run for stats
        // tag for stats runs: remember which public segment this is (reuse chan_stride, unused by kernels)
        seg.chan_stride = (long long)i;
        c.host.push_back(seg);
        c.tiles += g.tiles;
    }
    for (auto& c : out) {
        const long long gpc = (long long)(kThreads / c.group) * tiles_per_group(c.group);
        c.grid = (c.tiles + gpc - 1) / gpc;
        if (c.resident && c.grid > (long long)tn.sm_count * rowring::kCtasPerSm) c.grid = (long long)tn.sm_count * rowring::kCtasPerSm;
    }
    return 0;
}

// one 32-byte entry per row for lsq_rowstats3_kernel: row pointer, output slot relative to the class's first segment, length, denominator
int upload_row_table(lsqb200_plan::Class& c, const lsqb200_plan* p) {
    if (c.dev_tile_seg || c.tiles <= 0) return 0;
    std::vector<RowEntry> rows((size_t)c.tiles);
    const long long base = p->stats_offset[(size_t)c.host[0].chan_stride];
    for (size_t i = 0; i < c.host.size(); i++) {
        const Seg& h = c.host[i];
        const lsqb200_segment& pub = p->segs[(size_t)h.chan_stride];
        const long long b = h.tile_begin, e = (i + 1 < c.host.size()) ? c.host[i + 1].tile_begin : c.tiles;
        const long long es = elem_size(pub.xdtype);
        for (long long t = b; t < e; t++) {
            RowEntry& r = rows[(size_t)t];
            r.row = static_cast<const char*>(h.x) + (t - b) * h.inner * es;
            r.out_rel = p->stats_offset[(size_t)h.chan_stride] + (t - b) - base;
            r.inner = (int)h.inner; r.denom = h.stats_denom; r.pad[0] = r.pad[1] = 0;
        }
    }
    cudaError_t e = cudaMalloc(&c.dev_tile_seg, rows.size() * sizeof(RowEntry));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(plan row table)");
    e = cudaMemcpy(c.dev_tile_seg, rows.data(), rows.size() * sizeof(RowEntry), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(plan row table)");
    return 0;
}

int upload_tile_map(lsqb200_plan::Class& c) {
    if (c.dev_tile_seg || c.tiles <= 0 || c.tiles > (1LL << 26)) return 0;   // huge plans keep the in-kernel search
    std::vector<int> map((size_t)c.tiles);
    for (size_t i = 0; i < c.host.size(); i++) {
        const long long b = c.host[i].tile_begin, e = (i + 1 < c.host.size()) ? c.host[i + 1].tile_begin : c.tiles;
        for (long long t = b; t < e; t++) map[(size_t)t] = (int)i;
    }
    cudaError_t e = cudaMalloc(&c.dev_tile_seg, map.size() * sizeof(int));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(plan tile map)");
    e = cudaMemcpy(c.dev_tile_seg, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(plan tile map)");
    return 0;
}

int upload_classes(std::vector<lsqb200_plan::Class>& cls) {
    for (auto& c : cls) {
        if (c.host.empty()) continue;
        cudaError_t e = cudaMalloc(&c.dev, c.host.size() * sizeof(Seg));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(plan table)");
        e = cudaMemcpy(c.dev, c.host.data(), c.host.size() * sizeof(Seg), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(plan table)");
        if (int r = upload_tile_map(c)) return r;
    }
    return 0;
}

int run_classes(std::vector<lsqb200_plan::Class>& cls, cudaStream_t st) {
    for (auto& c : cls) {
        if (c.host.empty()) continue;
        if (int r = launch(c.kernel, c.host[0], c.dev, c.dev_tile_seg, (int)c.host.size(), c.tiles, c.grid, st, c.smem)) return r;
    }
    return 0;
}

// ---- re-pointing a plan's per-step tensors: the new pointers travel as KERNEL ARGUMENTS of a one-CTA patch kernel (truly
//      asynchronous, stream ordered, no pinned staging and no pageable cudaMemcpyAsync whose host side may wait for the stream).
//      It is launched without the PDL attribute and never triggers its dependents early, so it starts after every earlier
//      kernel that reads the table has finished and the next plan kernel starts after its writes are visible.
constexpr int kPatchMax = 300;
enum { PF_Y = 0, PF_G = 1, PF_GX = 2, PF_GSCALE = 3, PF_GSHIFT = 4 };
struct PatchArgs {
    Seg* table;
    int n;
    unsigned short seg[kPatchMax];
    unsigned char field[kPatchMax];
    const void* ptr[kPatchMax];
};
static_assert(sizeof(PatchArgs) <= 4000, "patch arguments must fit the 4 KB kernel parameter space");

__global__ void lsq_plan_patch_kernel(const PatchArgs a) {
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        Seg& s = a.table[a.seg[i]];
        void* p = const_cast<void*>(a.ptr[i]);
        switch (a.field[i]) {
            case PF_Y: s.y = p; break;
            case PF_G: s.g = p; break;
            case PF_GX: s.gx = p; break;
            case PF_GSCALE: s.gscale = p; break;
            default: s.gshift = p; break;
        }
    }
}

struct Patcher {
    PatchArgs a;
    cudaStream_t st;
    int rc = 0;
    Patcher(Seg* table, cudaStream_t s) : st(s) { a.table = table; a.n = 0; }
    void flush() {
        if (a.n == 0 || rc) return;
        lsq_plan_patch_kernel<<<1, 128, 0, st>>>(a);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "plan patch launch");
        a.n = 0;
    }
    void add(size_t seg, int field, const void* p) {
        a.seg[a.n] = (unsigned short)seg; a.field[a.n] = (unsigned char)field; a.ptr[a.n] = p;
        if (++a.n == kPatchMax) flush();
    }
};

// same-or-better alignment than the pointer the geometry (unit width) was planned with
bool keeps_alignment(const void* now, const void* before) { return common_alignment({now}) >= common_alignment({before}); }

}  // namespace

extern "C" {

int lsqb200_plan_rebind(lsqb200_plan* plan, void* const* y, const void* const* grad, void* const* gx, void* const* gscale,
                        void* const* gshift, void* stream) {
    if (!plan) return fail(LSQB200_ERR_PLAN, "NULL plan");
    const size_t n = plan->segs.size();
    for (size_t i = 0; i < n; i++) {
        lsqb200_segment& s = plan->segs[i];
        if (s.outer * s.C * s.inner == 0) continue;
        if ((y && (!y[i] || !keeps_alignment(y[i], s.y))) || (grad && (!grad[i] || !keeps_alignment(grad[i], s.grad))) ||
            (gx && (!gx[i] || !s.gx || !keeps_alignment(gx[i], s.gx))) ||
            (gscale && (!gscale[i] || !keeps_alignment(gscale[i], s.gscale))) || (gshift && (!gshift[i] || !keeps_alignment(gshift[i], s.gshift))))
            return fail(LSQB200_ERR_PLAN, "rebind: NULL pointer, or less aligned than the pointer the plan was created with");
    }
    for (int pass = 0; pass < 2; pass++) {
        auto& classes = pass == 0 ? plan->fwd : plan->bwd;
        if (pass == 0 ? !y : !(grad || gx || gscale || gshift)) continue;
        for (auto& c : classes) {
            if (c.host.empty() || !c.dev) continue;
            if (c.host.size() > 65535) return fail(LSQB200_ERR_PLAN, "rebind: more than 65535 segments in one launch class");
            Patcher pt(c.dev, (cudaStream_t)stream);
            for (size_t j = 0; j < c.host.size(); j++) {
                Seg& h = c.host[j];
                const size_t i = (size_t)h.chan_stride;      // public segment index (see build_classes)
                if (pass == 0) { h.y = y[i]; pt.add(j, PF_Y, y[i]); continue; }
                if (grad) { h.g = grad[i]; pt.add(j, PF_G, grad[i]); }
                if (gx) { h.gx = gx[i]; pt.add(j, PF_GX, gx[i]); }
                if (gscale) { h.gscale = gscale[i]; pt.add(j, PF_GSCALE, gscale[i]); }
                if (gshift) { h.gshift = gshift[i]; pt.add(j, PF_GSHIFT, gshift[i]); }
            }
            pt.flush();
            if (pt.rc) return pt.rc;
        }
    }
    for (size_t i = 0; i < n; i++) {
        lsqb200_segment& s = plan->segs[i];
        if (y) s.y = y[i];
        if (grad) s.grad = grad[i];
        if (gx) s.gx = gx[i];
        if (gscale) s.gscale = gscale[i];
        if (gshift) s.gshift = gshift[i];
    }
    return 0;
}

int lsqb200_abi_version(void) { return LSQB200_ABI_VERSION; }
int64_t lsqb200_cuda_version(void) { return (int64_t)CUDA_VERSION; }
const char* lsqb200_last_error(void) { return g_last_error.c_str(); }
size_t lsqb200_workspace_bytes(void) { return kWorkspaceBytes; }

int lsqb200_set_tuning(const char* spec) {
    const int sms = g_tuning.sm_count;
    g_tuning = Tuning();
    g_tuning.sm_count = sms;
    if (!spec || !*spec) return 0;
    std::string s(spec);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(',', pos);
        if (end == std::string::npos) end = s.size();
        const std::string kv = s.substr(pos, end - pos);
        const size_t eq = kv.find('=');
        if (eq == std::string::npos) return fail(LSQB200_ERR_ARG, "tuning spec: expected key=value");
        const std::string k = kv.substr(0, eq);
        const int v = std::atoi(kv.c_str() + eq + 1);
        if (k == "fwd_tile_kb") g_tuning.fwd_tile_kb = v;
        else if (k == "bwd_tile_kb") g_tuning.bwd_tile_kb = v;
        else if (k == "stats_tile_kb") g_tuning.stats_tile_kb = v;
        else if (k == "fwd_min_tiles_per_sm") g_tuning.fwd_min_tiles_per_sm = v;
        else if (k == "bwd_min_tiles_per_sm") g_tuning.bwd_min_tiles_per_sm = v;
        else if (k == "warp_units") g_tuning.warp_units = v;
        else if (k == "min_iters") g_tuning.min_iters = v;
        else if (k == "sm_count") g_tuning.sm_count = v;
        else if (k == "max_unit_bytes") g_tuning.max_unit_bytes = v;
        else if (k == "interleave") g_tuning.interleave = v;
        else if (k == "pdl") g_tuning.pdl = v;
        else if (k == "l2_prefetch") g_tuning.l2_prefetch = (v >= 0 && v <= 8) ? v : 0;
        else if (k == "whole_waves") g_tuning.whole_waves = v;
        else if (k == "column_path") g_tuning.column_path = v;
        else if (k == "col_variant") g_tuning.col_variant = g_tuning.col_variant_fwd = (v >= 0 && v < kColVariants) ? v : 0;   // both directions
        else if (k == "col_variant_fwd") g_tuning.col_variant_fwd = (v >= 0 && v < kColVariants) ? v : 0;
        else if (k == "col_variant_bwd") g_tuning.col_variant = (v >= 0 && v < kColVariants) ? v : 0;
        else if (k == "col_waves") g_tuning.col_waves = v > 0 ? v : 2;
        else if (k == "col_waves_bwd") g_tuning.col_waves_bwd = v > 0 ? v : 1;
        else if (k == "col_tma") g_tuning.col_tma = (v >= 0 && v <= 3) ? v : 0;
        else if (k == "column_max_row_bytes") g_tuning.column_max_row_bytes = v;
        else if (k == "rowkernels") g_tuning.rowkernels = v;
        else if (k == "flatkernels") g_tuning.flatkernels = v;
        else if (k == "rowstats") g_tuning.rowstats = (v >= 0 && v <= 7) ? v : 4;
        else return fail(LSQB200_ERR_ARG, "tuning spec: unknown key");
        pos = end + 1;
    }
    if (g_tuning.fwd_tile_kb < 1 || g_tuning.bwd_tile_kb < 1 || g_tuning.stats_tile_kb < 1 || g_tuning.min_iters < 1 ||
        g_tuning.sm_count < 1 || g_tuning.warp_units < 0 || g_tuning.fwd_min_tiles_per_sm < 1 || g_tuning.bwd_min_tiles_per_sm < 1) {
        g_tuning = Tuning();
        g_tuning.sm_count = sms;
        return fail(LSQB200_ERR_ARG, "tuning spec: value out of range");
    }
    return 0;
}

int lsqb200_query_launch(int64_t outer, int64_t C, int64_t inner, int xdtype, int backward, int aligned16,
                         lsqb200_launch_info* out) {
    if (!out || outer < 0 || C < 0 || inner < 0 || xdtype < 0 || xdtype > 2) return fail(LSQB200_ERR_ARG, "bad argument");
    const Geometry g = plan_geometry(outer, C, inner, xdtype, backward ? K_BWD : K_FWD, aligned16 ? 32 : 2, tuning());
    out->regime = g.regime; out->vec = g.vec; out->threads = kThreads; out->splits = g.splits;
    out->grid = g.grid; out->units_per_split = g.units_per_split;
    return 0;
}

int lsqb200_fwd_tensor(const void* x, void* y, const void* scale, const void* shift, int64_t numel, int xdtype,
                       int pdtype, const lsqb200_qargs* q, void* stream) {
    return forward_common(x, nullptr, y, scale, shift, 1, 1, numel, xdtype, pdtype, 0, q, LSQB200_PRE_NONE, stream);
}
int lsqb200_fwd_tensor_pre(const void* x, const void* x2, void* y, const void* scale, const void* shift, int64_t numel, int xdtype,
                           int pdtype, const lsqb200_qargs* q, int prologue, void* stream) {
    return forward_common(x, x2, y, scale, shift, 1, 1, numel, xdtype, pdtype, 0, q, prologue, stream);
}

int lsqb200_bwd_tensor(const void* grad, const void* x, void* gx, const void* scale, const void* shift, void* gscale,
                       void* gshift, int64_t numel, int xdtype, int pdtype, const lsqb200_qargs* q, void* workspace,
                       size_t workspace_bytes, void* stream) {
    return backward_common(grad, x, nullptr, gx, scale, shift, gscale, gshift, 1, 1, numel, xdtype, pdtype, 0, q, LSQB200_PRE_NONE, workspace,
                           workspace_bytes, stream);
}
int lsqb200_bwd_tensor_pre(const void* grad, const void* x, const void* x2, void* gx, const void* scale, const void* shift, void* gscale,
                           void* gshift, int64_t numel, int xdtype, int pdtype, const lsqb200_qargs* q, int prologue,
                           void* workspace, size_t workspace_bytes, void* stream) {
    return backward_common(grad, x, x2, gx, scale, shift, gscale, gshift, 1, 1, numel, xdtype, pdtype, 0, q, prologue, workspace,
                           workspace_bytes, stream);
}

int lsqb200_fwd_channel(const void* x, void* y, const void* scale, const void* shift, int64_t outer, int64_t C,
                        int64_t inner, int xdtype, int pdtype, const lsqb200_qargs* q, void* stream) {
    return forward_common(x, nullptr, y, scale, shift, outer, C, inner, xdtype, pdtype, 1, q, LSQB200_PRE_NONE, stream);
}
int lsqb200_fwd_channel_pre(const void* x, const void* x2, void* y, const void* scale, const void* shift, int64_t outer, int64_t C,
                            int64_t inner, int xdtype, int pdtype, const lsqb200_qargs* q, int prologue, void* stream) {
    return forward_common(x, x2, y, scale, shift, outer, C, inner, xdtype, pdtype, 1, q, prologue, stream);
}

int lsqb200_bwd_channel(const void* grad, const void* x, void* gx, const void* scale, const void* shift, void* gscale,
                        void* gshift, int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                        const lsqb200_qargs* q, void* workspace, size_t workspace_bytes, void* stream) {
    return backward_common(grad, x, nullptr, gx, scale, shift, gscale, gshift, outer, C, inner, xdtype, pdtype, 1, q, LSQB200_PRE_NONE, workspace,
                           workspace_bytes, stream);
}
int lsqb200_bwd_channel_pre(const void* grad, const void* x, const void* x2, void* gx, const void* scale, const void* shift, void* gscale,
                            void* gshift, int64_t outer, int64_t C, int64_t inner, int xdtype, int pdtype,
                            const lsqb200_qargs* q, int prologue, void* workspace, size_t workspace_bytes, void* stream) {
    return backward_common(grad, x, x2, gx, scale, shift, gscale, gshift, outer, C, inner, xdtype, pdtype, 1, q, prologue, workspace,
                           workspace_bytes, stream);
}

int lsqb200_weight_init_stats(const void* w, float* scale_out, int64_t outer, int64_t C, int64_t inner, int xdtype,
                              int64_t quant_min, int64_t quant_max, void* workspace, size_t workspace_bytes,
                              void* stream) {
    if (outer < 0 || C < 0 || inner < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (xdtype < 0 || xdtype > 2) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    if (quant_max <= quant_min) return fail(LSQB200_ERR_ARG, "quant_max must exceed quant_min");
    if (outer * C * inner == 0) return 0;
    if (!w || !scale_out) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    const Geometry g = plan_geometry(outer, C, inner, xdtype, K_STATS, common_alignment({w}), tuning());
    double* partials = nullptr;
    unsigned* counters = nullptr;
    if (g.splits > 1) {
        if (!workspace || workspace_bytes < kWorkspaceBytes || (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0)
            return fail(LSQB200_ERR_WORKSPACE, "workspace missing, misaligned or smaller than lsqb200_workspace_bytes()");
        counters = reinterpret_cast<unsigned*>(workspace);
        partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + kMaxCounters * 4);
    }
    lsqb200_qargs q{};
    q.quant_min = quant_min; q.quant_max = quant_max; q.type_min = quant_min; q.type_max = quant_max; q.grad_scaler = 1.0;
    SegArgs a = seg_args(w, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, outer, C, inner, xdtype, DT_F32,
                         C > 1, &q);
    a.stats_out = scale_out;
    Seg seg = make_seg(a, g, partials, counters, 0);
    seg.flags = tuning().l2_prefetch;
    const bool rows = row_kernels_eligible(g, xdtype) && tuning().rowstats;
    // the row-entry kernels (variants >= 4) need a plan's table: single calls take the descriptor form
    KernelFn k = rows ? get_rowstats_kernel(xdtype, tuning().rowstats >= 4 ? 2 : tuning().rowstats) : get_stats_kernel(xdtype, g.nw, g.group);
    const long long grid = g.grid;
    return launch(k, seg, nullptr, nullptr, 0, g.tiles, grid, (cudaStream_t)stream);
}

int lsqb200_observe(const void* x, int64_t outer, int64_t C, int64_t inner, int xdtype, int per_channel, float* min_val,
                    float* max_val, float* scale_out, float* shift_out, const lsqb200_observer_args* oa, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (!oa) return fail(LSQB200_ERR_ARG, "observer args are NULL");
    if (outer < 0 || C < 0 || inner < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (xdtype < 0 || xdtype > 2) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    if (oa->quant_min >= oa->quant_max) return fail(LSQB200_ERR_ARG, "quant_min must be smaller than quant_max");
    if (outer * C * inner == 0) return 0;                       // torch observers ignore empty inputs too
    if (!x || !min_val || !max_val) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    if (!per_channel) { inner *= outer * C; outer = 1; C = 1; }
    const Geometry g = plan_geometry(outer, C, inner, xdtype, K_STATS, common_alignment({x}), tuning());
    double* partials = nullptr;
    unsigned* counters = nullptr;
    if (g.splits > 1) {
        if (!workspace || workspace_bytes < kWorkspaceBytes || (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0)
            return fail(LSQB200_ERR_WORKSPACE, "workspace missing, misaligned or smaller than lsqb200_workspace_bytes()");
        counters = reinterpret_cast<unsigned*>(workspace);
        partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + kMaxCounters * 4);
    }
    lsqb200_qargs q{};
    q.quant_min = oa->quant_min; q.quant_max = oa->quant_max; q.type_min = oa->quant_min; q.type_max = oa->quant_max; q.grad_scaler = 1.0;
    SegArgs a = seg_args(x, scale_out, nullptr, shift_out, nullptr, nullptr, min_val, max_val, outer, C, inner, xdtype, DT_F32,
                         per_channel ? 1 : 0, &q);
    Seg seg = make_seg(a, g, partials, counters, 0);
    seg.obs_c = (float)oa->averaging_constant;
    seg.obs_eps = (float)oa->eps;
    seg.obs_flags = (oa->symmetric ? 1 : 0) | (oa->moving_average ? 2 : 0);
    seg.obs_zp_sym = oa->zero_point_sym;
    seg.flags = tuning().l2_prefetch;
    KernelFn k = get_observe_kernel(xdtype, g.nw, g.group);
    return launch(k, seg, nullptr, nullptr, 0, g.tiles, g.grid, (cudaStream_t)stream);
}

int lsqb200_quantize(const void* x, void* codes, const void* scale, const void* shift, int64_t outer, int64_t C, int64_t inner,
                     int xdtype, int pdtype, int per_channel, const lsqb200_qargs* q, int codes_signed, int semantics,
                     void* stream) {
    if (!per_channel) { inner *= outer * C; outer = 1; C = 1; }
    return export_common(DIR_QUANT, x, codes, scale, shift, outer, C, inner, xdtype, pdtype, per_channel ? 1 : 0, q,
                         codes_signed, semantics, stream);
}

int lsqb200_dequantize(const void* codes, void* y, const void* scale, const void* shift, int64_t outer, int64_t C,
                       int64_t inner, int xdtype, int pdtype, int per_channel, const lsqb200_qargs* q, int codes_signed,
                       int semantics, void* stream) {
    if (!per_channel) { inner *= outer * C; outer = 1; C = 1; }
    return export_common(DIR_DEQUANT, y, const_cast<void*>(codes), scale, shift, outer, C, inner, xdtype, pdtype,
                         per_channel ? 1 : 0, q, codes_signed, semantics, stream);
}

int lsqb200_qparams(const void* scale, const void* shift, float* scale_out, int64_t* zero_point_out, int64_t n, int pdtype,
                    int64_t type_min, int64_t type_max, void* stream) {
    if (n < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (pdtype < 0 || pdtype > 2) return fail(LSQB200_ERR_DTYPE, "unknown dtype code");
    if (type_min > type_max) return fail(LSQB200_ERR_ARG, "type_min must not exceed type_max");
    if (n == 0) return 0;
    if (!scale || !shift || !scale_out) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    const int e = launch_qparams(scale, shift, scale_out, reinterpret_cast<long long*>(zero_point_out), n, pdtype,
                                 (float)type_min, (float)type_max, tuning().pdl != 0, (cudaStream_t)stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "kernel launch");
    return 0;
}

int lsqb200_flat_optimizer_step(float* params, const float* grads, float* state1, float* state2, int64_t n,
                                const lsqb200_optim_args* o, void* stream) {
    if (!o) return fail(LSQB200_ERR_ARG, "optimizer args are NULL");
    if (n < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (o->kind != LSQB200_OPT_SGD && o->kind != LSQB200_OPT_ADAM) return fail(LSQB200_ERR_ARG, "unknown optimizer kind");
    if (n == 0) return 0;
    if (!params || !grads) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    if (o->kind == LSQB200_OPT_ADAM && (!state1 || !state2)) return fail(LSQB200_ERR_ARG, "Adam needs exp_avg and exp_avg_sq buffers");
    if (o->kind == LSQB200_OPT_SGD && o->momentum != 0.0 && !state1) return fail(LSQB200_ERR_ARG, "SGD with momentum needs a momentum buffer");
    if (o->kind == LSQB200_OPT_ADAM && !(o->beta1 >= 0.0 && o->beta1 < 1.0 && o->beta2 >= 0.0 && o->beta2 < 1.0))
        return fail(LSQB200_ERR_ARG, "Adam betas must lie in [0, 1)");
    const int e = launch_flat_optim(params, grads, state1, state2, n, o, tuning().pdl != 0, (cudaStream_t)stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "kernel launch");
    return 0;
}

int lsqb200_flat_optimizer_step_sites(float* params, const float* grads, float* state1, float* state2, int32_t* steps,
                                      const uint8_t* active, int64_t n, const lsqb200_optim_args* o, void* stream) {
    if (!o) return fail(LSQB200_ERR_ARG, "optimizer args are NULL");
    if (n < 0) return fail(LSQB200_ERR_ARG, "negative size");
    if (o->kind != LSQB200_OPT_SGD && o->kind != LSQB200_OPT_ADAM) return fail(LSQB200_ERR_ARG, "unknown optimizer kind");
    if (n == 0) return 0;
    if (!params || !grads || !steps) return fail(LSQB200_ERR_ARG, "NULL tensor pointer");
    if (o->kind == LSQB200_OPT_ADAM && (!state1 || !state2)) return fail(LSQB200_ERR_ARG, "Adam needs exp_avg and exp_avg_sq buffers");
    if (o->kind == LSQB200_OPT_SGD && o->momentum != 0.0 && !state1) return fail(LSQB200_ERR_ARG, "SGD with momentum needs a momentum buffer");
    if (o->kind == LSQB200_OPT_ADAM && !(o->beta1 >= 0.0 && o->beta1 < 1.0 && o->beta2 >= 0.0 && o->beta2 < 1.0))
        return fail(LSQB200_ERR_ARG, "Adam betas must lie in [0, 1)");
    const int e = launch_flat_optim_sites(params, grads, state1, state2, steps, active, n, o, tuning().pdl != 0, (cudaStream_t)stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "kernel launch");
    return 0;
}

int lsqb200_plan_create(const lsqb200_segment* segs, int32_t nseg, lsqb200_plan** out) {
    if (!segs || nseg <= 0 || !out) return fail(LSQB200_ERR_PLAN, "empty segment list");
    lsqb200_plan* p = new lsqb200_plan();
    p->segs.assign(segs, segs + nseg);
    cudaGetDevice(&p->device);
    long long off = 0;
    for (const auto& s : p->segs) {
        if (s.outer < 0 || s.C < 0 || s.inner < 0 || s.xdtype < 0 || s.xdtype > 3 || s.pdtype < 0 || s.pdtype > 3 ||
            !prologue_known(s.prologue) || (prologue_adds(s.prologue) && !s.x2 && s.outer * s.C * s.inner != 0)) {
            delete p;
            return fail(LSQB200_ERR_PLAN, "bad segment (negative size, unknown dtype, unknown prologue or missing x2)");
        }
        p->stats_offset.push_back(off);
        off += s.per_channel ? s.C : 1;
    }
    // pass 1: size the private workspace (backward and stats never run concurrently on one plan)
    size_t need_b = 0, need_s = 0;
    {
        std::vector<lsqb200_plan::Class> tmp;
        int r = build_classes(p, K_BWD, tmp, nullptr, need_b, false);
        if (!r) { tmp.clear(); r = build_classes(p, K_STATS, tmp, nullptr, need_s, false); }
        if (r) { delete p; return r; }
    }
    const size_t ws = need_b > need_s ? need_b : need_s;
    if (ws) {
        cudaError_t e = cudaMalloc(&p->workspace, ws);
        if (e == cudaSuccess) e = cudaMemset(p->workspace, 0, ws);
        if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(plan workspace)"); }
    }
    size_t used = 0;
    int r = build_classes(p, K_FWD, p->fwd, nullptr, used, false);
    used = 0;
    if (!r) r = build_classes(p, K_BWD, p->bwd, (char*)p->workspace, used, true);
    used = 0;
    if (!r) r = build_classes(p, K_STATS, p->stats, (char*)p->workspace, used, true);
    if (!r) r = upload_classes(p->fwd);
    if (!r) r = upload_classes(p->bwd);
    if (r) { lsqb200_plan_destroy(p); return r; }
    cudaDeviceSynchronize();
    *out = p;
    return 0;
}

int lsqb200_plan_forward(lsqb200_plan* plan, void* stream) {
    if (!plan) return fail(LSQB200_ERR_PLAN, "NULL plan");
    return run_classes(plan->fwd, (cudaStream_t)stream);
}

int lsqb200_plan_backward(lsqb200_plan* plan, void* stream) {
    if (!plan) return fail(LSQB200_ERR_PLAN, "NULL plan");
    return run_classes(plan->bwd, (cudaStream_t)stream);
}

int lsqb200_plan_weight_init_stats(lsqb200_plan* plan, float* scale_out, void* stream) {
    if (!plan || !scale_out) return fail(LSQB200_ERR_PLAN, "NULL plan or output");
    // tables are uploaded with the output pointers patched in (again only when the output buffer changes); this is an init-time call
    const bool same_out = plan->stats_uploaded_for == scale_out;
    plan->stats_uploaded_for = scale_out;
    for (auto& c : plan->stats) {
        if (c.host.empty()) continue;
        if (same_out && c.dev) continue;
        for (auto& s : c.host) s.stats_out = scale_out + plan->stats_offset[(size_t)s.chan_stride];
        if (!c.dev) {
            cudaError_t e = cudaMalloc(&c.dev, c.host.size() * sizeof(Seg));
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(plan table)");
            if (int r = c.rowtable ? upload_row_table(c, plan) : upload_tile_map(c)) return r;
        }
        cudaError_t e = cudaMemcpyAsync(c.dev, c.host.data(), c.host.size() * sizeof(Seg), cudaMemcpyHostToDevice,
                                        (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(plan table)");
    }
    return run_classes(plan->stats, (cudaStream_t)stream);
}

int lsqb200_plan_launches(const lsqb200_plan* plan, int backward) {
    if (!plan) return 0;
    int n = 0;
    for (const auto& c : (backward ? plan->bwd : plan->fwd)) n += !c.host.empty();
    return n;
}

int lsqb200_plan_destroy(lsqb200_plan* plan) {
    if (!plan) return 0;
    for (auto* v : {&plan->fwd, &plan->bwd, &plan->stats})
        for (auto& c : *v)
            { if (c.dev) cudaFree(c.dev); if (c.dev_tile_seg) cudaFree(c.dev_tile_seg); }
    if (plan->workspace) cudaFree(plan->workspace);
    delete plan;
    return 0;
}

}  // extern "C"
