// tools/tune.cu -- standalone sweep of compile-time launch shapes for the streaming kernels.
// Not part of the product library: it includes the same device code (lsq_device.cuh) and the
// same host planner (lsq_host.h) and instantiates ALTERNATIVE (threads, unroll, min-blocks,
// load policy, store policy) variants, times them with CUDA events on big L2-defeating buffers
// and prints one CSV line per (variant, tile size).  Build + run:  make -C tools && tools/tune
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../lsqfakequantize-pytorch_b200/csrc/lsq_host.h"

using namespace lsqb200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Variant {
    const char* name;
    KernelFn fn;
    int kind, threads, unroll, xdtype, unit_bytes;
};

template <typename T> struct DtOf;
template <> struct DtOf<float> { static constexpr int v = DT_F32; };
template <> struct DtOf<__half> { static constexpr int v = DT_F16; };
template <> struct DtOf<__nv_bfloat16> { static constexpr int v = DT_BF16; };

#define FWD(T, NW, TH, UN, MB, LD, ST) {"fwd," #T "," #NW "," #TH "," #UN "," #MB "," #LD "," #ST, \
    lsq_fwd_kernel<T, M_FP32, NW, false, TH, TH, UN, LD, ST, MB>, K_FWD, TH, UN, DtOf<T>::v, NW * 4}
#define BWD(T, NW, TH, UN, MB, LD, ST) {"bwd," #T "," #NW "," #TH "," #UN "," #MB "," #LD "," #ST, \
    lsq_bwd_kernel<T, M_FP32, NW, B_NORMAL, TH, TH, UN, LD, ST, MB>, K_BWD, TH, UN, DtOf<T>::v, NW * 4}
typedef __nv_bfloat16 bf16;

static std::vector<Variant> variants() {
    return {
        // forward, bf16: unit width x unroll x occupancy, then cache policies
        FWD(bf16, 4, 256, 4, 4, 1, 0), FWD(bf16, 8, 256, 2, 4, 1, 0), FWD(bf16, 8, 256, 4, 2, 1, 0), FWD(bf16, 8, 256, 1, 8, 1, 0),
        FWD(bf16, 4, 256, 2, 8, 1, 0), FWD(bf16, 4, 256, 8, 2, 1, 0), FWD(bf16, 8, 512, 2, 2, 1, 0), FWD(bf16, 8, 128, 2, 8, 1, 0),
        FWD(bf16, 8, 1024, 1, 2, 1, 0), FWD(bf16, 8, 256, 2, 4, 0, 0), FWD(bf16, 8, 256, 2, 4, 2, 0), FWD(bf16, 8, 256, 2, 4, 1, 1),
        FWD(bf16, 8, 256, 2, 4, 1, 2), FWD(bf16, 8, 256, 2, 4, 2, 1), FWD(bf16, 4, 256, 4, 4, 0, 0), FWD(bf16, 4, 256, 4, 4, 1, 1),
        // forward, fp32
        FWD(float, 4, 256, 4, 4, 1, 0), FWD(float, 8, 256, 2, 4, 1, 0), FWD(float, 8, 256, 4, 2, 1, 0), FWD(float, 8, 256, 2, 4, 2, 1),
        // backward, bf16
        BWD(bf16, 4, 256, 4, 3, 1, 0), BWD(bf16, 8, 256, 2, 3, 1, 0), BWD(bf16, 8, 256, 1, 4, 1, 0), BWD(bf16, 8, 256, 2, 2, 1, 0),
        BWD(bf16, 4, 256, 2, 4, 1, 0), BWD(bf16, 8, 256, 1, 6, 1, 0), BWD(bf16, 8, 512, 1, 2, 1, 0), BWD(bf16, 8, 128, 2, 6, 1, 0),
        BWD(bf16, 8, 128, 1, 8, 1, 0), BWD(bf16, 8, 1024, 1, 1, 1, 0), BWD(bf16, 8, 256, 4, 1, 1, 0),
        BWD(bf16, 8, 256, 2, 3, 0, 0), BWD(bf16, 8, 256, 2, 3, 2, 0), BWD(bf16, 8, 256, 2, 3, 1, 1), BWD(bf16, 8, 256, 2, 3, 1, 2),
        BWD(bf16, 8, 256, 2, 3, 2, 1), BWD(bf16, 4, 256, 4, 3, 0, 0), BWD(bf16, 4, 256, 4, 3, 1, 1),
        // backward, fp32
        BWD(float, 4, 256, 4, 3, 1, 0), BWD(float, 8, 256, 2, 3, 1, 0), BWD(float, 8, 256, 1, 4, 1, 0), BWD(float, 8, 256, 2, 2, 1, 0),
        BWD(float, 8, 256, 2, 3, 2, 1), BWD(float, 8, 512, 1, 2, 1, 0),
    };
}

__global__ void fill_kernel(uint32_t* p, size_t n, uint32_t seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // two bf16 / one fp32 value(s) in roughly [-2, 2]: exponent 0x3f/0x40, random mantissa and sign
        const uint32_t lo = 0x3f00u | (h & 0x80ffu), hi = 0x3f00u | ((h >> 16) & 0x80ffu);
        p[i] = (hi << 16) | lo;
    }
}
__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 1024) * (size_t)(1 << 20);   // per buffer, MiB
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const char* filter = argc > 3 ? argv[3] : "";
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void *x, *g, *y, *flush;
    CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&g, bytes)); CK(cudaMalloc(&y, bytes));
    const size_t flush_bytes = 512u << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    float hp[2] = {0.03f, -1.7f};
    float* params; float* grads; void* ws;
    CK(cudaMalloc(&params, 8)); CK(cudaMalloc(&grads, 8)); CK(cudaMalloc(&ws, kWorkspaceBytes));
    CK(cudaMemcpy(params, hp, 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(ws, 0, kWorkspaceBytes));
    fill_kernel<<<sms * 8, 256>>>((uint32_t*)x, bytes / 4, 1u);
    fill_kernel<<<sms * 8, 256>>>((uint32_t*)g, bytes / 4, 2u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("# device sms=%d buffer=%zu MiB reps=%d\n", sms, bytes >> 20, reps);
    printf("kind,dtype,nw,threads,unroll,minb,ld,st,interleave,tile_kb,occ_ctas_per_sm,regs,grid,ms_best,ms_med,GBps_best,GBps_med\n");

    // reference points: cudaMemcpy D2D and a plain uint4 grid-stride copy
    {
        std::vector<float> t;
        for (int r = 0; r < reps; r++) {
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        printf("memcpy_d2d,-,-,-,-,-,-,-,-,-,-,-,-,%.4f,%.4f,%.1f,%.1f\n", t[0], t[t.size() / 2], 2.0 * bytes / t[0] / 1e6, 2.0 * bytes / t[t.size() / 2] / 1e6);
        for (int mult : {4, 8, 16, 32}) {
            t.clear();
            for (int r = 0; r < reps; r++) {
                CK(cudaEventRecord(e0));
                copy_kernel<<<sms * mult, 256>>>((const uint4*)x, (uint4*)y, bytes / 16);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
            }
            std::sort(t.begin(), t.end());
            printf("copy_kernel,-,4,256,-,-,-,-,-,%d,-,-,%d,%.4f,%.4f,%.1f,%.1f\n", mult, sms * mult, t[0], t[t.size() / 2], 2.0 * bytes / t[0] / 1e6, 2.0 * bytes / t[t.size() / 2] / 1e6);
        }
    }

    for (const Variant& v : variants()) {
        if (*filter && !strstr(v.name, filter)) continue;
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, (const void*)v.fn));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)v.fn, v.threads, 0));
        const int es = v.xdtype == DT_F32 ? 4 : 2;
        const long long n = (long long)(bytes / es);
        for (int il : {1, 0})
        for (int tps : {8, 16, 32, 64, 128, 256, 512, 2048}) {   // tile size in KB of one operand
            if (il == 0 && tps != 32 && tps != 256) continue;
            Tuning tn;
            tn.interleave = il;
            tn.sm_count = sms;
            tn.fwd_tile_kb = tn.bwd_tile_kb = tps;
            tn.max_unit_bytes = v.unit_bytes;
            Geometry geo = plan_geometry(1, 1, n, v.xdtype, v.kind, 32, tn, v.threads, v.unroll);
            SegArgs a{};
            a.x = x; a.y = y; a.g = g; a.gx = y; a.scale = params; a.shift = params + 1; a.gscale = grads; a.gshift = grads + 1;
            a.outer = 1; a.C = 1; a.inner = n; a.xdtype = v.xdtype; a.pdtype = DT_F32; a.per_channel = 0;
            a.qmin = 0; a.qmax = 127; a.tmin = 0; a.tmax = 255; a.grad_scaler = 1.0; a.use_grad_scaling = 1; a.sym = 0;
            Seg seg = make_seg(a, geo, (double*)((char*)ws + kMaxCounters * 4), (unsigned*)ws, 0);
            std::vector<float> t;
            for (int r = 0; r < reps + 2; r++) {
                CK(cudaMemsetAsync(flush, r, flush_bytes));   // defeat L2 between iterations
                CK(cudaEventRecord(e0));
                v.fn<<<(unsigned)geo.grid, v.threads>>>(seg, nullptr, nullptr, 0, geo.tiles);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 2) t.push_back(ms);
            }
            std::sort(t.begin(), t.end());
            const double alg = (v.kind == K_FWD ? 2.0 : 3.0) * bytes;
            printf("%s,%d,%d,%d,%d,%lld,%.4f,%.4f,%.1f,%.1f\n", v.name, il, tps, occ, fa.numRegs, geo.grid, t[0], t[t.size() / 2],
                   alg / t[0] / 1e6, alg / t[t.size() / 2] / 1e6);
            fflush(stdout);
        }
    }
    return 0;
}
