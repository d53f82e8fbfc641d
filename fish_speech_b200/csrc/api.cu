// Library-level C-ABI entry points + the unit-test hooks declared in include/fishb200.h.
#include "../../include/fishb200.h"
#include "gemm_tc.cuh"
#include "lm_kernels.cuh"

using namespace fsb;

extern "C" {

const char* fsb_last_error(void) { return get_error(); }
long long fsb_launch_count(void) { return g_launch_count; }

int fsb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaDeviceProp prop;
    FSB_CUDA(cudaGetDevice(&dev));
    FSB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

int fsb_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FSB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
int fsb_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FSB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    FSB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Reduce stream-K partial sums: out[j][i] = sum_s ws[s][j][i]
static __global__ void reduce_parts_kernel(const float* ws, long long slot_stride, const int* nparts,
                                           float* out, int m, int n) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= m || j >= n) return;
    const int np = nparts[i >> 7];
    float s = 0.f;
    for (int q = 0; q < np; ++q) s += ws[q * slot_stride + static_cast<long long>(j) * m + i];
    out[static_cast<long long>(j) * m + i] = s;
}

int fsb_op_gemm(const void* d_a, const void* d_b, float* d_out, int m, int n, int k, int bn,
                int streamk_ctas, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    GemmPlan plan;
    memset(&plan, 0, sizeof(plan));
    GemmOperand A{reinterpret_cast<const __nv_bfloat16*>(d_a), k, m, 1, k, static_cast<long long>(m) * k};
    GemmOperand B{reinterpret_cast<const __nv_bfloat16*>(d_b), k, n, 1, k, static_cast<long long>(n) * k};
    const int kblocks = cdiv(k, 64);
    plan.p.kb_per_tap = kblocks;
    plan.p.num_taps = 1;
    plan.p.rows_i = m;
    plan.p.rows_j = n;
    const int tiles_i = cdiv(m, 128), tiles_j = cdiv(n, bn);
    FSB_TRY(gemm_plan_init(&plan, A, B, bn, 6, tiles_i, tiles_j, 1));
    float* ws = nullptr;
    if (streamk_ctas > 0) {
        FSB_CHECK(tiles_j == 1, "stream-K hook needs n <= bn");
        FSB_TRY(gemm_plan_streamk(&plan, tiles_i, kblocks, streamk_ctas));
        const size_t slot = static_cast<size_t>(n) * m;
        FSB_CUDA(cudaMalloc(&ws, slot * plan.max_parts * sizeof(float)));
        plan.p.mode = 0;
        plan.p.ws = ws;
        plan.p.ws_ld = m;
        plan.p.ws_slot_stride = static_cast<long long>(slot);
        int rc = gemm_launch(plan, st);
        if (rc == 0) {
            reduce_parts_kernel<<<dim3(cdiv(m, 256), n), 256, 0, st>>>(ws, plan.p.ws_slot_stride,
                                                                       plan.nparts_dev, d_out, m, n);
            if (cudaGetLastError() != cudaSuccess) rc = 1;
        }
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(ws);
        gemm_plan_free(&plan);
        FSB_CHECK(rc == 0, "%s", get_error());
        FSB_CUDA(e);
        return 0;
    }
    // direct fp32 epilogue: out[j*m + i]
    plan.p.mode = 1;
    plan.p.out0 = d_out;
    plan.p.out_f32 = 1;
    plan.p.o_zs = 0;
    plan.p.o_is = 1;
    plan.p.o_js = m;
    plan.p.chan_on_i = 1;
    int rc = gemm_launch(plan, st);
    cudaError_t e = cudaStreamSynchronize(st);
    gemm_plan_free(&plan);
    if (rc != 0) return rc;
    FSB_CUDA(e);
    return 0;
}

}  // extern "C"
