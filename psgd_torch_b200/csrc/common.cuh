// common.cuh -- shared types of the PSGD engine kernels: dtypes, the fused GEMM epilogue spec, reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psgd_b200.h"

namespace psgd {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt) { return dt == PSGD_BF16 ? 2 : 4; }

__device__ __forceinline__ float ld_as_float(const void* p, int dt, size_t idx) {
  return dt == PSGD_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[idx])
                         : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void st_from_float(void* p, int dt, size_t idx, float v) {
  if (dt == PSGD_BF16)
    reinterpret_cast<bf16*>(p)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p)[idx] = v;
}
// value after rounding to the storage dtype (what a consumer of the stored tensor would read)
__device__ __forceinline__ float round_to(int dt, float v) {
  return dt == PSGD_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__host__ __device__ inline float dtype_eps(int dt) { return dt == PSGD_BF16 ? 0.0078125f : 1.1920928955078125e-07f; }
// torch.finfo(dtype).smallest_normal: identical for bf16 and fp32 (same exponent range)
__host__ __device__ inline float dtype_tiny(int) { return 1.1754943508222875e-38f; }

// ---------------------------------------------------------------------------------------------
// Fused GEMM epilogue, shared by the SIMT and the tcgen05 kernels:
//   v        = acc * alpha * (*alpha_ptr) * row_scale[i] * col_scale[j]  +  beta * (*beta_ptr) * D[i,j] * d_row_scale[i] * d_col_scale[j]
//   C[i,j]   = round_to(out_dtype, v)
//   reductions below are taken over the ROUNDED values (the reference computes its norms / traces on the
//   materialised low-precision tensors), accumulated with atomics into fp32 device buffers that the caller
//   zeroed beforehand.
// ---------------------------------------------------------------------------------------------
// slots of the per-factor update scalars (device floats written by one kernel's epilogue / finish and read by the next kernel's epilogue)
enum { FS_ALPHA = 0, FS_BETA = 1, FS_INV_SR = 2, FS_TR1 = 3, FS_TR2 = 4, FS_TR3 = 5, FS_DOT = 6, FS_COUNT = 8 };

struct Epi {
  void* C;
  int ldc;
  int out_dtype;
  float alpha;
  const float* alpha_ptr;
  const void* D;
  int ldd;
  int d_dtype;
  float beta;
  const float* beta_ptr;
  const float* row_scale;  // length M (fp32) or null
  const float* col_scale;  // length N (fp32) or null
  const float* d_row_scale;  // optional per-row / per-column factors of the D term (fp32): used to add back the bf16 rounding residual of
  const float* d_col_scale;  // the diagonal of P = Q^T Q, (P_bf16 + diag(resid)) X = P_bf16 X + resid_i X_ij
  float* diag_resid;       // [M]: (unrounded - rounded) value of C[i,i]
  const void* D2;          // optional second addend: + beta2 * D2[i,j] (dtype d_dtype)
  int ldd2;
  float beta2;
  // procrustes_step2 finish (psgd.py:121-124) folded into the epilogue of the product R (RQ): with tr1 = pro_fs[FS_TR1] = tr(RQ) and
  // tr2 = tr(R RQ) / |R| = -pro_fs[FS_INV_SR] * pro_fs[FS_DOT] (R is skew: tr(R X) = -<R, X>, accumulated by the previous product's `dot`
  // reduction), a = tr2 < 0 ? min(-tr1 / tr2, pro_max_step) : pro_max_step; then alpha *= a^2 / 2 and beta2 = a, so that with D = Q and
  // D2 = RQ the product writes Q + a (RQ + a/2 RRQ) directly -- RRQ is never materialised and needs no extra pass
  const float* pro_fs;
  float pro_max_step;
  const void* dotm;        // optional: dot_out += sum_ij C[i,j] * dotm[i,j] (dtype d_dtype, over the rounded C)
  int ld_dot;
  float* dot_out;
  // row (axis 1) / column (axis 2) normalisation folded with a device scalar: factor = min(*norm_inv_nf / (sqrt(norm_sumsq[idx]) + norm_tiny), 3e38)
  // (psgd.py:66 `V /= |V| + tiny` followed by the /nf of the next product; the clamp keeps 0 * factor = 0 when a probe row is exactly 0)
  const float* norm_sumsq;
  const float* norm_inv_nf;
  float norm_tiny;
  int norm_axis;
  float* row_sumsq;        // [M] += sum_j C[i,j]^2
  float* col_sumsq;        // [N] += sum_i C[i,j]^2
  float* diag_max;         // max_i C[i,i]   (values assumed >= 0; buffer zero-initialised)
  float* abs_max;          // max |C[i,j]|
  float* trace;            // += sum_i C[i,i]
  float* total_sumsq;      // += sum_ij C[i,j]^2
};

__host__ inline Epi make_epi(void* C, int ldc, int out_dtype) {
  Epi e;
  e.C = C; e.ldc = ldc; e.out_dtype = out_dtype;
  e.alpha = 1.f; e.alpha_ptr = nullptr;
  e.D = nullptr; e.ldd = 0; e.d_dtype = out_dtype; e.beta = 0.f; e.beta_ptr = nullptr;
  e.row_scale = nullptr; e.col_scale = nullptr; e.d_row_scale = nullptr; e.d_col_scale = nullptr; e.diag_resid = nullptr;
  e.D2 = nullptr; e.ldd2 = 0; e.beta2 = 0.f; e.pro_fs = nullptr; e.pro_max_step = 0.f; e.dotm = nullptr; e.ld_dot = 0; e.dot_out = nullptr;
  e.norm_sumsq = nullptr; e.norm_inv_nf = nullptr; e.norm_tiny = 0.f; e.norm_axis = 0;
  e.row_sumsq = nullptr; e.col_sumsq = nullptr; e.diag_max = nullptr; e.abs_max = nullptr; e.trace = nullptr;
  e.total_sumsq = nullptr;
  return e;
}

// non-negative float max via integer atomics (buffer initialised to 0)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  if (v > 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

// step length of procrustes_step2 (psgd.py:121-123) from the device scalars described at Epi::pro_fs
__device__ __forceinline__ float epi_procrustes_step(const Epi& e) {
  const float tr1 = e.pro_fs[FS_TR1];
  const float tr2 = -e.pro_fs[FS_INV_SR] * e.pro_fs[FS_DOT];
  return (tr2 < 0.f) ? fminf(-tr1 / tr2, e.pro_max_step) : e.pro_max_step;
}

__device__ __forceinline__ float epi_norm_factor(const Epi& e, int idx) {
  return fminf(*e.norm_inv_nf / (sqrtf(e.norm_sumsq[idx]) + e.norm_tiny), 3.0e38f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? smem32[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}
__device__ __forceinline__ float block_max(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = -INFINITY;
  if (w == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? smem32[lane] : -INFINITY;
    r = warp_max(r);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// engine context
// ---------------------------------------------------------------------------------------------
struct Ctx {
  int device;
  int num_sms;         // SMs the persistent kernels size their grids for (psgd_set_sm_limit; default: all)
  int num_sms_hw;
  int gemm_path;       // 0 auto, 1 simt, 2 tc
  int mn_lbo, mn_sbo;  // MN-major UMMA descriptor byte offsets
  int force_bn;        // debug: 0 = automatic tile width, else 128 / 256
  int debug_flags;     // bit0: no symmetric (upper-blocks-only) products, bit1: no P-first chain, bit2: no tensor-core norm bounds
  int64_t launches;
  char last_error[256];
  void* encode_tiled;  // cuTensorMapEncodeTiled entry point
  // optional per-launch timing of the tcgen05 GEMM kernel (bench.py roofline): CUDA events on the launching stream
  int timing_on;
  int timing_count;      // event pairs recorded since the last reset
  int64_t timing_seen;   // tcgen05 GEMM launches since the last reset (recorded or not: the event pool is finite)
  double timing_flops;   // algorithmic FLOPs (2*M*N*K, true sizes) of the recorded launches
  double timing_flops_exec;  // FLOPs the tensor cores actually executed for them (tile-padded, upper-blocks-only for symmetric products)
  cudaEvent_t* ev_begin;
  cudaEvent_t* ev_end;
  int ev_capacity;
  // split-K workspace of the tcgen05 GEMM (fp32 tile slots + arrival counters), zero between launches
  float* ws;
  int* ws_count;
  int ws_slots;
  // operands of an fp32 product re-expressed as bf16 triples (hi + mid + lo, concatenated along K) for the tensor cores; grown on demand
  void* x3_buf[2];
  size_t x3_cap[2];
  unsigned* nb_sync; // [0] grid-barrier counter, [1] completion counter of the fused norm-bound kernel (bounds.cuh); null: not available
  int fp32_tensor;   // opt-in (psgd_set_fp32_tensor_cores): big fp32 products as bf16 triples on the tensor cores
};

// cudaFuncSetAttribute acts on the current device's context: one "already done" flag per (call site, device) for processes that drive
// several GPUs (the supported layout is one process per GPU, where this is a single flag)
struct PerDeviceOnce {
  bool done[64];
  bool need(int dev) {
    if (dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// one GEMM problem: C = epi(op(A) op(B)), op(A) M x K, op(B) K x N
struct GemmDesc {
  const void* A;
  const void* B;
  int lda, ldb;
  int ta, tb;  // ta: A stored K x M; tb: B stored N x K
  int M, N, K;
  int in_dtype;
  int sym;  // output is symmetric (C = X X^T or X^T X, M == N): the tcgen05 path computes the upper 128-blocks only and mirrors them
  Epi epi;
};

// launchers implemented in gemm_simt.cu / gemm_tc.cu / api.cu
int launch_gemm_simt(Ctx* ctx, const GemmDesc& g, cudaStream_t st);
bool tc_eligible(const GemmDesc& g);
int launch_gemm_tc_group(Ctx* ctx, const GemmDesc* g, int n, cudaStream_t st);
int launch_gemm(Ctx* ctx, const GemmDesc& g, cudaStream_t st);  // picks the path
int launch_gemm_pair(Ctx* ctx, const GemmDesc& g0, const GemmDesc& g1, cudaStream_t st);

int check_cuda(Ctx* ctx, cudaError_t e, const char* what);

}  // namespace psgd
