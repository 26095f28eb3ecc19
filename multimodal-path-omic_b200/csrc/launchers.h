// C++-side launch functions shared between the kernel translation units and the extern "C" surface.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "mpo_common.cuh"

namespace mpo {

extern thread_local char g_err[512];
int fail(int code, const char* fmt, const char* detail);
int check_cuda(cudaError_t e, const char* where);
int num_sms();
extern unsigned long long g_launches;   // kernels launched by this library since the last reset
inline void count_launch(int n = 1) { g_launches += n; }

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
int launch_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int32_t* step_dev, bool zero_grad, bool bump, cudaStream_t st);
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows);
int make_tmap_f32_rows32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t ld, uint32_t kblocks);
int make_tmap_f32_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
                        bool swizzle128);
int fwd_cluster_size();   // CTAs per cluster of the forward bag kernel (1, 2 or 4; env MPO_FWD_CLUSTER)
cudaError_t launch_bag_fwd(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const CUtensorMap& tm_h,
                           const BagFwdParams& prm, int num_sms, cudaStream_t stream);
cudaError_t launch_bag_merge(const int* tile_prefix, const float* part_ml, int ml_stride, const float* part_pool,
                             float* pooled, float* lse, float* suma, int B, cudaStream_t stream,
                             const float* part_pool2 = nullptr, float* pooled2 = nullptr);
cudaError_t launch_bag_gate(const CUtensorMap& tm_h, const CUtensorMap& tm_hlo, const CUtensorMap& tm_w,
                            const BagGateParams& prm, int num_sms, cudaStream_t stream);
cudaError_t launch_bag_dhk(const CUtensorMap& tm_dkg, const CUtensorMap& tm_w, const CUtensorMap& tm_dz,
                           const BagDhkParams& prm, int num_sms, cudaStream_t stream);
cudaError_t launch_bag_bwd_dz(int mode, const CUtensorMap& tm_in, const CUtensorMap& tm_out, const BagBwdDzParams& prm,
                              int num_sms, cudaStream_t stream);
cudaError_t launch_bag_bwd_reduce(const int* tile_prefix, const float* part_dqk, const float* part_db, float* dqk,
                                  float* grad_bias, int B, int num_tiles, cudaStream_t stream);
cudaError_t launch_bag_bwd_dw(const CUtensorMap& tm_dz, const CUtensorMap& tm_x, float* grad_w, int total_rows,
                              int ncols, int ld, bool f16, const uint32_t* dg_max, int num_sms, cudaStream_t stream);
cudaError_t launch_bag_bwd_dwz(const CUtensorMap& tm_x, const CUtensorMap& tm_scr, const BagBwdDwzParams& prm, int num_sms,
                               cudaStream_t stream);
int bag_bwd_dwz_max_clusters(int num_sms);
size_t bag_bwd_dwz_scratch_bytes(int clusters);
// fp32-accurate tcgen05 GEMM on bf16 (hi, lo) operand pairs (tc_gemm.cu)
struct DropSpec;      // tail_dev.cuh
cudaError_t launch_split_bf16(const float* src, long long ld, int rows, int cols, void* hi, void* lo, int pitch,
                              cudaStream_t stream, const DropSpec* drop = nullptr, uint32_t base = 0);
int launch_tc_gemm(const void* a_hi, const void* a_lo, int a_rows, int a_pitch, bool a_mn, const void* b_hi, const void* b_lo,
                   int b_rows, int b_pitch, bool b_mn, float* C, long long ldc, int M, int N, int K, float alpha, bool accumulate,
                   cudaStream_t stream);
cudaError_t launch_bag_bwd_dkc(const int* tile_prefix, const float* part_dkc, float* dkc, int B, cudaStream_t stream);

}  // namespace mpo
