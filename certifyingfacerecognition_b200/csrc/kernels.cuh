// Launchers for the HBM-bound / small kernels of the certification path (implemented in kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cfr {

// certificate.py:64-67 + smoothing_model.py:63-67 + stylegan_generator_model.py:322-328
//   noise = N(0,1)*sigma (Philox, or injected), w = z + (x+noise)@dir_mat, wp2[b][0]=w_avg+(w-w_avg)*psi, wp2[b][1]=w
int launch_noise_project(const float* z, const float* x, const float* sigma, int sigma_len, const float* noise_in,
                         const float* dir_mat, const float* w_avg, float psi, unsigned long long seed,
                         unsigned long long sample_offset, int b, float* noise_out, float* wp2, cudaStream_t st);
// plain latents (gallery building): wp2 from w[b,512] directly
int launch_truncate(const float* w, const float* w_avg, float psi, int b, float* wp2, cudaStream_t st);

// MappingModule (stylegan_generator_model.py:265-295): PixelNorm (:398-406) + 8 x {Linear(512,512) * scale + b, LeakyReLU(0.2)}
//   wt [8][512 in][512 out] = W^T * (sqrt(2)/sqrt(512) * lr_mul), bias [8][512] = b * lr_mul;  z [b,512] -> w [b,512] (fp32)
int launch_mapping(const float* z, const float* wt, const float* bias, int b, float* w_out, cudaStream_t st);

// stylegan_generator_model.py:503 via DenseBlock :811-815  -- all 18 layers' style vectors at once
int launch_styles(const float* wp2, const float* w_style, const float* b_style, int rows, int rows_trunc, int b,
                  float* styles, cudaStream_t st);

// FirstConvBlock :581-584 + epilogue: sample-independent normalised const, then AdaIN
int launch_layer0(const float* xhat0, const float* styles, int style_stride, int style_off, int b, __half* out,
                  cudaStream_t st);

int launch_layer0_split(const float* xhat0, const float* styles, int style_stride, int style_off, int b, __half* out,
                        cudaStream_t st);

// BlurLayer :463 + EpilogueBlock :560-562 (+noise*w +bias, lrelu) + per-(n,c) sums for IN.  mode 0: blur+act+write;
// mode 1: statistics of an existing tensor only.
int launch_blur_act_stats(const __half* raw, __half* y, int n, int h, int w, int c, const float* noise,
                          const float* noise_w, const float* bias, void* sum, void* sq, int mode, cudaStream_t st);

// InstanceNormLayer :420-422 + StyleModulationLayer :505 ->  x = y*A + B,  A = rstd*(s0+1),  B = s1 - mean*A
int launch_finalize_stats(const void* sum, const void* sq, const float* styles, int style_stride, int style_off,
                          int n, int c, float inv_count, float* A, float* B, cudaStream_t st);
int launch_affine(const __half* y, const float* A, const float* B, int n, int hw, int c, __half* x, cudaStream_t st);
// split-precision early layers: fp32 activations between kernels, fp16 hi/lo operand pairs into the tensor cores
int launch_blur_act_stats_f32(const float* raw, float* y, int n, int h, int w, int c, const float* noise,
                              const float* noise_w, const float* bias, void* sum, void* sq, int mode, cudaStream_t st);
int launch_affine_f32(const float* y, const float* A, const float* B, int n, int hw, int c, __half* x, int split,
                      cudaStream_t st);

// LastConvBlock :759-762 + postprocess mod_stylegan_generator.py:303-307 + get_transform gen_utils.py:77-85
// x [n,H,W,C] fp16 (optionally still un-normalised: per-(n,c) A,B applied on load) -> img [n,R,R,16] fp16 (ch 0..2)
int launch_torgb_resize(const __half* x, const float* A, const float* B, int n, int hin, int c, const float* w_rgb,
                        const float* b_rgb, int rout, float mean, float stdv, __half* out, float* out_planar,
                        const int* slot, cudaStream_t st, const int* keep_map = nullptr, int keep_dim = 0);
int launch_set_int(int* p, int v, cudaStream_t st);
// tensor-core gallery match helpers (see kernels.cu)
int launch_split_hilo(const float* src, int rows, int rows_pad, int mode, __half* dst, float* bias, cudaStream_t st);
int launch_vote_argmax(unsigned long long* keys, int b, int n, int* pred, long long* counts, cudaStream_t st);

// smoothing_model.py:56-61 + smooth.py:135,140-146: argmin_j ||e - g_j||_2 (exact differences, fp32), votes
int launch_match_vote(const float* emb, int b, const float* gallery, int n, unsigned long long* keys, int* pred,
                      long long* counts, cudaStream_t st);

// split-K partial sums [n][parts][c] (+ bias) -> [n][c], fixed summation order
int launch_sum_partials(const float* in, const float* bias, int n, int parts, int c, float* out, cudaStream_t st);
// InceptionResnetV1 glue: MaxPool2d(3,2) into a channel slice, global average pool, row-wise L2 normalisation
int launch_maxpool3s2(const __half* in, int n, int h, int w, int c, __half* out, int out_c_total, int c_off, cudaStream_t st);
int launch_avgpool(const __half* in, int n, int hw, int c, __half* out, cudaStream_t st);
int launch_l2norm(const float* in, int n, int c, float* out, cudaStream_t st);
// gallery sharded over ranks (partition C): per-rank 64-bit keys whose unsigned minimum over ranks is the global winner
int launch_match_keys(const float* emb, int b, const float* gallery, int n, unsigned row0, unsigned long long* keys,
                      cudaStream_t st);
int launch_export_argmax_keys(unsigned long long* keys, int b, unsigned row0, unsigned long long* out, cudaStream_t st);
int launch_vote_keys(const unsigned long long* keys, int b, int* pred, long long* counts, cudaStream_t st);

}  // namespace cfr
