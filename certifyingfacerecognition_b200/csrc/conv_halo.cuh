// Halo-resident implicit-GEMM convolution for the HBM-bound high-resolution StyleGAN layers (Cin <= 64).
//
// conv_igemm.cu re-fetches the activation tile once per filter tap; with 16..64 channels a tap is 128 rows of
// only 32..128 bytes, so that kernel is TMA-issue / L2 bound on these layers (profiles/ops_r01.tsv).  Here one
// TMA box brings a (TH+2) x 130 pixel halo band of the NHWC input into shared memory ONCE; every tap of every
// output row in the band is then just a different start address in the UMMA shared-memory descriptor (the
// 32/64/128-byte swizzle is a function of absolute smem address bits, so row-shifted starts stay consistent
// with what TMA wrote).  Weights for all taps/phases stay resident in smem for the CTA's lifetime.
//
// The previous layer's InstanceNorm + AdaIN (x = y*A[n,c] + B[n,c], stylegan_generator_model.py:420-422,:505) is
// fused in one of two ways.  FOLD (what the synthesis program uses for all five layers): A goes into per-sample fp16
// weights, and the shift B, this layer's bias and the noise gain ride on one auxiliary 16-wide row per output pixel
// {noise, inside-image indicators of the 3x3 input neighbourhood} consumed by ONE extra MMA -- exact at the borders,
// nothing left for the loaders or the epilogue to do.  Non-FOLD: the loader warps apply the affine to the band in
// shared memory (in-bounds pixels only, so the conv's zero padding stays zero) and the epilogue adds noise*w_c + b_c.
// Epilogue in both: LeakyReLU(0.2), fp16 store (one 256-bit store per 16 channels), per-(n,c) sum / sumsq.
//
// MMA issue orders (FOLD): 3x3 convs are row-stationary -- input row i times [W(dy=+1,dx)|W(dy=0,dx)|W(dy=-1,dx)]
// (N = 3*Cout) accumulates into output rows i-2, i-1, i, whose accumulators are adjacent TMEM columns (`rowmma`);
// the nearest-x2 up-conv issues one MMA per input position for all the sub-pixel phases that read it (`upshare`);
// the composite blur o up-conv shares the A operand between its four phases (one N = 4*Cout MMA per tap).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/cfr_b200.h"

namespace cfr {

constexpr int kLoaderWarps = 8;
constexpr int kMmaWarps = 4;          // tcgen05.mma issue is per-thread serial (~30 SASS instr per MMA incl. R2UR moves);
                                      // with N = 16..64 the MMAs are tiny, so several warps issue alternate tiles
#ifndef CFR_HALO_EPI_GROUPS
#define CFR_HALO_EPI_GROUPS 2
#endif
constexpr int kEpiGroups = CFR_HALO_EPI_GROUPS;       // each group = 4 warps (one per TMEM lane quarter); tile t -> group t % kEpiGroups
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kHaloMmaWarp0 = kEpiWarps + kLoaderWarps;   // [epilogue warps][loader/transform warps][MMA issuers]
constexpr int kHaloThreads = (kEpiWarps + kLoaderWarps + kMmaWarps) * 32;
// Register re-allocation between the roles (setmaxnreg; each role is a whole number of 4-warp groups): the kernel starts
// with 96 registers per thread (640 threads); the MMA issuers and loaders hand registers to the epilogue warps, whose
// per-channel accumulators otherwise spill.  8*32*kRegsEpi + 8*32*kRegsLoader + 4*32*kRegsMma <= 640 * 96.
#ifndef CFR_HALO_REGS_EPI
#define CFR_HALO_REGS_EPI 128
#define CFR_HALO_REGS_LOADER 72
#define CFR_HALO_REGS_MMA 64
#endif
constexpr int kRegsEpi = CFR_HALO_REGS_EPI, kRegsLoader = CFR_HALO_REGS_LOADER, kRegsMma = CFR_HALO_REGS_MMA;
constexpr int kHaloW = 130;         // 128 output columns + 1 halo column each side

struct HaloParams {
  const __half* in;                 // NHWC fp16 input [N,H,W,Cin]
  CUtensorMap tmW;                  // weights (Cin, [n *] phases*taps*Cout), box {Cin, wBoxRows}
  CUtensorMap tmIn;                 // tmaBand: input (Cin, W, H, N), box {Cin, 130, TH+2, 1}, OOB zero fill == conv padding
  int tmaBand;                      // FOLD: the band is fetched by ONE TMA box instead of ~3.6-7 K cp.async per band
  CUtensorMap tmWa;                 // FOLD: aux weight tiles (16, n*phases*taps*Cout), box {16, wBoxRows}
  int fold;                         // per-sample folded weights + aux band
  int composite;                    // blur o up-conv: 4 phases x 9 taps, 8 weight sets (first/last-row variants), 4 noise slots
  int wsets;                        // weight sets resident in smem (numPhases, or 8 when composite)
  int upshare;                      // FOLD 4-phase up-conv: phases share the A operand per input position (weights in layout 3)
  int rowmma;                       // 3x3 FOLD conv, Cout <= 32: row-stationary MMA order (1 per group, 2 per band), weights in layout 2
  const float* corr;                // composite: border-column correction [N][2 sides][outH][Cout] fp32
  int dbg;                          // ablation bits for profiling only (env CFR_HALO_DBG): 1 skip epilogue math/store,
                                    // 2 skip band cp.async, 4 skip aux rows, 8 skip MMAs (results are then garbage)
  int auxBytes, wAuxBytes;          // aux band bytes per stage / aux weight bytes (0 unless fold)
  int N, H, W;                      // conv output grid == input grid (stride 1)
  int Cin, Cout;
  int TH;                           // output rows per band
  int bandsX, bandsY;               // per image
  uint32_t magicX, magicY;          // floor(2^32 / bands) + 1: divisions in decode_band as multiply-high
  int numPhases, ntaps;
  int8_t tap_dy[4][9], tap_dx[4][9];
  int rowBytes;                     // Cin*2 == swizzle width of the halo band and of the weight rows
  int haloBytes;                    // (TH+2)*130*rowBytes rounded up to 1024
  int haloStages;                   // band buffers in the ring (2 or 3)
  int wBytes, wRows, wBoxRows;
  int accStages;                    // TMEM accumulator stages (each Cout columns, padded to >=16)
  // affine on load (may be null)
  const float* inA; const float* inB;   // [N, Cin]
  // output
  __half* out; int outH, outW, outC, oscale;
  int8_t ooff_y[4], ooff_x[4];
  const float* bias; const float* noise; const float* noise_w;
  const void* noise_tab;            // FOLD: packed per-pixel aux-row head (see launch_pack_noise), filled by cp.async
  int act; float slope;
  unsigned long long* stat_sum; unsigned long long* stat_sq;  // [N, Cout] Q43.20 fixed point, or null
  // sparse store (one phase, oscale 1, square output): only pixels whose row AND column have keep_map[.] >= 0 are written,
  // into a compact [N][keepDim][keepDim][Cout] buffer; statistics still cover every pixel (cfr_conv_desc.keepMap)
  const int* keep_map; int keepDim;
};

struct HaloOp {
  HaloParams p;
  int grid, smemBytes;
  double flops;
  double bytes;   // algorithmic HBM traffic per launch
};

int halo_build(const cfr_conv_desc& s, const float* inA, const float* inB, const void* w_aux, int composite,
               const float* corr, HaloOp* op);
// per-sample weight folding for the FOLD variant (see conv_halo.cu)
int launch_fold_weights(const float* base_w, const float* inA, const float* inB, const float* bias, const float* noise_w,
                        const int8_t* tap_dy, const int8_t* tap_dx, int n, int phases, int ntaps, int cout, int cin,
                        int layout, __half* w_main, __half* w_aux, cudaStream_t st);   // 0 plain, 1 composite, 2 row-stationary, 3 shared-A up-conv
// composite up-conv+blur: exact values for the first / last hi-res column (see engine.composite_upconv_weights)
int launch_upblur_corr(const __half* y, const float* inA, const float* inB, const float* corr_d, int n, int h, int w,
                       int cin, int cout, float* corr, cudaStream_t st);
// FOLD aux rows: the part of an aux row that comes from memory, pre-packed once per run so the loader can cp.async it
// (no register-staged noise loads on its critical path).  mode 0: out[y][x] = {half noise(y,x), half (y>0 && x>0)};
// mode 1 (composite, low-res grid h x w, noise is 2h x 2w): out[y][x] = 4 halves noise(2y+a, 2x+b), (a,b) row-major.
int launch_pack_noise(const float* noise, int h, int w, int mode, void* out, cudaStream_t st);
// FOLD 4-phase (nearest x2 + 3x3) up-conv whose phases can share A operands: 18 weight tiles per sample instead of 16
constexpr int kUpShareTiles = 18;
bool halo_upshare_ok(const cfr_conv_desc& s);
int halo_launch(const HaloOp& op, cudaStream_t stream);

}  // namespace cfr
