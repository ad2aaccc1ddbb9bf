// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA.  One persistent, warp-specialised kernel serves every dense contraction
// on the certification path:
//   * StyleGAN ConvBlock 3x3 (stylegan_generator_model.py:738-741) with the fused epilogue
//     +noise*w_c +b_c -> LeakyReLU(0.2) and per-(n,c) sum / sum-of-squares for InstanceNorm (:559-562,:420-422)
//   * StyleGAN UpConvBlock (:665-676) as 4 sub-pixel phases x (2x2 taps) on the low-res grid
//   * iresnet50 3x3 / strided 3x3 / 1x1 convs with folded BN, PReLU and residual add (iresnet.py:46-57)
//   * the 25088->512 FC (iresnet.py:153) as a 1x1 conv over a 1x1 "image"
//
// Layout: activations NHWC fp16; weights [rows = (sample?, phase, Cout)][K = (tap, Cin)] fp16, K-major.
// GEMM view: M = 128 output pixels (a TW x TH x TN box of the output grid), N = BN output channels,
// K = taps * Cin walked in 64-element stages.  A tiles come from a 4-D TMA box whose start
// coordinate is shifted by the tap offset: out-of-bounds rows/cols are zero-filled by TMA, which
// *is* the conv zero padding; element strides give stride-2 convs.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/cfr_b200.h"

namespace cfr {

// warps 0 .. 8*MT-1: epilogue groups of four warps (one per TMEM lane quarter), then the TMA producer warp, then the MMA
// issuer + TMEM owner.  Every M tile is drained by TWO groups, each taking half of the tile's BN columns:
// the epilogue is latency-bound (tcgen05.ld, __ldg of the per-channel vectors, shuffle reduces: ncu shows its warps
// issuing 12 % of the time and never waiting on the fused-statistics layers), so twice the warps is twice its rate.
// MT == 1: groups {0: half 0, 1: half 1};  MT == 2: groups {0: tile 0 half 0, 1: tile 1 half 0, 2: tile 0 half 1, 3: tile 1 half 1}
constexpr int kConvThreads = 320;
constexpr int kConvThreadsMT2 = 576;
constexpr int kBM = 128;
constexpr int kMaxPhases = 4;
constexpr int kMaxTaps = 9;
constexpr int kStatsMaxC = 512;

enum Act : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_PRELU = 2, ACT_RELU_POST = 3 };   // 3: ReLU AFTER the residual add

struct ConvParams {
  CUtensorMap tmA;   // activations: dims (C, W, H, N)
  CUtensorMap tmB;   // weights: dims (Kpad, rows)
  // problem
  int N, Hout, Wout;             // conv output grid (for up-conv phases: the low-res grid)
  int TW, TH, TN;                // output-grid box per M tile (TW*TH*TN == 128)
  int tilesX, tilesY, tilesN;    // ceil-div tile counts
  int numPhases, numNTiles;
  int stride;                    // 1 or 2 (input coordinate = output coordinate * stride + tap offset)
  int ntaps;
  int8_t tap_dy[kMaxPhases][kMaxTaps];
  int8_t tap_dx[kMaxPhases][kMaxTaps];
  int CB, nCB, G;                // channel block (<=64), blocks per tap, chunks per 64-wide K stage
  int BN;                        // output channels per tile (multiple of 16, <= 256)
  int swizzleA;                  // bytes: 32 / 64 / 128 (= CB*2)
  int numStages, stageBytes;
  int MT;                        // M tiles per CTA step (1 or 2): two adjacent tiles share every weight stage
  int nAcc;                      // TMEM accumulator sets (of MT x BN columns) in flight: 2, or 1 when MT*BN == 512
  int dbg;                       // ablation bits for profiling only (env CFR_IGEMM_DBG; results are then garbage): 1 skip the
                                 // epilogue's loads / math / stores, 2 skip the MMAs, 4 skip the A-tile loads
  int CG;                        // 1, or 2: a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) works on two adjacent M tiles
                                 // (one per CTA) against a 256-wide N tile of which each CTA stages half the weight rows
  int tilesPerItem;              // M tiles per work item: MT, or 2 for a CTA pair
  int ctrlOffset;                // bytes from the (1024-aligned) shared-memory base to the barrier block
  // ROWS mode (128-pixel-wide layers: an M tile is one image row): per 64-channel K chunk ONE box of `aRows` input rows
  // x 130 pixels serves every tap of two output rows -- taps are row / pixel shifts of the A descriptor, as in the halo
  // kernel -- and, for up-convs, both column phases of a row phase; weights stream through their own ring
  int rows;                      // 0 / 1
  int aRows;                     // input rows per box: 4 (3x3) or 3 (nearest-x2 up-conv row phase)
  int aBoxBytes;                 // aRows * 130 * 128 rounded up to 1024
  int bStages;                   // weight ring depth
  int tapsPerStage;              // weight tiles behind one barrier: all taps of a phase when the ring still holds >= 3 such
                                 // stages (fewer, longer pipeline round trips), else 1
  int rowsPG, rowsNPG;           // phases per work item (1 or 2: same row phase), work-item phase groups
  int rowsMerge;                 // 1 (up-conv): the input column both column phases read feeds ONE N = 2*BN MMA (3 instead of 4
                                 // per row); 2 (3x3): an input row both output rows read feeds ONE N = 256 MMA (12 instead of 18)
  int rowsTap[9];                // rowsMerge == 2: tap index of (dx, dy) in the order dx = -1, 0, 1; dy = +1, 0, -1
  int8_t rowsDyMin[2];           // first input row of the box relative to the item's first output row, per phase group
  CUtensorMap tmA2;              // activations (C, W, H, N), box {64, 130, aRows, 1}
  int wRowsPerSample;            // 0: weights shared by all samples
  int wRowsPerPhase;
  // output
  __half* out;
  float* out32;                  // if non-null, fp32 output instead of fp16
  int outH, outW, outC;
  int out256;                       // fp16 output rows are 32-byte aligned per 16-channel chunk: 256-bit stores
  int oscale;                    // output pixel = grid pixel * oscale + ooff[phase]
  int8_t ooff_y[kMaxPhases], ooff_x[kMaxPhases];
  // epilogue
  const float* bias;             // [Cout] or null
  const float* cbias;            // [(sample?) x phase x 9 x Cout] border-class bias or null
  int cbiasPerSample;
  const float* noise;            // [outH*outW] or null
  const float* noise_w;          // [Cout]
  int act;
  float slope;
  const float* alpha;            // PReLU [Cout]
  const __half* resid;           // [N,outH,outW,residC] or null
  int residC;
  unsigned long long* stat_sum;  // [N, Cout] Q43.20 fixed point, or null (requires TN == 1, numPhases == 1)
  unsigned long long* stat_sq;
  unsigned long long* argmax_keys;   // [N] or null: no store; per-"image" argmax over output channels (gallery match)
  int CoutTotal;
};

// Host-side op: everything needed to (re)launch one conv.
struct ConvOp {
  ConvParams p;
  int grid;
  int smemBytes;
  double flops;   // algorithmic (2*MAC, un-padded dims)
};

// Host-side description == the public C struct (include/cfr_b200.h).
typedef ::cfr_conv_desc ConvSpec;

int conv_build(const ConvSpec& s, ConvOp* op);
// after conv_build: switch the epilogue to argmax mode (keys[n] = max over channels of (score, first index))
void conv_set_argmax(ConvOp* op, unsigned long long* keys);
int conv_launch(const ConvOp& op, cudaStream_t stream);
void set_error(const char* fmt, ...);
const char* last_error();
int num_sms();
bool pdl_enabled();   // programmatic dependent launch of the conv kernels (env CFR_PDL=0 switches it off)
void count_launch(int n = 1);
unsigned long long launch_count();
void profile_enable(int on);
int profile_read(int kind, double* ms, double* work, long long* launches);
bool profile_on();
cudaEvent_t profile_event(int kind);
void profile_account(int kind, double work);

}  // namespace cfr
