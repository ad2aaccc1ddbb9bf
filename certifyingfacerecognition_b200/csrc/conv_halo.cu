// Halo-resident tcgen05 convolution kernel -- see conv_halo.cuh.
#include "conv_halo.cuh"
#include <type_traits>
#include "conv_igemm.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace cfr {

__device__ __forceinline__ float warp_reduce16h(float (&v)[16], uint32_t lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step;
    const int cnt = 8 >> step;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      float send = upper ? v[i] : v[i + cnt];
      float keep = upper ? v[i + cnt] : v[i];
      float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = keep + recv;
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
__device__ __forceinline__ int reduce16_channel_h(uint32_t lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

struct Band {
  int n, y0, x0, rows;
};
__device__ __forceinline__ Band decode_band(const HaloParams& p, int b) {
  // divisions by the (runtime) band counts as multiply-high with precomputed reciprocals (exact: halo_build checks
  // b < 2^21, divisors <= 2^10); every role decodes every band, and two integer divisions cost ~60 instructions each time
  Band r;
  // (a divisor of 1 has no 32-bit reciprocal: magic == 0 marks it)
  const int t = p.magicX ? static_cast<int>(__umulhi(static_cast<uint32_t>(b), p.magicX)) : b;
  const int bx = b - t * p.bandsX;
  r.n = p.magicY ? static_cast<int>(__umulhi(static_cast<uint32_t>(t), p.magicY)) : t;
  const int by = t - r.n * p.bandsY;
  r.x0 = bx * 128;
  r.y0 = by * p.TH;
  r.rows = min(p.TH, p.H - r.y0);
  return r;
}

// FOLD: the previous layer's IN+AdaIN scale is folded into per-sample weights, and its shift, this layer's bias and
// noise ride on an auxiliary 16-channel row per OUTPUT pixel {noise, inside-image indicators of its 3x3 input
// neighbourhood, 0..} consumed by ONE extra MMA per tile,
// so neither the loaders nor the epilogue touch them.
// COMP: composite blur o up-conv (4 phases x 9 taps, first/last-row weight sets, border-column correction); a separate
// instantiation so that the plain conv paths keep their register allocation / schedule.
template <int COUT, bool FOLD, bool COMP = false>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int NCH = COUT / 16;                 // 16-column epilogue chunks
  // COUT == 64: both epilogue warp groups drain EVERY tile, each its own half of the channels, so that a thread's
  // per-channel statistics (2 x 32 registers) stay in registers like in the narrower variants -- the per-chunk
  // shuffle transpose-reduce it replaces was more than half of that epilogue's instructions
  constexpr bool SPLIT = COUT == 64;
  constexpr int NCH_T = SPLIT ? NCH / 2 : NCH;   // 16-column chunks per epilogue thread
  constexpr bool REG_STATS = true;               // keep per-thread channel sums in registers across a band
  constexpr int ACC_COLS = COUT;                 // TMEM columns per accumulator
  constexpr int G = COUT == 64 ? 2 : 4;          // tiles per accumulator group (one mbarrier handshake per group)
  constexpr int AS = 512 / (G * COUT);           // accumulator groups resident in TMEM (8 / 4 / 4)
  constexpr int asLog = AS == 8 ? 3 : 2;
#ifndef CFR_HALO_KB16
#define CFR_HALO_KB16 2
#endif
#ifndef CFR_HALO_KBC
#define CFR_HALO_KBC 1
#endif
  constexpr int KB = COUT == 16 ? (COMP ? CFR_HALO_KBC : CFR_HALO_KB16) : 1;   // tiles / 16-column chunks per tile whose TMEM loads the epilogue
  constexpr int CBN = 1;                               // keeps in flight together (registers: KB * CBN * 16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NS = p.haloStages;                                   // 2 or 3 halo band buffers
  uint8_t* aux = smem + NS * p.haloBytes;                        // [NS] aux bands (FOLD only; auxBytes == 0 otherwise)
  uint8_t* wsm = aux + NS * p.auxBytes;
  uint8_t* wasm = wsm + p.wBytes;                                // aux weight tiles (FOLD only)
  uint8_t* ctrl = wasm + p.wAuxBytes;
  uint64_t* hready = reinterpret_cast<uint64_t*>(ctrl);    // [3] band loaded (+ affine applied)
  uint64_t* hempty = hready + 3;                           // [3]
  uint64_t* wbar = hempty + 3;                             // [1]
  uint64_t* wdrain = wbar + 1;                             // [1] all MMAs reading the current weights are done
  uint64_t* tfull = wdrain + 1;                            // [16]
  uint64_t* tempty = tfull + 16;                           // [16]
  uint64_t* hland = tempty + 16;                           // [3] (+1 pad) non-FOLD tmaBand: band landed (TMA complete_tx)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hland + 4);
  float* sA = reinterpret_cast<float*>(tmem_slot + 4);     // [64]
  float* sB = sA + 64;                                     // [64]
  // per-(epilogue warp, channel) float partials; each warp owns its slice (64-bit shared atomics are CAS spin loops)
  float* s_sum = sB + 64;                                  // [kEpiWarps][COUT]
  float* s_sq = s_sum + kEpiWarps * COUT;                  // [kEpiWarps][COUT]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const int totalBands = p.N * p.bandsY * p.bandsX;
  const int per = (totalBands + gridDim.x - 1) / gridDim.x;
  const int band0 = blockIdx.x * per;
  const int band1 = min(totalBands, band0 + per);
  const bool affine = p.inA != nullptr;
  constexpr uint32_t tmemCols = 512;

  if (warp == kHaloMmaWarp0 && lane == 0) {
    tma_prefetch_desc(&p.tmW);
    if (FOLD) tma_prefetch_desc(&p.tmWa);
  }
  if (warp == kHaloMmaWarp0) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) {
        mbar_init(&hready[i], (FOLD && p.tmaBand) ? 2 : 1);   // FOLD tmaBand: + the expect_tx arrive of the TMA issuer
        mbar_init(&hland[i], 1);
        mbar_init(&hempty[i], kMmaWarps);
      }
      mbar_init(wbar, 1);
      mbar_init(wdrain, kMmaWarps);
      for (int i = 0; i < 16; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], SPLIT ? 8 : 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * kEpiWarps * COUT; i += kHaloThreads) s_sum[i] = 0.f;
  if constexpr (FOLD) {                      // aux rows are {noise, indicator, 0 x14}: zero everything once
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < (NS * p.auxBytes) >> 4; i += kHaloThreads) reinterpret_cast<uint4*>(aux)[i] = z;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp >= kHaloMmaWarp0) {
    // ================================================================ MMA issuers (kMmaWarps warps, alternate tiles)
    // The whole warp runs the loop (so descriptor words stay warp-uniform); one elected lane issues
    // tcgen05.mma / commit.  Warp m issues the tiles with (tile index mod kMmaWarps) == m.
    setmaxnreg_dec<kRegsMma>();
    const uint32_t mw = warp - kHaloMmaWarp0;
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, COUT);
    [[maybe_unused]] const uint32_t idesc2 = make_idesc_f16(128, 2 * COUT), idesc4 = make_idesc_f16(128, 4 * COUT);
    [[maybe_unused]] const uint32_t idesc3 = make_idesc_f16(128, 3 * COUT);
    const uint32_t dhi = smem_desc_hi(8 * p.rowBytes, p.rowBytes);
    const int kPer = p.Cin / 16;
    const uint32_t rb16 = p.rowBytes >> 4;                 // row pitch in 16-byte units
    const uint32_t w_lo = smem_desc_lo(smem_u32(wsm));
    const uint32_t w_tap = COUT * rb16;                    // B tile pitch per (phase, tap)
    uint32_t toff[4][9];                                   // A offsets of every (phase, tap), 16-byte units
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int t = 0; t < 9; ++t)
        toff[ph][t] = (ph < p.numPhases && t < p.ntaps)
                          ? ((1 + p.tap_dy[ph][t]) * kHaloW + 1 + p.tap_dx[ph][t]) * rb16 : 0u;
    const uint32_t row_step = kHaloW * rb16;
    const uint32_t dhi_aux = smem_desc_hi(256, 32);        // aux operands: 32-byte rows, SWIZZLE_32B
    const uint32_t wa_lo = smem_desc_lo(smem_u32(wasm));
    const uint32_t aux_row_step = 128 * 2;                 // one aux row (32 B) per output pixel, 128 pixels per tile row
    auto load_weights = [&](int n) {      // every (phase, tap) tile; FOLD: the per-sample set of image n
      const int row_base = FOLD ? n * p.wRows : 0;
      const int aux_rows = p.wsets * COUT;                 // one aux B tile per weight set
      mbar_expect_tx(wbar, p.wRows * p.rowBytes + (FOLD ? aux_rows * 32 : 0));
      for (int r0 = 0; r0 < p.wRows; r0 += p.wBoxRows)
        tma_load_2d(wsm + static_cast<size_t>(r0) * p.rowBytes, &p.tmW, wbar, 0, row_base + r0);
      if (FOLD) tma_load_2d(wasm, &p.tmWa, wbar, 0, n * aux_rows);
    };
    uint32_t wphase = 0, dphase = 0;
    int cur_n = -1;
    if (!FOLD) {
      if (mw == 0 && leader) load_weights(0);
      __syncwarp();
      mbar_wait(wbar, 0);
    }
    int hs = 0;
    uint32_t hphase = 0;
    uint32_t gbase = 0;                    // accumulator-group counter at the start of the band
    // One tfull/tempty handshake covers a GROUP of G tiles (G accumulators of COUT columns side by side in TMEM):
    // with 4 KB tiles the per-tile mbarrier round trips were ~3/4 of the kernel time (profiles/ablation_r01.log).
    // numPhases == 4: group = the 4 sub-pixel phases of one low-res row;  numPhases == 1: group = G consecutive rows.
    // One tile = ntaps x (Cin/16) MMAs (+1 aux MMA when FOLD).  The (ntaps, Cin/16) pair is dispatched ONCE per tile to
    // a fully unrolled, branch-free sequence of predicated MMAs (only the elected lane issues): the generic loop with
    // per-MMA run-time checks cost ~40 SASS instructions per MMA and kept the issuing warps 70 % busy (ncu, r2r).
    const uint32_t lead = (leader && !(p.dbg & 8)) ? 1u : 0u;
    auto issue_tile_t = [&](auto NTc, auto KPc, uint32_t d_tmem, uint32_t a_row, uint32_t x_row, uint32_t b_lo,
                            uint32_t ba_lo, const uint32_t (&to)[9]) {
      constexpr int NT = decltype(NTc)::value, KP = decltype(KPc)::value;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const uint32_t a_lo = a_row + to[t];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (t == 0 && k == 0) umma_f16_pred<false>(d_tmem, a_lo, dhi, b_lo, dhi, idesc, lead);
          else umma_f16_pred<true>(d_tmem, a_lo + 2 * k, dhi, b_lo + t * w_tap + 2 * k, dhi, idesc, lead);
        }
      }
      if constexpr (FOLD) umma_f16_pred<true>(d_tmem, x_row, dhi_aux, ba_lo, dhi_aux, idesc, leader ? 1u : 0u);
    };
    using std::integral_constant;
    auto issue_tile = [&](uint32_t d_tmem, uint32_t a_row, uint32_t x_row, uint32_t b_lo, uint32_t ba_lo,
                          const uint32_t (&to)[9]) {
      if (p.ntaps == 9) {
        if (kPer == 1) issue_tile_t(integral_constant<int, 9>{}, integral_constant<int, 1>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
        else if (kPer == 2) issue_tile_t(integral_constant<int, 9>{}, integral_constant<int, 2>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
        else issue_tile_t(integral_constant<int, 9>{}, integral_constant<int, 4>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
      } else if (p.ntaps == 4) {
        if (kPer == 1) issue_tile_t(integral_constant<int, 4>{}, integral_constant<int, 1>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
        else if (kPer == 2) issue_tile_t(integral_constant<int, 4>{}, integral_constant<int, 2>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
        else issue_tile_t(integral_constant<int, 4>{}, integral_constant<int, 4>{}, d_tmem, a_row, x_row, b_lo, ba_lo, to);
      }                                     // (halo_build only admits 4 or 9 taps)
    };
    for (int b = band0; b < band1; ++b) {
      const Band bd = decode_band(p, b);
      if (FOLD && bd.n != cur_n) {
        if (cur_n >= 0) {                  // drain: every MMA that reads the old weights must have completed
          if (leader) umma_commit(wdrain);
          __syncwarp();
          mbar_wait(wdrain, dphase);
          dphase ^= 1;
        }
        if (mw == 0 && leader) load_weights(bd.n);
        __syncwarp();
        mbar_wait(wbar, wphase);
        wphase ^= 1;
        cur_n = bd.n;
      }
      mbar_wait(&hready[hs], hphase);
      tc_fence_after();
      const uint32_t a_band = smem_desc_lo(smem_u32(smem + hs * p.haloBytes));
      const uint32_t x_band = smem_desc_lo(smem_u32(aux + hs * p.auxBytes));
      const int ngroups = p.numPhases == 4 ? bd.rows : (bd.rows + G - 1) / G;
      if (FOLD && !COMP && p.rowmma == 2) {
        // Row-stationary order (see the per-group variant below) over the WHOLE band: one warp owns a band (same-thread
        // MMAs execute in order, which the overlapping accumulate chains need) and walks its input rows once, 3 x kPer
        // MMAs of N = 3*COUT each; accumulator groups are acquired / handed to the epilogue as the walk reaches them.
        if (((b - band0) & (kMmaWarps - 1)) == static_cast<int>(mw)) {
          const int rows = bd.rows;
          auto tile_d = [&](int o) {
            const uint32_t gc = gbase + o / G;
            return tmem_base + ((gc & (AS - 1)) * G + (o % G)) * ACC_COLS;
          };
          for (int i = 0; i < rows + 2; ++i) {
            if (i < rows) {
              if (i % G == 0) {
                const uint32_t gc = gbase + i / G;
                mbar_wait(&tempty[gc & (AS - 1)], ((gc >> asLog) & 1) ^ 1);
                tc_fence_after();
              }
              umma_f16_pred<false>(tile_d(i), x_band + i * aux_row_step, dhi_aux, wa_lo, dhi_aux, idesc, lead);
            }
            const int o_hi = min(i, rows - 1);
            const uint32_t a_row = a_band + i * row_step;
            for (int o = max(i - 2, 0); o <= o_hi;) {
              int e = o;                                       // longest run of TMEM-contiguous accumulators (ring wrap)
              while (e < o_hi && tile_d(e + 1) == tile_d(e) + ACC_COLS) ++e;
              const int nt = e - o + 1;
              const uint32_t id = nt == 1 ? idesc : (nt == 2 ? idesc2 : idesc3);
              const uint32_t d = tile_d(o);
              const uint32_t bt = static_cast<uint32_t>(2 - i + o);   // first dy tile of the triple: dy = 1 - bt
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < kPer)
                    umma_f16_pred<true>(d, a_row + dx * rb16 + 2 * k, dhi, w_lo + (dx * 3 + bt) * w_tap + 2 * k, dhi, id, lead);
              }
              o = e + 1;
            }
            const int done = i - 2;                            // output row i-2 has all nine taps now
            if (done >= 0 && ((done % G) == G - 1 || done == rows - 1)) {
              if (leader) umma_commit(&tfull[(gbase + done / G) & (AS - 1)]);
              __syncwarp();
            }
          }
        }
      } else
      for (int j = 0; j < ngroups; ++j) {
        const uint32_t gc = gbase + j;
        if ((gc & (kMmaWarps - 1)) != mw) continue;
        const uint32_t slot = gc & (AS - 1);
        mbar_wait(&tempty[slot], ((gc >> asLog) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + slot * (G * ACC_COLS);
        if constexpr (COMP) {
          // composite blur o up-conv: all four phases are 3x3 convs of the SAME low-res window, so they share the A
          // operand: one N = 4*COUT MMA per (tap, k-step) fills the four side-by-side accumulators of the group
          // (weights are tap-major, sets ordered [4 5 | 0 1 2 3 | 6 7]).  The first / last image row swap in their
          // own sets for phases 0,1 / 2,3: two N = 2*COUT MMAs there.
          const uint32_t a_row = a_band + j * row_step, x_row = x_band + j * aux_row_step;
          const int gy = bd.y0 + j;
          const uint32_t set_u = COUT * rb16;                       // one weight set of one tap, 16-byte units
          const uint32_t tap_u = 8 * set_u;
          if (gy != 0 && gy != p.H - 1) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t a_lo = a_row + toff[0][t], b_lo = w_lo + t * tap_u + 2 * set_u;
              if (t == 0) umma_f16_pred<false>(d0, a_lo, dhi, b_lo, dhi, idesc4, lead);
              else umma_f16_pred<true>(d0, a_lo, dhi, b_lo, dhi, idesc4, lead);
              if (kPer > 1) umma_f16_pred<true>(d0, a_lo + 2, dhi, b_lo + 2, dhi, idesc4, lead);
            }
            umma_f16_pred<true>(d0, x_row, dhi_aux, wa_lo + 2 * (COUT * 2), dhi_aux, idesc4, leader ? 1u : 0u);
          } else {
            const uint32_t s_lo = gy == 0 ? 0u : 2u, s_hi = gy == 0 ? 4u : 6u;   // first set of the (0,1) / (2,3) pair
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const uint32_t so = hf == 0 ? s_lo : s_hi, dd = d0 + hf * 2 * ACC_COLS;
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = a_row + toff[0][t], b_lo = w_lo + t * tap_u + so * set_u;
                if (t == 0) umma_f16_pred<false>(dd, a_lo, dhi, b_lo, dhi, idesc2, lead);
                else umma_f16_pred<true>(dd, a_lo, dhi, b_lo, dhi, idesc2, lead);
                if (kPer > 1) umma_f16_pred<true>(dd, a_lo + 2, dhi, b_lo + 2, dhi, idesc2, lead);
              }
              umma_f16_pred<true>(dd, x_row, dhi_aux, wa_lo + so * (COUT * 2), dhi_aux, idesc2, leader ? 1u : 0u);
            }
          }
        } else if (FOLD && COUT <= 32 && p.upshare) {
          // 4-phase up-conv (nearest x2 + 3x3): phase (py,px) reads the low-res taps {py-1,py} x {px-1,px}, so each of the
          // nine input positions serves 1, 2 or 4 phases.  One MMA per (position, k-step) with the weights of the phases
          // that use it side by side (layout 3; the (0,+-1) positions span three accumulators with a zero tile in the
          // middle) and ONE N = 4*COUT aux MMA: 1 + 9*kPer MMAs per low-res row tile instead of 4 + 16*kPer.
          const uint32_t a_row = a_band + j * row_step, x_row = x_band + j * aux_row_step;
          umma_f16_pred<false>(d0, x_row, dhi_aux, wa_lo, dhi_aux, idesc4, lead);
          constexpr int P_DY[9] = {0, -1, 1, 0, 0, -1, -1, 1, 1}, P_DX[9] = {0, 0, 0, -1, 1, -1, 1, -1, 1};
          constexpr int P_D[9] = {0, 0, 2, 0, 1, 0, 1, 2, 3}, P_N[9] = {4, 2, 2, 3, 3, 1, 1, 1, 1};
          constexpr int P_B[9] = {0, 4, 6, 8, 11, 14, 15, 16, 17};
#pragma unroll
          for (int P = 0; P < 9; ++P) {
            const uint32_t a_lo = a_row + ((1 + P_DY[P]) * kHaloW + 1 + P_DX[P]) * rb16, b_lo = w_lo + P_B[P] * w_tap;
            const uint32_t id = P_N[P] == 4 ? idesc4 : (P_N[P] == 3 ? idesc3 : (P_N[P] == 2 ? idesc2 : idesc));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < kPer) umma_f16_pred<true>(d0 + P_D[P] * ACC_COLS, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, id, lead);
          }
        } else if (p.numPhases == 4) {     // G == 4: tile k of the group == phase k of low-res row j
          const uint32_t a_row = a_band + j * row_step, x_row = x_band + j * aux_row_step;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            issue_tile(d0 + k * ACC_COLS, a_row, x_row, w_lo + k * p.ntaps * w_tap, wa_lo + k * (COUT * 2), toff[k]);
        } else if (FOLD && !COMP && COUT <= 32 && p.rowmma == 1) {
          // Row-stationary issue order for 3x3 convs with few output channels.  The MMA stream of these layers is bound
          // by the issuing warps' instruction overhead and the 128-row A operand read, both per MMA and independent of
          // N (profiles/ncu_r01_notes.md section 9), so instead of 9 MMAs per OUTPUT row this issues 3 per INPUT row:
          // input row i times [W(dy=+1,dx) | W(dy=0,dx) | W(dy=-1,dx)] (N = 3*COUT; weights in layout 2) lands in the
          // accumulators of output rows i-2, i-1, i, which sit side by side in the group's TMEM columns.  The aux MMA
          // (shift / bias / noise row) of an output row goes first and initialises its accumulator.  A group of 4
          // output rows reads 6 input rows: 4 + 18 MMAs instead of 40.
          const int r0 = j * G, gr = min(G, bd.rows - r0);
          const uint32_t a0 = a_band + r0 * row_step, x0 = x_band + r0 * aux_row_step;
          auto issue_rows = [&](auto GRc) {
            constexpr int GR = decltype(GRc)::value, NI = GR + 2, H = (NI + 1) / 2;
#pragma unroll
            for (int k = 0; k < GR; ++k)
              umma_f16_pred<false>(d0 + k * ACC_COLS, x0 + k * aux_row_step, dhi_aux, wa_lo, dhi_aux, idesc, lead);
            // input rows a and a + H touch disjoint accumulators: alternate them to keep two accumulate chains going
#pragma unroll
            for (int a = 0; a < H; ++a) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  if (kk < kPer) {
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                      const int ii = a + s2 * H;
                      if (ii < NI) {
                        const int o_lo = ii - 2 > 0 ? ii - 2 : 0, o_hi = ii < GR - 1 ? ii : GR - 1;
                        const int nt = o_hi - o_lo + 1, bt = 2 - ii + o_lo;       // first dy tile of the triple: dy = 1 - bt
                        const uint32_t id = nt == 1 ? idesc : (nt == 2 ? idesc2 : idesc3);
                        umma_f16_pred<true>(d0 + o_lo * ACC_COLS, a0 + ii * row_step + dx * rb16 + 2 * kk, dhi,
                                            w_lo + (dx * 3 + bt) * w_tap + 2 * kk, dhi, id, lead);
                      }
                    }
                  }
                }
              }
            }
          };
          if (gr == 4) issue_rows(integral_constant<int, 4>{});
          else if (gr == 3) issue_rows(integral_constant<int, 3>{});
          else if (gr == 2) issue_rows(integral_constant<int, 2>{});
          else issue_rows(integral_constant<int, 1>{});
        } else if (FOLD && p.ntaps == 9 && kPer == 1 && j * G + G <= bd.rows && !(p.dbg & 16)) {
          // tile k of the group == output row j*G + k.  Tap-major over the G tiles: consecutive MMAs write DIFFERENT
          // accumulators, so the tensor pipe is not serialised on one accumulate chain (N = 16 MMAs are latency-,
          // not throughput-limited)
          const uint32_t a0 = a_band + j * G * row_step, x0 = x_band + j * G * aux_row_step;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int k = 0; k < G; ++k) {
              if (t == 0) umma_f16_pred<false>(d0 + k * ACC_COLS, a0 + k * row_step + toff[0][t], dhi, w_lo, dhi, idesc, lead);
              else umma_f16_pred<true>(d0 + k * ACC_COLS, a0 + k * row_step + toff[0][t], dhi, w_lo + t * w_tap, dhi, idesc, lead);
            }
          }
#pragma unroll
          for (int k = 0; k < G; ++k)
            umma_f16_pred<true>(d0 + k * ACC_COLS, x0 + k * aux_row_step, dhi_aux, wa_lo, dhi_aux, idesc, leader ? 1u : 0u);
        } else {                           // tile k of the group == output row j*G + k
#pragma unroll
          for (int k = 0; k < G; ++k) {
            const int r = j * G + k;
            if (r < bd.rows)
              issue_tile(d0 + k * ACC_COLS, a_band + r * row_step, x_band + r * aux_row_step, w_lo, wa_lo, toff[0]);
          }
        }
        if (leader) umma_commit(&tfull[slot]);
        __syncwarp();
      }
      gbase += ngroups;
      if (leader) umma_commit(&hempty[hs]);          // halo stage free once every MMA of this band has read it
      __syncwarp();
      if (++hs == NS) {
        hs = 0;
        hphase ^= 1;
      }
    }
  } else if (warp >= kEpiWarps) {
    // ================================================================ band loader + affine-on-load (warps 8..15)
    // cp.async 16-byte chunks straight into the swizzled operand layout (TMA boxes with 32..128-byte rows are
    // row-rate bound: profiles/ncu_r01_halo_tma.txt), zero-filling out-of-image pixels == the conv padding.
    setmaxnreg_dec<kRegsLoader>();
    constexpr int LT = kLoaderWarps * 32;
    const int tt = threadIdx.x - kEpiWarps * 32;
    const int nchLog = p.rowBytes == 32 ? 1 : (p.rowBytes == 64 ? 2 : 3);
    const int nch = 1 << nchLog;
    const int RC = kHaloW * nch;               // 16-byte chunks per band row (260 / 520 / 1040)
    const size_t imgStride = static_cast<size_t>(p.H) * p.W * p.Cin;
    const size_t rowStride = static_cast<size_t>(p.W) * p.Cin;
    const int lc = tt & (nch - 1);             // this thread's 8-channel group (LT % nch == 0: constant)

    // Thread t owns chunks q = t, t + LT, .. (< RC) of EVERY band row: per chunk the inner loop is an address add,
    // a swizzle XOR and one cp.async (consecutive threads -> consecutive 16-byte chunks of one image row).
    auto issue = [&](const Band& bd, int hs) {
      const uint32_t hb_addr = smem_u32(smem + hs * p.haloBytes);
      const __half* img = p.in + static_cast<size_t>(bd.n) * imgStride + lc * 8;
      if (p.tmaBand) {
        if (tt == 0) {                        // one 4-D box: (TH + 2) x 130 pixels, out-of-image parts zero-filled
          // FOLD: nothing touches the band before the MMAs, so the box signals `hready` itself; otherwise the loader
          // warps wait for it on `hland`, apply the affine in place and then arrive on `hready`
          uint64_t* bar = FOLD ? &hready[hs] : &hland[hs];
          mbar_expect_tx(bar, static_cast<uint32_t>((p.TH + 2) * kHaloW * p.rowBytes));
          tma_load_4d(smem + hs * p.haloBytes, &p.tmIn, bar, 0, bd.x0 - 1, bd.y0 - 1, bd.n);
        }
      } else if (!(p.dbg & 2)) {
        // rows of the band that lie inside the image: [r_lo, r_hi) of TH + 2; the others are zero-filled
        const int R = p.TH + 2;
        const int r_lo = bd.y0 == 0 ? 1 : 0;
        const int r_hi = min(R, p.H - (bd.y0 - 1));
        const uint32_t rowB = static_cast<uint32_t>(RC) << 4;             // smem bytes per band row
        const uint32_t swm = static_cast<uint32_t>(nch - 1);
        // (a) whole multiples of LT chunks per row: a thread keeps its column(s), walks down the rows with an address
        //     add, a swizzle XOR and one cp.async per chunk (all checks hoisted out of the row loop)
        const int nfull = RC / LT;
        for (int c = 0; c < nfull; ++c) {
          const int q = tt + c * LT;
          const int gx = bd.x0 - 1 + (q >> nchLog);
          const bool colok = static_cast<unsigned>(gx) < static_cast<unsigned>(p.W);
          const uint32_t sz = colok ? 16u : 0u;
          const __half* src = img + static_cast<size_t>(colok ? gx : 0) * p.Cin +
                              static_cast<long long>(bd.y0 - 1 + r_lo) * static_cast<long long>(rowStride);
          uint32_t lin = hb_addr + (static_cast<uint32_t>(q) << 4);
          int row = 0;
          for (; row < r_lo; ++row, lin += rowB) cp_async16(lin ^ (((lin >> 7) & swm) << 4), p.in, 0u);
#pragma unroll 2
          for (; row < r_hi; ++row, lin += rowB, src += rowStride) cp_async16(lin ^ (((lin >> 7) & swm) << 4), src, sz);
          for (; row < R; ++row, lin += rowB) cp_async16(lin ^ (((lin >> 7) & swm) << 4), p.in, 0u);
        }
        // (b) the RC % LT left-over chunks of every row (4 / 8 / 16), spread as (row, chunk) pairs over the threads
        const int left = RC - nfull * LT, leftLog = nchLog + 1;            // left == 2 * nch
        for (int idx = tt; idx < R * left; idx += LT) {
          const int row = idx >> leftLog, q = nfull * LT + (idx & (left - 1));
          const int gx = bd.x0 - 1 + (q >> nchLog), gy = bd.y0 - 1 + row;
          const bool ok = static_cast<unsigned>(gx) < static_cast<unsigned>(p.W) && row >= r_lo && row < r_hi;
          const __half* src = img + static_cast<size_t>(ok ? gx : 0) * p.Cin +
                              static_cast<long long>(ok ? gy : 0) * static_cast<long long>(rowStride);
          const uint32_t lin = hb_addr + static_cast<uint32_t>(row) * rowB + (static_cast<uint32_t>(q) << 4);
          cp_async16(lin ^ (((lin >> 7) & swm) << 4), src, ok ? 16u : 0u);
        }
      }
      if (FOLD && !(p.dbg & 4)) {
        // aux rows, one per OUTPUT pixel of the band: k0 = noise, k(1 + 3*(dy+1) + (dx+1)) = 1 if input pixel
        // (y+dy, x+dx) lies inside the image (so the folded IN/AdaIN shift respects the zero padding), rest 0.
        // The head of the row that depends on memory ({noise, k1} / composite: 4 noise values) is cp.async'ed from
        // the packed table; the position-only indicators are plain shared stores to the other bytes of the row.
        const uint32_t xb_addr = smem_u32(aux + hs * p.auxBytes);
        const int cx = tt & 127;
        const int gx = bd.x0 + cx;
        const float cl = gx > 0 ? 1.f : 0.f, cr = gx < p.W - 1 ? 1.f : 0.f;
        const bool colin = gx < p.W;
        if constexpr (COMP) {
          // aux row of a low-res pixel: k0..k3 = noise at its 4 hi-res phases, k4..k12 = inside-image indicators
          const uint2* tab = static_cast<const uint2*>(p.noise_tab);
          for (int r = tt >> 7; r < p.TH; r += LT >> 7) {
            const int gy = bd.y0 + r;
            const bool ok = colin && gy < p.H;
            const float ru = gy > 0 ? 1.f : 0.f, rd = gy < p.H - 1 ? 1.f : 0.f;
            const __half2 q2 = __floats2half2_rn(ru * cl, ru), q3 = __floats2half2_rn(ru * cr, cl);      // k4 k5 | k6 k7
            const __half2 q4 = __floats2half2_rn(1.f, cr), q5 = __floats2half2_rn(rd * cl, rd);          // k8 k9 | k10 k11
            const __half2 q6 = __floats2half2_rn(rd * cr, 0.f);                                          // k12
            const uint32_t lin = xb_addr + (static_cast<uint32_t>(r * 128 + cx) << 5);
            const uint32_t sw = ((lin >> 7) & 1u) << 4;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(lin ^ sw),
                         "l"(ok ? tab + static_cast<size_t>(gy) * p.W + gx : tab), "r"(ok ? 8u : 0u) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"((lin ^ sw) + 8),
                         "r"(*reinterpret_cast<const uint32_t*>(&q2)), "r"(*reinterpret_cast<const uint32_t*>(&q3))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((lin + 16) ^ sw),
                         "r"(*reinterpret_cast<const uint32_t*>(&q4)), "r"(*reinterpret_cast<const uint32_t*>(&q5)),
                         "r"(*reinterpret_cast<const uint32_t*>(&q6)), "r"(0u)
                         : "memory");
          }
        } else {
          const uint32_t* tab = static_cast<const uint32_t*>(p.noise_tab);
          const __half2 h45 = __floats2half2_rn(cl, 1.f);                  // k4 (0,-1), k5 (0,0)
          for (int r = tt >> 7; r < p.TH; r += LT >> 7) {
            const int gy = bd.y0 + r;
            const bool ok = colin && gy < p.H;
            const float ru = gy > 0 ? 1.f : 0.f, rd = gy < p.H - 1 ? 1.f : 0.f;
            const __half2 h23 = __floats2half2_rn(ru, ru * cr);           // k2 (-1,0), k3 (-1,+1)
            const __half2 h67 = __floats2half2_rn(cr, rd * cl);           // k6 (0,+1), k7 (+1,-1)
            const __half2 h89 = __floats2half2_rn(rd, rd * cr);           // k8 (+1,0), k9 (+1,+1)
            const uint32_t lin = xb_addr + (static_cast<uint32_t>(r * 128 + cx) << 5);
            const uint32_t sw = ((lin >> 7) & 1u) << 4;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(lin ^ sw),      // k0 noise, k1 (-1,-1)
                         "l"(ok ? tab + static_cast<size_t>(gy) * p.W + gx : tab), "r"(ok ? 4u : 0u) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"((lin ^ sw) + 4), "r"(*reinterpret_cast<const uint32_t*>(&h23))
                         : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"((lin ^ sw) + 8),
                         "r"(*reinterpret_cast<const uint32_t*>(&h45)), "r"(*reinterpret_cast<const uint32_t*>(&h67))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"((lin + 16) ^ sw),
                         "r"(*reinterpret_cast<const uint32_t*>(&h89)), "r"(0u)
                         : "memory");
          }
        }
      }
      cp_async_commit();
    };

    int cur_n = -1;
    float ca[8], cb[8];                       // this thread's A/B for the current image
#pragma unroll
    for (int i = 0; i < 8; ++i) { ca[i] = 1.f; cb[i] = 0.f; }
    // prefetch distance NS-1: band b+NS-1 is requested right after band b has been handed to the MMA warps
    for (int i = 0; i < NS - 1; ++i)
      if (band0 + i < band1) issue(decode_band(p, band0 + i), i);
    int hs = 0;                               // stage of band b
    for (int b = band0; b < band1; ++b) {
      const Band bd = decode_band(p, b);
      const int ahead = min(NS - 1, band1 - b) - 1;      // younger groups still allowed in flight
      if (ahead >= 1) cp_async_wait<1>(); else cp_async_wait<0>();
      if (!FOLD && p.tmaBand) mbar_wait(&hland[hs], (static_cast<uint32_t>(b - band0) / NS) & 1);   // band landed
      if (!FOLD && affine) {
        if (bd.n != cur_n) {
          cur_n = bd.n;
          const float4* ap = reinterpret_cast<const float4*>(p.inA + bd.n * p.Cin + lc * 8);
          const float4* bp = reinterpret_cast<const float4*>(p.inB + bd.n * p.Cin + lc * 8);
          const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
          ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
          cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
        }
        uint8_t* hb = smem + hs * p.haloBytes;
        const uint32_t hb_addr = smem_u32(hb);
        auto xform = [&](uint32_t lin) {                      // y*A + B on one 16-byte chunk, in place
          const uint32_t off = (lin ^ (((lin >> 7) & (nch - 1)) << 4)) - hb_addr;
          uint4 v = *reinterpret_cast<uint4*>(hb + off);
          __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h2[i]);
            f.x = fmaf(f.x, ca[2 * i], cb[2 * i]);
            f.y = fmaf(f.y, ca[2 * i + 1], cb[2 * i + 1]);
            h2[i] = __floats2half2_rn(f.x, f.y);
          }
          *reinterpret_cast<uint4*>(hb + off) = v;
        };
        // exactly the chunks this thread copied in issue() (so its own cp.async wait suffices: no barrier needed)
        const int R = p.TH + 2;
        const int r_lo = bd.y0 == 0 ? 1 : 0, r_hi = min(R, p.H - (bd.y0 - 1));
        const uint32_t rowB = static_cast<uint32_t>(RC) << 4;
        const int nfull = RC / LT;
        for (int c = 0; c < nfull; ++c) {
          const int q = tt + c * LT;
          const int gx = bd.x0 - 1 + (q >> nchLog);
          if (static_cast<unsigned>(gx) >= static_cast<unsigned>(p.W)) continue;     // zero padding stays zero
          uint32_t lin = hb_addr + (static_cast<uint32_t>(q) << 4) + r_lo * rowB;
          for (int row = r_lo; row < r_hi; ++row, lin += rowB) xform(lin);
        }
        const int left = RC - nfull * LT, leftLog = nchLog + 1;
        for (int idx = tt; idx < R * left; idx += LT) {
          const int row = idx >> leftLog, q = nfull * LT + (idx & (left - 1));
          const int gx = bd.x0 - 1 + (q >> nchLog);
          if (static_cast<unsigned>(gx) < static_cast<unsigned>(p.W) && row >= r_lo && row < r_hi)
            xform(hb_addr + static_cast<uint32_t>(row) * rowB + (static_cast<uint32_t>(q) << 4));
        }
      }
      fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
      named_bar_sync(2, LT);
      if (tt == 0) mbar_arrive(&hready[hs]);
      if (b + NS - 1 < band1) {
        // band j = (b - band0) + NS - 1 goes into stage j % NS, whose previous tenant was band j - NS: wait for
        // that band's MMAs (the (j/NS - 1)-th completion of hempty[stage]) before overwriting it.
        const int j = (b - band0) + NS - 1;
        const int js = j % NS, jt = j / NS;
        if (jt >= 1) mbar_wait(&hempty[js], (jt & 1) ^ 1);
        issue(decode_band(p, b + NS - 1), js);
      }
      if (++hs == NS) hs = 0;
    }
  } else {
    // ================================================================ epilogue (warps 0..7)
    setmaxnreg_inc<kRegsEpi>();
    constexpr bool HOIST = COUT == 16 && !FOLD;     // per-channel noise gain / bias live in registers
    const int q = warp & 3;                // TMEM lane quarter
    const int grp = warp >> 2;             // handles tiles with tcount % kEpiGroups == grp
    const int et = threadIdx.x;            // 0..255
    const bool do_stats = p.stat_sum != nullptr;
    const bool has_noise = !FOLD && p.noise != nullptr, has_bias = !FOLD && p.bias != nullptr;
    // LeakyReLU as max(v, v * slope) (slope < 1); no activation: slope 1.  All epilogue math is plain scalar fp32: the
    // packed .f32x2 forms saved a third of the arithmetic but cost ~50 register moves per 16-channel chunk (the pairs
    // never coalesced with the tcgen05.ld destinations / the FMNMX results), 160 SASS per chunk instead of ~85
    float slope = p.act == CFR_ACT_LRELU ? p.slope : 1.0f;
    asm volatile("" : "+f"(slope));                // (pinned: not re-derived from the parameters per tile)
    const int ci0 = SPLIT ? grp * NCH_T : 0;       // first chunk this thread handles
    float racc[NCH_T * 16], racc2[NCH_T * 16];      // per-thread channel sums
#pragma unroll
    for (int i = 0; i < NCH_T * 16; ++i) { racc[i] = 0.f; racc2[i] = 0.f; }
    // element offset of phase ph inside the output row pair (compile-time phase index after unrolling)
    int ph_off[4];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
      ph_off[ph] = ph < p.numPhases ? (static_cast<int>(p.ooff_y[ph]) * p.outW + p.ooff_x[ph]) * COUT : 0;
    int row_el = p.oscale * p.outW * COUT;         // elements between consecutive band rows in the output
    int kd_el = p.keepDim * COUT;                  // ... between consecutive rows of the compact (sparse-store) output
    asm volatile("" : "+r"(row_el), "+r"(kd_el));
    float hnw[HOIST ? COUT : 1], hbs[HOIST ? COUT : 1];
    if constexpr (HOIST) {
#pragma unroll
      for (int i = 0; i < COUT; ++i) {
        hnw[i] = has_noise ? p.noise_w[i] : 0.f;
        hbs[i] = has_bias ? p.bias[i] : 0.f;
      }
    }
    auto flush_channel = [&](int img) {      // thread et < COUT: all warps' partials (fixed order) -> Q43.20 -> global
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kEpiWarps; ++w) {
        a += s_sum[w * COUT + et];
        b += s_sq[w * COUT + et];
        s_sum[w * COUT + et] = 0.f;
        s_sq[w * COUT + et] = 0.f;
      }
      atomicAdd(&p.stat_sum[img * COUT + et], stat_fx(a));
      atomicAdd(&p.stat_sq[img * COUT + et], stat_fx(b));
    };
    // per-thread channel sums -> this warp's shared-memory slice: once per IMAGE (not per band: the two 16-value
    // transpose-reduces were a third of the epilogue's instructions on the 16-channel layers); fp32 partials of at
    // most a few thousand O(1) values per thread, fixed order => still bit-reproducible
    auto regs_to_smem = [&]() {
#pragma unroll
      for (int cl = 0; cl < NCH_T; ++cl) {
        float a[16], a2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          a[i] = racc[cl * 16 + i];
          a2[i] = racc2[cl * 16 + i];
          racc[cl * 16 + i] = 0.f;
          racc2[cl * 16 + i] = 0.f;
        }
        const float ssum = warp_reduce16h(a, lane);
        const float ssq = warp_reduce16h(a2, lane);
        if ((lane & 1) == 0) {
          const int ch = (ci0 + cl) * 16 + reduce16_channel_h(lane);
          s_sum[warp * COUT + ch] += ssum;
          s_sq[warp * COUT + ch] += ssq;
        }
      }
    };
    uint32_t gbase = 0;                     // accumulator-group counter at the start of the band
    int cur_n = -1;
    for (int b = band0; b < band1; ++b) {
      Band bd = decode_band(p, b);
      asm volatile("" : "+r"(bd.n), "+r"(bd.y0), "+r"(bd.x0), "+r"(bd.rows));   // (pinned: not re-decoded per tile)
      if (do_stats && cur_n >= 0 && bd.n != cur_n) {
        regs_to_smem();
        named_bar_sync(1, kEpiWarps * 32);
        if (et < COUT) {
          flush_channel(cur_n);
        }
        named_bar_sync(1, kEpiWarps * 32);
      }
      cur_n = bd.n;
      const int gx = bd.x0 + q * 32 + static_cast<int>(lane);
      const bool colok = gx < p.W;
      const int ngroups = p.numPhases == 4 ? bd.rows : (bd.rows + G - 1) / G;
      // output pixel of (row r, phase ph): band origin + r * row pitch + per-phase offset (all precomputed)
      __half* const out_band =
          p.out + ((static_cast<size_t>(bd.n) * p.outH + bd.y0 * p.oscale) * p.outW + gx * p.oscale) * COUT;
      constexpr bool SPARSE_OK = COUT == 16 && FOLD && !COMP;      // (only instantiated where it is used: registers)
      const bool sparse = SPARSE_OK && p.keep_map != nullptr;
      // sparse store: compact column of this lane's pixel; the compact rows of the band's (<= 16) output rows are fetched
      // once per band, one per lane, and broadcast by shuffle (a dependent global load per row stalls the warp)
      const int kcol = (sparse && colok) ? __ldg(p.keep_map + gx) : -1;
      const int krow_l = (sparse && static_cast<int>(lane) < bd.rows) ? __ldg(p.keep_map + bd.y0 + lane) : -1;
      __half* const sp_base = p.out + (static_cast<size_t>(bd.n) * p.keepDim * p.keepDim + (kcol < 0 ? 0 : kcol)) * COUT;
      // (pinned: under register pressure ptxas otherwise re-derives these 64-bit bases from the kernel parameters for
      //  every tile, ~40 instructions per 16-channel chunk)
      unsigned long long out_band_u = reinterpret_cast<unsigned long long>(out_band);
      unsigned long long sp_base_u = reinterpret_cast<unsigned long long>(sp_base);
      asm volatile("" : "+l"(out_band_u), "+l"(sp_base_u));
      // composite: this lane's border-column correction rows of the band (null unless it sits on hi-res column 0 /
      // 2W-1: the left border corrects the phases with b == 0, the right border those with b == 1)
      [[maybe_unused]] unsigned long long corr_u = 0;
      [[maybe_unused]] int corr_par = -1;
      if constexpr (COMP) {
        const bool cleft = gx == 0, cright = gx == p.W - 1;
        if (cleft || cright) {
          corr_par = cright ? 1 : 0;
          corr_u = reinterpret_cast<unsigned long long>(
              p.corr + ((static_cast<size_t>(bd.n) * 2 + corr_par) * p.outH + bd.y0 * p.oscale) * COUT);
        }
        asm volatile("" : "+l"(corr_u), "+r"(corr_par));
      }
      // this warp group's accumulator groups of the band: index gbase + j with (gbase + j) % kEpiGroups == grp
      for (int j = SPLIT ? 0 : ((grp - static_cast<int>(gbase)) & (kEpiGroups - 1)); j < ngroups;
           j += SPLIT ? 1 : kEpiGroups) {
        const uint32_t gc = gbase + j;
        const uint32_t slot = gc & (AS - 1);
        mbar_wait(&tfull[slot], (gc >> asLog) & 1);
        tc_fence_after();
        // TMEM -> registers: the loads of KB tiles (of CBN chunks each) are issued back to back and waited for once (with
        // a wait per 16-column load the two epilogue warps of an SM sub-partition spend most of their time on that latency)
#pragma unroll
        for (int kb = 0; kb < G; kb += KB) {
#pragma unroll
        for (int cb = 0; cb < NCH_T; cb += CBN) {
        uint32_t vraw[KB][CBN][16];
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int kk = 0; kk < KB; ++kk)
#pragma unroll
            for (int cc = 0; cc < CBN; ++cc)
              tmem_ld16_issue(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (slot * G + kb + kk) * ACC_COLS +
                              (ci0 + cb + cc) * 16, vraw[kk][cc]);
          tmem_ld_wait();
#pragma unroll
          for (int kk = 0; kk < KB; ++kk)
#pragma unroll
            for (int cc = 0; cc < CBN; ++cc) tmem_ld16_fence(vraw[kk][cc]);
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int k = kb + kk;
          const int r = p.numPhases == 4 ? j : j * G + k;
          const int ph = p.numPhases == 4 ? k : 0;
          if (r >= bd.rows) break;
          __half* optr;
          bool st_ok = colok;
          if (sparse) {                      // warp-uniform row lookup: most rows skip the pack + store altogether
            const int krow = __shfl_sync(0xffffffffu, krow_l, r);
            st_ok = krow >= 0 && kcol >= 0;
            optr = reinterpret_cast<__half*>(sp_base_u) + (krow < 0 ? 0 : krow) * kd_el;
          } else {
            optr = reinterpret_cast<__half*>(out_band_u) + r * row_el + (p.numPhases == 4 ? ph_off[k] : ph_off[0]);
          }
          float nz = 0.f;
          if (has_noise && colok) {
            const int oy = (bd.y0 + r) * p.oscale + p.ooff_y[ph];
            const int ox = gx * p.oscale + p.ooff_x[ph];
            nz = __ldg(&p.noise[oy * p.outW + ox]);
          }
          if (!(p.dbg & 1))
#pragma unroll
          for (int cc = 0; cc < CBN; ++cc) {
            const int cl = cb + cc;
            const int ci = ci0 + cl;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vraw[kk][cc][i]);
            const int ch0 = ci * 16;
            if constexpr (HOIST) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaf(nz, hnw[i], v[i]) + hbs[i];
            } else if constexpr (!FOLD) {
              if (has_bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i));
                  v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
              }
              if (has_noise) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.noise_w + ch0 + i));
                  v[i] = fmaf(nz, w4.x, v[i]); v[i + 1] = fmaf(nz, w4.y, v[i + 1]);
                  v[i + 2] = fmaf(nz, w4.z, v[i + 2]); v[i + 3] = fmaf(nz, w4.w, v[i + 3]);
                }
              }
            }
            if constexpr (COMP) {             // composite: exact first / last hi-res column
              if (corr_par == (k & 1)) {      // (composite phases are (a, b) row-major: k & 1 == b, k >> 1 == a)
                const float4* cp = reinterpret_cast<const float4*>(
                    reinterpret_cast<const float*>(corr_u) + (r * 2 + (k >> 1)) * COUT + ch0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 c4 = __ldg(cp + i);
                  v[4 * i] += c4.x; v[4 * i + 1] += c4.y; v[4 * i + 2] += c4.z; v[4 * i + 3] += c4.w;
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], v[i] * slope);
            if (st_ok) {
              uint4 o[2];
              __half2* h2 = reinterpret_cast<__half2*>(o);
#pragma unroll
              for (int i = 0; i < 8; ++i) h2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
              st_global_256(optr + ch0, o[0], o[1]);     // 16 channels = one 32-byte sector (halo_build checks alignment)
            }
            if (do_stats) {
              if constexpr (REG_STATS) {
                if (colok) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    racc[cl * 16 + i] += v[i];
                    racc2[cl * 16 + i] = fmaf(v[i], v[i], racc2[cl * 16 + i]);
                  }
                }
              } else {
                float sq[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  v[i] = colok ? v[i] : 0.f;
                  sq[i] = v[i] * v[i];
                }
                const float ssum = warp_reduce16h(v, lane);
                const float ssq = warp_reduce16h(sq, lane);
                if ((lane & 1) == 0) {
                  const int ch = ch0 + reduce16_channel_h(lane);
                  s_sum[warp * COUT + ch] += ssum;
                  s_sq[warp * COUT + ch] += ssq;
                }
              }
            }
          }
        }
        }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[slot]);
      }
      gbase += ngroups;
    }
    if (do_stats && cur_n >= 0) {
      regs_to_smem();
      named_bar_sync(1, kEpiWarps * 32);
      if (et < COUT) {
        flush_channel(cur_n);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kHaloMmaWarp0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmemCols);
  }
}

// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void* get_encode_tiled();

static CUtensorMapSwizzle swz(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// ---------------------------------------------------------------------------------------------------------
// Per-sample weight folding (FOLD variant).  For image n, phase ph, tap t (input offset (dy,dx)):
//   w_main[n][ph][t][co][ci]  = fp16( W[ph][t][co][ci] * A[n][ci] )                  -- IN+AdaIN scale of the input
//   w_aux [n][ph][co][0]      = noise_w[co]                                            x aux k0 (noise at the pixel)
//   w_aux [n][ph][co][1+3(dy+1)+(dx+1)] = sum_ci W[ph][t][co][ci]*B[n][ci] (+ bias[co] at (0,0))
//                                                                                      x aux "input (y+dy,x+dx) inside?"
// so that  conv_W(A*y + B, zero padded) + noise*w + bias  ==  conv_wmain(y) + <aux row, w_aux>   exactly at borders.
// ---------------------------------------------------------------------------------------------------------
struct FoldTaps {
  int8_t k[8][9];      // aux K index of the "input (y+dy, x+dx) inside?" indicator of (weight set, tap), -1 = unused
  int8_t noise_k[8];   // aux K index of the noise value used by the weight set
  int8_t center_k;     // aux K index of the (0,0) indicator (carries the bias)
  int8_t tile[8][9];   // non-composite layouts: destination weight tile of (weight set, tap)
};
__global__ void k_fold_weights(const float* __restrict__ base_w, const float* __restrict__ inA,
                               const float* __restrict__ inB, const float* __restrict__ bias,
                               const float* __restrict__ noise_w, FoldTaps ft, int wsets, int ntaps, int ntiles, int cout,
                               int cin, int composite, __half* __restrict__ w_main, __half* __restrict__ w_aux) {
  const int n = blockIdx.y, ws = blockIdx.x;
  // composite: tap-major output with the weight sets ordered [4 5 | 0 1 2 3 | 6 7] inside a tap, so that the four
  // interior phases are ONE contiguous N = 4*cout operand (and the first / last-row variants contiguous pairs)
  const int pos = composite ? (ws < 4 ? ws + 2 : (ws < 6 ? ws - 4 : ws)) : ws;
  // blockDim.x = min(256, cout*cin) threads = co_step whole output channels; cin in {16, 32, 64}
  __shared__ float part[8][9];                                       // cin == 64: per-warp halves of a channel row, per tap
  const int ci = threadIdx.x % cin, co_step = blockDim.x / cin;
  const float a = inA != nullptr ? inA[n * cin + ci] : 1.f;
  const float b = inB != nullptr ? inB[n * cin + ci] : 0.f;
  const int co = blockIdx.z * co_step + threadIdx.x / cin;             // gridDim.z * co_step == cout
  __half* aux_row = w_aux + ((static_cast<size_t>(n) * wsets + pos) * cout + co) * 16;
  if (ci < 16) aux_row[ci] = __float2half_rn((ci == ft.noise_k[ws] && noise_w != nullptr) ? noise_w[co] : 0.f);
  float w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t)                                          // all taps in flight before the first use
    w[t] = t < ntaps ? base_w[((static_cast<size_t>(ws) * ntaps + t) * cout + co) * cin + ci] : 0.f;
  float sh[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (t < ntaps) {
      const size_t oidx = composite ? ((static_cast<size_t>(t) * wsets + pos) * cout + co) * cin + ci
                                    : (static_cast<size_t>(ft.tile[ws][t]) * cout + co) * cin + ci;
      w_main[static_cast<size_t>(n) * ntiles * cout * cin + oidx] = __float2half_rn(w[t] * a);
    }
    sh[t] = w[t] * b;
    for (int o = (cin < 32 ? cin : 32) >> 1; o > 0; o >>= 1) sh[t] += __shfl_xor_sync(0xffffffffu, sh[t], o);
  }
  if (cin == 64) {                                                     // channel row = two warps: combine through smem
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int t = 0; t < 9; ++t) part[threadIdx.x >> 5][t] = sh[t];
    }
    __syncthreads();
    if (ci == 0) {
#pragma unroll
      for (int t = 0; t < 9; ++t) sh[t] = part[threadIdx.x >> 5][t] + part[(threadIdx.x >> 5) + 1][t];
    }
  }
  __syncwarp();                                                        // the zero / noise fill of aux_row above
  if (ci == 0) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int k = t < ntaps ? ft.k[ws][t] : -1;
      if (k >= 0) aux_row[k] = __float2half_rn(sh[t] + ((k == ft.center_k && bias != nullptr) ? bias[co] : 0.f));
    }
  }
}

int launch_fold_weights(const float* base_w, const float* inA, const float* inB, const float* bias, const float* noise_w,
                        const int8_t* tap_dy, const int8_t* tap_dx, int n, int phases, int ntaps, int cout, int cin,
                        int layout, __half* w_main, __half* w_aux, cudaStream_t st) {
  const int composite = layout == 1 ? 1 : 0;
  const int threads = cout * cin < 256 ? cout * cin : 256;       // small blocks: the kernel is a latency chain, not bandwidth
  if ((cin != 16 && cin != 32 && cin != 64) || cout % (threads / cin) != 0) { set_error("fold_weights: Cout=%d Cin=%d unsupported", cout, cin); return 2; }
  FoldTaps ft;
  const int wsets = composite ? 8 : phases;
  const int m_off = composite ? 4 : 1;            // indicators follow the noise slots in the aux row
  for (int ws = 0; ws < 8; ++ws) {
    const int ph = ws < 4 ? ws : ws - 4;          // weight sets 4..7 are the first/last-row variants of phases 0..3
    ft.noise_k[ws] = static_cast<int8_t>(composite ? ph : 0);
    for (int t = 0; t < 9; ++t)
      ft.k[ws][t] = (ws < wsets && t < ntaps)
                        ? static_cast<int8_t>(m_off + 3 * (tap_dy[(ph % phases) * 9 + t] + 1) + (tap_dx[(ph % phases) * 9 + t] + 1)) : -1;
  }
  ft.center_k = static_cast<int8_t>(m_off + 4);
  // layout 2 (row-stationary MMAs, see conv_halo_kernel): tiles ordered dx-major, dy descending, so the three dy
  // variants of one dx are ONE contiguous N = 3*cout operand [dy=+1 | dy=0 | dy=-1]
  // layout 3 (shared-A up-conv, see conv_halo_kernel): one B operand per input position (dy,dx), holding the phases
  // that use it side by side in TMEM order
  static const int8_t up_base[3][3] = {{14, 4, 15}, {8, 0, 11}, {16, 6, 17}};   // first tile of position [dy+1][dx+1]
  static const int8_t up_doff[3][3] = {{0, 0, 1}, {0, 0, 1}, {2, 2, 3}};        // first phase (accumulator) it writes
  for (int ws = 0; ws < 8; ++ws)
    for (int t = 0; t < 9; ++t) {
      int tile = ws * ntaps + t;
      if (ws < wsets && t < ntaps) {
        const int dy = tap_dy[(ws % phases) * 9 + t], dx = tap_dx[(ws % phases) * 9 + t];
        if (layout == 2) tile = ws * ntaps + (dx + 1) * 3 + (1 - dy);
        if (layout == 3) tile = up_base[dy + 1][dx + 1] + ws - up_doff[dy + 1][dx + 1];
      }
      ft.tile[ws][t] = static_cast<int8_t>(tile);
    }
  const int ntiles = layout == 3 ? kUpShareTiles : wsets * ntaps;
  k_fold_weights<<<dim3(wsets, n, cout / (threads / cin)), threads, 0, st>>>(base_w, inA, inB, bias, noise_w, ft, wsets, ntaps, ntiles, cout, cin, composite, w_main, w_aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("fold_weights launch: %s", cudaGetErrorString(e)); return 4; }
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// composite blur o up-conv: exact-value correction for hi-res column 0 (side 0) and 2W-1 (side 1).
//   corr[n][side][Y][co] = sum_{dy,ci} D[side][a][rc][dy][co][ci] * x(i+dy, col),  Y = 2i+a, col = 0 / W-1,
//   x = A[n][ci]*y + B[n][ci] inside the image, 0 outside; rc = first / last hi-res row variant.
// One thread per hi-res row Y (all 16 output channels); a few MFLOP per image.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCorrRows = 64;       // low-res rows per block
__global__ void __launch_bounds__(2 * kCorrRows) k_upblur_corr(const __half* __restrict__ y, const float* __restrict__ inA,
                                                     const float* __restrict__ inB, const float* __restrict__ corr_d,
                                                     int h, int w, int cin, int cout, float* __restrict__ corr) {
  extern __shared__ float csm[];
  // row pitch cin + 1: the 16 rows a warp reads at once (lanes = consecutive Y) must not share a bank
  const int xp = cin + 1;
  float* xcol = csm;                                   // [kCorrRows + 2][cin + 1] transformed border column (0 outside)
  float* dsm = csm + (((kCorrRows + 2) * xp + 3) & ~3);  // [a][dy][ci][co] interior-row coefficients (co fastest)
  const int n = blockIdx.z, side = blockIdx.y, i0 = blockIdx.x * kCorrRows;
  const int col = side == 0 ? 0 : w - 1;
  // border column: 8 channels (one 16-byte load) per thread-iteration
  for (int t = threadIdx.x; t < (kCorrRows + 2) * (cin >> 3); t += blockDim.x) {
    const int r = i0 - 1 + t / (cin >> 3), c8 = (t % (cin >> 3)) * 8;
    float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (r >= 0 && r < h) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(y + ((static_cast<size_t>(n) * h + r) * w + col) * cin + c8));
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h2[k]);
        xv[2 * k] = f.x;
        xv[2 * k + 1] = f.y;
      }
      if (inA != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          xv[k] = __half2float(__float2half_rn(fmaf(xv[k], inA[n * cin + c8 + k], inB[n * cin + c8 + k])));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) xcol[(t / (cin >> 3)) * xp + c8 + k] = xv[k];
  }
  // interior-row coefficients: read [a][dy][co][ci] contiguously (ci fastest), store transposed (co fastest)
  for (int t = threadIdx.x; t < 2 * 3 * cin * cout; t += blockDim.x) {
    const int ci = t % cin, co = (t / cin) % cout, dy = (t / (cout * cin)) % 3, a = t / (cout * cin * 3);
    dsm[((a * 3 + dy) * cin + ci) * cout + co] =
        corr_d[((((static_cast<size_t>(side) * 2 + a) * 3 + 0) * 3 + dy) * cout + co) * cin + ci];
  }
  // first / last hi-res row of the image (Y = 0: a = 0, rc = 1; Y = 2h-1: a = 1, rc = 2) use their own coefficient set:
  // the one block that owns such a row stages it too (a single thread streaming 1.5 K dependent global loads was the
  // long pole of this kernel: ~170 us per launch)
  float* dsm2 = dsm + 2 * 3 * cin * cout;              // [which: 0 = first row, 1 = last row][dy][ci][co]
  const bool has_first = i0 == 0, has_last = 2 * (i0 + kCorrRows) >= 2 * h;
  for (int t = threadIdx.x; t < 2 * 3 * cin * cout; t += blockDim.x) {
    const int ci = t % cin, co = (t / cin) % cout, dy = (t / (cout * cin)) % 3, which = t / (cout * cin * 3);
    if (which == 0 ? has_first : has_last) {
      const int a = which, rc = which + 1;
      dsm2[((which * 3 + dy) * cin + ci) * cout + co] =
          corr_d[((((static_cast<size_t>(side) * 2 + a) * 3 + rc) * 3 + dy) * cout + co) * cin + ci];
    }
  }
  __syncthreads();
  // one thread per hi-res row Y, all 16 output channels in registers: one broadcast load of x and four 128-bit loads
  // of D per 16 FMAs (one thread per (Y, co) was shared-memory-load bound: two loads per FMA)
  for (int Yl = threadIdx.x; Yl < 2 * kCorrRows; Yl += blockDim.x) {
    const int Y = 2 * i0 + Yl, il = Yl >> 1, a = Yl & 1;
    if (Y >= 2 * h) break;
    const int rc = (Y == 0) ? 1 : (Y == 2 * h - 1 ? 2 : 0);
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    const float* dbase = rc == 0 ? dsm + (a * 3) * cin * 16 : dsm2 + ((rc - 1) * 3) * cin * 16;
    for (int dy = 0; dy < 3; ++dy) {
      const float* xr = xcol + (il + dy) * xp;
      const float4* dd = reinterpret_cast<const float4*>(dbase + dy * cin * 16);
      for (int ci = 0; ci < cin; ++ci) {
        const float xv = xr[ci];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 d4 = dd[ci * 4 + q];
          acc[4 * q] = fmaf(d4.x, xv, acc[4 * q]);
          acc[4 * q + 1] = fmaf(d4.y, xv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(d4.z, xv, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(d4.w, xv, acc[4 * q + 3]);
        }
      }
    }
    float4* op = reinterpret_cast<float4*>(corr + ((static_cast<size_t>(n) * 2 + side) * (2 * h) + Y) * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) op[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

int launch_upblur_corr(const __half* y, const float* inA, const float* inB, const float* corr_d, int n, int h, int w,
                       int cin, int cout, float* corr, cudaStream_t st) {
  dim3 grid((h + kCorrRows - 1) / kCorrRows, 2, n);
  const size_t smem = ((static_cast<size_t>(kCorrRows + 2) * (cin + 1) + 3) / 4 * 4 + 2 * 2 * 3 * cin * cout) * sizeof(float);
  if (smem > 48 * 1024 || cout != 16) { set_error("upblur_corr: needs Cout == 16 and a small Cin"); return 2; }
  k_upblur_corr<<<grid, 2 * kCorrRows, smem, st>>>(y, inA, inB, corr_d, h, w, cin, cout, corr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("upblur_corr launch: %s", cudaGetErrorString(e)); return 4; }
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// packed aux-row heads for the FOLD variants (see conv_halo.cuh)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_pack_noise(const float* __restrict__ noise, int h, int w, int mode, void* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  const int y = i / w, x = i - y * w;
  if (mode == 0) {
    const __half2 v = __floats2half2_rn(noise != nullptr ? noise[i] : 0.f, (y > 0 && x > 0) ? 1.f : 0.f);
    static_cast<uint32_t*>(out)[i] = *reinterpret_cast<const uint32_t*>(&v);
  } else {
    float n00 = 0.f, n01 = 0.f, n10 = 0.f, n11 = 0.f;
    if (noise != nullptr) {
      const float* np = noise + static_cast<size_t>(2 * y) * (2 * w) + 2 * x;
      n00 = np[0]; n01 = np[1]; n10 = np[2 * w]; n11 = np[2 * w + 1];
    }
    const __half2 a = __floats2half2_rn(n00, n01), b = __floats2half2_rn(n10, n11);
    static_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}
int launch_pack_noise(const float* noise, int h, int w, int mode, void* out, cudaStream_t st) {
  k_pack_noise<<<(h * w + 255) / 256, 256, 0, st>>>(noise, h, w, mode, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("pack_noise launch: %s", cudaGetErrorString(e)); return 4; }
  count_launch();
  return 0;
}

bool halo_upshare_ok(const cfr_conv_desc& s) {
  if (const char* e = getenv("CFR_HALO_UPSHARE")) if (atoi(e) == 0) return false;
  if (s.numPhases != 4 || s.ntaps != 4 || s.Cout > 32 || s.oscale != 2) return false;
  for (int ph = 0; ph < 4; ++ph) {                     // phase (py,px) = nearest x2 + 3x3: taps {py-1,py} x {px-1,px}
    const int py = ph >> 1, px = ph & 1;
    if (s.ooff_y[ph] != py || s.ooff_x[ph] != px) return false;
    unsigned seen = 0;
    for (int t = 0; t < 4; ++t) {
      const int dy = s.tap_dy[ph][t] - (py - 1), dx = s.tap_dx[ph][t] - (px - 1);
      if (dy < 0 || dy > 1 || dx < 0 || dx > 1) return false;
      seen |= 1u << (dy * 2 + dx);
    }
    if (seen != 0xfu) return false;
  }
  return true;
}

int halo_build(const cfr_conv_desc& s, const float* inA, const float* inB, const void* w_aux, int composite,
               const float* corr, HaloOp* op) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)"); return 1; }
  HaloParams& p = op->p;
  memset(&p, 0, sizeof(p));
  if (!(s.Cin == 16 || s.Cin == 32 || s.Cin == 64) || !(s.Cout == 16 || s.Cout == 32 || s.Cout == 64)) {
    set_error("halo conv: Cin/Cout must be 16, 32 or 64 (got %d -> %d)", s.Cin, s.Cout);
    return 2;
  }
  if (s.stride != 1 || s.Hout != s.Hin || s.Wout != s.Win || s.outIsF32 || s.cbias || s.resid || s.wRowsPerSample) {
    set_error("halo conv: unsupported option (stride/out grid/f32/cbias/resid/per-sample weights)");
    return 2;
  }
  p.N = s.N; p.H = s.Hout; p.W = s.Wout; p.Cin = s.Cin; p.Cout = s.Cout;
  p.numPhases = s.numPhases; p.ntaps = s.ntaps;
  memcpy(p.tap_dy, s.tap_dy, sizeof(p.tap_dy));
  memcpy(p.tap_dx, s.tap_dx, sizeof(p.tap_dx));
  p.fold = w_aux != nullptr;
  p.composite = composite;
  p.corr = corr;
  if (composite && (!p.fold || s.numPhases != 4 || s.ntaps != 9 || s.Cout != 16 || s.Cin > 32)) { set_error("halo conv: composite needs fold, 4 phases, 9 taps"); return 2; }
  p.wsets = composite ? 8 : s.numPhases;
  if (const char* e = getenv("CFR_HALO_DBG")) p.dbg = atoi(e);
  p.upshare = p.fold && !composite && halo_upshare_ok(s);
  p.rowmma = 0;
  if (p.fold && !composite && s.numPhases == 1 && s.ntaps == 9) {
    unsigned seen = 0;                   // a full 3x3 stencil, any tap order
    for (int t = 0; t < 9; ++t)
      if (s.tap_dy[0][t] >= -1 && s.tap_dy[0][t] <= 1 && s.tap_dx[0][t] >= -1 && s.tap_dx[0][t] <= 1)
        seen |= 1u << ((s.tap_dy[0][t] + 1) * 3 + s.tap_dx[0][t] + 1);
    // 1: per accumulator group (4 issuing warps; N = 16 tiles are issue-bound), 2: per band (fewest MMAs; wins once
    // the MMAs are N = 96 wide).  Measured: profiles/ncu_r01_notes.md section 13.
    if (seen == 0x1ffu) p.rowmma = s.Cout >= 32 ? 2 : 1;
    if (const char* e = getenv("CFR_HALO_ROWMMA")) p.rowmma = p.rowmma ? atoi(e) : 0;
    if (s.Cout > 32 && p.rowmma == 1) p.rowmma = 2;     // (the per-group variant is only instantiated for Cout <= 32)
  }
  if (p.fold && s.Cin > 64) { set_error("halo conv: folded variant needs Cin <= 64"); return 2; }
  p.rowBytes = s.Cin * 2;
  p.wRows = p.upshare ? kUpShareTiles * s.Cout : p.wsets * s.ntaps * s.Cout;
  p.wAuxBytes = p.fold ? (p.wsets * s.Cout * 32 + 1023) / 1024 * 1024 : 0;
  p.wBytes = (p.wRows * p.rowBytes + 1023) / 1024 * 1024;
  p.wBoxRows = p.wRows;
  while (p.wBoxRows > 256 || p.wRows % p.wBoxRows != 0) --p.wBoxRows;
  const int ctrl = 8 * 44 + 16 + 2 * 64 * 4 + 2 * kEpiWarps * s.Cout * 4 + 64;
  const int budget = 227 * 1024 - 1024 - ctrl - p.wBytes - p.wAuxBytes;
  auto aux_bytes = [&](int th) { return p.fold ? (th * 128 * 32 + 1023) / 1024 * 1024 : 0; };
  auto halo_bytes = [&](int th) { return ((th + 2) * kHaloW * p.rowBytes + 1023) / 1024 * 1024 + aux_bytes(th); };
  // three band buffers (prefetch distance 2) with the tallest band that fits, but at least 4 rows per band;
  // otherwise fall back to two buffers
  // two band buffers with tall bands beat three with short ones (per-band overhead and halo re-reads dominate):
  // profiles/ops_r01_*.tsv.  CFR_HALO_STAGES=3 switches back for experiments.
  int ns = 2, th = 16;
  if (const char* e = getenv("CFR_HALO_STAGES")) ns = atoi(e) == 3 ? 3 : 2;
  while (th > 1 && ns * halo_bytes(th) > budget) --th;
  if (th < 4) {
    ns = 2;
    th = 16;
    while (th > 1 && ns * halo_bytes(th) > budget) --th;
  }
  if (th > s.Hout) th = s.Hout;
  if (ns * halo_bytes(th) > budget) { set_error("halo conv: smem budget"); return 2; }
  p.TH = th;
  p.haloStages = ns;
  p.auxBytes = aux_bytes(th);
  p.haloBytes = halo_bytes(th) - p.auxBytes;
  p.bandsX = (s.Wout + 127) / 128;
  p.bandsY = (s.Hout + th - 1) / th;
  p.magicX = p.bandsX > 1 ? static_cast<uint32_t>((1ull << 32) / p.bandsX + 1) : 0u;   // decode_band: b / bandsX == umulhi(b, magicX)
  p.magicY = p.bandsY > 1 ? static_cast<uint32_t>((1ull << 32) / p.bandsY + 1) : 0u;
  if (static_cast<long long>(s.N) * p.bandsX * p.bandsY >= (1ll << 21) || p.bandsX > 1024 || p.bandsY > 1024) {
    set_error("halo conv: too many bands for the reciprocal band decode");
    return 2;
  }
  p.accStages = 0;   // (fixed per template: 512 / (G * Cout) groups of G accumulators)
  if (s.numPhases != 1 && s.numPhases != 4) { set_error("halo conv: 1 or 4 phases"); return 2; }
  if (s.ntaps != 4 && s.ntaps != 9) { set_error("halo conv: 4 or 9 taps per phase"); return 2; }
  if (s.numPhases == 4 && s.Cout == 64) { set_error("halo conv: 4-phase up-conv needs Cout <= 32"); return 2; }
  if ((reinterpret_cast<uintptr_t>(s.out) & 31) != 0) { set_error("halo conv: output must be 32-byte aligned"); return 2; }
  p.inA = inA; p.inB = inB;
  p.out = static_cast<__half*>(s.out);
  p.outH = s.outH; p.outW = s.outW; p.outC = s.outC; p.oscale = s.oscale;
  memcpy(p.ooff_y, s.ooff_y, sizeof(p.ooff_y));
  memcpy(p.ooff_x, s.ooff_x, sizeof(p.ooff_x));
  p.bias = s.bias; p.noise = s.noise; p.noise_w = s.noise_w; p.act = s.act; p.slope = s.slope;
  p.stat_sum = reinterpret_cast<unsigned long long*>(s.stat_sum); p.stat_sq = reinterpret_cast<unsigned long long*>(s.stat_sq);
  if (s.outC != s.Cout) { set_error("halo conv: outC must equal Cout"); return 2; }
  if (s.keepMap != nullptr) {
    if (s.numPhases != 1 || s.oscale != 1 || s.outH != s.outW || s.outH != s.Hout || s.keepDim <= 0 || s.keepDim > s.outH ||
        s.Cout != 16 || !p.fold || composite) {
      set_error("halo conv: sparse store needs the folded Cout = 16 variant, one phase, oscale 1, a square output and "
                "0 < keepDim <= outH");
      return 2;
    }
    p.keep_map = s.keepMap;
    p.keepDim = s.keepDim;
  }
  p.in = static_cast<const __half*>(s.in);
  {
    const int totalRows = p.fold ? s.N * p.wRows : p.wRows;
    if (s.Kpad != s.Cin || s.wRows != totalRows) { set_error("halo conv: weights must be [(n x) phases*taps*Cout][Cin] (got [%d][%d], want %d rows)", s.wRows, s.Kpad, totalRows); return 2; }
    cuuint64_t dims[2] = {(cuuint64_t)s.Cin, (cuuint64_t)totalRows};
    cuuint64_t strides[1] = {(cuuint64_t)s.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)s.Cin, (cuuint32_t)p.wBoxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(s.w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.rowBytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("halo conv: encode(W) failed: %d", (int)r); return 3; }
    static const bool want_tma = getenv("CFR_HALO_TMA") == nullptr || atoi(getenv("CFR_HALO_TMA")) != 0;
    p.tmaBand = want_tma;
    if (p.tmaBand) {
      cuuint64_t idims[4] = {(cuuint64_t)s.Cin, (cuuint64_t)s.Win, (cuuint64_t)s.Hin, (cuuint64_t)s.N};
      cuuint64_t istr[3] = {(cuuint64_t)s.Cin * 2, (cuuint64_t)s.Win * s.Cin * 2, (cuuint64_t)s.Hin * s.Win * s.Cin * 2};
      cuuint32_t ibox[4] = {(cuuint32_t)s.Cin, (cuuint32_t)kHaloW, (cuuint32_t)(p.TH + 2), 1};
      cuuint32_t iestr[4] = {1, 1, 1, 1};
      r = enc(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.in), idims, istr, ibox, iestr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz(p.rowBytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("halo conv: encode(input band) failed: %d", (int)r); return 3; }
    }
    if (p.fold) {
      cuuint64_t adims[2] = {16, (cuuint64_t)s.N * p.wsets * s.Cout};
      cuuint64_t astr[1] = {32};
      cuuint32_t abox[2] = {16, (cuuint32_t)(p.wsets * s.Cout)};
      r = enc(&p.tmWa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_aux), adims, astr, abox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("halo conv: encode(Waux) failed: %d", (int)r); return 3; }
    }
  }
  const int total = p.N * p.bandsX * p.bandsY;
  op->grid = total < num_sms() ? total : num_sms();
  if (const char* e = getenv("CFR_MAX_CTAS")) {      // tests: few CTAs => many work items per CTA (ring wrap, phase flips)
    const int m = atoi(e);
    if (m > 0 && op->grid > m) op->grid = m;
  }
  op->smemBytes = p.haloStages * (p.haloBytes + p.auxBytes) + p.wBytes + p.wAuxBytes + ctrl + 1024;
  op->flops = 2.0 * s.N * s.Hout * s.Wout * s.numPhases * s.ntaps * static_cast<double>(s.Cin) * s.Cout;
  // algorithmic HBM bytes: every input element read once, every output element written once (fp16)
  // (sparse store: only the kept keepDim^2 pixels are written)
  const double outPix = s.keepMap != nullptr ? static_cast<double>(s.keepDim) * s.keepDim : static_cast<double>(s.outH) * s.outW;
  op->bytes = 2.0 * s.N * (static_cast<double>(s.Hin) * s.Win * s.Cin + outPix * s.Cout);
  return 0;
}

int halo_launch(const HaloOp& op, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    const void* fns[] = {(const void*)conv_halo_kernel<16, false>, (const void*)conv_halo_kernel<32, false>,
                         (const void*)conv_halo_kernel<64, false>, (const void*)conv_halo_kernel<16, true>,
                         (const void*)conv_halo_kernel<32, true>, (const void*)conv_halo_kernel<16, true, true>,
                         (const void*)conv_halo_kernel<64, true>};
    for (const void* f : fns)
      if (attr_err == cudaSuccess)
        attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err)); return 4; }
  cudaEvent_t e1 = nullptr;
  if (profile_on()) {
    cudaEventRecord(profile_event(1), stream);
    e1 = profile_event(1);
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see pdl_trigger() / pdl_wait() in ptx.cuh
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(op.grid);
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = op.smemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  switch (op.p.Cout) {
    case 16:
      if (op.p.composite) cudaLaunchKernelEx(&cfg, conv_halo_kernel<16, true, true>, op.p);
      else if (op.p.fold) cudaLaunchKernelEx(&cfg, conv_halo_kernel<16, true, false>, op.p);
      else cudaLaunchKernelEx(&cfg, conv_halo_kernel<16, false, false>, op.p);
      break;
    case 32:
      if (op.p.fold) cudaLaunchKernelEx(&cfg, conv_halo_kernel<32, true, false>, op.p);
      else cudaLaunchKernelEx(&cfg, conv_halo_kernel<32, false, false>, op.p);
      break;
    default:
      if (op.p.fold) cudaLaunchKernelEx(&cfg, conv_halo_kernel<64, true, false>, op.p);
      else cudaLaunchKernelEx(&cfg, conv_halo_kernel<64, false, false>, op.p);
      break;
  }
  if (profile_on()) {
    cudaEventRecord(e1, stream);
    profile_account(1, op.bytes);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("halo conv launch: %s", cudaGetErrorString(e)); return 4; }
  count_launch();
  return 0;
}

}  // namespace cfr
