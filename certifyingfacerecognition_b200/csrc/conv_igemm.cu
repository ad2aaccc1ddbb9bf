// tcgen05/TMEM/TMA implicit-GEMM convolution kernel -- see conv_igemm.cuh for the design.
#include "conv_igemm.cuh"
#include "ptx.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace cfr {

// --------------------------------------------------------------------------------------------
// device
// --------------------------------------------------------------------------------------------
// Warp roles: epilogue warps first (groups of four, warp id & 3 == TMEM lane quarter), then the TMA producer, and the
// single MMA-issuing thread in the HIGHEST warp (round 1: the SM sub-partition arbiter favours higher warp ids and a starved
// issuer stalls the whole pipeline; re-measured with 16 epilogue warps -- issuer in warp 5 of 18 vs warp 17: no difference,
// CFR_IGEMM_MMA_WARP_LAST=0 builds the old layout).
#ifndef CFR_IGEMM_MMA_WARP_LAST
#define CFR_IGEMM_MMA_WARP_LAST 1
#endif
template <int MT> struct Roles {
  static constexpr int kEpi = 8 * MT;                                         // epilogue warps
  static constexpr int kProducer = CFR_IGEMM_MMA_WARP_LAST ? kEpi : 4;
  static constexpr int kMma = CFR_IGEMM_MMA_WARP_LAST ? kEpi + 1 : 5;
};

struct TileCoord {
  int n0, y0, x0, phase, ntile;
};

// MT == 2: `item` indexes PAIRS of adjacent M tiles; `sub` selects the tile of the pair
__device__ __forceinline__ TileCoord decode_item(const ConvParams& p, int item, int sub) {
  TileCoord t;
  t.ntile = item % p.numNTiles;
  int r = item / p.numNTiles;
  t.phase = r % p.numPhases;
  int m = (r / p.numPhases) * p.tilesPerItem + sub;
  int tx = m % p.tilesX;
  m /= p.tilesX;
  int ty = m % p.tilesY;
  int tn = m / p.tilesY;
  t.x0 = tx * p.TW;
  t.y0 = ty * p.TH;
  t.n0 = tn * p.TN;
  return t;
}

// Sum v[0..15] over the 32 lanes of a warp, 16 channels at once (transpose-reduce, 16 shuffles).
// On return lane l (l even) holds in v[0] the warp total of channel ((l>>1)&15) bit-reversed as below.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], uint32_t lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step;      // 16, 8, 4, 2
    const int cnt = 8 >> step;       // 8, 4, 2, 1
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
// channel (0..15) whose total lane `lane` holds after warp_reduce16
__device__ __forceinline__ int reduce16_channel(uint32_t lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// channel c of image img: the four epilogue warps' partial sums (fixed order) -> Q43.20 fixed point -> global accumulate
// (s_sum / s_sq point at the four slices of ONE epilogue group; slices are CoutTotal floats apart)
__device__ __forceinline__ void flush_stats(const ConvParams& p, float* s_sum, float* s_sq, int img, int c) {
  // (both column halves of a tile write disjoint channels of the same four lane-quarter slices)
  const int C = p.CoutTotal;
  const float a = (s_sum[c] + s_sum[C + c]) + (s_sum[2 * C + c] + s_sum[3 * C + c]);
  const float b = (s_sq[c] + s_sq[C + c]) + (s_sq[2 * C + c] + s_sq[3 * C + c]);
  atomicAdd(&p.stat_sum[img * C + c], stat_fx(a));
  atomicAdd(&p.stat_sq[img * C + c], stat_fx(b));
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    s_sum[w * C + c] = 0.f;
    s_sq[w * C + c] = 0.f;
  }
}

// MT == 2: the CTA works on two adjacent M tiles at once.  Both tiles' A boxes and ONE weight box make a pipeline
// stage, and every K step issues two MMAs (one per accumulator) against the same B operand: the bytes the SM pulls from
// L2 per FLOP drop by 1/4 - 1/3.  These layers are bound by the chip-wide L2 -> SM feed (ncu: ~12 TB/s delivered,
// tensor pipe 17-55 % busy), not by the tensor cores.  A second group of four epilogue warps drains the second tile.
//
// CG == 2 (with MT == 1): the two CTAs of a 2-CTA cluster (two SMs of a TPC) work on two adjacent M tiles -- one each --
// against ONE 256-wide N tile.  Each CTA stages its own A tile and HALF of the weight rows; the even CTA issues
// tcgen05.mma.cta_group::2 (M = 256: rows 0..127 from its shared memory / into its TMEM, rows 128..255 the peer's), every
// TMA load signals the even CTA's `full` barrier, tcgen05.commit multicasts `empty` / `tfull` to both CTAs, and both
// CTAs' epilogue warps arrive on the even CTA's `tempty`.  Bytes pulled from L2 per FLOP: 32 KB per 128 x 256 x 64 per SM
// instead of 48 KB (MT == 2, BN == 128) -- these layers sit on the chip-wide L2 -> SM cap -- with both accumulator sets
// (2 x 256 columns) still in flight.
//
// ROWS (with MT == 2, CG == 1; layers whose output grid is exactly 128 pixels wide: StyleGAN L11 / L12): a work item is two
// consecutive image rows (x one or two column phases of an up-conv's row phase).  Per 64-channel K chunk ONE TMA box of
// aRows input rows x 130 pixels lands in a two-deep A ring; every tap of both rows is that box read through a shifted
// UMMA descriptor ((row, pixel) offset in units of the 128-byte swizzled rows -- the swizzle is a function of the
// absolute shared-memory address, so shifted starts stay consistent with what TMA wrote, as in conv_halo.cu).  Weight
// tiles stream through their own ring, one per (phase, tap, K chunk).  A bytes per K chunk: 66 KB instead of 18 x 16 KB
// (3x3) and 50 KB instead of 16 x 16 KB for both column phases of an up-conv row phase: these layers sit on the
// chip-wide L2 -> SM cap (profiles/ncu_r02_notes.md sections 4, 10).  rowsMerge (section 15): operands that two accumulators
// share are issued as ONE wider MMA -- up-conv: the input column both column phases read, B = the two phases' tiles side by
// side (N = 2 * BN); 3x3: the input rows both output rows read, B = the tiles of (dy, dy - 1) in consecutive ring slots
// (N = 256 across both rows' accumulators; two N = 128 MMAs when the pair straddles the ring's wrap-around).
template <int MT, int CG = 1, bool ROWS = false>
__global__ void __launch_bounds__(MT == 2 ? kConvThreadsMT2 : kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(CG == 1 || MT == 1, "a CTA pair works on one M tile per CTA");
  static_assert(!ROWS || (MT == 2 && CG == 1), "ROWS: two image rows per CTA step");
  constexpr int TPI = CG == 2 ? 2 : MT;           // M tiles per work item
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.numStages;                      // (ROWS: bStages + 2 -- the last two barrier pairs belong to the A ring)
  uint8_t* ctrl = smem + p.ctrlOffset;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // per-(epilogue warp, channel) float partial sums: every warp owns its slice, so no shared-memory atomics (64-bit
  // shared atomics compile to ATOMS.CAST.SPIN loops); the tile -> warp order is static, hence still bit-reproducible
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);      // [MT][4][CoutTotal]
  float* s_sq = s_sum + MT * 4 * p.CoutTotal;                  // [MT][4][CoutTotal]

  const int warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();

  // (an odd tile count leaves the last pair with a dummy second tile: its sample index is >= N, so TMA zero-fills its
  //  A box and the epilogue only keeps the barriers in step)
  const int totalItems = ROWS ? p.N * (p.Hout / 2) * p.rowsNPG * p.numNTiles
                              : ((p.tilesX * p.tilesY * p.tilesN + TPI - 1) / TPI) * p.numPhases * p.numNTiles;
  const int accW = ROWS ? 128 : p.BN;             // TMEM columns per M tile (ROWS: rowsPG phases x BN)
  const int nWorkers = CG == 2 ? gridDim.x / 2 : gridDim.x;       // a CTA pair walks one item list together
  const int per = (totalItems + nWorkers - 1) / nWorkers;
  const int item0 = (CG == 2 ? blockIdx.x / 2 : blockIdx.x) * per;
  const int item1 = min(totalItems, item0 + per);

  const int chunksTotal = p.ntaps * p.nCB;
  const int stagesPerTile = (chunksTotal + p.G - 1) / p.G;
  const uint32_t subBytes = kBM * p.CB * 2;
  uint32_t tmemCols = 32;
  while (tmemCols < static_cast<uint32_t>(p.nAcc * MT * accW)) tmemCols <<= 1;

  if (warp == Roles<MT>::kProducer && lane == 0) {
    tma_prefetch_desc(ROWS ? &p.tmA2 : &p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  // ROWS work item -> (image, first output row, phase group, N tile)
  auto rows_item = [&](int item, int& n, int& y0, int& pg, int& ntile) {
    ntile = item % p.numNTiles;
    int r = item / p.numNTiles;
    pg = r % p.rowsNPG;
    r /= p.rowsNPG;
    const int hp = p.Hout / 2;
    y0 = 2 * (r % hp);
    n = r / hp;
  };
  if (warp == Roles<MT>::kMma) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 8 * MT * CG);          // every warp of the 2 * MT epilogue groups (of both CTAs) arrives
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, tmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, tmemCols);
      tmem_relinquish();
    }
  }
  if (p.stat_sum != nullptr) {
    for (int i = threadIdx.x; i < MT * 8 * p.CoutTotal; i += blockDim.x) s_sum[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == Roles<MT>::kProducer) {
    // ===================================================================== TMA producer
    // (whole warp runs the loop so coordinates / descriptors stay in uniform registers; one elected lane issues)
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (ROWS) {
      const int SB = p.bStages;
      uint8_t* bring = smem + 2 * p.aBoxBytes;
      const uint32_t bBytes = static_cast<uint32_t>(p.BN) * 128u;
      int ab = 0;
      uint32_t abphase = 0;
      for (int item = item0; item < item1; ++item) {
        int n, y0, pg, ntile;
        rows_item(item, n, y0, pg, ntile);
        for (int cb = 0; cb < p.nCB; ++cb) {
          mbar_wait(&empty_bar[SB + ab], abphase ^ 1);
          if (leader) {
            const bool skip = (p.dbg & 4) != 0;
            mbar_expect_tx(&full_bar[SB + ab], skip ? 0u : static_cast<uint32_t>(p.aRows) * 130u * 128u);
            if (!skip)
              tma_load_4d(smem + ab * p.aBoxBytes, &p.tmA2, &full_bar[SB + ab], cb * 64, -1, y0 + p.rowsDyMin[pg], n);
          }
          __syncwarp();
          if (++ab == 2) {
            ab = 0;
            abphase ^= 1;
          }
          if (p.rowsMerge == 2) {
            // one tile per stage, column by column of the 3x3, rows bottom-up: the tiles of (dy, dy - 1) are consecutive
            for (int u = 0; u < 9; ++u) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (leader) {
                mbar_expect_tx(&full_bar[stage], bBytes);
                tma_load_2d(bring + stage * bBytes, &p.tmB, &full_bar[stage], (p.rowsTap[u] * p.nCB + cb) * 64, ntile * p.BN);
              }
              __syncwarp();
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
            }
          } else if (p.rowsMerge) {
            // one stage per input row i: [ph0 (i, left) | ph0 (i, centre) | ph1 (i, centre) | ph1 (i, right)] -- the two
            // centre tiles are adjacent, so they form the 2*BN-row B operand of the shared MMA
            for (int i = 0; i < 2; ++i) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (leader) {
                mbar_expect_tx(&full_bar[stage], bBytes * 4);
                for (int u = 0; u < 4; ++u)
                  tma_load_2d(bring + (stage * 4 + u) * bBytes, &p.tmB, &full_bar[stage],
                              ((2 * i + (u & 1)) * p.nCB + cb) * 64, (pg * 2 + (u >> 1)) * p.wRowsPerPhase + ntile * p.BN);
              }
              __syncwarp();
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
            }
          } else
          for (int j = 0; j < p.rowsPG; ++j) {
            const int ph = pg * p.rowsPG + j;
            const int wrow = ph * p.wRowsPerPhase + ntile * p.BN;
            for (int t = 0; t < p.ntaps; t += p.tapsPerStage) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (leader) {
                mbar_expect_tx(&full_bar[stage], bBytes * p.tapsPerStage);
                for (int u = 0; u < p.tapsPerStage; ++u)
                  tma_load_2d(bring + (stage * p.tapsPerStage + u) * bBytes, &p.tmB, &full_bar[stage],
                              ((t + u) * p.nCB + cb) * 64, wrow);
              }
              __syncwarp();
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else
    for (int item = item0; item < item1; ++item) {
      const TileCoord t = decode_item(p, item, CG == 2 ? static_cast<int>(crank) : 0);
      const TileCoord t1 = decode_item(p, item, MT - 1);          // second tile of the pair (== t when MT == 1)
      // (CTA pair: this CTA's half of the N tile's weight rows)
      const int wrow = t.n0 * p.wRowsPerSample + t.phase * p.wRowsPerPhase + t.ntile * p.BN +
                       (CG == 2 ? static_cast<int>(crank) * (p.BN / 2) : 0);
      const int xb = t.x0 * p.stride, yb = t.y0 * p.stride;
      const int xb1 = t1.x0 * p.stride, yb1 = t1.y0 * p.stride;
      int chunk = 0;
      for (int s = 0; s < stagesPerTile; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int nch = min(p.G, chunksTotal - chunk);
        uint8_t* a_dst = smem + static_cast<size_t>(stage) * p.stageBytes;
        if (leader) {
          if constexpr (CG == 2) {
            // both CTAs' boxes complete on the EVEN CTA's barrier, which expects the bytes of both
            const int na = (p.dbg & 4) ? 0 : nch;
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * (na * subBytes + (p.BN / 2) * 128));
            for (int g = 0; g < na; ++g) {
              const int c = chunk + g;
              const int tap = c / p.nCB;
              const int cb = c - tap * p.nCB;
              tma_load_4d_pair(a_dst + g * subBytes, &p.tmA, &full_bar[stage], cb * p.CB, xb + p.tap_dx[t.phase][tap],
                               yb + p.tap_dy[t.phase][tap], t.n0);
            }
            tma_load_2d_pair(a_dst + kBM * 128, &p.tmB, &full_bar[stage], s * 64, wrow);
          } else {
          const int na = (p.dbg & 4) ? 0 : nch;
          mbar_expect_tx(&full_bar[stage], MT * na * subBytes + p.BN * 128);
          for (int g = 0; g < na; ++g) {
            const int c = chunk + g;
            const int tap = c / p.nCB;
            const int cb = c - tap * p.nCB;
            tma_load_4d(a_dst + g * subBytes, &p.tmA, &full_bar[stage], cb * p.CB, xb + p.tap_dx[t.phase][tap],
                        yb + p.tap_dy[t.phase][tap], t.n0);
            if (MT == 2)
              tma_load_4d(a_dst + kBM * 128 + g * subBytes, &p.tmA, &full_bar[stage], cb * p.CB,
                          xb1 + p.tap_dx[t.phase][tap], yb1 + p.tap_dy[t.phase][tap], t1.n0);
          }
          tma_load_2d(a_dst + MT * kBM * 128, &p.tmB, &full_bar[stage], s * 64, wrow);
          }
        }
        __syncwarp();
        chunk += nch;
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == Roles<MT>::kMma) {
    // ===================================================================== MMA issuer (CTA pair: the even CTA's only)
    const bool leader = elect_one() && crank == 0;
    const uint32_t idesc = make_idesc_f16(CG == 2 ? 2 * kBM : kBM, p.BN);
    const uint32_t a_hi = smem_desc_hi(8 * p.swizzleA, p.swizzleA);
    const uint32_t b_hi = smem_desc_hi(1024, 128);
    const int kPer = p.CB / 16;
    const uint32_t sub16 = subBytes >> 4;
    const uint32_t smem_lo = smem_desc_lo(smem_u32(smem));
    const uint32_t stage16 = static_cast<uint32_t>(p.stageBytes) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    if constexpr (ROWS) {
      const int SB = p.bStages;
      const uint32_t bring16 = smem_lo + ((2u * static_cast<uint32_t>(p.aBoxBytes)) >> 4);
      const uint32_t b16 = (static_cast<uint32_t>(p.BN) * 128u) >> 4;
      const uint32_t abox16 = static_cast<uint32_t>(p.aBoxBytes) >> 4;
      int ab = 0;
      uint32_t abphase = 0;
      for (int item = item0; item < item1; ++item) {
        int n, y0, pg, ntile;
        rows_item(item, n, y0, pg, ntile);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * MT * 128;
        for (int cb = 0; cb < p.nCB; ++cb) {
          mbar_wait(&full_bar[SB + ab], abphase);
          tc_fence_after();
          const uint32_t a_base = smem_lo + ab * abox16;
          if (p.rowsMerge == 2) {
            // 3x3, output rows y0 and y0 + 1: input row q (relative to y0) feeds row 0 with W(dy = q) and row 1 with
            // W(dy = q - 1).  For q = 0, 1 both exist and their tiles are consecutive ring slots: ONE N = 256 MMA writes both
            // rows' accumulators (columns [0,128) | [128,256)) unless the pair straddles the ring's wrap-around.
            const uint32_t idesc2 = make_idesc_f16(kBM, 256);
            for (int g = 0; g < 3; ++g) {
              int st[3];
              for (int v = 0; v < 3; ++v) {
                st[v] = stage;
                mbar_wait(&full_bar[stage], phase);
                if (++stage == SB) {
                  stage = 0;
                  phase ^= 1;
                }
              }
              tc_fence_after();
              if (leader && !(p.dbg & 2)) {
                const uint32_t aq = a_base + static_cast<uint32_t>(g) * 8u;       // box row 0 (q = -1), pixel offset dx + 1 = g
                constexpr uint32_t row = 130u * 8u;
                const uint32_t b0 = bring16 + st[0] * b16, b1 = bring16 + st[1] * b16, b2 = bring16 + st[2] * b16;
                const uint32_t d0 = d_base, d1 = d_base + 128;
#pragma unroll
                for (uint32_t k = 0; k < 8; k += 2) {
                  const uint32_t acc = (cb == 0 && g == 0 && k == 0) ? 0u : 1u;   // the first (q = 1) MMAs initialise both rows
                  if (st[1] == st[0] + 1) {
                    umma_f16_lohi(d0, aq + 2 * row + k, a_hi, b0 + k, b_hi, idesc2, acc);
                  } else {
                    umma_f16_lohi(d0, aq + 2 * row + k, a_hi, b0 + k, b_hi, idesc, acc);
                    umma_f16_lohi(d1, aq + 2 * row + k, a_hi, b1 + k, b_hi, idesc, acc);
                  }
                  if (st[2] == st[1] + 1) {
                    umma_f16_lohi(d0, aq + row + k, a_hi, b1 + k, b_hi, idesc2, 1u);
                  } else {
                    umma_f16_lohi(d0, aq + row + k, a_hi, b1 + k, b_hi, idesc, 1u);
                    umma_f16_lohi(d1, aq + row + k, a_hi, b2 + k, b_hi, idesc, 1u);
                  }
                  umma_f16_lohi(d1, aq + 3 * row + k, a_hi, b0 + k, b_hi, idesc, 1u);     // q = 2: row 1 only, W(dy = +1)
                  umma_f16_lohi(d0, aq + k, a_hi, b2 + k, b_hi, idesc, 1u);               // q = -1: row 0 only, W(dy = -1)
                }
              }
              if (leader) {
                umma_commit(&empty_bar[st[0]]);
                umma_commit(&empty_bar[st[1]]);
                umma_commit(&empty_bar[st[2]]);
              }
              __syncwarp();
            }
          } else if (p.rowsMerge) {
            const uint32_t idesc2 = make_idesc_f16(kBM, 2 * p.BN);
            const int ph0 = pg * 2, ph1 = ph0 + 1;
            for (int i = 0; i < 2; ++i) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              if (leader && !(p.dbg & 2)) {
                const int r0 = p.tap_dy[ph0][2 * i] - p.rowsDyMin[pg];
                const uint32_t arow = a_base + static_cast<uint32_t>(r0 * 130 + 1) * 8u;
                const uint32_t aL = arow + p.tap_dx[ph0][2 * i] * 8, aC = arow + p.tap_dx[ph0][2 * i + 1] * 8,
                               aR = arow + p.tap_dx[ph1][2 * i + 1] * 8;
                const uint32_t bL = bring16 + (stage * 4) * b16, bC = bL + b16, bR = bL + 3 * b16;
                const uint32_t dL = d_base, dR = d_base + p.BN;
                constexpr uint32_t row2 = 130u * 8u;                          // second output row
#pragma unroll
                for (uint32_t k = 0; k < 8; k += 2) {
                  const uint32_t acc = (cb == 0 && i == 0 && k == 0) ? 0u : 1u;    // the wide MMA initialises both phases
                  umma_f16_lohi(dL, aC + k, a_hi, bC + k, b_hi, idesc2, acc);
                  umma_f16_lohi(dL + 128, aC + row2 + k, a_hi, bC + k, b_hi, idesc2, acc);
                  umma_f16_lohi(dL, aL + k, a_hi, bL + k, b_hi, idesc, 1u);
                  umma_f16_lohi(dL + 128, aL + row2 + k, a_hi, bL + k, b_hi, idesc, 1u);
                  umma_f16_lohi(dR, aR + k, a_hi, bR + k, b_hi, idesc, 1u);
                  umma_f16_lohi(dR + 128, aR + row2 + k, a_hi, bR + k, b_hi, idesc, 1u);
                }
              }
              if (leader) umma_commit(&empty_bar[stage]);
              __syncwarp();
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
            }
          } else
          for (int j = 0; j < p.rowsPG; ++j) {
            const int ph = pg * p.rowsPG + j;
            for (int t0 = 0; t0 < p.ntaps; t0 += p.tapsPerStage) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              for (int u = 0; u < p.tapsPerStage; ++u)
              if (leader && !(p.dbg & 2)) {
                const int t = t0 + u;
                // input (row, pixel) of output pixel 0 of tile 0 inside the box, in 128-byte rows (16-byte units: x 8)
                const int r0 = p.tap_dy[ph][t] - p.rowsDyMin[pg];
                const int c0 = p.tap_dx[ph][t] + 1;
                const uint32_t a0 = a_base + static_cast<uint32_t>(r0 * 130 + c0) * 8u;
                const uint32_t a1 = a0 + 130u * 8u;                        // second output row
                const uint32_t b0 = bring16 + (stage * p.tapsPerStage + u) * b16;
                const uint32_t d0 = d_base + j * p.BN, d1 = d0 + 128;
                const uint32_t acc = (cb == 0 && t == 0) ? 0u : 1u;
                umma_f16_lohi(d0, a0, a_hi, b0, b_hi, idesc, acc);
                umma_f16_lohi(d1, a1, a_hi, b0, b_hi, idesc, acc);
                umma_f16_lohi(d0, a0 + 2, a_hi, b0 + 2, b_hi, idesc, 1u);
                umma_f16_lohi(d1, a1 + 2, a_hi, b0 + 2, b_hi, idesc, 1u);
                umma_f16_lohi(d0, a0 + 4, a_hi, b0 + 4, b_hi, idesc, 1u);
                umma_f16_lohi(d1, a1 + 4, a_hi, b0 + 4, b_hi, idesc, 1u);
                umma_f16_lohi(d0, a0 + 6, a_hi, b0 + 6, b_hi, idesc, 1u);
                umma_f16_lohi(d1, a1 + 6, a_hi, b0 + 6, b_hi, idesc, 1u);
              }
              if (leader) umma_commit(&empty_bar[stage]);
              __syncwarp();
              if (++stage == SB) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          if (leader) umma_commit(&empty_bar[SB + ab]);                   // this K chunk's input rows are free again
          __syncwarp();
          if (++ab == 2) {
            ab = 0;
            abphase ^= 1;
          }
        }
        if (leader) umma_commit(&tfull_bar[as]);
        __syncwarp();
        if (++as == p.nAcc) {
          as = 0;
          aphase ^= 1;
        }
      }
    } else
    for (int item = (CG == 2 && crank != 0) ? item1 : item0; item < item1; ++item) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * MT * p.BN;
      const uint32_t d_tmem1 = d_tmem + p.BN;                      // accumulator of the second tile (MT == 2)
      constexpr uint32_t a1off = kBM * 128 >> 4;
      uint32_t acc = 0;
      int remaining = chunksTotal;
      for (int s = 0; s < stagesPerTile; ++s) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const int nk = (remaining < p.G ? remaining : p.G) * kPer;     // K=16 steps in this stage
        remaining -= p.G;
        const uint32_t a0 = smem_lo + stage * stage16;
        const uint32_t b0 = a0 + MT * (kBM * 128 >> 4);
        if (leader) {
          if (p.dbg & 2) {
          } else if constexpr (CG == 2) {  // (conv_build: CB == 64 only) one M = 256 MMA per K step across the pair
            umma_f16_lohi_pair(d_tmem, a0, a_hi, b0, b_hi, idesc, acc);
            umma_f16_lohi_pair(d_tmem, a0 + 2, a_hi, b0 + 2, b_hi, idesc, 1u);
            umma_f16_lohi_pair(d_tmem, a0 + 4, a_hi, b0 + 4, b_hi, idesc, 1u);
            umma_f16_lohi_pair(d_tmem, a0 + 6, a_hi, b0 + 6, b_hi, idesc, 1u);
          } else if (kPer == 4) {          // one 64-channel chunk per stage: 4 K-steps inside the 128B swizzle row
            umma_f16_lohi(d_tmem, a0, a_hi, b0, b_hi, idesc, acc);
            if (MT == 2) umma_f16_lohi(d_tmem1, a0 + a1off, a_hi, b0, b_hi, idesc, acc);
            umma_f16_lohi(d_tmem, a0 + 2, a_hi, b0 + 2, b_hi, idesc, 1u);
            if (MT == 2) umma_f16_lohi(d_tmem1, a0 + a1off + 2, a_hi, b0 + 2, b_hi, idesc, 1u);
            umma_f16_lohi(d_tmem, a0 + 4, a_hi, b0 + 4, b_hi, idesc, 1u);
            if (MT == 2) umma_f16_lohi(d_tmem1, a0 + a1off + 4, a_hi, b0 + 4, b_hi, idesc, 1u);
            umma_f16_lohi(d_tmem, a0 + 6, a_hi, b0 + 6, b_hi, idesc, 1u);
            if (MT == 2) umma_f16_lohi(d_tmem1, a0 + a1off + 6, a_hi, b0 + 6, b_hi, idesc, 1u);
          } else {
            uint32_t a_lo = a0;
            uint32_t accl = acc;
            for (int k = 0, j = 0; k < nk; ++k) {
              umma_f16_lohi(d_tmem, a_lo + 2 * j, a_hi, b0 + 2 * k, b_hi, idesc, accl);
              if (MT == 2) umma_f16_lohi(d_tmem1, a_lo + a1off + 2 * j, a_hi, b0 + 2 * k, b_hi, idesc, accl);
              accl = 1;
              if (++j == kPer) {
                j = 0;
                a_lo += sub16;
              }
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
        }
        acc = 1;
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) {                                   // accumulator(s) complete -> epilogue (of both CTAs of a pair)
        if constexpr (CG == 2) umma_commit_pair(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      if (++as == p.nAcc) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 0..3 and 6..)
    const int egrp = CFR_IGEMM_MMA_WARP_LAST ? (warp >> 2) : (warp < 4 ? 0 : (warp - 2) >> 2);   // group 0 .. 2 * MT - 1
    const int slice = MT == 2 ? (egrp & 1) : 0;  // which tile of this CTA's (MT) tiles this group drains
    const int sub = CG == 2 ? static_cast<int>(crank) : slice;     // ... = which tile of the work item
    const int half = MT == 2 ? (egrp >> 1) : egrp;               // which half of the tile's BN columns
    const int hcols = (accW % 32 == 0) ? accW / 2 : (half == 0 ? accW : 0);   // columns of this half (BN = 16 / 48 / ..: half 0 takes all)
    const int col_lo = half * (accW / 2);
    // ROWS with two phases per item: the column halves ARE the two phases (BN = 64 each)
    const int chan_off = (ROWS && p.rowsPG == 2) ? col_lo : 0;
    const int q = warp & 3;                      // TMEM lane quarter this warp may read (== warp id % 4)
    const int row = q * 32 + lane;               // M row == pixel index inside the tile box
    const int tx = row % p.TW;
    const int ty = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const int et = half * 128 + q * 32 + static_cast<int>(lane); // 0..255 among the threads draining this tile
    const int gbar = 1 + slice;                  // named barrier of this tile's two groups
    s_sum += slice * 4 * p.CoutTotal;            // this group's four slices
    s_sq += slice * 4 * p.CoutTotal;
    int as = 0;
    uint32_t aphase = 0;
    int cur_img = -1;
    const bool do_stats = p.stat_sum != nullptr;
    for (int item = item0; item < item1; ++item) {
      TileCoord t;
      if constexpr (ROWS) {
        int n, y0, pg, ntile;
        rows_item(item, n, y0, pg, ntile);
        t.n0 = n; t.y0 = y0 + sub; t.x0 = 0; t.ntile = ntile;
        t.phase = pg * p.rowsPG + (p.rowsPG == 2 ? half : 0);
      } else {
        t = decode_item(p, item, sub);
      }
      if (!ROWS && TPI == 2 && t.n0 >= p.N) {    // dummy tile of an odd tile count
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_remote(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]);
        }
        if (++as == p.nAcc) {
          as = 0;
          aphase ^= 1;
        }
        continue;
      }
      if (do_stats && cur_img >= 0 && t.n0 != cur_img) {
        named_bar_sync(gbar, 256);
        for (int c = et; c < p.CoutTotal; c += 256) {
          flush_stats(p, s_sum, s_sq, cur_img, c);
        }
        named_bar_sync(gbar, 256);
      }
      cur_img = t.n0;
      const int gx = t.x0 + tx, gy = t.y0 + ty, n = t.n0 + tn;
      const bool valid = gx < p.Wout && gy < p.Hout && n < p.N;
      const int oy = gy * p.oscale + p.ooff_y[t.phase];
      const int ox = gx * p.oscale + p.ooff_x[t.phase];
      const size_t pix = (static_cast<size_t>(n) * p.outH + oy) * p.outW + ox;
      const int cls = (gy == 0 ? 0 : (gy == p.Hout - 1 ? 2 : 1)) * 3 + (gx == 0 ? 0 : (gx == p.Wout - 1 ? 2 : 1));
      const float* cb_row = nullptr;
      if (p.cbias != nullptr && valid) {
        cb_row = p.cbias +
                 ((static_cast<size_t>(p.cbiasPerSample ? n : 0) * p.numPhases + t.phase) * 9 + cls) * p.CoutTotal;
      }
      const float nz = (p.noise != nullptr && valid) ? __ldg(&p.noise[oy * p.outW + ox]) : 0.f;

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (as * MT + slice) * accW;
      if (p.argmax_keys != nullptr) {
        // gallery match: score = acc + bias[j] (= 2 e.g_j - |g_j|^2); keep the best column of this tile per row and
        // fold it into the global per-query key (max score, then lowest index == torch.argmax tie-break)
        // -inf start + first column of the tile: a NaN score (fp16 overflow upstream) never wins a comparison, so the
        // pick stays a real row (padded rows carry bias -inf and tile 0 starts at row 0) instead of a padded one
        float best = -INFINITY;
        int best_j = t.ntile * p.BN + col_lo;
        for (int c0 = col_lo; c0 < col_lo + hcols; c0 += 16) {
          float v[16];
          tmem_ld16(t_row + c0, v);
          const int ch0 = t.ntile * p.BN + c0 - chan_off;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (v[i] > best) {
              best = v[i];
              best_j = ch0 + i;
            }
          }
        }
        if (valid && hcols > 0) {
          uint32_t u = __float_as_uint(best);
          u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;                      // order-preserving float -> uint
          const unsigned long long key = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(best_j));
          atomicMax(&p.argmax_keys[n], key);
        }
      } else
      for (int c0 = col_lo; c0 < ((p.dbg & 1) ? col_lo : col_lo + hcols); c0 += 16) {
        const int ch0 = t.ntile * p.BN + c0 - chan_off;
        // the chunk's residual (the one per-pixel global read) is requested BEFORE the TMEM load is waited for: the epilogue
        // is latency-bound, and behind the (asm volatile) tcgen05.ld / wait it would start only afterwards.  (Doing the
        // same for the per-channel vectors costs 48 more registers: the 18-warp kernel is capped at 96 and spills;
        // they are L1-resident after the first tile.)
        const bool has_cb = cb_row != nullptr, has_bias = p.bias != nullptr, has_nz = p.noise != nullptr;
        const bool has_alpha = p.act == ACT_PRELU, has_res = p.resid != nullptr && valid;
        uint4 rv[2];
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.resid + pix * p.residC + ch0);
          rv[0] = __ldg(rp);
          rv[1] = __ldg(rp + 1);
        }
        float v[16];
        tmem_ld16(t_row + c0, v);
        if (has_cb) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(cb_row + ch0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
        if (has_nz) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.noise_w + ch0 + i));
            v[i] += nz * w4.x; v[i + 1] += nz * w4.y; v[i + 2] += nz * w4.z; v[i + 3] += nz * w4.w;
          }
        }
        if (p.act == ACT_LRELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : v[i] * p.slope;
        } else if (has_alpha) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.alpha + ch0 + i));
            v[i] = v[i] >= 0.f ? v[i] : v[i] * a4.x;
            v[i + 1] = v[i + 1] >= 0.f ? v[i + 1] : v[i + 1] * a4.y;
            v[i + 2] = v[i + 2] >= 0.f ? v[i + 2] : v[i + 2] * a4.z;
            v[i + 3] = v[i + 3] >= 0.f ? v[i + 3] : v[i + 3] * a4.w;
          }
        }
        if (has_res) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&rv[h]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(h2[i]);
              v[h * 8 + 2 * i] += f.x;
              v[h * 8 + 2 * i + 1] += f.y;
            }
          }
        }
        if (p.act == ACT_RELU_POST) {             // Inception-ResNet blocks: relu(x + scale * conv(cat))
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (valid) {
          if (p.out32 != nullptr) {
            float4* op = reinterpret_cast<float4*>(p.out32 + pix * p.outC + ch0);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            uint4 o[2];
            __half2* h2 = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) h2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            __half* op = p.out + pix * p.outC + ch0;
            if (p.out256) {
              st_global_256(op, o[0], o[1]);                 // 16 channels = one 32-byte sector per lane
            } else {
              reinterpret_cast<uint4*>(op)[0] = o[0];
              reinterpret_cast<uint4*>(op)[1] = o[1];
            }
          }
        }
        if (do_stats) {
          float sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = valid ? v[i] : 0.f;
            sq[i] = v[i] * v[i];
          }
          const float ssum = warp_reduce16(v, lane);
          const float ssq = warp_reduce16(sq, lane);
          if ((lane & 1) == 0) {
            const int ch = ch0 + reduce16_channel(lane);
            s_sum[q * p.CoutTotal + ch] += ssum;
            s_sq[q * p.CoutTotal + ch] += ssq;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(&tempty_bar[as], 0); else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == p.nAcc) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (do_stats && cur_img >= 0) {
      named_bar_sync(gbar, 256);
      for (int c = et; c < p.CoutTotal; c += 256) {
        flush_stats(p, s_sum, s_sq, cur_img, c);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();      // no CTA leaves while its peer may still read its shared memory / TMEM
  if (warp == Roles<MT>::kMma) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, tmemCols); else tmem_dealloc(tmem_base, tmemCols);
  }
}

// --------------------------------------------------------------------------------------------
// host
// --------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
static unsigned long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
unsigned long long launch_count() { return g_launches; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

void conv_set_argmax(ConvOp* op, unsigned long long* keys) { op->p.argmax_keys = keys; }

void* get_encode_tiled() { return reinterpret_cast<void*>(get_encode()); }

static CUtensorMapSwizzle swz_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                      : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                     : (bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

int conv_build(const ConvSpec& s, ConvOp* op) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return 1;
  }
  ConvParams& p = op->p;
  memset(&p, 0, sizeof(p));
  if (s.TW * s.TH * s.TN != kBM) { set_error("conv: TW*TH*TN must be 128 (got %d*%d*%d)", s.TW, s.TH, s.TN); return 2; }
  if (s.Cin % 16 != 0 || (s.Cin > 64 && s.Cin % 64 != 0)) { set_error("conv: Cin=%d unsupported", s.Cin); return 2; }
  if (s.Cout % 16 != 0) { set_error("conv: Cout=%d must be a multiple of 16", s.Cout); return 2; }
  if (s.ntaps < 1 || s.ntaps > kMaxTaps || s.numPhases < 1 || s.numPhases > kMaxPhases) { set_error("conv: bad taps/phases"); return 2; }
  if (s.stride != 1 && s.stride != 2) { set_error("conv: stride %d unsupported", s.stride); return 2; }
  if (s.stat_sum != nullptr && (s.TN != 1 || s.numPhases != 1 || s.Cout > kStatsMaxC)) { set_error("conv: fused stats need TN==1, one phase, Cout<=512"); return 2; }

  p.N = s.N; p.Hout = s.Hout; p.Wout = s.Wout;
  p.TW = s.TW; p.TH = s.TH; p.TN = s.TN;
  p.tilesX = (s.Wout + s.TW - 1) / s.TW;
  p.tilesY = (s.Hout + s.TH - 1) / s.TH;
  p.tilesN = (s.N + s.TN - 1) / s.TN;
  p.numPhases = s.numPhases;
  p.stride = s.stride;
  p.ntaps = s.ntaps;
  memcpy(p.tap_dy, s.tap_dy, sizeof(p.tap_dy));
  memcpy(p.tap_dx, s.tap_dx, sizeof(p.tap_dx));
  p.CB = s.Cin >= 64 ? 64 : s.Cin;
  p.nCB = s.Cin / p.CB;
  p.G = 64 / p.CB;
  p.swizzleA = p.CB * 2;
  // N tile: largest of 256/128/.. dividing Cout
  int bn = s.Cout;
  // N tile cap.  256 halves the A re-fetch per output channel, but two 256-column accumulators of a tile pair fill
  // TMEM, so the epilogue cannot overlap the next tile; measured (profiles/ncu_r01_notes.md section 10): plain convs
  // are faster with 128 (two accumulator sets in flight, twice the work items on the small ArcFace maps), the
  // 4-phase up-convs (short K, Cin >= 256: ingest-bound) with 256.  CFR_IGEMM_BN_MAX overrides for A/B runs.
  static const int bn_env = getenv("CFR_IGEMM_BN_MAX") != nullptr ? atoi(getenv("CFR_IGEMM_BN_MAX")) : 0;
  // split-precision convs have 3x the K loop (the epilogue share is small) and 3x the A bytes: the wider tile wins
  const int bn_max = bn_env > 0 ? bn_env : ((s.numPhases == 1 && s.kSplit != 3) ? 128 : 256);
  if (bn > bn_max) {
    bn = bn_max;
    while (s.Cout % bn != 0) bn -= 16;
  }
  p.BN = bn;
  p.numNTiles = s.Cout / bn;
  p.CoutTotal = s.Cout;
  // two M tiles per CTA step (shared weight stage) whenever the tile list pairs up and still fills the machine
  {
    const int mTiles = p.tilesX * p.tilesY * p.tilesN;
    static const int force = getenv("CFR_IGEMM_MT") != nullptr ? atoi(getenv("CFR_IGEMM_MT")) : 0;   // A/B knob: 1 or 2
    const bool same_rows = s.wRowsPerSample == 0 || (p.tilesX * p.tilesY) % 2 == 0;   // a pair reads ONE weight tile
    const long long pairItems = static_cast<long long>(mTiles / 2) * s.numPhases * (s.Cout / bn);
    bool mt2 = mTiles >= 2 && same_rows && 2 * bn <= 512 && pairItems >= num_sms() / 2;
    if (force == 1) mt2 = false;
    p.MT = mt2 ? 2 : 1;
    p.nAcc = (2 * p.MT * bn <= 512) ? 2 : 1;
    p.CG = 1;
    // CTA pair (tcgen05 cta_group::2) for plain convs with Cout % 256 == 0: N tile 256 split over the pair's two SMs, one
    // M tile per CTA, both accumulator sets in flight.  Not for 1x1-grid GEMMs (FC, gallery matcher: few M tiles), not
    // for the split-precision convs (their 3x K loop already hides the epilogue at BN = 256).  CFR_IGEMM_CG2=0: A/B runs.
    static const int cg_env = getenv("CFR_IGEMM_CG2") != nullptr ? atoi(getenv("CFR_IGEMM_CG2")) : 1;
    const long long pairItems256 = static_cast<long long>(mTiles / 2) * s.numPhases * (s.Cout / 256);
    // Measured (profiles/ncu_r02_notes.md section 9): L7 3.87 -> 3.56 us, L9 4.10 -> 3.85 us; short-K layers (1x1
    // down-samples, 28x28 Cin 128) lose 10 % to the coarser work items, hence K >= 2304.
    if (cg_env != 0 && s.numPhases == 1 && s.kSplit != 3 && s.Cout % 256 == 0 && s.Cin % 64 == 0 && mTiles >= 2 &&
        same_rows && s.Hout * s.Wout > 1 && pairItems256 >= num_sms() / 4 && (cg_env == 2 || s.ntaps * s.Cin >= 2304)) {
      p.CG = 2;
      p.MT = 1;
      p.nAcc = 2;
      bn = 256;
      p.BN = bn;
      p.numNTiles = s.Cout / bn;
    }
    p.tilesPerItem = p.CG == 2 ? 2 : p.MT;
    // ROWS mode (see the kernel comment): exactly 128 output pixels per row, stride 1, 64-channel K chunks, shared weights
    static const int rows_env = getenv("CFR_IGEMM_ROWS") != nullptr ? atoi(getenv("CFR_IGEMM_ROWS")) : 1;
    p.rows = 0;
    if (rows_env != 0 && p.CG == 1 && s.Wout == 128 && s.Win == 128 && s.Hin == s.Hout && s.stride == 1 && s.TW == 128 &&
        s.TH == 1 && s.TN == 1 && s.Hout % 2 == 0 && s.Cin % 64 == 0 && s.kSplit != 3 && s.wRowsPerSample == 0) {
      bool ok = false, merge = false, merge3 = false;
      if (s.numPhases == 1 && s.ntaps == 9 && s.Cout % 128 == 0) {          // 3x3: rows y-1 .. y+2 serve two output rows
        ok = true;
        for (int t = 0; t < 9; ++t)
          ok = ok && s.tap_dy[0][t] >= -1 && s.tap_dy[0][t] <= 1 && s.tap_dx[0][t] >= -1 && s.tap_dx[0][t] <= 1;
        p.aRows = 4; p.rowsPG = 1; p.rowsNPG = 1; p.rowsDyMin[0] = -1; p.rowsDyMin[1] = -1;
        bn = 128;
        static const int merge3_env = getenv("CFR_IGEMM_ROWS_MERGE") != nullptr ? atoi(getenv("CFR_IGEMM_ROWS_MERGE")) : 1;
        for (int u = 0; u < 9; ++u) p.rowsTap[u] = -1;
        for (int t = 0; ok && t < 9; ++t) p.rowsTap[(s.tap_dx[0][t] + 1) * 3 + (1 - s.tap_dy[0][t])] = t;
        merge3 = ok && merge3_env != 0;
        for (int u = 0; u < 9; ++u) merge3 = merge3 && p.rowsTap[u] >= 0;
      } else if (s.numPhases == 4 && s.ntaps == 4 && s.Cout % 64 == 0 && s.wRowsPerPhase == s.Cout) {
        // nearest-x2 up-conv: phases (a, b) = 2a + b; both column phases of row phase a read input rows {a-1, a}
        ok = true;
        for (int ph = 0; ph < 4; ++ph)
          for (int t = 0; t < 4; ++t) {
            const int a = ph >> 1;
            ok = ok && s.tap_dy[ph][t] >= a - 1 && s.tap_dy[ph][t] <= a && s.tap_dx[ph][t] >= -1 && s.tap_dx[ph][t] <= 1;
          }
        p.aRows = 3; p.rowsPG = 2; p.rowsNPG = 2; p.rowsDyMin[0] = -1; p.rowsDyMin[1] = 0;
        bn = 64;
        // taps of a phase ordered (row, column) and the two column phases of a row phase sharing their middle input column
        static const int merge_env = getenv("CFR_IGEMM_ROWS_MERGE") != nullptr ? atoi(getenv("CFR_IGEMM_ROWS_MERGE")) : 1;
        merge = merge_env != 0;
        for (int a = 0; a < 2; ++a)
          for (int i = 0; i < 2; ++i) {
            const int p0 = 2 * a, p1 = p0 + 1, dy = s.tap_dy[p0][2 * i];
            merge = merge && s.tap_dy[p0][2 * i + 1] == dy && s.tap_dy[p1][2 * i] == dy && s.tap_dy[p1][2 * i + 1] == dy &&
                    s.tap_dx[p0][2 * i + 1] == s.tap_dx[p1][2 * i];
          }
      }
      if (ok) {
        p.rows = 1;
        p.MT = 2;
        p.nAcc = 2;
        p.tilesPerItem = 2;
        p.BN = bn;
        p.numNTiles = s.Cout / bn;
        p.rowsMerge = merge ? 1 : (merge3 ? 2 : 0);
      }
    }
  }
  if (const char* e = getenv("CFR_IGEMM_DBG")) p.dbg = atoi(e);
  p.stageBytes = p.MT * kBM * 128 + (p.CG == 2 ? bn / 2 : bn) * 128;
  const int statBytes = s.stat_sum != nullptr ? p.MT * 8 * s.Cout * 4 : 0;
  const int ctrlBytes = 8 * (2 * 8 + 4) + 16 + statBytes;
  const int budget = 227 * 1024 - 1024 - ctrlBytes;
  p.numStages = budget / p.stageBytes;
  if (p.numStages > 8) p.numStages = 8;
  if (const char* e = getenv("CFR_IGEMM_STAGES_MAX")) {          // experiment: pipeline depth vs throughput (Little's law)
    const int m = atoi(e);
    if (m >= 2 && p.numStages > m) p.numStages = m;
  }
  if (p.rows) {                                        // two input-row boxes + a ring of weight tiles
    p.aBoxBytes = (p.aRows * 130 * 128 + 1023) / 1024 * 1024;
    p.tapsPerStage = ((budget - 2 * p.aBoxBytes) / (s.ntaps * bn * 128) >= 3) ? s.ntaps : 1;
    p.bStages = (budget - 2 * p.aBoxBytes) / (p.tapsPerStage * bn * 128);
    if (p.bStages > 6) p.bStages = 6;
    if (p.rowsMerge == 1 && p.tapsPerStage != s.ntaps) p.rowsMerge = 0;   // stages four tiles (one input row of both phases)
    if (p.rowsMerge == 2 && p.tapsPerStage != 1) p.rowsMerge = 0;         // streams single tiles
    if (p.bStages < 3) p.rows = 0;                     // (cannot happen for BN <= 128; falls back to the tile-pair kernel)
    else {
      p.numStages = p.bStages + 2;
      p.stageBytes = 0;
    }
  }
  if (p.numStages < 3 && p.MT == 2) {                   // not enough stages to pipeline: fall back to one tile per step
    p.MT = 1;
    p.tilesPerItem = 1;
    p.nAcc = 2;
    p.stageBytes = kBM * 128 + bn * 128;
    p.numStages = (227 * 1024 - 1024 - (8 * (2 * 8 + 4) + 16 + (s.stat_sum != nullptr ? 8 * s.Cout * 4 : 0))) / p.stageBytes;
    if (p.numStages > 8) p.numStages = 8;
  }
  if (p.numStages < 2) { set_error("conv: not enough shared memory for 2 stages"); return 2; }
  p.ctrlOffset = p.rows ? 2 * p.aBoxBytes + p.bStages * p.tapsPerStage * bn * 128 : p.numStages * p.stageBytes;
  op->smemBytes = p.ctrlOffset + (8 * (2 * 8 + 4) + 16 + (s.stat_sum != nullptr ? p.MT * 8 * s.Cout * 4 : 0)) + 1024;
  p.wRowsPerSample = s.wRowsPerSample;
  p.wRowsPerPhase = s.wRowsPerPhase;
  if (s.outIsF32) p.out32 = static_cast<float*>(s.out); else p.out = static_cast<__half*>(s.out);
  p.outH = s.outH; p.outW = s.outW; p.outC = s.outC; p.oscale = s.oscale;
  p.out256 = (reinterpret_cast<uintptr_t>(s.out) & 31) == 0 && s.outC % 16 == 0;   // every 16-channel chunk is a 32-byte sector
  memcpy(p.ooff_y, s.ooff_y, sizeof(p.ooff_y));
  memcpy(p.ooff_x, s.ooff_x, sizeof(p.ooff_x));
  p.bias = s.bias; p.cbias = s.cbias; p.cbiasPerSample = s.cbiasPerSample;
  p.noise = s.noise; p.noise_w = s.noise_w;
  p.act = s.act; p.slope = s.slope; p.alpha = s.alpha;
  p.resid = static_cast<const __half*>(s.resid); p.residC = s.residC;
  p.stat_sum = reinterpret_cast<stat_t*>(s.stat_sum); p.stat_sq = reinterpret_cast<stat_t*>(s.stat_sq);

  // activations: (C, W, H, N), fp16
  {
    cuuint64_t dims[4] = {(cuuint64_t)s.Cin, (cuuint64_t)s.Win, (cuuint64_t)s.Hin, (cuuint64_t)s.N};
    cuuint64_t strides[3] = {(cuuint64_t)s.Cin * 2, (cuuint64_t)s.Win * s.Cin * 2, (cuuint64_t)s.Hin * s.Win * s.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.CB, (cuuint32_t)(s.TW * s.stride), (cuuint32_t)(s.TH * s.stride), (cuuint32_t)s.TN};
    cuuint32_t estr[4] = {1, (cuuint32_t)s.stride, (cuuint32_t)s.stride, 1};
    CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(p.swizzleA), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d N=%d box=%u,%u,%u,%u)", (int)r, s.Cin, s.Win, s.Hin, s.N, box[0], box[1], box[2], box[3]); return 3; }
  }
  if (p.rows) {                                        // input rows: (C, W, H, N), box {64, 130, aRows, 1}
    cuuint64_t dims[4] = {(cuuint64_t)s.Cin, (cuuint64_t)s.Win, (cuuint64_t)s.Hin, (cuuint64_t)s.N};
    cuuint64_t strides[3] = {(cuuint64_t)s.Cin * 2, (cuuint64_t)s.Win * s.Cin * 2, (cuuint64_t)s.Hin * s.Win * s.Cin * 2};
    cuuint32_t box[4] = {64, 130, (cuuint32_t)p.aRows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tmA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(s.in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A rows) failed: %d", (int)r); return 3; }
  }
  // weights: (Kpad, rows), fp16
  {
    cuuint64_t dims[2] = {(cuuint64_t)s.Kpad, (cuuint64_t)s.wRows};
    cuuint64_t strides[1] = {(cuuint64_t)s.Kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(p.CG == 2 ? bn / 2 : bn)};     // CTA pair: each CTA stages half of the N tile
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(s.w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d (K=%d rows=%d bn=%d)", (int)r, s.Kpad, s.wRows, bn); return 3; }
  }
  if (s.Kpad % 64 != 0 || s.Kpad < s.ntaps * s.Cin) { set_error("conv: Kpad=%d must be a multiple of 64 and >= taps*Cin", s.Kpad); return 2; }

  const int total = p.rows ? s.N * (s.Hout / 2) * p.rowsNPG * p.numNTiles
                           : ((p.tilesX * p.tilesY * p.tilesN + p.tilesPerItem - 1) / p.tilesPerItem) * p.numPhases * p.numNTiles;
  op->grid = total < num_sms() ? total : num_sms();
  if (p.CG == 2) op->grid = 2 * (total < num_sms() / 2 ? total : num_sms() / 2);      // whole CTA pairs
  if (const char* e = getenv("CFR_MAX_CTAS")) {      // tests: few CTAs => many work items per CTA (ring wrap, phase flips)
    const int m = atoi(e);
    if (m > 0 && op->grid > m) op->grid = m;
    if (p.CG == 2) op->grid = op->grid < 2 ? 2 : (op->grid & ~1);
  }
  // algorithmic FLOPs of this launch: 2 * (valid output-grid pixels) * phases * taps * Cin * Cout
  // (split-precision convs issue 3x the MMAs for the same algorithmic contraction: counted on the logical Cin / 3)
  if (s.kSplit != 0 && s.kSplit != 1 && s.kSplit != 3) { set_error("conv: kSplit=%d unsupported", s.kSplit); return 2; }
  if (s.kSplit == 3 && s.Cin % 3 != 0) { set_error("conv: kSplit=3 needs Cin %% 3 == 0"); return 2; }
  const int cinLogical = s.kSplit == 3 ? s.Cin / 3 : s.Cin;
  op->flops = 2.0 * s.N * s.Hout * s.Wout * s.numPhases * s.ntaps * static_cast<double>(cinLogical) * s.Cout;
  return 0;
}

// ---- optional per-launch timing (bench.py roofline): CUDA events around every conv launch -------------
// kind 0: conv_igemm_kernel (work = algorithmic FLOPs), kind 1: conv_halo_kernel (work = algorithmic bytes)
struct ProfState {
  std::vector<cudaEvent_t> ev;     // pairs (start, stop)
  size_t used = 0;
  double work = 0.0;
  long long launches = 0;
};
static bool g_prof_on = false;
static ProfState g_prof[2];

void profile_enable(int on) {
  g_prof_on = on != 0;
  for (auto& p : g_prof) {
    p.used = 0;
    p.work = 0.0;
    p.launches = 0;
  }
}
int profile_read(int kind, double* ms, double* work, long long* launches) {
  ProfState& p = g_prof[kind & 1];
  double total = 0.0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    float t = 0.f;
    cudaError_t e = cudaEventSynchronize(p.ev[i + 1]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]);
    if (e != cudaSuccess) { set_error("profile_read: %s", cudaGetErrorString(e)); return 5; }
    total += t;
  }
  *ms = total;
  *work = p.work;
  *launches = p.launches;
  return 0;
}
bool profile_on() { return g_prof_on; }
cudaEvent_t profile_event(int kind) {
  ProfState& p = g_prof[kind & 1];
  if (p.used == p.ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p.ev.push_back(e);
  }
  return p.ev[p.used++];
}
void profile_account(int kind, double work) {
  g_prof[kind & 1].work += work;
  g_prof[kind & 1].launches += 1;
}

bool pdl_enabled() {
  static const bool on = getenv("CFR_PDL") == nullptr || atoi(getenv("CFR_PDL")) != 0;     // CFR_PDL=0: A/B runs
  return on;
}

int conv_launch(const ConvOp& op, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_igemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_igemm_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv_igemm_kernel<2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err)); return 4; }
  cudaEvent_t e1 = nullptr;
  if (profile_on()) {
    cudaEventRecord(profile_event(0), stream);
    e1 = profile_event(0);
  }
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (op.p.CG == 2) {                                                   // the CTA pair is a 2-CTA cluster
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_trigger() / pdl_wait() in ptx.cuh
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(op.grid);
  cfg.blockDim = dim3(op.p.MT == 2 ? kConvThreadsMT2 : kConvThreads);
  cfg.dynamicSmemBytes = op.smemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (op.p.rows) cudaLaunchKernelEx(&cfg, conv_igemm_kernel<2, 1, true>, op.p);
  else if (op.p.CG == 2) cudaLaunchKernelEx(&cfg, conv_igemm_kernel<1, 2>, op.p);
  else if (op.p.MT == 2) cudaLaunchKernelEx(&cfg, conv_igemm_kernel<2>, op.p);
  else cudaLaunchKernelEx(&cfg, conv_igemm_kernel<1>, op.p);
  if (profile_on()) {
    cudaEventRecord(e1, stream);
    profile_account(0, op.flops);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv launch: %s", cudaGetErrorString(e)); return 4; }
  count_launch();
  return 0;
}

}  // namespace cfr
