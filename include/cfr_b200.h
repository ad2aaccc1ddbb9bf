/* C ABI of libcfr_b200.so -- the B200 (sm_100a) implementation of the Monte-Carlo certification hot path of
 * juancprzs/certifyingFaceRecognition.
 *
 * The reference has no FFI: its seam is the duck-typed Python pair Smooth / base_classifier
 * (smoothing/smooth.py:21-37,135).  The entry points below are what a binding for that path needs; the
 * reference symbol each one replaces is cited per function.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: plain C, every pointer is a DEVICE pointer owned by the caller unless the name ends in
 * `_host`; every call enqueues on the given cudaStream_t and returns 0 on success (non-zero: see
 * cfr_last_error()).  No hidden host synchronisation except in the *_host entry points.
 * Activations are NHWC fp16, accumulation fp32.
 */
#ifndef CFR_B200_H
#define CFR_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CFR_API __attribute__((visibility("default")))
#else
#define CFR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cfr_stream_t; /* == cudaStream_t */
typedef struct cfr_program cfr_program;   /* ordered list of kernel launches with baked-in tensor maps */
typedef struct cfr_sampler cfr_sampler;   /* the whole Smooth._sample_noise body */
typedef struct cfr_matcher cfr_matcher;   /* tensor-core gallery match for large galleries */

#define CFR_MAX_PHASES 4
#define CFR_MAX_TAPS 9
enum { CFR_ACT_NONE = 0, CFR_ACT_LRELU = 1, CFR_ACT_PRELU = 2,
       CFR_ACT_RELU_POST = 3 /* ReLU applied AFTER the residual add (Inception-ResNet blocks) */ };

/* One implicit-GEMM convolution (tcgen05 / TMEM / TMA).  Replaces the F.conv2d / F.conv_transpose2d calls of
 * stylegan_generator_model.py:667-675,739 and iresnet.py:49,52,55,142,153 together with the element-wise ops
 * fused into the epilogue (see csrc/conv_igemm.cuh). */
typedef struct cfr_conv_desc {
  const void* in; int32_t N, Hin, Win, Cin;   /* NHWC fp16 input; Cin in {16,32} or a multiple of 64 */
  const void* w; int32_t wRows, Kpad;         /* fp16 [wRows][Kpad], K = (tap, cin), Kpad % 64 == 0 */
  int32_t Cout;                               /* multiple of 16 */
  int32_t Hout, Wout;                         /* conv output grid (low-res grid for up-conv phases) */
  int32_t TW, TH, TN;                         /* M-tile box on the output grid, TW*TH*TN == 128 */
  int32_t stride, ntaps, numPhases;
  int8_t tap_dy[CFR_MAX_PHASES][CFR_MAX_TAPS];
  int8_t tap_dx[CFR_MAX_PHASES][CFR_MAX_TAPS];
  int32_t wRowsPerSample, wRowsPerPhase;      /* weight row = n*wRowsPerSample + phase*wRowsPerPhase + cout */
  void* out; int32_t outIsF32, outH, outW, outC, oscale;
  int8_t ooff_y[CFR_MAX_PHASES], ooff_x[CFR_MAX_PHASES];
  const float* bias;                          /* [Cout] or NULL */
  const float* cbias; int32_t cbiasPerSample; /* [(n?) phase][9 border classes][Cout] or NULL */
  const float* noise; const float* noise_w;   /* [outH*outW], [Cout] or NULL */
  int32_t act; float slope; const float* alpha;
  const void* resid; int32_t residC;          /* fp16 [N,outH,outW,residC] or NULL */
  int64_t* stat_sum; int64_t* stat_sq;        /* [N,Cout] per-(n,c) sum / sum of squares, Q43.20 fixed point
                                                 (integer atomics: bit-reproducible), or NULL */
  int32_t kSplit;                             /* 0 / 1: plain fp16 operands.  3: split-precision conv -- the Cin input
                                                 channels are [x_hi | x_lo | x_hi] of Cin/3 logical channels and the
                                                 weights [w_hi | w_hi | w_lo] per tap (hi = fp16(v), lo = fp16(v - hi)),
                                                 so the fp32 accumulator holds x.w to ~2^-21 (the x_lo.w_lo term is
                                                 dropped); algorithmic FLOPs are counted on Cin/3.  Used for the early
                                                 StyleGAN layers, whose rounding errors dominate the embedding error. */
  const int32_t* keepMap;                     /* halo convs, one phase, square output (may be NULL): sparse store.
                                                 keepMap[y] (device, outH entries) = index of output row / column y in a
                                                 COMPACT [N][keepDim][keepDim][Cout] output, or -1: pixel (y, x) is stored
                                                 only when both keepMap[y] and keepMap[x] are >= 0.  The statistics still
                                                 cover every pixel.  Used for the last StyleGAN layer, of whose 1024^2
                                                 pixels the bilinear resize (gen_utils.py:77-85) reads 224^2. */
  int32_t keepDim;
} cfr_conv_desc;

CFR_API const char* cfr_last_error(void);
CFR_API int cfr_version(void);
CFR_API int cfr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- programs: record once, replay per chunk -------------------------------------------------------- */
CFR_API int cfr_program_create(cfr_program** out);
CFR_API void cfr_program_destroy(cfr_program* p);
CFR_API int cfr_program_run(cfr_program* p, cfr_stream_t stream);
CFR_API int cfr_program_num_launches(const cfr_program* p);
/* replay ops [first, last) only -- diagnostics (per-layer comparison with the oracle, tools/diag_precision.py) */
CFR_API int cfr_program_run_range(cfr_program* p, int first, int last, cfr_stream_t stream);
/* per-op introspection / timing (tools/profile_program.py): label, algorithmic FLOPs, and one timed replay with a
 * CUDA event between consecutive ops (synchronises) */
CFR_API const char* cfr_program_op_label(const cfr_program* p, int i);
CFR_API double cfr_program_op_flops(const cfr_program* p, int i);
CFR_API int cfr_program_run_timed(cfr_program* p, cfr_stream_t stream, float* ms_out, int n);

CFR_API int cfr_program_add_conv(cfr_program* p, const cfr_conv_desc* d);
/* Same contract as add_conv for the HBM-bound high-resolution layers (Cin, Cout in {16,32,64}, stride 1): the
 * input band + halo is staged in shared memory once and shared by all taps; weights are [phase*tap*Cout][Cin]
 * (Kpad == Cin).  inA/inB (may be NULL): per-(n, cin) affine x = y*A + B applied on load -- the previous layer's
 * InstanceNorm + AdaIN (stylegan_generator_model.py:420-422,:505) -- with the conv's zero padding left at zero. */
CFR_API int cfr_program_add_conv_halo(cfr_program* p, const cfr_conv_desc* d, const float* inA, const float* inB);
/* Folded variant (Cin <= 64): records (1) a per-sample weight-folding kernel -- w_main[n] = fp16(base_w * A[n]); one aux
 * tile per phase carrying the noise gain, and per 3x3-neighbourhood position sum_ci base_w*B[n] (+ bias at the centre) --
 * and (2) the halo conv, which feeds one extra 16-channel row per output pixel {noise, inside-image indicators} through
 * ONE more MMA per tile, so IN+AdaIN, noise and bias cost nothing on the CUDA cores and stay exact at the borders.
 * base_w: fp32 [phases*taps][Cout][Cin]; w_main_f16: [N][phases*taps*Cout][Cin]; w_aux_f16: [N][phases*Cout][16] (zeroed).
 * d->bias / d->noise / d->noise_w are consumed by the fold; d->w is ignored.  The layout INSIDE w_main is the library's
 * business (3x3 convs: taps reordered for row-stationary MMAs; nearest-x2 4-phase up-convs: one operand per shared input
 * position, held in a library-owned buffer of 18 tiles per sample -- w_main_f16 is then left untouched). */
CFR_API int cfr_program_add_conv_halo_folded(cfr_program* p, const cfr_conv_desc* d, const float* base_w, const float* inA,
                                     const float* inB, void* w_main_f16, void* w_aux_f16);
/* UpConvBlock INCLUDING its BlurLayer and epilogue (stylegan_generator_model.py:665-678, :463, :559-562) as ONE
 * halo conv: nearest-x2, 3x3 conv and the [1,2,1]^2 blur are linear, so each sub-pixel phase is a 3x3 conv on the
 * low-res grid with composite weights (engine.composite_upconv_weights); the blur's zero padding of the raw image is
 * reproduced exactly by separate weight sets for the first / last hi-res row and a small per-column correction.
 * d: 4 phases x 9 taps (3x3), noise/bias/act/stats as for add_conv.  base_w: fp32 [8 sets * 9 taps * Cout][Cin];
 * corr_d: fp32 [2][2][3][3][Cout][Cin]; w_main_f16 [N][8*9*Cout][Cin]; w_aux_f16 [N][8*Cout][16]; corr_buf
 * fp32 [N][2][outH][Cout]. */
CFR_API int cfr_program_add_upconv_blur_folded(cfr_program* p, const cfr_conv_desc* d, const float* base_w,
                                       const float* corr_d, const float* inA, const float* inB, void* w_main_f16,
                                       void* w_aux_f16, float* corr_buf);
CFR_API int cfr_program_add_memset(cfr_program* p, void* ptr, int value, size_t bytes);
/* StyleModulationLayer dense, stylegan_generator_model.py:503 (all 18 layers): styles[b][rows] */
CFR_API int cfr_program_add_styles(cfr_program* p, const float* wp2, const float* w_style, const float* b_style,
                           int rows, int rows_trunc, int b, float* styles);
/* FirstConvBlock + epilogue, :581-584 */
CFR_API int cfr_program_add_layer0(cfr_program* p, const float* xhat0, const float* styles, int style_stride,
                           int style_off, int b, void* out_f16);
/* split-precision variant: out [b,4,4,3*512] = [hi | lo | hi] (see cfr_conv_desc.kSplit) */
CFR_API int cfr_program_add_layer0_split(cfr_program* p, const float* xhat0, const float* styles, int style_stride,
                                 int style_off, int b, void* out_f16_split);
/* BlurLayer :463 + noise/bias/LeakyReLU :560-562 + InstanceNorm sums :420-422.  mode 1 = sums only */
CFR_API int cfr_program_add_blur_act_stats(cfr_program* p, const void* raw_f16, void* y_f16, int n, int h, int w, int c,
                                   const float* noise, const float* noise_w, const float* bias, int64_t* sum,
                                   int64_t* sq, int mode);
/* InstanceNorm + AdaIN coefficients: x = y*A + B  (:420-422, :505) */
CFR_API int cfr_program_add_finalize_stats(cfr_program* p, const int64_t* sum, const int64_t* sq, const float* styles,
                                   int style_stride, int style_off, int n, int c, float inv_count, float* A,
                                   float* B);
CFR_API int cfr_program_add_affine(cfr_program* p, const void* y_f16, const float* A, const float* B, int n, int hw, int c,
                           void* x_f16);
/* fp32 activations of the split-precision layers: blur/act/stats fp32 -> fp32, and x = y*A + B written as fp16
 * operands, split == 3: [.., 3c] = [hi | lo | hi] for the next split conv; split == 1: plain [.., c] */
CFR_API int cfr_program_add_blur_act_stats_f32(cfr_program* p, const float* raw, float* y, int n, int h, int w, int c,
                                       const float* noise, const float* noise_w, const float* bias, int64_t* sum,
                                       int64_t* sq, int mode);
CFR_API int cfr_program_add_affine_f32(cfr_program* p, const float* y, const float* A, const float* B, int n, int hw, int c,
                               void* x_f16, int split);
/* out[n][c] = bias[c] + sum_p in[n][p][c], fixed order: the reduction of a split-K GEMM recorded as a multi-phase conv
 * (the ArcFace FC 25088 -> 512, iresnet.py:153, runs as 4 K-slices so that it fills more than 8 SMs) */
CFR_API int cfr_program_add_sum_partials(cfr_program* p, const float* in, const float* bias, int n, int parts, int c,
                                 float* out);
/* facenet_pytorch.InceptionResnetV1 glue (main_attack.py:126-129; no source under /root/reference: parity unpinned):
 * MaxPool2d(3, stride 2) NHWC fp16 -> channel slice [c_off, c_off + c) of a wider NHWC buffer (torch.cat of the Mixed
 * blocks); AdaptiveAvgPool2d(1); F.normalize(p=2, dim=1) on fp32 rows. */
CFR_API int cfr_program_add_maxpool3s2(cfr_program* p, const void* in_f16, int n, int h, int w, int c, void* out_f16,
                                       int out_c_total, int c_off);
CFR_API int cfr_program_add_avgpool(cfr_program* p, const void* in_f16, int n, int hw, int c, void* out_f16);
CFR_API int cfr_program_add_l2norm(cfr_program* p, const float* in, int n, int c, float* out);
CFR_API int cfr_program_add_torgb_resize(cfr_program* p, const void* x_f16, const float* A, const float* B, int n, int hin,
                                 int c, const float* w_rgb, const float* b_rgb, int rout, float mean, float stdv,
                                 void* out_f16_nhwc16, float* out_planar_f32, const int32_t* out_slot);
/* out_slot (device int, may be NULL): the n images are written at group *out_slot of a [groups*n, R, R, 16] buffer */
/* the same reading a COMPACT source [n][keep_dim][keep_dim][c] written by a sparse-store conv (cfr_conv_desc.keepMap) */
CFR_API int cfr_program_add_torgb_resize_sparse(cfr_program* p, const void* x_f16, const float* A, const float* B, int n,
                                        int hin, int c, const float* w_rgb, const float* b_rgb, int rout, float mean,
                                        float stdv, void* out_f16_nhwc16, float* out_planar_f32, const int32_t* out_slot,
                                        const int32_t* keep_map, int keep_dim);

/* ---- immediate ops ------------------------------------------------------------------------------------ */
/* L2Certificate.sample_noise (certificate.py:64-67) + WrappedModel.forward latent perturbation
 * (smoothing_model.py:63-67) + TruncationModule (stylegan_generator_model.py:322-328).
 * noise_in == NULL: Philox4x32-10 normals, counter = sample_offset + i, key = seed, scaled by sigma[1|5]. */
CFR_API int cfr_noise_project(const float* z, const float* x, const float* sigma, int sigma_len, const float* noise_in,
                      const float* dir_mat, const float* w_avg, float psi, uint64_t seed, uint64_t sample_offset,
                      int b, float* noise_out, float* wp2, cfr_stream_t stream);
CFR_API int cfr_truncate(const float* w, const float* w_avg, float psi, int b, float* wp2, cfr_stream_t stream);
/* MappingModule.forward (stylegan_generator_model.py:265-295; PixelNormLayer :398-406, DenseBlock :765-815, WScaleLayer
 * :508-535) -- the Z -> W step of generate_data.py:58-123 / ModStyleGANGenerator.synthesize 'Z' (mod_stylegan_generator.py
 * :228-236).  wt: [8][512 in][512 out] fp32 = W_l^T * (sqrt(2)/sqrt(512) * 0.01); bias: [8][512] = b_l * 0.01. */
CFR_API int cfr_mapping(const float* z, const float* wt, const float* bias, int b, float* w_out, cfr_stream_t stream);
/* WrappedModel.compute_probs + .argmax(1) + Smooth._count_arr (smoothing_model.py:56-61, smooth.py:135-146).
 * keys: b uint64 scratch, all-ones before the first call (re-armed by the call). */
CFR_API int cfr_match_vote(const float* emb, int b, const float* gallery, int n, uint64_t* keys, int32_t* pred,
                   int64_t* counts, cfr_stream_t stream);

/* Large-gallery variant of cfr_match_vote (BASELINE config 5, up to 1 M identities): argmax_j (2 e.g_j - |g_j|^2) as one
 * tcgen05 GEMM over fp16 hi/lo splits of both operands (e.g = e_h.g_h + e_h.g_l + e_l.g_h, i.e. ~2^-22 relative --
 * fp32-class scores), with the running argmax (first index on ties) fused into the GEMM epilogue; nothing of size
 * [b, N] is ever written.  create() splits the gallery once ([N,1536] fp16 + |g|^2). */
CFR_API int cfr_matcher_create(const float* gallery, int n_gallery, int max_b, cfr_stream_t stream, cfr_matcher** out);
CFR_API void cfr_matcher_destroy(cfr_matcher* m);
CFR_API int cfr_matcher_run(cfr_matcher* m, const float* emb, int b, int32_t* pred, int64_t* counts, cfr_stream_t stream);

/* Gallery sharded over ranks (partition C; the reference keeps the whole gallery on one device,
 * smoothing_model.py:56-61).  Each rank matches the replicated queries against ITS rows [row_offset, row_offset + n) and
 * emits one 64-bit key per query; the UNSIGNED MINIMUM of the keys over ranks (all-gather + min, done by the host side)
 * is the global winner with torch.argmax's first-index tie-break, its low 32 bits the global row.  cfr_match_keys is the
 * exact fp32 matcher, cfr_matcher_keys the tensor-core one (shards of 32 768+ rows); cfr_vote_keys tallies merged keys. */
CFR_API int cfr_match_keys(const float* emb, int b, const float* gallery, int n, uint32_t row_offset, uint64_t* keys,
                           cfr_stream_t stream);
CFR_API int cfr_matcher_keys(cfr_matcher* m, const float* emb, int b, uint32_t row_offset, uint64_t* keys,
                             cfr_stream_t stream);
CFR_API int cfr_vote_keys(const uint64_t* keys, int b, int32_t* pred, int64_t* counts, cfr_stream_t stream);

/* ---- Smooth._sample_noise (smooth.py:109-138) as one call --------------------------------------------- */
typedef struct cfr_sampler_desc {
  cfr_program* synth;     /* wp2 -> image at FRM resolution (one chunk) */
  cfr_program* frm;       /* image -> embeddings */
  int32_t chunk;          /* samples per program run */
  float* wp2;             /* [chunk,2,512] program input */
  const float* emb;       /* [chunk,512] program output */
  const float* dir_mat;   /* [5,512] */
  const float* w_avg;     /* [512] */
  float psi;
  const float* gallery;   /* [n_gallery,512] */
  int32_t n_gallery;
  /* optional: run the FRM once per frm_group synthesis chunks (better SM fill for the small ArcFace layers) */
  int32_t frm_group;      /* K >= 1 */
  cfr_program* frm_big;   /* image [K*chunk] -> embeddings, or NULL */
  const float* emb_big;   /* [K*chunk,512] */
  int32_t* out_slot;      /* device int read by the synthesis program's torgb_resize op */
  cfr_matcher* matcher;   /* optional: tensor-core match (max_b >= frm_group*chunk) instead of the exact SIMT kernel */
  cfr_sampler* tail;      /* optional: a sampler recorded for a SMALLER chunk (it may have a tail of its own: a chain in
                             descending chunk size).  A program run always costs a whole chunk, so the remainder of a call
                             (num % chunk samples) goes to the smallest sampler of the chain that still holds it -- what
                             keeps 13-sample selection passes (N0 = 100 split over 8 ranks) from costing 125 samples */
  /* optional: overlap.  When img_frm != NULL the FRM programs were recorded on img_frm (a second image buffer) and the
   * sampler runs them, the gallery match and the vote on an internal second stream while the caller's stream already
   * synthesises the next group: after a group's synthesis the images are copied img_src -> img_frm (img_chunk_bytes per
   * chunk, device to device), which is all the two streams share.  The persistent conv kernels of the two programs then
   * fill each other's tail waves and launch gaps.  Results are identical to the serial order (same kernels, same data);
   * on return everything the call enqueued is ordered before later work on `stream`. */
  const void* img_src;
  void* img_frm;
  uint64_t img_chunk_bytes;
} cfr_sampler_desc;
CFR_API int cfr_sampler_create(const cfr_sampler_desc* d, cfr_sampler** out);
CFR_API void cfr_sampler_destroy(cfr_sampler* s);
/* switch the two-stream overlap off / on again (per-kernel CUDA-event timing is only meaningful without it) */
CFR_API int cfr_sampler_set_overlap(cfr_sampler* s, int on);
/* counts[n_gallery] (int64) is ACCUMULATED into.  noise_in: NULL or [num,5] already-scaled noise.
 * pred_out / emb_out / noise_out: optional per-sample outputs ([num], [num,512], [num,5]). */
CFR_API int cfr_sample_votes(cfr_sampler* s, const float* z, const float* x, const float* sigma, int sigma_len,
                     const float* noise_in, int64_t num, uint64_t seed, uint64_t sample_offset, int64_t* counts,
                     int32_t* pred_out, float* emb_out, float* noise_out, cfr_stream_t stream);
/* Several identities in ONE call: identity g draws num_host[g] samples around (z[g], x[g]) with Philox offsets
 * sample_offset_host[g] .. and tallies them into counts[g][n_gallery] (ACCUMULATED).  The samples of consecutive
 * identities share program runs (a chunk is filled across identity boundaries), so the 12-13 samples per rank of a
 * selection pass split over 8 ranks (N0 = 100, smooth.py:64) cost 1/8 of a chunk instead of a whole one when 8 identities
 * are certified together (Smooth.certify_many).  z: [n_ids,512], x: [n_ids,5] device; num_host / sample_offset_host: HOST
 * arrays [n_ids].  Per-sample results are those of cfr_sample_votes with the same (seed, offset). */
CFR_API int cfr_sample_votes_multi(cfr_sampler* s, int n_ids, const float* z, const float* x, const float* sigma,
                           int sigma_len, const int64_t* num_host, uint64_t seed, const uint64_t* sample_offset_host,
                           int64_t* counts, cfr_stream_t stream);
/* Same with HOST buffers (z[512], x[5], sigma[sigma_len] in, counts_host[n_gallery] out, overwritten):
 * H2D copies, the MC loop, the D2H copy of the counts and a stream sync all inside the call. */
CFR_API int cfr_sample_votes_host(cfr_sampler* s, const float* z_host, const float* x_host, const float* sigma_host,
                          int sigma_len, int64_t num, uint64_t seed, uint64_t sample_offset, int64_t* counts_host,
                          cfr_stream_t stream);
/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
CFR_API uint64_t cfr_launch_count(void);
/* Per-launch CUDA-event timing of the two tcgen05 conv kernels (bench.py roofline).  enable(1) resets the
 * counters; read() synchronises on the recorded events and returns, since enable: summed device time, summed
 * algorithmic work and launch count.  kind 0 = conv_igemm_kernel (work = FLOPs, 2*MAC on un-padded dims),
 * kind 1 = conv_halo_kernel (work = bytes: input read once + output written once). */
CFR_API int cfr_profile_enable(int on);
CFR_API int cfr_profile_read(int kind, double* ms, double* work, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* CFR_B200_H */
