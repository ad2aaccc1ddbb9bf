// C ABI (include/cfr_b200.h): programs (recorded launch lists), immediate ops and the MC sampler.
#include "../../include/cfr_b200.h"
#include "conv_igemm.cuh"
#include "conv_halo.cuh"
#include "kernels.cuh"

#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>

using namespace cfr;

struct cfr_program {
  std::vector<std::function<int(cudaStream_t)>> ops;
  std::vector<std::string> labels;
  std::vector<double> flops;           // algorithmic FLOPs per op (convs only, else 0)
  std::vector<std::unique_ptr<ConvOp>> convs;
  std::vector<std::unique_ptr<HaloOp>> halos;
  std::vector<cudaEvent_t> events;
  std::vector<void*> owned;            // device buffers the program allocated itself (packed noise tables)
  void add(std::function<int(cudaStream_t)> f, std::string label, double fl = 0.0) {
    ops.push_back(std::move(f));
    labels.push_back(std::move(label));
    flops.push_back(fl);
  }
  ~cfr_program() {
    for (auto e : events) cudaEventDestroy(e);
    for (auto q : owned) cudaFree(q);
  }
};

struct cfr_matcher {
  int n_gallery = 0, n_pad = 0, max_b = 0;
  __half* g_split = nullptr;            // [n_pad][1536]  = [g_h, g_l, g_h]
  float* g_bias = nullptr;              // [n_pad]        = -|g|^2
  __half* q_split = nullptr;            // [max_b][1536]  = [2e_h, 2e_h, 2e_l]
  unsigned long long* keys = nullptr;   // [max_b]
  ConvOp op;
  // owns its device buffers: an early return from cfr_matcher_create (e.g. out of memory on a 1 M-row gallery) frees
  // whatever had been allocated through the unique_ptr
  ~cfr_matcher() {
    cudaFree(g_split);
    cudaFree(g_bias);
    cudaFree(q_split);
    cudaFree(keys);
  }
};

struct cfr_sampler {
  cfr_sampler_desc d;
  unsigned long long* keys = nullptr;   // [chunk]
  float* dev_in = nullptr;              // z[512] | x[5] (pad 8) | sigma[5] (pad 8)   (host-entry staging)
  long long* dev_counts = nullptr;      // [n_gallery]
  float* pin_in = nullptr;              // pinned mirror of dev_in
  long long* pin_counts = nullptr;      // pinned [n_gallery]
  // two-stream overlap (cfr_sampler_desc.img_frm): FRM + match + vote of group i on `s2` under the synthesis of group i+1
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_start = nullptr, ev_synth = nullptr, ev_copied = nullptr, ev_done = nullptr;
  bool overlap = false;                 // enabled (cfr_sampler_set_overlap)
  bool active = false;                  // ... and in use by the current call (calls of a single group stay serial)
  bool copied_pending = false;          // an img_src -> img_frm copy the caller's stream has not waited for yet
  // owns its buffers / stream / events (early returns from cfr_sampler_create free what had been created)
  ~cfr_sampler() {
    cudaFree(keys);
    cudaFree(dev_in);
    cudaFree(dev_counts);
    cudaFreeHost(pin_in);
    cudaFreeHost(pin_counts);
    if (s2 != nullptr) {
      cudaStreamSynchronize(s2);
      cudaStreamDestroy(s2);
    }
    for (cudaEvent_t e : {ev_start, ev_synth, ev_copied, ev_done})
      if (e != nullptr) cudaEventDestroy(e);
  }
};

static inline cudaStream_t S(cfr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define CFR_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      set_error("%s: %s", #call, cudaGetErrorString(e__));                    \
      return 5;                                                               \
    }                                                                         \
  } while (0)

// ---- the FRM side of one group: serial on `stream`, or on the sampler's second stream behind an image copy ----------
static int frm_stream_begin(cfr_sampler* s, cudaStream_t stream, int chunks, cudaStream_t* out) {
  const size_t bytes = static_cast<size_t>(s->d.img_chunk_bytes) * chunks;
  if (!s->active) {
    if (s->d.img_frm != nullptr)        // programs are recorded on img_frm: serial mode still needs the copy
      CFR_CUDA(cudaMemcpyAsync(s->d.img_frm, s->d.img_src, bytes, cudaMemcpyDeviceToDevice, stream));
    *out = stream;
    return 0;
  }
  CFR_CUDA(cudaEventRecord(s->ev_synth, stream));
  CFR_CUDA(cudaStreamWaitEvent(s->s2, s->ev_synth, 0));
  CFR_CUDA(cudaMemcpyAsync(s->d.img_frm, s->d.img_src, bytes, cudaMemcpyDeviceToDevice, s->s2));
  CFR_CUDA(cudaEventRecord(s->ev_copied, s->s2));
  s->copied_pending = true;
  *out = s->s2;
  return 0;
}
// before the caller's stream overwrites img_src again
static int synth_may_overwrite(cfr_sampler* s, cudaStream_t stream) {
  if (s->active && s->copied_pending) {
    CFR_CUDA(cudaStreamWaitEvent(stream, s->ev_copied, 0));
    s->copied_pending = false;
  }
  return 0;
}
static int overlap_begin(cfr_sampler* s, cudaStream_t stream, bool several_groups) {
  s->active = s->overlap && several_groups;
  s->copied_pending = false;            // a call that failed half-way may have left it set
  if (!s->active) return 0;
  CFR_CUDA(cudaEventRecord(s->ev_start, stream));          // e.g. the caller's zeroing of `counts` precedes the votes
  CFR_CUDA(cudaStreamWaitEvent(s->s2, s->ev_start, 0));
  return 0;
}
static int overlap_join(cfr_sampler* s, cudaStream_t stream) {
  if (!s->active) return 0;
  s->active = false;
  s->copied_pending = false;
  CFR_CUDA(cudaEventRecord(s->ev_done, s->s2));
  CFR_CUDA(cudaStreamWaitEvent(stream, s->ev_done, 0));
  return 0;
}

extern "C" {

CFR_API const char* cfr_last_error(void) { return last_error(); }
CFR_API int cfr_version(void) { return 100; }
CFR_API uint64_t cfr_launch_count(void) { return launch_count(); }
CFR_API int cfr_profile_enable(int on) { profile_enable(on); return 0; }
CFR_API int cfr_profile_read(int kind, double* ms, double* work, int64_t* launches) {
  long long l = 0;
  int r = profile_read(kind, ms, work, &l);
  *launches = l;
  return r;
}

CFR_API int cfr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CFR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CFR_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

CFR_API int cfr_program_create(cfr_program** out) {
  *out = new cfr_program();
  return 0;
}
CFR_API void cfr_program_destroy(cfr_program* p) { delete p; }
CFR_API int cfr_program_num_launches(const cfr_program* p) { return static_cast<int>(p->ops.size()); }
CFR_API const char* cfr_program_op_label(const cfr_program* p, int i) {
  return (i >= 0 && i < static_cast<int>(p->labels.size())) ? p->labels[i].c_str() : "";
}
CFR_API double cfr_program_op_flops(const cfr_program* p, int i) {
  return (i >= 0 && i < static_cast<int>(p->flops.size())) ? p->flops[i] : 0.0;
}
// Run once with a CUDA event between consecutive ops; ms_out[i] = device time of op i.  Synchronises.
CFR_API int cfr_program_run_timed(cfr_program* p, cfr_stream_t stream, float* ms_out, int n) {
  const size_t need = p->ops.size() + 1;
  while (p->events.size() < need) {
    cudaEvent_t e;
    CFR_CUDA(cudaEventCreate(&e));
    p->events.push_back(e);
  }
  CFR_CUDA(cudaEventRecord(p->events[0], S(stream)));
  for (size_t i = 0; i < p->ops.size(); ++i) {
    int r = p->ops[i](S(stream));
    if (r != 0) return r;
    CFR_CUDA(cudaEventRecord(p->events[i + 1], S(stream)));
  }
  CFR_CUDA(cudaStreamSynchronize(S(stream)));
  for (size_t i = 0; i < p->ops.size() && static_cast<int>(i) < n; ++i)
    CFR_CUDA(cudaEventElapsedTime(&ms_out[i], p->events[i], p->events[i + 1]));
  return 0;
}

CFR_API int cfr_program_run(cfr_program* p, cfr_stream_t stream) {
  for (auto& op : p->ops) {
    int r = op(S(stream));
    if (r != 0) return r;
  }
  return 0;
}

// Replay ops [first, last) only (diagnostics: tools/diag_precision.py compares every layer with the oracle).
CFR_API int cfr_program_run_range(cfr_program* p, int first, int last, cfr_stream_t stream) {
  if (first < 0 || last > static_cast<int>(p->ops.size()) || first > last) {
    set_error("cfr_program_run_range: [%d, %d) outside [0, %d)", first, last, static_cast<int>(p->ops.size()));
    return 2;
  }
  for (int i = first; i < last; ++i) {
    int r = p->ops[i](S(stream));
    if (r != 0) return r;
  }
  return 0;
}

CFR_API int cfr_program_add_conv(cfr_program* p, const cfr_conv_desc* d) {
  std::unique_ptr<ConvOp> op(new ConvOp());
  int r = conv_build(*d, op.get());
  if (r != 0) return r;
  ConvOp* raw = op.get();
  p->convs.push_back(std::move(op));
  char lab[160];
  snprintf(lab, sizeof(lab), "conv %dx%d s%d taps%dx%d Cin%d%s Cout%d grid%dx%d n%d BN%d%s", d->Hout, d->Wout, d->stride,
           d->numPhases, d->ntaps, d->kSplit == 3 ? d->Cin / 3 : d->Cin, d->kSplit == 3 ? "(x3 hi/lo)" : "", d->Cout,
           d->Hout, d->Wout, d->N, raw->p.BN, raw->p.rows ? (raw->p.rowsMerge ? " rows-merged" : " rows") : (raw->p.CG == 2 ? " cta-pair" : ""));
  p->add([raw](cudaStream_t st) { return conv_launch(*raw, st); }, lab, raw->flops);
  return 0;
}

CFR_API int cfr_program_add_conv_halo(cfr_program* p, const cfr_conv_desc* d, const float* inA, const float* inB) {
  std::unique_ptr<HaloOp> op(new HaloOp());
  int r = halo_build(*d, inA, inB, nullptr, 0, nullptr, op.get());
  if (r != 0) return r;
  HaloOp* raw = op.get();
  p->halos.push_back(std::move(op));
  char lab[160];
  snprintf(lab, sizeof(lab), "halo %dx%d taps%dx%d Cin%d Cout%d n%d TH%d%s", d->Hout, d->Wout, d->numPhases, d->ntaps,
           d->Cin, d->Cout, d->N, raw->p.TH, inA ? " +affine" : "");
  p->add([raw](cudaStream_t st) { return halo_launch(*raw, st); }, lab, raw->flops);
  return 0;
}

static int add_folded(cfr_program* p, const cfr_conv_desc* d, const float* base_w, const float* inA, const float* inB,
                      void* w_main_f16, void* w_aux_f16, int composite, const float* corr_d, float* corr_buf) {
  std::unique_ptr<HaloOp> op(new HaloOp());
  cfr_conv_desc dd = *d;
  dd.w = w_main_f16;
  const int wsets = composite ? 8 : d->numPhases;
  dd.wRows = d->N * wsets * d->ntaps * d->Cout;
  dd.Kpad = d->Cin;
  const bool upshare = !composite && halo_upshare_ok(*d);
  if (upshare) {                                 // shared-A up-conv: 18 weight tiles per sample (two of them zero) -- the
    const size_t bytes = static_cast<size_t>(d->N) * kUpShareTiles * d->Cout * d->Cin * 2;   // library owns that buffer
    void* own = nullptr;
    if (cudaMalloc(&own, bytes) != cudaSuccess) { set_error("folded conv: cudaMalloc(per-sample weights) failed"); return 5; }
    p->owned.push_back(own);
    if (cudaMemset(own, 0, bytes) != cudaSuccess) { set_error("folded conv: cudaMemset failed"); return 5; }
    w_main_f16 = own;
    dd.w = own;
    dd.wRows = d->N * kUpShareTiles * d->Cout;
  }
  int r = halo_build(dd, nullptr, nullptr, w_aux_f16, composite, composite ? corr_buf : nullptr, op.get());
  if (r != 0) return r;
  HaloOp* raw = op.get();
  p->halos.push_back(std::move(op));
  {                                              // packed aux-row heads (noise [+ corner indicator]) for the loader
    void* tab = nullptr;
    const size_t tab_bytes = static_cast<size_t>(d->Hout) * d->Wout * (composite ? 8 : 4);
    if (cudaMalloc(&tab, tab_bytes) != cudaSuccess) { set_error("folded conv: cudaMalloc(noise table) failed"); return 5; }
    p->owned.push_back(tab);
    raw->p.noise_tab = tab;
    const float* nz = d->noise;
    const int h = d->Hout, w = d->Wout;
    // the noise maps are weights (fixed per program): packed ONCE, now, instead of on every run
    int pr = launch_pack_noise(nz, h, w, composite, tab, nullptr);
    if (pr != 0) return pr;
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("folded conv: packing the noise table failed"); return 5; }
  }
  const int n = d->N, cout = d->Cout, cin = d->Cin, phases = d->numPhases, ntaps = d->ntaps;
  const float* bias = d->bias;
  const float* noise_w = d->noise_w;
  cfr_conv_desc keep = *d;                       // tap tables live in the closure
  const int layout = composite ? 1 : (raw->p.upshare ? 3 : (raw->p.rowmma ? 2 : 0));
  p->add([=](cudaStream_t st) {
    return launch_fold_weights(base_w, inA, inB, bias, noise_w, &keep.tap_dy[0][0], &keep.tap_dx[0][0], n, phases, ntaps,
                               cout, cin, layout, static_cast<__half*>(w_main_f16), static_cast<__half*>(w_aux_f16), st);
  }, "fold_weights");
  if (composite) {
    const __half* yin = static_cast<const __half*>(d->in);
    const int h = d->Hin, w = d->Win;
    p->add([=](cudaStream_t st) { return launch_upblur_corr(yin, inA, inB, corr_d, n, h, w, cin, cout, corr_buf, st); },
           "upblur_corr");
  }
  char lab[160];
  snprintf(lab, sizeof(lab), "halo-%s %dx%d taps%dx%d Cin%d Cout%d n%d TH%d%s", composite ? "upblur" : "folded", d->Hout,
           d->Wout, d->numPhases, d->ntaps, d->Cin, d->Cout, d->N, raw->p.TH, d->keepMap ? " sparse-store" : "");
  p->add([raw](cudaStream_t st) { return halo_launch(*raw, st); }, lab, raw->flops);
  return 0;
}
CFR_API int cfr_program_add_conv_halo_folded(cfr_program* p, const cfr_conv_desc* d, const float* base_w, const float* inA,
                                             const float* inB, void* w_main_f16, void* w_aux_f16) {
  return add_folded(p, d, base_w, inA, inB, w_main_f16, w_aux_f16, 0, nullptr, nullptr);
}
CFR_API int cfr_program_add_upconv_blur_folded(cfr_program* p, const cfr_conv_desc* d, const float* base_w,
                                               const float* corr_d, const float* inA, const float* inB, void* w_main_f16,
                                               void* w_aux_f16, float* corr_buf) {
  return add_folded(p, d, base_w, inA, inB, w_main_f16, w_aux_f16, 1, corr_d, corr_buf);
}

CFR_API int cfr_program_add_memset(cfr_program* p, void* ptr, int value, size_t bytes) {
  p->add([=](cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ptr, value, bytes, st);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return 5; }
    return 0;
  }, "memset");
  return 0;
}

CFR_API int cfr_program_add_styles(cfr_program* p, const float* wp2, const float* w_style, const float* b_style, int rows,
                           int rows_trunc, int b, float* styles) {
  p->add([=](cudaStream_t st) { return launch_styles(wp2, w_style, b_style, rows, rows_trunc, b, styles, st); }, "styles");
  return 0;
}

CFR_API int cfr_program_add_layer0(cfr_program* p, const float* xhat0, const float* styles, int style_stride, int style_off,
                           int b, void* out_f16) {
  p->add([=](cudaStream_t st) {
    return launch_layer0(xhat0, styles, style_stride, style_off, b, static_cast<__half*>(out_f16), st);
  }, "layer0");
  return 0;
}

CFR_API int cfr_program_add_layer0_split(cfr_program* p, const float* xhat0, const float* styles, int style_stride,
                                 int style_off, int b, void* out_f16_split) {
  p->add([=](cudaStream_t st) {
    return launch_layer0_split(xhat0, styles, style_stride, style_off, b, static_cast<__half*>(out_f16_split), st);
  }, "layer0_split");
  return 0;
}

CFR_API int cfr_program_add_blur_act_stats_f32(cfr_program* p, const float* raw, float* y, int n, int h, int w, int c,
                                       const float* noise, const float* noise_w, const float* bias, int64_t* sum,
                                       int64_t* sq, int mode) {
  p->add([=](cudaStream_t st) {
    return launch_blur_act_stats_f32(raw, y, n, h, w, c, noise, noise_w, bias, sum, sq, mode, st);
  }, "blur_act_stats_f32");
  return 0;
}

CFR_API int cfr_program_add_affine_f32(cfr_program* p, const float* y, const float* A, const float* B, int n, int hw, int c,
                               void* x_f16, int split) {
  p->add([=](cudaStream_t st) {
    return launch_affine_f32(y, A, B, n, hw, c, static_cast<__half*>(x_f16), split, st);
  }, split == 3 ? "affine_f32_split" : "affine_f32");
  return 0;
}

CFR_API int cfr_program_add_blur_act_stats(cfr_program* p, const void* raw_f16, void* y_f16, int n, int h, int w, int c,
                                   const float* noise, const float* noise_w, const float* bias, int64_t* sum, int64_t* sq,
                                   int mode) {
  p->add([=](cudaStream_t st) {
    return launch_blur_act_stats(static_cast<const __half*>(raw_f16), static_cast<__half*>(y_f16), n, h, w, c, noise,
                                 noise_w, bias, sum, sq, mode, st);
  }, "blur_act_stats");
  return 0;
}

CFR_API int cfr_program_add_finalize_stats(cfr_program* p, const int64_t* sum, const int64_t* sq, const float* styles,
                                   int style_stride, int style_off, int n, int c, float inv_count, float* A, float* B) {
  p->add([=](cudaStream_t st) {
    return launch_finalize_stats(sum, sq, styles, style_stride, style_off, n, c, inv_count, A, B, st);
  }, "finalize_stats");
  return 0;
}

CFR_API int cfr_program_add_affine(cfr_program* p, const void* y_f16, const float* A, const float* B, int n, int hw, int c,
                           void* x_f16) {
  p->add([=](cudaStream_t st) {
    return launch_affine(static_cast<const __half*>(y_f16), A, B, n, hw, c, static_cast<__half*>(x_f16), st);
  }, "affine");
  return 0;
}

CFR_API int cfr_program_add_sum_partials(cfr_program* p, const float* in, const float* bias, int n, int parts, int c,
                                 float* out) {
  p->add([=](cudaStream_t st) { return launch_sum_partials(in, bias, n, parts, c, out, st); }, "sum_partials");
  return 0;
}

CFR_API int cfr_program_add_maxpool3s2(cfr_program* p, const void* in_f16, int n, int h, int w, int c, void* out_f16,
                                       int out_c_total, int c_off) {
  p->add([=](cudaStream_t st) {
    return launch_maxpool3s2(static_cast<const __half*>(in_f16), n, h, w, c, static_cast<__half*>(out_f16), out_c_total, c_off, st);
  }, "maxpool3s2");
  return 0;
}
CFR_API int cfr_program_add_avgpool(cfr_program* p, const void* in_f16, int n, int hw, int c, void* out_f16) {
  p->add([=](cudaStream_t st) {
    return launch_avgpool(static_cast<const __half*>(in_f16), n, hw, c, static_cast<__half*>(out_f16), st);
  }, "avgpool");
  return 0;
}
CFR_API int cfr_program_add_l2norm(cfr_program* p, const float* in, int n, int c, float* out) {
  p->add([=](cudaStream_t st) { return launch_l2norm(in, n, c, out, st); }, "l2norm");
  return 0;
}

CFR_API int cfr_program_add_torgb_resize(cfr_program* p, const void* x_f16, const float* A, const float* B, int n, int hin,
                                 int c, const float* w_rgb, const float* b_rgb, int rout, float mean, float stdv,
                                 void* out_f16_nhwc16, float* out_planar_f32, const int32_t* out_slot) {
  p->add([=](cudaStream_t st) {
    return launch_torgb_resize(static_cast<const __half*>(x_f16), A, B, n, hin, c, w_rgb, b_rgb, rout, mean, stdv,
                               static_cast<__half*>(out_f16_nhwc16), out_planar_f32, out_slot, st);
  }, "torgb_resize");
  return 0;
}

CFR_API int cfr_program_add_torgb_resize_sparse(cfr_program* p, const void* x_f16, const float* A, const float* B, int n,
                                        int hin, int c, const float* w_rgb, const float* b_rgb, int rout, float mean,
                                        float stdv, void* out_f16_nhwc16, float* out_planar_f32, const int32_t* out_slot,
                                        const int32_t* keep_map, int keep_dim) {
  if (keep_map == nullptr || keep_dim <= 0) { set_error("torgb_resize_sparse: keep_map / keep_dim missing"); return 2; }
  p->add([=](cudaStream_t st) {
    return launch_torgb_resize(static_cast<const __half*>(x_f16), A, B, n, hin, c, w_rgb, b_rgb, rout, mean, stdv,
                               static_cast<__half*>(out_f16_nhwc16), out_planar_f32, out_slot, st, keep_map, keep_dim);
  }, "torgb_resize_sparse");
  return 0;
}

CFR_API int cfr_noise_project(const float* z, const float* x, const float* sigma, int sigma_len, const float* noise_in,
                      const float* dir_mat, const float* w_avg, float psi, uint64_t seed, uint64_t sample_offset, int b,
                      float* noise_out, float* wp2, cfr_stream_t stream) {
  if (sigma_len != 1 && sigma_len != 5) { set_error("sigma_len must be 1 or 5"); return 2; }
  return launch_noise_project(z, x, sigma, sigma_len, noise_in, dir_mat, w_avg, psi, seed, sample_offset, b, noise_out,
                              wp2, S(stream));
}
CFR_API int cfr_truncate(const float* w, const float* w_avg, float psi, int b, float* wp2, cfr_stream_t stream) {
  return launch_truncate(w, w_avg, psi, b, wp2, S(stream));
}
CFR_API int cfr_mapping(const float* z, const float* wt, const float* bias, int b, float* w_out, cfr_stream_t stream) {
  if (z == nullptr || wt == nullptr || bias == nullptr || w_out == nullptr) { set_error("cfr_mapping: null pointer"); return 2; }
  return launch_mapping(z, wt, bias, b, w_out, S(stream));
}
CFR_API int cfr_match_vote(const float* emb, int b, const float* gallery, int n, uint64_t* keys, int32_t* pred, int64_t* counts,
                   cfr_stream_t stream) {
  return launch_match_vote(emb, b, gallery, n, reinterpret_cast<unsigned long long*>(keys), pred,
                           reinterpret_cast<long long*>(counts), S(stream));
}

CFR_API int cfr_matcher_create(const float* gallery, int n_gallery, int max_b, cfr_stream_t stream, cfr_matcher** out) {
  if (n_gallery <= 0 || max_b <= 0) { set_error("matcher: bad sizes"); return 2; }
  std::unique_ptr<cfr_matcher> m(new cfr_matcher());
  m->n_gallery = n_gallery;
  m->n_pad = (n_gallery + 255) / 256 * 256;
  m->max_b = max_b;
  CFR_CUDA(cudaMalloc(&m->g_split, sizeof(__half) * 1536 * static_cast<size_t>(m->n_pad)));
  CFR_CUDA(cudaMalloc(&m->g_bias, sizeof(float) * m->n_pad));
  CFR_CUDA(cudaMalloc(&m->q_split, sizeof(__half) * 1536 * static_cast<size_t>(max_b)));
  CFR_CUDA(cudaMalloc(&m->keys, sizeof(unsigned long long) * max_b));
  CFR_CUDA(cudaMemsetAsync(m->keys, 0, sizeof(unsigned long long) * max_b, S(stream)));
  int r = launch_split_hilo(gallery, n_gallery, m->n_pad, 0, m->g_split, m->g_bias, S(stream));
  if (r) return r;
  cfr_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.in = m->q_split; d.N = max_b; d.Hin = 1; d.Win = 1; d.Cin = 1536;
  d.w = m->g_split; d.wRows = m->n_pad; d.Kpad = 1536; d.Cout = m->n_pad;
  d.Hout = 1; d.Wout = 1; d.TW = 1; d.TH = 1; d.TN = 128;
  d.stride = 1; d.ntaps = 1; d.numPhases = 1;
  d.out = m->q_split;            // never written in argmax mode
  d.outIsF32 = 0; d.outH = 1; d.outW = 1; d.outC = m->n_pad; d.oscale = 1;
  d.bias = m->g_bias;
  r = conv_build(d, &m->op);
  if (r) return r;
  conv_set_argmax(&m->op, m->keys);
  *out = m.release();
  return 0;
}
CFR_API void cfr_matcher_destroy(cfr_matcher* m) { delete m; }
CFR_API int cfr_matcher_run(cfr_matcher* m, const float* emb, int b, int32_t* pred, int64_t* counts, cfr_stream_t stream) {
  if (b > m->max_b) { set_error("matcher: b=%d exceeds max_b=%d", b, m->max_b); return 2; }
  if (b <= 0) return 0;
  int r = launch_split_hilo(emb, b, m->max_b, 1, m->q_split, nullptr, S(stream));
  if (r) return r;
  if ((r = conv_launch(m->op, S(stream))) != 0) return r;
  // rows >= b are zero queries: their keys are reset below without being counted
  r = launch_vote_argmax(m->keys, b, m->n_gallery, pred, reinterpret_cast<long long*>(counts), S(stream));
  if (r) return r;
  if (b < m->max_b) CFR_CUDA(cudaMemsetAsync(m->keys + b, 0, sizeof(unsigned long long) * (m->max_b - b), S(stream)));
  return 0;
}

// ---- gallery sharded over ranks (partition C) ----
CFR_API int cfr_match_keys(const float* emb, int b, const float* gallery, int n, uint32_t row_offset, uint64_t* keys,
                           cfr_stream_t stream) {
  if (n <= 0) { set_error("match_keys: empty gallery shard"); return 2; }
  return launch_match_keys(emb, b, gallery, n, row_offset, reinterpret_cast<unsigned long long*>(keys), S(stream));
}
CFR_API int cfr_matcher_keys(cfr_matcher* m, const float* emb, int b, uint32_t row_offset, uint64_t* keys,
                             cfr_stream_t stream) {
  if (b > m->max_b) { set_error("matcher: b=%d exceeds max_b=%d", b, m->max_b); return 2; }
  if (b <= 0) return 0;
  int r = launch_split_hilo(emb, b, m->max_b, 1, m->q_split, nullptr, S(stream));
  if (r) return r;
  if ((r = conv_launch(m->op, S(stream))) != 0) return r;
  r = launch_export_argmax_keys(m->keys, b, row_offset, reinterpret_cast<unsigned long long*>(keys), S(stream));
  if (r) return r;
  if (b < m->max_b) CFR_CUDA(cudaMemsetAsync(m->keys + b, 0, sizeof(unsigned long long) * (m->max_b - b), S(stream)));
  return 0;
}
CFR_API int cfr_vote_keys(const uint64_t* keys, int b, int32_t* pred, int64_t* counts, cfr_stream_t stream) {
  return launch_vote_keys(reinterpret_cast<const unsigned long long*>(keys), b, pred, reinterpret_cast<long long*>(counts),
                          S(stream));
}

CFR_API int cfr_sampler_create(const cfr_sampler_desc* d, cfr_sampler** out) {
  if (d->chunk <= 0 || d->n_gallery <= 0) { set_error("sampler: bad chunk / gallery size"); return 2; }
  std::unique_ptr<cfr_sampler> s(new cfr_sampler());
  s->d = *d;
  const size_t nkeys = static_cast<size_t>(d->chunk) * (d->frm_group > 1 ? d->frm_group : 1);
  CFR_CUDA(cudaMalloc(&s->keys, sizeof(unsigned long long) * nkeys));
  CFR_CUDA(cudaMemset(s->keys, 0xFF, sizeof(unsigned long long) * nkeys));
  CFR_CUDA(cudaMalloc(&s->dev_in, sizeof(float) * 528));
  CFR_CUDA(cudaMalloc(&s->dev_counts, sizeof(long long) * d->n_gallery));
  CFR_CUDA(cudaMallocHost(&s->pin_in, sizeof(float) * 528));
  CFR_CUDA(cudaMallocHost(&s->pin_counts, sizeof(long long) * d->n_gallery));
  if (d->img_frm != nullptr) {
    if (d->img_src == nullptr || d->img_chunk_bytes == 0) { set_error("sampler: img_frm needs img_src and img_chunk_bytes"); return 2; }
    int lo = 0, hi = 0;
    CFR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));            // hi = numerically lowest = highest priority
    // the FRM stream gets the higher priority: its (shorter) kernels take SMs as they free up, synthesis fills the rest
    CFR_CUDA(cudaStreamCreateWithPriority(&s->s2, cudaStreamNonBlocking, hi));
    for (cudaEvent_t* e : {&s->ev_start, &s->ev_synth, &s->ev_copied, &s->ev_done})
      CFR_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    s->overlap = true;
  }
  *out = s.release();
  return 0;
}
CFR_API int cfr_sampler_set_overlap(cfr_sampler* s, int on) {
  if (on && s->s2 == nullptr) { set_error("sampler: created without img_frm, no overlap available"); return 2; }
  if (s->s2 != nullptr) CFR_CUDA(cudaStreamSynchronize(s->s2));
  s->overlap = on != 0;
  s->copied_pending = false;
  if (s->d.tail != nullptr && s->d.tail->s2 != nullptr) return cfr_sampler_set_overlap(s->d.tail, on);
  return 0;
}
CFR_API void cfr_sampler_destroy(cfr_sampler* s) { delete s; }

CFR_API int cfr_sample_votes(cfr_sampler* s, const float* z, const float* x, const float* sigma, int sigma_len,
                     const float* noise_in, int64_t num, uint64_t seed, uint64_t sample_offset, int64_t* counts,
                     int32_t* pred_out, float* emb_out, float* noise_out, cfr_stream_t stream) {
  const cfr_sampler_desc& d = s->d;
  if (sigma_len != 1 && sigma_len != 5) { set_error("sigma_len must be 1 or 5"); return 2; }
  const int K = (d.frm_group > 1 && d.frm_big != nullptr && d.out_slot != nullptr) ? d.frm_group : 1;
  int64_t done = 0;
  if (int r0 = overlap_begin(s, S(stream), num > static_cast<int64_t>(K) * d.chunk)) return r0;
  while (done < num) {
    const int64_t rem = num - done;
    if (rem < d.chunk && d.tail != nullptr) {
      // remainder: hand it to the smallest sampler down the chain whose chunk still holds it
      cfr_sampler* best = nullptr;
      for (cfr_sampler* t = d.tail; t != nullptr; t = t->d.tail)
        if (t->d.chunk >= rem && (best == nullptr || t->d.chunk < best->d.chunk)) best = t;
      if (best != nullptr) {
        if (int rj = overlap_join(s, S(stream))) return rj;
        return cfr_sample_votes(best, z, x, sigma, sigma_len, noise_in ? noise_in + done * 5 : nullptr, rem, seed,
                                sample_offset + done, counts, pred_out ? pred_out + done : nullptr,
                                emb_out ? emb_out + done * 512 : nullptr, noise_out ? noise_out + done * 5 : nullptr, stream);
      }
    }
    // a full group of K chunks goes through the big ArcFace program (better SM fill); the tail chunk by chunk
    const int g = (num - done >= static_cast<int64_t>(K) * d.chunk) ? K : 1;
    const int b = static_cast<int>(num - done < static_cast<int64_t>(g) * d.chunk ? num - done : static_cast<int64_t>(g) * d.chunk);
    if (int rw = synth_may_overwrite(s, S(stream))) return rw;
    for (int k = 0; k < g; ++k) {
      const int64_t off = done + static_cast<int64_t>(k) * d.chunk;
      const int bk = static_cast<int>(num - off < d.chunk ? num - off : d.chunk);
      int r = 0;
      if (d.out_slot != nullptr && (r = launch_set_int(d.out_slot, k, S(stream))) != 0) return r;
      r = launch_noise_project(z, x, sigma, sigma_len, noise_in ? noise_in + off * 5 : nullptr, d.dir_mat, d.w_avg, d.psi,
                               seed, sample_offset + off, bk, noise_out ? noise_out + off * 5 : nullptr, d.wp2, S(stream));
      if (r) return r;
      if ((r = cfr_program_run(d.synth, stream)) != 0) return r;
    }
    cudaStream_t fs = nullptr;
    int r = frm_stream_begin(s, S(stream), g, &fs);
    if (r) return r;
    cfr_stream_t fst = reinterpret_cast<cfr_stream_t>(fs);
    r = cfr_program_run(g == K && K > 1 ? d.frm_big : d.frm, fst);
    if (r) return r;
    const float* emb = (g == K && K > 1) ? d.emb_big : d.emb;
    if (emb_out) {
      CFR_CUDA(cudaMemcpyAsync(emb_out + done * 512, emb, sizeof(float) * 512 * b, cudaMemcpyDeviceToDevice, fs));
    }
    if (d.matcher != nullptr) {
      r = cfr_matcher_run(d.matcher, emb, b, pred_out ? pred_out + done : nullptr, counts, fst);
    } else {
      r = launch_match_vote(emb, b, d.gallery, d.n_gallery, s->keys, pred_out ? pred_out + done : nullptr,
                            reinterpret_cast<long long*>(counts), fs);
    }
    if (r) return r;
    done += b;
  }
  return overlap_join(s, S(stream));
}

CFR_API int cfr_sample_votes_multi(cfr_sampler* s, int n_ids, const float* z, const float* x, const float* sigma,
                           int sigma_len, const int64_t* num_host, uint64_t seed, const uint64_t* sample_offset_host,
                           int64_t* counts, cfr_stream_t stream) {
  if (sigma_len != 1 && sigma_len != 5) { set_error("sigma_len must be 1 or 5"); return 2; }
  if (n_ids < 0) { set_error("sample_votes_multi: n_ids < 0"); return 2; }
  int64_t total = 0;
  for (int g = 0; g < n_ids; ++g) {
    if (num_host[g] < 0) { set_error("sample_votes_multi: negative sample count"); return 2; }
    total += num_host[g];
  }
  int g = 0;                 // current identity and how many of its samples are done
  int64_t g_done = 0;
  int64_t left = total;
  // (the sampler in use runs its FRM side on its own second stream; it is joined before another sampler of the tail chain
  //  takes over -- they may share the matcher's scratch buffers -- and at the end)
  cfr_sampler* last = nullptr;
  while (left > 0) {
    // the sampler whose chunk this run uses: the main one while a whole chunk is left, else the smallest that fits
    cfr_sampler* cur = s;
    if (left < s->d.chunk)
      for (cfr_sampler* t = s->d.tail; t != nullptr; t = t->d.tail)
        if (t->d.chunk >= left && t->d.chunk < cur->d.chunk) cur = t;
    const cfr_sampler_desc& d = cur->d;
    const int cap = d.chunk;
    struct Piece { int id, slot, b; };
    Piece pieces[512];
    int np = 0, slot = 0;
    int r = 0;
    if (cur != last) {
      if (last != nullptr && (r = overlap_join(last, S(stream))) != 0) return r;
      if ((r = overlap_begin(cur, S(stream), left > cur->d.chunk)) != 0) return r;
      last = cur;
    }
    if ((r = synth_may_overwrite(cur, S(stream))) != 0) return r;
    if (d.out_slot != nullptr && (r = launch_set_int(d.out_slot, 0, S(stream))) != 0) return r;
    while (slot < cap && g < n_ids) {
      const int64_t rem = num_host[g] - g_done;
      if (rem == 0) { ++g; g_done = 0; continue; }
      const int b = static_cast<int>(rem < cap - slot ? rem : cap - slot);
      r = launch_noise_project(z + static_cast<size_t>(g) * 512, x + static_cast<size_t>(g) * 5, sigma, sigma_len, nullptr,
                               d.dir_mat, d.w_avg, d.psi, seed, sample_offset_host[g] + g_done, b, nullptr,
                               d.wp2 + static_cast<size_t>(slot) * 2 * 512, S(stream));
      if (r) return r;
      if (np == 512) { set_error("sample_votes_multi: more than 512 identities in one chunk"); return 2; }
      pieces[np++] = Piece{g, slot, b};
      slot += b;
      g_done += b;
    }
    if ((r = cfr_program_run(d.synth, stream)) != 0) return r;
    cudaStream_t fs = nullptr;
    if ((r = frm_stream_begin(cur, S(stream), 1, &fs)) != 0) return r;
    cfr_stream_t fst = reinterpret_cast<cfr_stream_t>(fs);
    if ((r = cfr_program_run(d.frm, fst)) != 0) return r;
    for (int i = 0; i < np; ++i) {
      const float* emb = d.emb + static_cast<size_t>(pieces[i].slot) * 512;
      int64_t* cnt = counts + static_cast<size_t>(pieces[i].id) * d.n_gallery;
      if (d.matcher != nullptr) r = cfr_matcher_run(d.matcher, emb, pieces[i].b, nullptr, cnt, fst);
      else r = launch_match_vote(emb, pieces[i].b, d.gallery, d.n_gallery, cur->keys, nullptr,
                                 reinterpret_cast<long long*>(cnt), fs);
      if (r) return r;
    }
    left -= slot;
  }
  return last != nullptr ? overlap_join(last, S(stream)) : 0;
}

CFR_API int cfr_sample_votes_host(cfr_sampler* s, const float* z_host, const float* x_host, const float* sigma_host,
                          int sigma_len, int64_t num, uint64_t seed, uint64_t sample_offset, int64_t* counts_host,
                          cfr_stream_t stream) {
  if (sigma_len != 1 && sigma_len != 5) { set_error("sigma_len must be 1 or 5"); return 2; }
  const int n = s->d.n_gallery;
  memcpy(s->pin_in, z_host, sizeof(float) * 512);
  memcpy(s->pin_in + 512, x_host, sizeof(float) * 5);
  memcpy(s->pin_in + 520, sigma_host, sizeof(float) * sigma_len);
  CFR_CUDA(cudaMemcpyAsync(s->dev_in, s->pin_in, sizeof(float) * 528, cudaMemcpyHostToDevice, S(stream)));
  CFR_CUDA(cudaMemsetAsync(s->dev_counts, 0, sizeof(long long) * n, S(stream)));
  int r = cfr_sample_votes(s, s->dev_in, s->dev_in + 512, s->dev_in + 520, sigma_len, nullptr, num, seed,
                           sample_offset, reinterpret_cast<int64_t*>(s->dev_counts), nullptr, nullptr, nullptr, stream);
  if (r) return r;
  CFR_CUDA(cudaMemcpyAsync(s->pin_counts, s->dev_counts, sizeof(long long) * n, cudaMemcpyDeviceToHost, S(stream)));
  CFR_CUDA(cudaStreamSynchronize(S(stream)));
  memcpy(counts_host, s->pin_counts, sizeof(long long) * n);
  return 0;
}

}  // extern "C"
