// bfl_abi.cu -- the C ABI of libbfl.so (see include/bfl.h): context / plan management,
// host<->device staging and kernel orchestration.  No CPU fallback: without a CUDA device
// every compute entry point fails with BFL_ERR_NODEVICE.
#include "../../include/bfl.h"
#include "bfl_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>

#include <cufft.h>

using namespace bfl;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      char buf__[512];                                                                             \
      snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
               __FILE__, __LINE__);                                                                \
      return fail(e__ == cudaErrorMemoryAllocation ? BFL_ERR_OOM : BFL_ERR_CUDA, buf__);           \
    }                                                                                              \
  } while (0)

struct bfl_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  LaunchStats stats;
  bool timing = false;
  void* counters_zeroed = nullptr;                          // the tile-counter allocation that has been cleared
  std::vector<double> sg_cached;                            // Savitzky-Golay taps currently on the device ...
  void* sg_cached_dev = nullptr;                            // ... and the buffer that holds them
  cudaStream_t s_in = nullptr, s_out = nullptr;           // copy streams of the pipelined host entry point
  std::vector<cudaEvent_t> pipe_ev;                        // cached events for it
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev;   // pending (start, stop) pairs of the window kernel
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tfree; // recycled event pairs
  std::map<std::string, std::pair<void*, size_t>> pool;   // grow-only named device buffers
  std::map<std::pair<int64_t, int>, cufftHandle> fft;     // cached batched D2Z plans keyed by (N, batch)

  int get(const char* name, size_t bytes, void** out) {
    auto& e = pool[name];
    if (e.second < bytes) {
      if (e.first) {
        cudaStreamSynchronize(stream);
        cudaFree(e.first);
        e.first = nullptr;
        e.second = 0;
      }
      size_t want = bytes + bytes / 4 + 256;
      cudaError_t err = cudaMalloc(&e.first, want);
      if (err != cudaSuccess) {
        e.first = nullptr;
        return fail(BFL_ERR_OOM, std::string("cudaMalloc failed for ") + name + ": " + cudaGetErrorString(err));
      }
      e.second = want;
    }
    *out = e.first;
    return BFL_OK;
  }
};

struct bfl_plan {
  bfl_ctx* ctx = nullptr;
  PlanView pv{};
  std::vector<int> ngrid;                 // per hypothesis, uncompacted
  std::vector<void*> allocs;
  std::vector<int> h_gorig;
};

extern "C" {

int bfl_version(void) { return BFL_VERSION; }
const char* bfl_last_error(void) { return g_err.c_str(); }

int bfl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int bfl_ctx_create(int device, void* stream, bfl_ctx** out) {
  if (!out) return fail(BFL_ERR_ARG, "ctx output pointer is null");
  *out = nullptr;
  if (bfl_device_count() <= 0)
    return fail(BFL_ERR_NODEVICE, "no CUDA device available: libbfl has no CPU fallback");
  if (device < 0) CK(cudaGetDevice(&device));
  CK(cudaSetDevice(device));
  bfl_ctx* c = new (std::nothrow) bfl_ctx();
  if (!c) return fail(BFL_ERR_OOM, "out of host memory");
  c->device = device;
  CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return fail(BFL_ERR_CUDA, cudaGetErrorString(e)); }
    c->own_stream = true;
  }
  *out = c;
  return BFL_OK;
}

int bfl_ctx_destroy(bfl_ctx* c) {
  if (!c) return BFL_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->pool)
    if (kv.second.first) cudaFree(kv.second.first);
  for (auto& kv : c->fft) cufftDestroy(kv.second);
  for (auto ev : c->pipe_ev) cudaEventDestroy(ev);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  for (auto& pr : c->tev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& pr : c->tfree) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return BFL_OK;
}

int bfl_ctx_sync(bfl_ctx* c) {
  if (!c) return fail(BFL_ERR_ARG, "null context");
  CK(cudaStreamSynchronize(c->stream));
  return BFL_OK;
}

int64_t bfl_ctx_launch_count(const bfl_ctx* c) { return c ? c->stats.launches : 0; }

int bfl_ctx_set_timing(bfl_ctx* c, int enable) {
  if (!c) return fail(BFL_ERR_ARG, "null context");
  c->timing = enable != 0;
  return BFL_OK;
}

int bfl_ctx_window_time(bfl_ctx* c, double* total_ms, int64_t* launches) {
  if (!c || !total_ms || !launches) return fail(BFL_ERR_ARG, "null argument to bfl_ctx_window_time");
  CK(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (auto& pr : c->tev) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
    tot += ms;
    c->tfree.push_back(pr);
  }
  *total_ms = tot;
  *launches = (int64_t)c->tev.size();
  c->tev.clear();
  return BFL_OK;
}

static int timing_begin(bfl_ctx* c, std::pair<cudaEvent_t, cudaEvent_t>& pr) {
  if (!c->tfree.empty()) { pr = c->tfree.back(); c->tfree.pop_back(); }
  else { CK(cudaEventCreate(&pr.first)); CK(cudaEventCreate(&pr.second)); }
  CK(cudaEventRecord(pr.first, c->stream));
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
static int plan_alloc(bfl_plan* p, size_t bytes, void** out) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes ? bytes : 8);
  if (e != cudaSuccess) return fail(BFL_ERR_OOM, std::string("cudaMalloc(plan): ") + cudaGetErrorString(e));
  p->allocs.push_back(d);
  *out = d;
  return BFL_OK;
}

int bfl_plan_destroy(bfl_plan* p) {
  if (!p) return BFL_OK;
  if (p->ctx) {
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
  }
  for (void* d : p->allocs) cudaFree(d);
  delete p;
  return BFL_OK;
}

int bfl_plan_create(bfl_ctx* c, int W, int P, const double* bgmodels, const double* mts,
                    const bfl_hypothesis* hyps, int nhyp, bfl_plan** out) {
  if (!c || !out || !bgmodels || !mts) return fail(BFL_ERR_ARG, "null argument to bfl_plan_create");
  *out = nullptr;
  if (W < 3 || W % 2 == 0)   // bayes.py:229-233
    return fail(BFL_ERR_ARG, "Error... Background length (bglen) must be an odd number");
  if (W > MAXW) return fail(BFL_ERR_UNSUPPORTED, "bglen > 255 is not supported");
  if (P < 1 || P > MAXP) return fail(BFL_ERR_UNSUPPORTED, "bgorder must be in 0..5");
  if (nhyp < 0 || nhyp > MAXH) return fail(BFL_ERR_UNSUPPORTED, "too many hypotheses in one plan (max 8)");
  if (nhyp > 0 && !hyps) return fail(BFL_ERR_ARG, "null hypothesis array");
  CK(cudaSetDevice(c->device));

  bfl_plan* p = new (std::nothrow) bfl_plan();
  if (!p) return fail(BFL_ERR_OOM, "out of host memory");
  p->ctx = c;
  const int WP = W + 1;
  PlanView& pv = p->pv;
  pv.W = W; pv.WP = WP; pv.P = P; pv.nhyp = nhyp;

  std::vector<TemplDesc> desc;
  std::vector<double> lp, lw, table_rows;
  std::vector<int> gorig, table_at;
  pv.hyp_begin[0] = 0;
  for (int h = 0; h < nhyp; h++) {
    const bfl_hypothesis& hy = hyps[h];
    if (hy.ngrid < 0 || !hy.logprior || !hy.logweight || (hy.kind != BFL_MODEL_TABLE && !hy.params) ||
        (hy.kind == BFL_MODEL_TABLE && !hy.table) || hy.kind < 0 || hy.kind > BFL_MODEL_TABLE) {
      delete p;
      return fail(BFL_ERR_ARG, "malformed hypothesis descriptor");
    }
    p->ngrid.push_back(hy.ngrid);
    pv.hyp_half[h] = hy.halfrange ? 1 : 0;
    for (int g = 0; g < hy.ngrid; g++) {
      const double pr = hy.logprior[g];
      if (!(pr > -INFINITY) || std::isnan(pr)) continue;   // masked grid point: row stays -inf (bayes.py:358-363)
      TemplDesc d{};
      d.kind = hy.kind; d.reverse = hy.reverse;
      for (int i = 0; i < 4; i++) d.p[i] = hy.params ? hy.params[(size_t)g * 4 + i] : 0.0;
      if (hy.kind == BFL_MODEL_TABLE) {
        table_at.push_back((int)desc.size());
        table_rows.insert(table_rows.end(), hy.table + (size_t)g * W, hy.table + (size_t)(g + 1) * W);
      }
      desc.push_back(d);
      lp.push_back(pr);
      lw.push_back(hy.logweight[g]);
      gorig.push_back(g);
    }
    pv.hyp_begin[h + 1] = (int)desc.size();
  }
  const int G = (int)desc.size();
  pv.G = G;
  p->h_gorig = gorig;

  // background basis, its pair products, and an orthonormalised copy (modified Gram-Schmidt
  // with re-orthogonalisation in extended precision; log det G = 2 sum log r_jj)
  std::vector<double> bg((size_t)P * WP, 0.0), bb((size_t)(P * (P + 1) / 2) * WP, 0.0), qb((size_t)P * WP, 0.0);
  for (int j = 0; j < P; j++)
    for (int w = 0; w < W; w++) bg[(size_t)j * WP + w] = bgmodels[(size_t)j * W + w];
  {
    int pq = 0;
    for (int a = 0; a < P; a++)
      for (int b = a; b < P; b++, pq++)
        for (int w = 0; w < W; w++) bb[(size_t)pq * WP + w] = bgmodels[(size_t)a * W + w] * bgmodels[(size_t)b * W + w];
  }
  long double logdet = 0.0L;
  {
    std::vector<std::vector<long double>> q(P, std::vector<long double>(W));
    for (int j = 0; j < P; j++) {
      std::vector<long double> v(W);
      for (int w = 0; w < W; w++) v[w] = (long double)bgmodels[(size_t)j * W + w];
      for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < j; i++) {
          long double r = 0.0L;
          for (int w = 0; w < W; w++) r += q[i][w] * v[w];
          for (int w = 0; w < W; w++) v[w] -= r * q[i][w];
        }
      long double nrm = 0.0L;
      for (int w = 0; w < W; w++) nrm += v[w] * v[w];
      nrm = sqrtl(nrm);
      if (!(nrm > 0.0L)) { delete p; return fail(BFL_ERR_ARG, "background basis is rank deficient"); }
      for (int w = 0; w < W; w++) q[j][w] = v[w] / nrm;
      logdet += 2.0L * logl(nrm);
      for (int w = 0; w < W; w++) qb[(size_t)j * WP + w] = (double)q[j][w];
    }
  }
  pv.logdetG = (double)logdet;

  // ---- separable-grid form ------------------------------------------------------------------
  // A Flare template is rise(tau_gauss) (+) centre sample (+) decay(tau_exp) on disjoint supports, so
  // m(tg, te) = m(tg, 0) + m(0, te) - m(0, 0) exactly, and the same holds after the (linear) projection
  // off the background basis.  For a grid of nA x nB time scales only nA + nB + 1 correlations per
  // sample are needed instead of nA * nB: shapes A_i = m_perp(tg_i, 0) - m_perp(0, 0), B_j = m_perp(0, te_j),
  // z_g = (A_i . d + B_j . d) / |m_perp_g|.  Other hypotheses stay dense (one shape per template).
  std::vector<TemplDesc> sdesc;            // shapes that must be generated (separable hypotheses)
  std::vector<int> sdesc_at;               // their row in the shape table
  std::vector<int> dense_src, dense_at;    // dense templates: row of `hat` -> row in the shape table
  std::vector<short> idxA(G, -1), idxB(G, 0);
  struct SubC { int a0, na, c; };
  std::vector<SubC> sub_c;                 // per separable hypothesis: first A row, number of A rows, C row
  int S = 0;
  bool sep_ok = split_window_supported(W);
  for (int h = 0; h < nhyp && sep_ok; h++) {
    const int g0 = pv.hyp_begin[h], g1 = pv.hyp_begin[h + 1];
    pv.hyp_A0[h] = 0; pv.hyp_nA[h] = 0;
    bool sep = hyps[h].kind == BFL_MODEL_FLARE && g1 - g0 >= 4;
    std::vector<double> tgs, tes;
    if (sep) {
      for (int g = g0; g < g1 && sep; g++) {
        if (desc[g].p[0] != desc[g0].p[0] || desc[g].p[3] != desc[g0].p[3]) sep = false;   // one t0, one amp
        if (std::find(tgs.begin(), tgs.end(), desc[g].p[2]) == tgs.end()) tgs.push_back(desc[g].p[2]);
        if (std::find(tes.begin(), tes.end(), desc[g].p[1]) == tes.end()) tes.push_back(desc[g].p[1]);
      }
      if ((int)tgs.size() > 1024 || (int)(tgs.size() + tes.size() + 1) >= g1 - g0) sep = false;
    }
    if (sep) {
      const int A0 = S, B0 = S + (int)tgs.size(), C = B0 + (int)tes.size();
      pv.hyp_A0[h] = A0; pv.hyp_nA[h] = (int)tgs.size();
      pv.hyp_B0[h] = B0; pv.hyp_nB[h] = (int)tes.size();
      for (size_t i = 0; i < tgs.size(); i++) {
        TemplDesc d = desc[g0]; d.p[1] = 0.0; d.p[2] = tgs[i];
        sdesc.push_back(d); sdesc_at.push_back(A0 + (int)i);
      }
      for (size_t j = 0; j < tes.size(); j++) {
        TemplDesc d = desc[g0]; d.p[1] = tes[j]; d.p[2] = 0.0;
        sdesc.push_back(d); sdesc_at.push_back(B0 + (int)j);
      }
      { TemplDesc d = desc[g0]; d.p[1] = 0.0; d.p[2] = 0.0; sdesc.push_back(d); sdesc_at.push_back(C); }
      for (int g = g0; g < g1; g++) {
        idxA[g] = (short)(std::find(tgs.begin(), tgs.end(), desc[g].p[2]) - tgs.begin());
        idxB[g] = (short)(B0 + (std::find(tes.begin(), tes.end(), desc[g].p[1]) - tes.begin()));
      }
      sub_c.push_back({A0, (int)tgs.size(), C});
      S = C + 1;
    } else {
      pv.hyp_B0[h] = S; pv.hyp_nB[h] = g1 - g0;
      for (int g = g0; g < g1; g++) { dense_src.push_back(g); dense_at.push_back(S); idxB[g] = (short)S; S++; }
    }
    if (S > 30000) sep_ok = false;
  }
  pv.S = S;

  void *d_desc, *d_mts, *d_raw, *d_hat, *d_Kg, *d_lp, *d_lw, *d_gorig, *d_bg, *d_bb, *d_qb;
  int rc;
#define PA(ptr, bytes) if ((rc = plan_alloc(p, (bytes), &(ptr))) != BFL_OK) { bfl_plan_destroy(p); return rc; }
  PA(d_desc, sizeof(TemplDesc) * (size_t)std::max(G, 1));
  PA(d_mts, sizeof(double) * (size_t)W);
  PA(d_raw, sizeof(double) * (size_t)std::max(G, 1) * WP);
  PA(d_hat, sizeof(double) * (size_t)std::max(G, 1) * WP);
  PA(d_Kg, sizeof(double) * (size_t)std::max(G, 1));
  PA(d_lp, sizeof(double) * (size_t)std::max(G, 1));
  PA(d_lw, sizeof(double) * (size_t)std::max(G, 1));
  PA(d_gorig, sizeof(int) * (size_t)std::max(G, 1));
  PA(d_bg, sizeof(double) * bg.size());
  PA(d_bb, sizeof(double) * bb.size());
  PA(d_qb, sizeof(double) * qb.size());
#undef PA
  cudaStream_t s = c->stream;
#define CKP(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { bfl_plan_destroy(p); \
    return fail(BFL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  if (G > 0) {
    CKP(cudaMemcpyAsync(d_desc, desc.data(), sizeof(TemplDesc) * (size_t)G, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_lp, lp.data(), sizeof(double) * (size_t)G, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_lw, lw.data(), sizeof(double) * (size_t)G, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_gorig, gorig.data(), sizeof(int) * (size_t)G, cudaMemcpyHostToDevice, s));
  }
  CKP(cudaMemcpyAsync(d_mts, mts, sizeof(double) * (size_t)W, cudaMemcpyHostToDevice, s));
  CKP(cudaMemcpyAsync(d_bg, bg.data(), sizeof(double) * bg.size(), cudaMemcpyHostToDevice, s));
  CKP(cudaMemcpyAsync(d_bb, bb.data(), sizeof(double) * bb.size(), cudaMemcpyHostToDevice, s));
  CKP(cudaMemcpyAsync(d_qb, qb.data(), sizeof(double) * qb.size(), cudaMemcpyHostToDevice, s));
  CKP(cudaMemsetAsync(d_raw, 0, sizeof(double) * (size_t)std::max(G, 1) * WP, s));
  CKP(launch_gen_templates((const TemplDesc*)d_desc, G, (const double*)d_mts, W, WP, (double*)d_raw, s, &c->stats));
  for (size_t i = 0; i < table_at.size(); i++)
    CKP(cudaMemcpyAsync((double*)d_raw + (size_t)table_at[i] * WP, table_rows.data() + i * W, sizeof(double) * (size_t)W,
                        cudaMemcpyHostToDevice, s));
  void *d_inorm = nullptr, *d_shapes = nullptr, *d_idxA = nullptr, *d_idxB = nullptr;
  void *d_rshapes = nullptr, *d_shlo = nullptr, *d_shhi = nullptr, *d_shg0 = nullptr, *d_shg1 = nullptr;
  bool wsep_ok = false;
  if ((rc = plan_alloc(p, sizeof(double) * (size_t)std::max(G, 1), &d_inorm))) { bfl_plan_destroy(p); return rc; }
  CKP(launch_plan_orthonormalise((const double*)d_raw, (const double*)d_qb, G, W, WP, P, (double*)d_hat,
                                 (double*)d_Kg, (double*)d_inorm, 1, s, &c->stats));
  if (sep_ok && G > 0) {
    void *d_sdesc, *d_sraw, *d_sperp;
    const int ns = (int)sdesc.size();
    if ((rc = plan_alloc(p, sizeof(double) * (size_t)S * WP, &d_shapes)) ||
        (rc = plan_alloc(p, sizeof(short) * (size_t)G, &d_idxA)) || (rc = plan_alloc(p, sizeof(short) * (size_t)G, &d_idxB)) ||
        (rc = plan_alloc(p, sizeof(TemplDesc) * (size_t)std::max(ns, 1), &d_sdesc)) ||
        (rc = plan_alloc(p, sizeof(double) * (size_t)std::max(ns, 1) * WP, &d_sraw)) ||
        (rc = plan_alloc(p, sizeof(double) * (size_t)std::max(ns, 1) * WP, &d_sperp))) { bfl_plan_destroy(p); return rc; }
    CKP(cudaMemcpyAsync(d_idxA, idxA.data(), sizeof(short) * (size_t)G, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_idxB, idxB.data(), sizeof(short) * (size_t)G, cudaMemcpyHostToDevice, s));
    if (ns > 0) {
      CKP(cudaMemcpyAsync(d_sdesc, sdesc.data(), sizeof(TemplDesc) * (size_t)ns, cudaMemcpyHostToDevice, s));
      CKP(launch_gen_templates((const TemplDesc*)d_sdesc, ns, (const double*)d_mts, W, WP, (double*)d_sraw, s, &c->stats));
      CKP(launch_plan_orthonormalise((const double*)d_sraw, (const double*)d_qb, ns, W, WP, P, (double*)d_sperp, nullptr,
                                     nullptr, 0, s, &c->stats));
      for (int i = 0; i < ns; i++)
        CKP(cudaMemcpyAsync((double*)d_shapes + (size_t)sdesc_at[i] * WP, (double*)d_sperp + (size_t)i * WP,
                            sizeof(double) * (size_t)WP, cudaMemcpyDeviceToDevice, s));
      for (auto& ac : sub_c) CKP(launch_subtract_row((double*)d_shapes, ac.a0, ac.na, ac.c, WP, s, &c->stats));
    }
    for (size_t i = 0; i < dense_src.size(); i++)
      CKP(cudaMemcpyAsync((double*)d_shapes + (size_t)dense_at[i] * WP, (double*)d_hat + (size_t)dense_src[i] * WP,
                          sizeof(double) * (size_t)WP, cudaMemcpyDeviceToDevice, s));
    // raw-template shapes for the weighted path: A_i = m(tg_i, 0) - m(0, 0) (rise only), B_j = m(0, te_j)
    if ((rc = plan_alloc(p, sizeof(double) * (size_t)S * WP, &d_rshapes)) ||
        (rc = plan_alloc(p, sizeof(short) * (size_t)S, &d_shlo)) || (rc = plan_alloc(p, sizeof(short) * (size_t)S, &d_shhi)) ||
        (rc = plan_alloc(p, sizeof(int) * (size_t)S, &d_shg0)) || (rc = plan_alloc(p, sizeof(int) * (size_t)S, &d_shg1))) {
      bfl_plan_destroy(p);
      return rc;
    }
    CKP(cudaMemsetAsync(d_rshapes, 0, sizeof(double) * (size_t)S * WP, s));
    for (int i = 0; i < ns; i++)
      CKP(cudaMemcpyAsync((double*)d_rshapes + (size_t)sdesc_at[i] * WP, (double*)d_sraw + (size_t)i * WP,
                          sizeof(double) * (size_t)WP, cudaMemcpyDeviceToDevice, s));
    for (auto& ac : sub_c) CKP(launch_subtract_row((double*)d_rshapes, ac.a0, ac.na, ac.c, WP, s, &c->stats));
    for (size_t i = 0; i < dense_src.size(); i++)
      CKP(cudaMemcpyAsync((double*)d_rshapes + (size_t)dense_at[i] * WP, (double*)d_raw + (size_t)dense_src[i] * WP,
                          sizeof(double) * (size_t)WP, cudaMemcpyDeviceToDevice, s));
    std::vector<double> hs((size_t)S * WP);
    CKP(cudaMemcpyAsync(hs.data(), d_rshapes, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, s));
    CKP(cudaStreamSynchronize(s));
    std::vector<short> lo(S, 0), hi(S, 0);
    for (int r = 0; r < S; r++) {
      int a = W, b = 0;
      for (int w = 0; w < W; w++)
        if (hs[(size_t)r * WP + w] != 0.0) { a = std::min(a, w); b = std::max(b, w + 1); }   // NaN counts as non-zero
      if (b <= a) { a = WP; b = 0; }   // empty shape (tau_gauss = 0 rise): contributes nothing to a group's span
      lo[r] = (short)(a & ~1); hi[r] = (short)b;
    }
    // the weighted norm separates only if rise and decay never overlap; templates must be grouped by B shape
    wsep_ok = (WP % 2) == 0;
    for (int h = 0; h < nhyp && wsep_ok; h++) {
      for (int g = pv.hyp_begin[h] + 1; g < pv.hyp_begin[h + 1]; g++)
        if (idxB[g] < idxB[g - 1]) wsep_ok = false;
      for (int ia = 0; ia < pv.hyp_nA[h] && wsep_ok; ia++)
        for (int ib = 0; ib < pv.hyp_nB[h] && wsep_ok; ib++)
          for (int w = 0; w < W; w++)
            if (hs[(size_t)(pv.hyp_A0[h] + ia) * WP + w] != 0.0 && hs[(size_t)(pv.hyp_B0[h] + ib) * WP + w] != 0.0) {
              wsep_ok = false;
              break;
            }
    }
    // run of templates per B shape (templates of a hypothesis are grouped by B shape, checked above)
    std::vector<int> sg0(S, 0), sg1(S, 0);
    for (int g = G - 1; g >= 0; g--) sg0[idxB[g]] = g;
    for (int g = 0; g < G; g++) sg1[idxB[g]] = g + 1;
    for (int r = 0; r < S && wsep_ok; r++)
      for (int g = sg0[r]; g < sg1[r]; g++)
        if (idxB[g] != r) wsep_ok = false;
    // split weighted form (bfl_wsplit.cu): every window sum of the weighted path is a correlation of one of the two
    // streams 1/v, d/v with a FIXED vector -- q_p q_q and m_s^2, m_s q_p on 1/v; q_p and m_s on d/v.  Vectors are
    // packed by support class (whole window / rise half / decay half) so the kernel's tap loops are compile-time.
    if (wsep_ok && wsplit_window_supported(W)) {
      const int H = W / 2, T0H = H & ~1, NTL = H + 1, NTLP = (NTL + 1) & ~1, NTHP = (W - T0H + 1) & ~1;
      std::vector<int> shrow(S, -1), cls(S, 0);
      int rows = P * (P + 1) / 2 + P;
      std::vector<char> used(S, 0);                   // the centre row of a separable hypothesis is folded into its A rows
      for (int h = 0; h < nhyp; h++) {
        for (int i = 0; i < pv.hyp_nA[h]; i++) used[pv.hyp_A0[h] + i] = 1;
        for (int i = 0; i < pv.hyp_nB[h]; i++) used[pv.hyp_B0[h] + i] = 1;
      }
      for (int r = 0; r < S; r++) {
        if (hi[r] <= lo[r] || !used[r]) continue;     // empty shape: no rows, its sums are zero
        shrow[r] = rows; rows += P + 2;
        cls[r] = (hi[r] <= NTL) ? 1 : (lo[r] >= T0H ? 2 : 0);
      }
      std::vector<double> tab;
      std::vector<int> joff, jrow;
      int njc[2][3] = {{0, 0, 0}, {0, 0, 0}};
      auto push = [&](int cl, int row, auto&& val) {     // val(w): the vector's value at tap w
        const int t0 = cl == 2 ? T0H : 0, nt = cl == 0 ? W : (cl == 1 ? NTL : W - T0H), np = cl == 0 ? WP : (cl == 1 ? NTLP : NTHP);
        joff.push_back((int)tab.size()); jrow.push_back(row);
        for (int w = 0; w < np; w++) tab.push_back(w < nt ? val(t0 + w) : 0.0);
      };
      for (int st = 0; st < 2; st++)
        for (int cl = 0; cl < 3; cl++) {
          const size_t before = joff.size();
          if (cl == 0) {
            if (st == 0) {
              int row = 0;
              for (int a = 0; a < P; a++)
                for (int b = a; b < P; b++, row++)
                  push(0, row, [&](int w) { return qb[(size_t)a * WP + w] * qb[(size_t)b * WP + w]; });
            } else {
              for (int a = 0; a < P; a++) push(0, P * (P + 1) / 2 + a, [&](int w) { return qb[(size_t)a * WP + w]; });
            }
          }
          for (int r = 0; r < S; r++) {
            if (shrow[r] < 0 || cls[r] != cl) continue;
            const double* m = hs.data() + (size_t)r * WP;
            if (st == 0) {
              for (int a = 0; a < P; a++) push(cl, shrow[r] + a, [&](int w) { return m[w] * qb[(size_t)a * WP + w]; });
              push(cl, shrow[r] + P, [&](int w) { return m[w] * m[w]; });
            } else {
              push(cl, shrow[r] + P + 1, [&](int w) { return m[w]; });
            }
          }
          njc[st][cl] = (int)(joff.size() - before);
        }
      void *d_wtab, *d_woff, *d_wrow, *d_wshrow;
      if ((rc = plan_alloc(p, sizeof(double) * tab.size(), &d_wtab)) || (rc = plan_alloc(p, sizeof(int) * joff.size(), &d_woff)) ||
          (rc = plan_alloc(p, sizeof(int) * jrow.size(), &d_wrow)) || (rc = plan_alloc(p, sizeof(int) * (size_t)S, &d_wshrow))) {
        bfl_plan_destroy(p);
        return rc;
      }
      CKP(cudaMemcpyAsync(d_wtab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, s));
      CKP(cudaMemcpyAsync(d_woff, joff.data(), sizeof(int) * joff.size(), cudaMemcpyHostToDevice, s));
      CKP(cudaMemcpyAsync(d_wrow, jrow.data(), sizeof(int) * jrow.size(), cudaMemcpyHostToDevice, s));
      CKP(cudaMemcpyAsync(d_wshrow, shrow.data(), sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, s));
      CKP(cudaStreamSynchronize(s));
      pv.wc_tab = (const double*)d_wtab; pv.wc_off = (const int*)d_woff; pv.wc_row = (const int*)d_wrow;
      pv.wc_shrow = (const int*)d_wshrow;
      pv.wc_tab_len = (int)tab.size(); pv.wc_nj = (int)joff.size(); pv.wc_rows = rows;
      for (int st = 0; st < 2; st++)
        for (int cl = 0; cl < 3; cl++) pv.wc_n[st][cl] = njc[st][cl];
    }
    CKP(cudaMemcpyAsync(d_shg0, sg0.data(), sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_shg1, sg1.data(), sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_shlo, lo.data(), sizeof(short) * (size_t)S, cudaMemcpyHostToDevice, s));
    CKP(cudaMemcpyAsync(d_shhi, hi.data(), sizeof(short) * (size_t)S, cudaMemcpyHostToDevice, s));
    CKP(cudaStreamSynchronize(s));
  }
  CKP(cudaStreamSynchronize(s));   // host vectors above go out of scope
#undef CKP
  pv.sep_ok = (sep_ok && G > 0) ? 1 : 0;
  pv.shapes = (const double*)d_shapes; pv.inorm = (const double*)d_inorm;
  pv.idxA = (const short*)d_idxA; pv.idxB = (const short*)d_idxB;
  pv.wsep_ok = (pv.sep_ok && wsep_ok) ? 1 : 0;
  pv.rshapes = (const double*)d_rshapes; pv.sh_lo = (const short*)d_shlo; pv.sh_hi = (const short*)d_shhi;
  pv.sh_g0 = (const int*)d_shg0; pv.sh_g1 = (const int*)d_shg1;
  pv.raw = (const double*)d_raw; pv.hat = (const double*)d_hat; pv.Kg = (const double*)d_Kg;
  pv.lp = (const double*)d_lp; pv.lw = (const double*)d_lw; pv.gorig = (const int*)d_gorig;
  pv.bg = (const double*)d_bg; pv.bb = (const double*)d_bb; pv.qb = (const double*)d_qb;
  pv.epi2_block[0] = pv.epi2_block[1] = nullptr;
  if (epi2_fits(pv)) {
    for (int np = 0; np < 2; np++) {
      void* blk = nullptr;
      if ((rc = plan_alloc(p, epi2_plan_block_bytes(pv), &blk))) { bfl_plan_destroy(p); return rc; }
      cudaError_t e = launch_epilogue2_prepare(pv, np, blk, s, &c->stats);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { bfl_plan_destroy(p); return fail(BFL_ERR_CUDA, std::string("k_epilogue2_prepare: ") + cudaGetErrorString(e)); }
      pv.epi2_block[np] = blk;
    }
  }
  *out = p;
  return BFL_OK;
}

int bfl_plan_shape_count(const bfl_plan* p) { return (p && p->pv.sep_ok) ? p->pv.S : -1; }

int bfl_plan_templates(bfl_plan* p, int ihyp, double* out) {
  if (!p || !out || ihyp < 0 || ihyp >= p->pv.nhyp) return fail(BFL_ERR_ARG, "bad argument to bfl_plan_templates");
  const PlanView& pv = p->pv;
  CK(cudaSetDevice(p->ctx->device));
  const int g0 = pv.hyp_begin[ihyp], g1 = pv.hyp_begin[ihyp + 1];
  std::vector<double> tmp((size_t)std::max(g1 - g0, 1) * pv.WP);
  std::memset(out, 0, sizeof(double) * (size_t)p->ngrid[ihyp] * pv.W);
  if (g1 > g0) {
    CK(cudaMemcpyAsync(tmp.data(), pv.raw + (size_t)g0 * pv.WP, sizeof(double) * (size_t)(g1 - g0) * pv.WP,
                       cudaMemcpyDeviceToHost, p->ctx->stream));
    CK(cudaStreamSynchronize(p->ctx->stream));
    for (int g = g0; g < g1; g++)
      std::memcpy(out + (size_t)p->h_gorig[g] * pv.W, tmp.data() + (size_t)(g - g0) * pv.WP, sizeof(double) * (size_t)pv.W);
  }
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
static int setup_scratch(bfl_ctx* c, const PlanView& pv, const SeriesView& sv, bool need_general, bool need_bg,
                         bool need_part, Scratch& sc, bool one_chunk = false) {
  const int64_t nk = sv.k1 - sv.k0;
  int KT = 1;
  sc.gchunk = pick_gchunk(std::max(pv.G, 1), pv.W, (int64_t)sv.B * nk, c->sm_count, &KT);
  // separable form: whole plan in one CTA -- only when there are enough warp tiles to fill the chip;
  // small problems keep the template-chunked dense form, which spreads one curve over all SMs
  if (sv.const_var && fast_sep_fits(pv, KT) && (int64_t)sv.B * nk >= (int64_t)32 * KT * 2 * c->sm_count)
    sc.gchunk = std::max(pv.G, 1);
  if (one_chunk) sc.gchunk = std::max(pv.G, 1);
  sc.KT = KT;
  sc.NC = (std::max(pv.G, 1) + sc.gchunk - 1) / sc.gchunk;
  sc.Y = nullptr;
  if (sv.const_var && sc.NC == 1 && KT <= 4 && sc.gchunk == std::max(pv.G, 1) && fast_split_fits(pv) &&
      (int64_t)sv.B * nk >= (int64_t)32 * KT * 2 * c->sm_count) {
    void* yp;
    int rcy = c->get("Y", sizeof(double) * (size_t)pv.S * sv.B * nk, &yp);
    if (rcy) return rcy;
    sc.Y = (double*)yp;       // large batch: split form (correlation kernel + epilogue kernel)
  }
  sc.sub = 1;
  int rc;
  void* ptr;
  if ((rc = c->get("sqrtu", sizeof(double) * (size_t)sv.B, &ptr))) return rc; sc.sqrtu = (double*)ptr;
  if ((rc = c->get("halflogv", sizeof(double) * (size_t)sv.B, &ptr))) return rc; sc.halflogv = (double*)ptr;
  if ((rc = c->get("uconst", sizeof(double) * (size_t)sv.B, &ptr))) return rc; sc.uconst = (double*)ptr;
  if ((rc = c->get("tile_counters", sizeof(unsigned long long) * (size_t)(sc.NC + 1), &ptr))) return rc;
  sc.tile_counters = (unsigned long long*)ptr;      // single-kernel paths: zeroed by k_prep_scalars before every launch
  // split form: [0] / [1] tile schedulers of the correlation and epilogue kernels, [2] / [3] their exit counters.
  // Both pairs re-arm themselves at the end of the kernel, so they are cleared once, when allocated.
  if ((rc = c->get("split_counters", sizeof(unsigned long long) * 4, &ptr))) return rc;
  if (ptr != c->counters_zeroed) {
    CK(cudaMemsetAsync(ptr, 0, sizeof(unsigned long long) * 4, c->stream));
    c->counters_zeroed = ptr;
  }
  sc.split_counters = (unsigned long long*)ptr;
  sc.dw = sc.uu = nullptr;
  if (need_general) {
    if ((rc = c->get("dw", sizeof(double) * (size_t)sv.B * sv.seglen, &ptr))) return rc; sc.dw = (double*)ptr;
    if ((rc = c->get("uu", sizeof(double) * (size_t)sv.B * sv.seglen, &ptr))) return rc; sc.uu = (double*)ptr;
  }
  sc.Yw = nullptr; sc.Yw_curves = 0; sc.Yw_lnO = nullptr; sc.Yw_noisepoly = 0;
  if (one_chunk && !sv.const_var && wsplit_fits(pv)) {   // weighted path, split form: window sums of a chunk of curves
    int ch = 1;
    const size_t bytes = wsplit_scratch_bytes(pv, sv.B, nk, &ch);
    if ((rc = c->get("Yw", bytes, &ptr))) return rc;
    sc.Yw = (double*)ptr; sc.Yw_curves = ch;
  }
  sc.part = nullptr;
  if (need_part) {
    if ((rc = c->get("part", sizeof(double2) * (size_t)std::max(pv.nhyp, 1) * sc.NC * sc.sub * sv.B * nk, &ptr))) return rc;
    sc.part = (double2*)ptr;
  }
  sc.bgabs = nullptr;
  if (need_bg) {
    if ((rc = c->get("bgabs", sizeof(double) * (size_t)sv.B * nk, &ptr))) return rc; sc.bgabs = (double*)ptr;
  }
  return BFL_OK;
}

static bool range_has_edges(const PlanView& pv, const SeriesView& sv) {
  const int H = pv.W / 2;
  return sv.k0 < H || sv.k1 > sv.N - H;
}

static int check_series(const PlanView& pv, const SeriesView& sv) {
  const int H = pv.W / 2;
  if (sv.B < 1 || sv.N < pv.W) return fail(BFL_ERR_ARG, "Error... bglen is greater than the data length!");
  if (sv.k0 < 0 || sv.k1 > sv.N || sv.k0 >= sv.k1) return fail(BFL_ERR_ARG, "bad sample range");
  const int64_t need_lo = std::max<int64_t>(0, sv.k0 - H), need_hi = std::min<int64_t>(sv.N, sv.k1 + H);
  if (sv.seg0 > need_lo || sv.seg0 + sv.seglen < need_hi)
    return fail(BFL_ERR_ARG, "segment does not cover the requested samples plus their halo");
  return BFL_OK;
}

static int detector_run(bfl_ctx* c, bfl_plan* p, int noisepoly, const SeriesView& sv, double* d_lnO, double* d_marg) {
  const PlanView& pv = p->pv;
  if (pv.nhyp < 1) return fail(BFL_ERR_ARG, "the detector needs a signal hypothesis (plan hypothesis 0)");
  int rc = check_series(pv, sv);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const bool edges = range_has_edges(pv, sv);
  const bool need_general = !sv.const_var || edges;
  const int nnoise = pv.nhyp - 1 + (noisepoly ? 1 : 0);
  const bool need_bg = (d_marg != nullptr) || nnoise == 0;
  Scratch sc{};
  // first pass without the partial-sum buffer: it is not needed when the kernel can fuse the combine
  // per-sample variances: the separable weighted kernel keeps the whole plan in one CTA
  const bool wsep = !sv.const_var && wsep_fits(pv);
  if ((rc = setup_scratch(c, pv, sv, need_general, need_bg, false, sc, wsep))) return rc;
  const bool fused = sv.const_var && !edges && !need_bg && fast_can_fuse(pv, sc, noisepoly);
  if (!fused && (rc = setup_scratch(c, pv, sv, need_general, need_bg, true, sc, wsep))) return rc;
  cudaStream_t s = c->stream;
  // the split form computes the per-curve scalars inside its correlation kernel and re-arms its own tile
  // schedulers: no preparation launch
  if (!(fused && sc.Y)) CK(launch_prep(sv, sc, need_general, s, &c->stats));
  if (fused) {
    std::pair<cudaEvent_t, cudaEvent_t> pr;
    if (c->timing && (rc = timing_begin(c, pr))) return rc;
    if (sc.Y) CK(launch_window_split(pv, sv, sc, 2, noisepoly, d_lnO, s, &c->stats));
    else CK(launch_window_fast_fused(pv, sv, sc, noisepoly, d_lnO, s, &c->stats));
    if (c->timing) { CK(cudaEventRecord(pr.second, s)); c->tev.push_back(pr); }
    return BFL_OK;
  }
  if (sv.const_var) {
    std::pair<cudaEvent_t, cudaEvent_t> pr;
    if (c->timing && (rc = timing_begin(c, pr))) return rc;
    if (sc.Y) CK(launch_window_split(pv, sv, sc, 0, noisepoly, nullptr, s, &c->stats));
    else CK(launch_window_fast_lse(pv, sv, sc, s, &c->stats));
    if (c->timing) { CK(cudaEventRecord(pr.second, s)); c->tev.push_back(pr); }
    if (need_bg) CK(launch_bgproj(pv, sv, sc, s, &c->stats));
    if (edges) CK(launch_window_general_lse(pv, sv, sc, true, s, &c->stats));
  } else if (wsep || weighted_fits(pv, sc)) {
    std::pair<cudaEvent_t, cudaEvent_t> pr;
    if (c->timing && (rc = timing_begin(c, pr))) return rc;
    // split weighted form without truncated windows or marginal outputs: k_wsolve folds the hypotheses itself
    const bool fold = wsep && sc.Yw && !edges && !d_marg;
    if (fold) { sc.Yw_lnO = d_lnO; sc.Yw_noisepoly = noisepoly; }
    if (wsep) CK(launch_window_wsep_lse(pv, sv, sc, s, &c->stats));
    else CK(launch_window_weighted_lse(pv, sv, sc, s, &c->stats));
    if (c->timing) { CK(cudaEventRecord(pr.second, s)); c->tev.push_back(pr); }
    if (fold) return BFL_OK;
    if (edges) CK(launch_window_general_lse(pv, sv, sc, true, s, &c->stats));
  } else {
    CK(launch_window_general_lse(pv, sv, sc, false, s, &c->stats));
  }
  CK(launch_combine(pv, sv, sc, noisepoly, d_lnO, d_marg, s, &c->stats));
  return BFL_OK;
}

int bfl_detector_odds_dev(bfl_ctx* c, bfl_plan* p, int noisepoly, const double* d_clc, const double* d_cle,
                          int const_var, const double* d_sk, int32_t B, int64_t N, int64_t seg0, int64_t seglen,
                          int64_t k0, int64_t k1, double* d_lnO) {
  if (!c || !p || !d_clc || !d_sk || !d_lnO) return fail(BFL_ERR_ARG, "null argument to bfl_detector_odds_dev");
  SeriesView sv{d_clc, d_cle, d_sk, B, N, seg0, seglen, k0, k1, (d_cle == nullptr || const_var) ? 1 : 0};
  return detector_run(c, p, noisepoly, sv, d_lnO, nullptr);
}


// --------------------------------------------------------------------------------------
// noise sigma on device
static int noise_run(bfl_ctx* c, const bfl_noise_spec* spec, const double* d_clc, int B, int64_t N, double* d_sk) {
  if (!spec) return fail(BFL_ERR_ARG, "noise specification is null");
  if (spec->method != 0 && spec->method != 1) {   // bayes.py:243-245
    return fail(BFL_ERR_ARG, "Noise estimation method must be 'powerspectrum' or 'tailveto'");
  }
  if (spec->bglen != 0 && (spec->bglen < 3 || spec->bglen % 2 == 0 || !spec->sgcoef || spec->bglen > MAXW))
    return fail(BFL_ERR_ARG, "window_size size must be a positive odd number");
  if (N < 4 || (spec->bglen && N <= spec->bglen)) return fail(BFL_ERR_ARG, "light curve too short for the noise estimate");
  if (spec->method == 0 && !(spec->param > 0.0 && spec->param <= 1.0)) return fail(BFL_ERR_ARG, "psestfrac must be in (0, 1]");
  if (spec->method == 1 && !(spec->param > 0.0)) return fail(BFL_ERR_ARG, "tvsigma must be positive");
  cudaStream_t s = c->stream;
  int rc;
  void* ptr;
  const double* d_res = d_clc;
  if (spec->bglen) {
    void* d_m;
    if ((rc = c->get("sg_coef", sizeof(double) * MAXW, &d_m))) return rc;
    if ((rc = c->get("sg_resid", sizeof(double) * (size_t)B * N, &ptr))) return rc;
    // the filter taps are uploaded only when they change (and never during a stream capture: the source is pageable)
    if (c->sg_cached.size() != (size_t)spec->bglen || c->sg_cached_dev != d_m ||
        std::memcmp(c->sg_cached.data(), spec->sgcoef, sizeof(double) * (size_t)spec->bglen) != 0) {
      c->sg_cached.assign(spec->sgcoef, spec->sgcoef + spec->bglen);
      c->sg_cached_dev = d_m;
      CK(cudaMemcpyAsync(d_m, c->sg_cached.data(), sizeof(double) * (size_t)spec->bglen, cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));
    }
    CK(launch_sg_resid(d_clc, N, B, (const double*)d_m, spec->bglen, (double*)ptr, s, &c->stats));
    d_res = (const double*)ptr;
  }
  if (spec->method == 0) {
    const int64_t L = N / 2 + 1;
    if ((rc = c->get("fft_out", sizeof(double2) * (size_t)B * L, &ptr))) return rc;
    auto key = std::make_pair(N, B);
    auto it = c->fft.find(key);
    if (it == c->fft.end()) {
      cufftHandle plan;
      int n[1] = {(int)N};
      cufftResult r = cufftPlanMany(&plan, 1, n, nullptr, 1, (int)N, nullptr, 1, (int)L, CUFFT_D2Z, B);
      if (r != CUFFT_SUCCESS) return fail(BFL_ERR_CUDA, "cufftPlanMany failed (" + std::to_string((int)r) + ")");
      it = c->fft.emplace(key, plan).first;
    }
    if (cufftSetStream(it->second, s) != CUFFT_SUCCESS) return fail(BFL_ERR_CUDA, "cufftSetStream failed");
    cufftResult r = cufftExecD2Z(it->second, (cufftDoubleReal*)d_res, (cufftDoubleComplex*)ptr);
    if (r != CUFFT_SUCCESS) return fail(BFL_ERR_CUDA, "cufftExecD2Z failed (" + std::to_string((int)r) + ")");
    c->stats.launches++;
    CK(launch_ps_sigma((const double2*)ptr, N, B, spec->param, d_sk, s, &c->stats));
  } else {
    void *d_cnt, *d_rng;
    if ((rc = c->get("tv_counts", sizeof(int) * (size_t)B * N, &d_cnt))) return rc;
    if ((rc = c->get("tv_range", sizeof(double) * 2 * (size_t)B, &d_rng))) return rc;
    CK(launch_tv_sigma(d_res, N, B, spec->param, (int*)d_cnt, (double*)d_rng, d_sk, s, &c->stats));
  }
  return BFL_OK;
}

int bfl_noise_sigma_dev(bfl_ctx* c, const bfl_noise_spec* spec, const double* d_clc, int32_t B, int64_t N, double* d_sk) {
  if (!c || !d_clc || !d_sk || B < 1) return fail(BFL_ERR_ARG, "bad argument to bfl_noise_sigma_dev");
  CK(cudaSetDevice(c->device));
  return noise_run(c, spec, d_clc, B, N, d_sk);
}

int bfl_noise_sigma(bfl_ctx* c, const bfl_noise_spec* spec, const double* clc, int32_t B, int64_t N, double* out_sk) {
  if (!c || !clc || !out_sk || B < 1 || N < 1) return fail(BFL_ERR_ARG, "bad argument to bfl_noise_sigma");
  CK(cudaSetDevice(c->device));
  void *d_clc, *d_sk;
  int rc;
  if ((rc = c->get("in_clc", sizeof(double) * (size_t)B * N, &d_clc))) return rc;
  if ((rc = c->get("in_sk", sizeof(double) * (size_t)B, &d_sk))) return rc;
  CK(cudaMemcpyAsync(d_clc, clc, sizeof(double) * (size_t)B * N, cudaMemcpyHostToDevice, c->stream));
  if ((rc = noise_run(c, spec, (const double*)d_clc, B, N, (double*)d_sk))) return rc;
  CK(cudaMemcpyAsync(out_sk, d_sk, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return BFL_OK;
}

static bool rows_constant(const double* a, int B, int64_t N) {
  for (int b = 0; b < B; b++) {
    const double* r = a + (size_t)b * N;
    const double v = r[0] * r[0];
    for (int64_t i = 1; i < N; i++)
      if (r[i] * r[i] != v) return false;
  }
  return true;
}

int bfl_detector_odds(bfl_ctx* c, bfl_plan* p, int noisepoly, const double* clc, const double* cle,
                      const double* sk, const bfl_noise_spec* noise, int32_t B, int64_t N, int64_t k0, int64_t k1,
                      double* out_lnO, double* out_marg, double* out_sk) {
  if (!c || !p || !clc || !out_lnO) return fail(BFL_ERR_ARG, "null argument to bfl_detector_odds");
  if (!sk && !noise) return fail(BFL_ERR_ARG, "either the noise sigmas or a noise specification is required");
  if (B < 1 || N < 1) return fail(BFL_ERR_ARG, "empty batch");
  if (k0 < 0 || k1 > N || k0 >= k1) return fail(BFL_ERR_ARG, "bad sample range");
  CK(cudaSetDevice(c->device));
  const bool cv = (cle == nullptr) || rows_constant(cle, B, N);
  const int64_t nk = k1 - k0;
  const size_t nb = sizeof(double) * (size_t)B * N;
  const size_t nbo = sizeof(double) * (size_t)B * (size_t)nk;
  void *d_clc, *d_cle = nullptr, *d_sk, *d_lnO, *d_marg = nullptr;
  int rc;
  if ((rc = c->get("in_clc", nb, &d_clc))) return rc;
  if (cle && (rc = c->get("in_cle", nb, &d_cle))) return rc;
  if ((rc = c->get("in_sk", sizeof(double) * (size_t)B, &d_sk))) return rc;
  if ((rc = c->get("out_lnO", nbo, &d_lnO))) return rc;
  if (out_marg && (rc = c->get("out_marg", nbo * (size_t)(p->pv.nhyp + 1), &d_marg))) return rc;
  cudaStream_t s = c->stream;

  // Large batches are pipelined in sub-batches of curves: host->device copies, the kernels and
  // the device->host copies of consecutive sub-batches overlap on three streams (the copy
  // engines are idle otherwise: 36 MB each way for 1,024 x 4,400 samples).
  // Sub-batch sizes 1:3:3:1 -- small first and last pieces keep the exposed copy-in / copy-out short,
  // large middle pieces keep the kernel's tail (one ~100 us warp tile) a small fraction of its run.
  int nsub = 1;
  if (!out_marg && B >= 512) nsub = 4;
  else if (!out_marg && B >= 256) nsub = 2;
  if (nsub > 1) {
    if (!c->s_in) CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    if (!c->s_out) CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    while ((int)c->pipe_ev.size() < 2 * nsub + 1) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      c->pipe_ev.push_back(ev);
    }
    std::vector<int> lo(nsub + 1);
    if (nsub == 4) { lo[0] = 0; lo[1] = B / 8; lo[2] = B / 2; lo[3] = B - B / 8; lo[4] = B; }
    else for (int j = 0; j <= nsub; j++) lo[j] = (int)((int64_t)B * j / nsub);
    // scratch is sized ONCE, for the largest sub-batch, before anything is in flight: growing a buffer
    // mid-pipeline would synchronise, free and reallocate under the copy streams
    {
      int bmax = 0;
      for (int j = 0; j < nsub; j++) bmax = std::max(bmax, lo[j + 1] - lo[j]);
      SeriesView svmax{(const double*)d_clc, cle ? (const double*)d_cle : nullptr, (const double*)d_sk, bmax, N, 0, N, k0, k1, cv ? 1 : 0};
      Scratch scmax{};
      const bool edges = range_has_edges(p->pv, svmax);
      if ((rc = check_series(p->pv, svmax))) return rc;
      if ((rc = setup_scratch(c, p->pv, svmax, !cv || edges, false, false, scmax, !cv && wsep_fits(p->pv)))) return rc;
      if (!sk && noise) {
        void* tmp;
        if (noise->bglen && (rc = c->get("sg_resid", sizeof(double) * (size_t)bmax * N, &tmp))) return rc;
        if (noise->method == 0 && (rc = c->get("fft_out", sizeof(double2) * (size_t)bmax * (N / 2 + 1), &tmp))) return rc;
      }
    }
    // an error below leaves asynchronous copies to / from CALLER memory in flight: drain all three streams first
    auto bail = [&](int code) {
      cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(s); cudaStreamSynchronize(c->s_out);
      return code;
    };
    if (sk) CK(cudaMemcpyAsync(d_sk, sk, sizeof(double) * (size_t)B, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(c->pipe_ev[2 * nsub], s));
    CK(cudaStreamWaitEvent(c->s_in, c->pipe_ev[2 * nsub], 0));   // previous users of the input buffers are done
    for (int j = 0; j < nsub; j++) {
      const size_t off = (size_t)lo[j] * N, cnt = (size_t)(lo[j + 1] - lo[j]) * N;
      CK(cudaMemcpyAsync((double*)d_clc + off, clc + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->s_in));
      if (cle) CK(cudaMemcpyAsync((double*)d_cle + off, cle + off, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->s_in));
      CK(cudaEventRecord(c->pipe_ev[j], c->s_in));
    }
    for (int j = 0; j < nsub; j++) {
      const int Bj = lo[j + 1] - lo[j];
      const size_t off = (size_t)lo[j] * N, ooff = (size_t)lo[j] * nk;
      CK(cudaStreamWaitEvent(s, c->pipe_ev[j], 0));
      if (!sk && (rc = noise_run(c, noise, (const double*)d_clc + off, Bj, N, (double*)d_sk + lo[j]))) return bail(rc);
      SeriesView sv{(const double*)d_clc + off, cle ? (const double*)d_cle + off : nullptr, (const double*)d_sk + lo[j],
                    Bj, N, 0, N, k0, k1, cv ? 1 : 0};
      if ((rc = detector_run(c, p, noisepoly, sv, (double*)d_lnO + ooff, nullptr))) return bail(rc);
      CK(cudaEventRecord(c->pipe_ev[nsub + j], s));
      CK(cudaStreamWaitEvent(c->s_out, c->pipe_ev[nsub + j], 0));
      CK(cudaMemcpyAsync(out_lnO + ooff, (double*)d_lnO + ooff, sizeof(double) * (size_t)Bj * nk, cudaMemcpyDeviceToHost,
                         c->s_out));
    }
    if (out_sk) CK(cudaMemcpyAsync(out_sk, d_sk, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(c->s_out));
    CK(cudaStreamSynchronize(s));
    return BFL_OK;
  }

  CK(cudaMemcpyAsync(d_clc, clc, nb, cudaMemcpyHostToDevice, s));
  if (cle) CK(cudaMemcpyAsync(d_cle, cle, nb, cudaMemcpyHostToDevice, s));
  if (sk) CK(cudaMemcpyAsync(d_sk, sk, sizeof(double) * (size_t)B, cudaMemcpyHostToDevice, s));
  else if ((rc = noise_run(c, noise, (const double*)d_clc, B, N, (double*)d_sk))) return rc;
  if (out_sk) CK(cudaMemcpyAsync(out_sk, d_sk, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
  SeriesView sv{(const double*)d_clc, (const double*)d_cle, (const double*)d_sk, B, N, 0, N, k0, k1, cv ? 1 : 0};
  if ((rc = detector_run(c, p, noisepoly, sv, (double*)d_lnO, (double*)d_marg))) return rc;
  CK(cudaMemcpyAsync(out_lnO, d_lnO, nbo, cudaMemcpyDeviceToHost, s));
  if (out_marg) CK(cudaMemcpyAsync(out_marg, d_marg, nbo * (size_t)(p->pv.nhyp + 1), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
// Pipelined batch search: bfl_pipe.  Every slot owns a child context (its own stream and scratch), so the
// host->device copy of job i+1, the kernels of job i and the device->host copy of job i-1 overlap, and the tail
// of one job's kernels is filled by the next job's.  The plan is shared (read-only on the device).
static_assert(sizeof(bfl_region) == sizeof(RegionRec), "bfl_region and RegionRec must have one layout");
struct bfl_pipe {
  bfl_ctx* parent = nullptr;
  bfl_plan* plan = nullptr;
  int noisepoly = 1;
  bool have_noise = false;
  bfl_noise_spec noise{};
  std::vector<double> sgcoef;
  int32_t B = 0, max_regions = 0;
  int64_t N = 0, k0 = 0, k1 = 0;
  struct Slot {
    bfl_ctx* ctx = nullptr;
    double *d_clc = nullptr, *d_sk = nullptr, *d_lnO = nullptr;
    RegionRec* d_rec = nullptr;
    int64_t* d_cnt = nullptr;        // [B + 3]
    cudaEvent_t done = nullptr;
    int64_t job = -1;
    const int64_t* host_cnt = nullptr;
  };
  std::vector<Slot> slots;
  int64_t submitted = 0;
  // A job's whole enqueue sequence (copies, ~10 kernels, copies) as ONE CUDA graph per (slot, host buffers, mode):
  // callers cycle through a few page-locked buffers, and on a box with 8 ranks the ~15 runtime calls per job were
  // what kept the GPUs waiting.  Captured the second time a key is seen on a warm slot.
  struct GraphKey {
    int slot; const void *clc, *sk, *lnO, *reg, *cnt, *osk; uint64_t thresh_bits;
    bool operator<(const GraphKey& o) const {
      return std::tie(slot, clc, sk, lnO, reg, cnt, osk, thresh_bits) <
             std::tie(o.slot, o.clc, o.sk, o.lnO, o.reg, o.cnt, o.osk, o.thresh_bits);
    }
  };
  struct GraphVal { cudaGraphExec_t exec = nullptr; int seen = 0; int64_t launches = 0; };
  std::map<GraphKey, GraphVal> graphs;
  bool use_graphs = true;
};

int bfl_pipe_destroy(bfl_pipe* p) {
  if (!p) return BFL_OK;
  if (p->parent) cudaSetDevice(p->parent->device);
  for (auto& kv : p->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (auto& sl : p->slots) {
    if (sl.ctx) cudaStreamSynchronize(sl.ctx->stream);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.d_clc) cudaFree(sl.d_clc);
    if (sl.d_sk) cudaFree(sl.d_sk);
    if (sl.d_lnO) cudaFree(sl.d_lnO);
    if (sl.d_rec) cudaFree(sl.d_rec);
    if (sl.d_cnt) cudaFree(sl.d_cnt);
    if (sl.ctx) bfl_ctx_destroy(sl.ctx);
  }
  delete p;
  return BFL_OK;
}

int bfl_pipe_create(bfl_ctx* c, bfl_plan* plan, int noisepoly, const bfl_noise_spec* noise, int32_t B, int64_t N,
                    int64_t k0, int64_t k1, int32_t depth, int32_t max_regions, bfl_pipe** out) {
  if (!c || !plan || !out) return fail(BFL_ERR_ARG, "null argument to bfl_pipe_create");
  *out = nullptr;
  if (B < 1 || N < plan->pv.W || k0 < 0 || k1 > N || k0 >= k1) return fail(BFL_ERR_ARG, "bad batch shape or sample range");
  if (depth < 1 || depth > 8) return fail(BFL_ERR_ARG, "pipeline depth must be in 1..8");
  if (max_regions < 0) return fail(BFL_ERR_ARG, "max_regions must not be negative");
  CK(cudaSetDevice(c->device));
  bfl_pipe* p = new (std::nothrow) bfl_pipe();
  if (!p) return fail(BFL_ERR_OOM, "out of host memory");
  p->parent = c; p->plan = plan; p->noisepoly = noisepoly ? 1 : 0;
  p->B = B; p->N = N; p->k0 = k0; p->k1 = k1; p->max_regions = max_regions;
  if (noise) {
    p->have_noise = true;
    p->noise = *noise;
    if (noise->sgcoef && noise->bglen > 0) {
      p->sgcoef.assign(noise->sgcoef, noise->sgcoef + noise->bglen);
      p->noise.sgcoef = p->sgcoef.data();
    }
  }
  const size_t nb = sizeof(double) * (size_t)B * N, nbo = sizeof(double) * (size_t)B * (size_t)(k1 - k0);
  p->slots.resize(depth);
  for (auto& sl : p->slots) {
    int rc = bfl_ctx_create(c->device, nullptr, &sl.ctx);
    if (rc) { bfl_pipe_destroy(p); return rc; }
    cudaError_t e = cudaMalloc(&sl.d_clc, nb);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_sk, sizeof(double) * (size_t)B);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_lnO, nbo);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_rec, sizeof(RegionRec) * (size_t)std::max(max_regions, 1));
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_cnt, sizeof(int64_t) * (size_t)(B + 3));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      bfl_pipe_destroy(p);
      return fail(e == cudaErrorMemoryAllocation ? BFL_ERR_OOM : BFL_ERR_CUDA, std::string("bfl_pipe_create: ") + cudaGetErrorString(e));
    }
  }
  *out = p;
  return BFL_OK;
}

// everything one job enqueues on its slot's stream (also what a captured graph holds)
static int pipe_enqueue(bfl_pipe* p, bfl_pipe::Slot& sl, const double* clc, const double* sk, double* out_lnO, double thresh,
                        bfl_region* out_regions, int64_t* out_counts, double* out_sk) {
  bfl_ctx* c = sl.ctx;
  cudaStream_t s = c->stream;
  const int32_t B = p->B;
  const int64_t N = p->N, nk = p->k1 - p->k0;
  const bool want_regions = !std::isnan(thresh);
  int rc;
  CK(cudaMemcpyAsync(sl.d_clc, clc, sizeof(double) * (size_t)B * N, cudaMemcpyHostToDevice, s));
  if (sk) CK(cudaMemcpyAsync(sl.d_sk, sk, sizeof(double) * (size_t)B, cudaMemcpyHostToDevice, s));
  else if ((rc = noise_run(c, &p->noise, sl.d_clc, B, N, sl.d_sk))) return rc;
  SeriesView sv{sl.d_clc, nullptr, sl.d_sk, B, N, 0, N, p->k0, p->k1, 1};
  if ((rc = detector_run(c, p->plan, p->noisepoly, sv, sl.d_lnO, nullptr))) return rc;
  if (out_sk) CK(cudaMemcpyAsync(out_sk, sl.d_sk, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
  if (want_regions) {
    CK(cudaMemsetAsync(sl.d_cnt, 0, sizeof(int64_t) * (size_t)(B + 3), s));
    CK(launch_batch_regions(sl.d_lnO, B, nk, thresh, sl.d_rec, p->max_regions, sl.d_cnt, sl.d_cnt + B, s, &c->stats));
    CK(cudaMemcpyAsync(out_counts, sl.d_cnt, sizeof(int64_t) * (size_t)(B + 3), cudaMemcpyDeviceToHost, s));
    if (out_regions && p->max_regions > 0)
      CK(cudaMemcpyAsync(out_regions, sl.d_rec, sizeof(RegionRec) * (size_t)p->max_regions, cudaMemcpyDeviceToHost, s));
  }
  if (out_lnO) CK(cudaMemcpyAsync(out_lnO, sl.d_lnO, sizeof(double) * (size_t)B * (size_t)nk, cudaMemcpyDeviceToHost, s));
  return BFL_OK;
}

static bool host_pinned(const void* ptr) {
  if (!ptr) return true;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

int bfl_pipe_submit(bfl_pipe* p, const double* clc, const double* sk, double* out_lnO, double thresh,
                    bfl_region* out_regions, int64_t* out_counts, double* out_sk, int64_t* job) {
  if (!p || !clc) return fail(BFL_ERR_ARG, "null argument to bfl_pipe_submit");
  if (!sk && !p->have_noise) return fail(BFL_ERR_ARG, "either the noise sigmas or a noise specification is required");
  const bool want_regions = !std::isnan(thresh);
  if (want_regions && !out_counts) return fail(BFL_ERR_ARG, "out_counts is required when a threshold is given");
  if (!want_regions && !out_lnO) return fail(BFL_ERR_ARG, "nothing to return: give out_lnO and/or a threshold");
  CK(cudaSetDevice(p->parent->device));
  const int slot = (int)(p->submitted % (int64_t)p->slots.size());
  bfl_pipe::Slot& sl = p->slots[(size_t)slot];
  if (sl.job >= 0) CK(cudaEventSynchronize(sl.done));   // the slot's previous job must have left the device
  cudaStream_t s = sl.ctx->stream;
  int rc = BFL_OK;
  bool done = false;
  if (p->use_graphs && sl.job >= 0 && !sl.ctx->timing) {   // warm slot: scratch is sized, plans exist
    uint64_t tb;
    std::memcpy(&tb, &thresh, sizeof(tb));
    bfl_pipe::GraphKey key{slot, clc, sk, out_lnO, out_regions, out_counts, out_sk, tb};
    bfl_pipe::GraphVal& gv = p->graphs[key];
    if (gv.exec) {
      CK(cudaGraphLaunch(gv.exec, s));
      sl.ctx->stats.launches += gv.launches;
      done = true;
    } else if (++gv.seen >= 2 && p->graphs.size() <= 64 && host_pinned(clc) && host_pinned(sk) && host_pinned(out_lnO) &&
               host_pinned(out_regions) && host_pinned(out_counts) && host_pinned(out_sk)) {
      const int64_t l0 = sl.ctx->stats.launches;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = pipe_enqueue(p, sl, clc, sk, out_lnO, thresh, out_regions, out_counts, out_sk);
        const cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (rc == BFL_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&gv.exec, graph, 0) == cudaSuccess) {
          gv.launches = sl.ctx->stats.launches - l0;
          CK(cudaGraphLaunch(gv.exec, s));
          done = true;
        } else {
          gv.exec = nullptr;
          gv.seen = -1000000;          // this key cannot be captured: enqueue directly from now on
          cudaGetLastError();
          rc = BFL_OK;
        }
        if (graph) cudaGraphDestroy(graph);
      } else {
        cudaGetLastError();
      }
    }
  }
  if (!done && (rc = pipe_enqueue(p, sl, clc, sk, out_lnO, thresh, out_regions, out_counts, out_sk))) return rc;
  CK(cudaEventRecord(sl.done, s));
  sl.job = p->submitted;
  sl.host_cnt = want_regions ? out_counts : nullptr;
  if (job) *job = p->submitted;
  p->submitted++;
  return BFL_OK;
}

int bfl_pipe_wait(bfl_pipe* p, int64_t job, int32_t* nregions) {
  if (!p) return fail(BFL_ERR_ARG, "null pipe");
  if (nregions) *nregions = 0;
  if (job < 0 || job >= p->submitted) return fail(BFL_ERR_ARG, "no such job");
  if (job < p->submitted - (int64_t)p->slots.size()) return BFL_OK;   // its slot was reused: it completed long ago
  bfl_pipe::Slot& sl = p->slots[(size_t)(job % (int64_t)p->slots.size())];
  CK(cudaEventSynchronize(sl.done));
  if (nregions && sl.host_cnt) {
    const int64_t n = sl.host_cnt[p->B + 2];
    if (n > p->max_regions) return fail(BFL_ERR_UNSUPPORTED, "more regions than max_regions (out_counts is complete; raise max_regions)");
    *nregions = (int32_t)n;
  }
  return BFL_OK;
}

int32_t bfl_pipe_graph_count(const bfl_pipe* p) {
  int32_t n = 0;
  if (p) for (auto& kv : p->graphs) n += kv.second.exec ? 1 : 0;
  return n;
}

int64_t bfl_pipe_launch_count(const bfl_pipe* p) {
  int64_t n = 0;
  if (p) for (auto& sl : p->slots) n += sl.ctx ? sl.ctx->stats.launches : 0;
  return n;
}

// --------------------------------------------------------------------------------------
__global__ void k_fill(double* p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int bfl_window_lnB(bfl_ctx* c, bfl_plan* p, int ihyp, const double* clc, const double* cle, double sk, int64_t N,
                   double sumlogsigma, double* out_lnB, double* out_bg) {
  if (!c || !p || !clc) return fail(BFL_ERR_ARG, "null argument to bfl_window_lnB");
  const PlanView& pv = p->pv;
  if (out_lnB && (ihyp < 0 || ihyp >= pv.nhyp)) return fail(BFL_ERR_ARG, "hypothesis index out of range");
  if (!out_lnB) ihyp = -1;
  CK(cudaSetDevice(c->device));
  const bool cv = (cle == nullptr) || rows_constant(cle, 1, N);
  const size_t nb = sizeof(double) * (size_t)N;
  void *d_clc, *d_cle = nullptr, *d_sk, *d_out = nullptr;
  int rc;
  if ((rc = c->get("in_clc", nb, &d_clc))) return rc;
  if (cle && (rc = c->get("in_cle", nb, &d_cle))) return rc;
  if ((rc = c->get("in_sk", sizeof(double), &d_sk))) return rc;
  const int ng = (ihyp >= 0) ? p->ngrid[ihyp] : 0;
  if (ng > 0 && (rc = c->get("out_full", nb * (size_t)ng, &d_out))) return rc;
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(d_clc, clc, nb, cudaMemcpyHostToDevice, s));
  if (cle) CK(cudaMemcpyAsync(d_cle, cle, nb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_sk, &sk, sizeof(double), cudaMemcpyHostToDevice, s));
  SeriesView sv{(const double*)d_clc, (const double*)d_cle, (const double*)d_sk, 1, N, 0, N, 0, N, cv ? 1 : 0};
  if ((rc = check_series(pv, sv))) return rc;
  Scratch sc{};
  if ((rc = setup_scratch(c, pv, sv, true, true, false, sc))) return rc;
  if (ng > 0) {
    const size_t n = (size_t)ng * N;
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((double*)d_out, n, -INFINITY);
    c->stats.launches++;
  }
  CK(launch_prep(sv, sc, true, s, &c->stats));
  if (cv) {
    CK(launch_bgproj(pv, sv, sc, s, &c->stats));
    if (ihyp >= 0) CK(launch_window_fast_full(pv, ihyp, sv, sc, sumlogsigma, (double*)d_out, s, &c->stats));
    CK(launch_window_general_full(pv, ihyp, sv, sc, true, sumlogsigma, (double*)d_out, s, &c->stats));
  } else if (weighted_fits(pv, sc)) {
    CK(launch_window_weighted_full(pv, ihyp, sv, sc, sumlogsigma, (double*)d_out, s, &c->stats));
    CK(launch_window_general_full(pv, ihyp, sv, sc, true, sumlogsigma, (double*)d_out, s, &c->stats));
  } else {
    CK(launch_window_general_full(pv, ihyp, sv, sc, false, sumlogsigma, (double*)d_out, s, &c->stats));
  }
  if (out_lnB && ng > 0) CK(cudaMemcpyAsync(out_lnB, d_out, nb * (size_t)ng, cudaMemcpyDeviceToHost, s));
  if (out_bg) CK(cudaMemcpyAsync(out_bg, sc.bgabs, nb, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (out_bg)
    for (int64_t i = 0; i < N; i++) out_bg[i] += sumlogsigma;   // general.pyx:363
  return BFL_OK;
}

int bfl_count_regions_dev(bfl_ctx* c, const double* d_lnO, int32_t B, int64_t n, double thresh, int64_t* d_counts) {
  if (!c || !d_lnO || !d_counts || B < 1 || n < 1) return fail(BFL_ERR_ARG, "bad argument to bfl_count_regions_dev");
  CK(cudaSetDevice(c->device));
  CK(launch_count_regions(d_lnO, B, n, thresh, d_counts, nullptr, c->stream, &c->stats));
  return BFL_OK;
}

int bfl_batch_regions_dev(bfl_ctx* c, const double* d_lnO, int32_t B, int64_t n, double thresh, bfl_region* d_regions,
                          int32_t max_regions, int64_t* d_counts) {
  if (!c || !d_lnO || !d_regions || !d_counts || B < 1 || n < 1 || max_regions < 1)
    return fail(BFL_ERR_ARG, "bad argument to bfl_batch_regions_dev");
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(d_counts, 0, sizeof(int64_t) * (size_t)(B + 3), c->stream));
  CK(launch_batch_regions(d_lnO, B, n, thresh, (RegionRec*)d_regions, max_regions, d_counts, d_counts + B, c->stream, &c->stats));
  return BFL_OK;
}

int bfl_sweep_simulate_dev(bfl_ctx* c, uint64_t seed, int64_t first, int32_t B, int32_t N, int32_t halfwin, double dt,
                           double nstd, double sinusoid_amp, double snr_lo, double snr_hi, int32_t inject, double* d_clc,
                           double* d_truth) {
  if (!c || !d_clc || !d_truth) return fail(BFL_ERR_ARG, "null argument to bfl_sweep_simulate_dev");
  if (B < 1 || N < 8 || N > 8192) return fail(BFL_ERR_UNSUPPORTED, "bfl_sweep_simulate_dev supports 8 <= N <= 8192 samples");
  if (halfwin < 0 || 2 * halfwin + 2 > N || !(dt > 0.0) || !(nstd > 0.0) || snr_hi < snr_lo)
    return fail(BFL_ERR_ARG, "bad argument to bfl_sweep_simulate_dev");
  CK(cudaSetDevice(c->device));
  CK(launch_sweep_simulate(seed, first, B, N, halfwin, dt, nstd, sinusoid_amp, snr_lo, snr_hi, inject, d_clc, d_truth,
                           c->stream, &c->stats));
  return BFL_OK;
}

int bfl_sweep_inject_dev(bfl_ctx* c, int32_t kind, int32_t B, int32_t N, double dt, const double* d_sk,
                         double* d_truth, double* d_clc) {
  if (!c || !d_sk || !d_truth || !d_clc) return fail(BFL_ERR_ARG, "null argument to bfl_sweep_inject_dev");
  if ((kind != 1 && kind != 2) || B < 1 || N < 2 || !(dt > 0.0)) return fail(BFL_ERR_ARG, "bad argument to bfl_sweep_inject_dev");
  CK(cudaSetDevice(c->device));
  CK(launch_sweep_inject(kind, B, N, dt, d_sk, d_truth, d_clc, c->stream, &c->stats));
  return BFL_OK;
}

int bfl_sweep_score_dev(bfl_ctx* c, const double* d_lnO, int32_t B, int64_t nk, int32_t halfwin, const double* d_truth,
                        double thresh, int32_t nbins, double snr_max, int64_t* d_injected, int64_t* d_detected,
                        int64_t* d_counters) {
  if (!c || !d_lnO || !d_truth || !d_injected || !d_detected || !d_counters || B < 1 || nk < 1 || nbins < 1 ||
      !(snr_max > 0.0))
    return fail(BFL_ERR_ARG, "bad argument to bfl_sweep_score_dev");
  CK(cudaSetDevice(c->device));
  CK(launch_sweep_score(d_lnO, B, nk, halfwin, d_truth, thresh, nbins, snr_max, d_injected, d_detected, d_counters,
                        c->stream, &c->stats));
  return BFL_OK;
}

int bfl_count_regions_total_dev(bfl_ctx* c, const double* d_lnO, int32_t B, int64_t n, double thresh, int64_t* d_counts,
                                int64_t* d_totals) {
  if (!c || !d_lnO || !d_counts || !d_totals || B < 1 || n < 1)
    return fail(BFL_ERR_ARG, "bad argument to bfl_count_regions_total_dev");
  CK(cudaSetDevice(c->device));
  CK(launch_count_regions(d_lnO, B, n, thresh, d_counts, d_totals, c->stream, &c->stats));
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
int bfl_threshold(bfl_ctx* c, const double* lnO, int64_t n, double thresh, int32_t expand, int64_t* segs,
                  double* maxval, int64_t* maxidx, int32_t max_segs, int32_t* nsegs) {
  if (!c || !lnO || !nsegs) return fail(BFL_ERR_ARG, "null argument to bfl_threshold");
  *nsegs = 0;
  if (n <= 0) return BFL_OK;
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const int cap = (int)std::min<int64_t>(n / 2 + 2, 1 << 24);
  void *d_l, *d_st, *d_sp, *d_cnt;
  int rc;
  if ((rc = c->get("thr_lnO", sizeof(double) * (size_t)n, &d_l))) return rc;
  if ((rc = c->get("thr_st", sizeof(int64_t) * (size_t)cap, &d_st))) return rc;
  if ((rc = c->get("thr_sp", sizeof(int64_t) * (size_t)cap, &d_sp))) return rc;
  if ((rc = c->get("thr_cnt", sizeof(int) * 2, &d_cnt))) return rc;
  CK(cudaMemcpyAsync(d_l, lnO, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(d_cnt, 0, sizeof(int) * 2, s));
  CK(launch_threshold_edges((const double*)d_l, n, thresh, (int64_t*)d_st, (int64_t*)d_sp, (int*)d_cnt, cap, s, &c->stats));
  int cnt[2] = {0, 0};
  CK(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (cnt[0] != cnt[1] || cnt[0] > cap) return fail(BFL_ERR_CUDA, "threshold edge extraction overflowed");
  const int nr = cnt[0];
  std::vector<int64_t> st(nr), sp(nr);
  if (nr) {
    CK(cudaMemcpyAsync(st.data(), d_st, sizeof(int64_t) * (size_t)nr, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(sp.data(), d_sp, sizeof(int64_t) * (size_t)nr, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  std::sort(st.begin(), st.end());
  std::sort(sp.begin(), sp.end());
  // expansion and merging of adjacent/overlapping regions: find.py:957-1006
  std::vector<std::pair<int64_t, int64_t>> fl(nr);
  for (int i = 0; i < nr; i++) fl[i] = {st[i], sp[i]};
  if (expand > 0 && nr > 0) {
    for (auto& sg : fl) {
      sg.first -= expand; sg.second += expand;
      if (sg.first < 0) sg.first = 0;
      if (sg.second >= n) sg.second = n;
    }
    if (nr > 1) {
      std::vector<std::pair<int64_t, int64_t>> merged;
      size_t j = 0;
      while (true) {
        auto cur = fl[j];
        j++;
        for (size_t k = j; k < fl.size(); k++) {
          if (cur.second >= fl[k].first) { cur.second = fl[k].second; j++; } else break;
        }
        merged.push_back(cur);
        if (j >= fl.size()) break;
      }
      fl.swap(merged);
    }
  }
  const int nf = (int)fl.size();
  *nsegs = nf;
  const int nout = std::min<int>(nf, max_segs);
  if (nout > 0 && segs) {
    for (int i = 0; i < nout; i++) { segs[2 * i] = fl[i].first; segs[2 * i + 1] = fl[i].second; }
  }
  if (nout > 0 && maxval && maxidx && segs) {
    void *d_sg, *d_mv, *d_mi;
    if ((rc = c->get("thr_sg", sizeof(int64_t) * 2 * (size_t)nout, &d_sg))) return rc;
    if ((rc = c->get("thr_mv", sizeof(double) * (size_t)nout, &d_mv))) return rc;
    if ((rc = c->get("thr_mi", sizeof(int64_t) * (size_t)nout, &d_mi))) return rc;
    CK(cudaMemcpyAsync(d_sg, segs, sizeof(int64_t) * 2 * (size_t)nout, cudaMemcpyHostToDevice, s));
    CK(launch_region_max((const double*)d_l, (const int64_t*)d_sg, nout, (double*)d_mv, (int64_t*)d_mi, s, &c->stats));
    CK(cudaMemcpyAsync(maxval, d_mv, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(maxidx, d_mi, sizeof(int64_t) * (size_t)nout, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
int bfl_pe_grid(bfl_ctx* c, int32_t kind, int32_t reverse, const double* cts, const double* clc, int64_t L,
                const double* params, const double* logprior, int64_t G, int32_t P, const double* polyms,
                double sigma, int32_t margbackground, const double* sgcoef, int32_t sgbins, double* out_post) {
  if (!c || !cts || !clc || !params || !logprior || !out_post) return fail(BFL_ERR_ARG, "null argument to bfl_pe_grid");
  if (sgcoef && (sgbins < 3 || sgbins % 2 == 0 || sgbins > MAXW || (sgbins - 1) / 2 + 1 > L))
    return fail(BFL_ERR_ARG, "window_size size must be a positive odd number");
  if (margbackground && !polyms) return fail(BFL_ERR_ARG, "polyms required when margbackground is set");
  if (kind < 0 || kind >= BFL_MODEL_TABLE) return fail(BFL_ERR_ARG, "bad model kind");
  if (L < 2 || L > 200) return fail(BFL_ERR_UNSUPPORTED, "chunk length must be in 2..200 samples");
  if (margbackground && (P < 1 || P > MAXP)) return fail(BFL_ERR_UNSUPPORTED, "bgorder must be in 0..5");
  if (!margbackground) P = 1;
  if (G < 1) return BFL_OK;
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  void *d_cts, *d_clc, *d_par, *d_lp, *d_poly = nullptr, *d_post;
  int rc;
  if ((rc = c->get("pe_cts", sizeof(double) * (size_t)L, &d_cts))) return rc;
  if ((rc = c->get("pe_clc", sizeof(double) * (size_t)L, &d_clc))) return rc;
  if ((rc = c->get("pe_par", sizeof(double) * 4 * (size_t)G, &d_par))) return rc;
  if ((rc = c->get("pe_lp", sizeof(double) * (size_t)G, &d_lp))) return rc;
  if ((rc = c->get("pe_post", sizeof(double) * (size_t)G, &d_post))) return rc;
  if ((rc = c->get("pe_poly", sizeof(double) * (size_t)P * L, &d_poly))) return rc;
  CK(cudaMemcpyAsync(d_cts, cts, sizeof(double) * (size_t)L, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_clc, clc, sizeof(double) * (size_t)L, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_par, params, sizeof(double) * 4 * (size_t)G, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_lp, logprior, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice, s));
  if (margbackground) CK(cudaMemcpyAsync(d_poly, polyms, sizeof(double) * (size_t)P * L, cudaMemcpyHostToDevice, s));
  void* d_sg = nullptr;
  if (sgcoef) {
    if ((rc = c->get("pe_sg", sizeof(double) * MAXW, &d_sg))) return rc;
    CK(cudaMemcpyAsync(d_sg, sgcoef, sizeof(double) * (size_t)sgbins, cudaMemcpyHostToDevice, s));
  }
  CK(launch_pe_grid(kind, reverse, (const double*)d_cts, (const double*)d_clc, (int)L, (const double*)d_par,
                    (const double*)d_lp, G, P, (const double*)d_poly, sigma, margbackground, (double*)d_post,
                    (const double*)d_sg, sgbins, s, &c->stats));
  CK(cudaMemcpyAsync(out_post, d_post, sizeof(double) * (size_t)G, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return BFL_OK;
}

// --------------------------------------------------------------------------------------
int bfl_whole_series_lnB(bfl_ctx* c, int32_t kind, int32_t reverse, int32_t halfrange, const double* cts, const double* clc,
                         const double* noisevar, int64_t N, const double* params, const double* logprior, int64_t G,
                         int32_t P, const double* bgmodels, const double* bgcross, const double* dbgr, double sumlogsigma,
                         double* out_lnB, double* out_bg) {
  if (!c || !cts || !clc || !noisevar || !bgmodels || !bgcross || !dbgr || (G > 0 && (!params || !logprior || !out_lnB)) ||
      (G <= 0 && !out_bg))
    return fail(BFL_ERR_ARG, "null argument to bfl_whole_series_lnB");
  if (G < 0) G = 0;
  if (kind < 0 || kind >= BFL_MODEL_TABLE) return fail(BFL_ERR_ARG, "bad model kind");
  if (P < 1 || P > MAXP) return fail(BFL_ERR_UNSUPPORTED, "bgorder must be in 0..5");
  if (N < 3 || N > (1 << 20)) return fail(BFL_ERR_UNSUPPORTED, "the whole-series background supports 3 .. 2^20 samples");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const int NP = (int)((N + 1) & ~(int64_t)1);
  std::vector<TemplDesc> desc((size_t)G);
  for (int64_t g = 0; g < G; g++) {
    desc[g].kind = kind; desc[g].reverse = reverse;
    for (int i = 0; i < 4; i++) desc[g].p[i] = params[(size_t)g * 4 + i];
  }
  std::vector<double> abg((size_t)P * N), dt((size_t)N);
  for (int j = 0; j < P; j++)
    for (int64_t n = 0; n < N; n++) abg[(size_t)j * N + n] = bgmodels[(size_t)j * N + n] / noisevar[n];   // bayes.py:394
  for (int64_t n = 0; n < N; n++) dt[n] = clc[n] / noisevar[n];                                              // bayes.py:403
  void *d_desc, *d_cts, *d_tm, *d_abg, *d_dt, *d_v, *d_bgc, *d_dbg, *d_lp, *d_out, *d_obg;
  int rc;
  if ((rc = c->get("ws_obg", sizeof(double), &d_obg))) return rc;
  if ((rc = c->get("ws_desc", sizeof(TemplDesc) * (size_t)std::max<int64_t>(G, 1), &d_desc)) || (rc = c->get("ws_cts", sizeof(double) * (size_t)N, &d_cts)) ||
      (rc = c->get("ws_tm", sizeof(double) * (size_t)std::max<int64_t>(G, 1) * NP, &d_tm)) || (rc = c->get("ws_abg", sizeof(double) * (size_t)P * N, &d_abg)) ||
      (rc = c->get("ws_dt", sizeof(double) * (size_t)N, &d_dt)) || (rc = c->get("ws_v", sizeof(double) * (size_t)N, &d_v)) ||
      (rc = c->get("ws_bgc", sizeof(double) * MAXP * MAXP, &d_bgc)) || (rc = c->get("ws_dbg", sizeof(double) * MAXP, &d_dbg)) ||
      (rc = c->get("ws_lp", sizeof(double) * (size_t)std::max<int64_t>(G, 1), &d_lp)) ||
      (rc = c->get("ws_out", sizeof(double) * (size_t)std::max<int64_t>(G, 1) * N, &d_out)))
    return rc;
  if (G > 0) CK(cudaMemcpyAsync(d_desc, desc.data(), sizeof(TemplDesc) * (size_t)G, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_cts, cts, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_abg, abg.data(), sizeof(double) * (size_t)P * N, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_dt, dt.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_v, noisevar, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_bgc, bgcross, sizeof(double) * (size_t)P * P, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_dbg, dbgr, sizeof(double) * (size_t)P, cudaMemcpyHostToDevice, s));
  if (G > 0) {
    CK(cudaMemcpyAsync(d_lp, logprior, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice, s));
    CK(launch_gen_templates((const TemplDesc*)d_desc, (int)G, (const double*)d_cts, (int)N, NP, (double*)d_tm, s, &c->stats));
  }
  CK(launch_whole_series(P, (const double*)d_tm, (const double*)d_abg, (const double*)d_dt, (const double*)d_v,
                         (const double*)d_bgc, (const double*)d_dbg, (const double*)d_lp, (int)N, NP, (int)G,
                         halfrange ? 1 : 0, sumlogsigma, (double*)d_out, out_bg ? (double*)d_obg : nullptr, s, &c->stats));
  if (G > 0) CK(cudaMemcpyAsync(out_lnB, d_out, sizeof(double) * (size_t)G * N, cudaMemcpyDeviceToHost, s));
  if (out_bg) CK(cudaMemcpyAsync(out_bg, d_obg, sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));   // also keeps the host staging vectors alive until the copies are done
  return BFL_OK;
}

int bfl_logtrapz(bfl_ctx* c, const double* lx, const double* t, int64_t n, int64_t m, double* out) {
  if (!c || !lx || !t || !out || n < 1 || m < 1) return fail(BFL_ERR_ARG, "bad argument to bfl_logtrapz");
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  void *d_lx, *d_t, *d_o;
  int rc;
  if ((rc = c->get("lt_lx", sizeof(double) * (size_t)n * m, &d_lx))) return rc;
  if ((rc = c->get("lt_t", sizeof(double) * (size_t)n, &d_t))) return rc;
  if ((rc = c->get("lt_o", sizeof(double) * (size_t)m, &d_o))) return rc;
  CK(cudaMemcpyAsync(d_lx, lx, sizeof(double) * (size_t)n * m, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_t, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CK(launch_logtrapz((const double*)d_lx, (const double*)d_t, n, m, (double*)d_o, s, &c->stats));
  CK(cudaMemcpyAsync(out, d_o, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return BFL_OK;
}

int bfl_fp64_peak_probe(bfl_ctx* c, int32_t iters, double* tflops, double* ms) {
  if (!c || !tflops) return fail(BFL_ERR_ARG, "null argument to bfl_fp64_peak_probe");
  if (iters < 1) iters = 2048;
  CK(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  void* d_sink;
  int rc;
  if ((rc = c->get("probe", 64, &d_sink))) return rc;
  const int blocks = c->sm_count * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(launch_fp64_probe((double*)d_sink, iters / 4 + 1, blocks, s, &c->stats));   // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0, s));
    CK(launch_fp64_probe((double*)d_sink, iters, blocks, s, &c->stats));
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    best = std::min(best, t);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return BFL_OK;
}

}  // extern "C"
