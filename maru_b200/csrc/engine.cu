// engine.cu — the evaluator behind the C ABI (include/maru_b200.h): weight pack -> device buffers,
// activation planes, the launch sequence of one forward, CUDA-graph capture per batch bucket,
// pinned staging and the blocking / device-resident entry points.
//
// Mirrors deepgo::Model (reference src/deepgo/native/cpp/Model.cpp:62-100): construct from a file +
// gpu id, forward(float* in, float* out, n) on host buffers.  There is NO CPU path: gpu_id < 0, a
// missing device or a failed launch is an error code, never a fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached
#include <stdlib.h>

namespace maru {

// NVTX range over a host-side phase (tracing).  Visible in Nsight Systems / ncu --nvtx; the kernels of a
// forward are captured into a graph, so the ranges mark capture / enqueue / blocking spans, not kernel time.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// attention_tc_kernel is written for 32-channel heads (every BASELINE.json config); the v1 kernel that also ran
// 64-channel heads exists in -DMARU_EXPERIMENTS builds only
static bool head_dim_supported(int d) {
#ifdef MARU_EXPERIMENTS
  return d == 32 || d == 64;
#else
  return d == 32;
#endif
}

thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }

#define CK(expr)                                                                         \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

// ------------------------------------------------------------------------------------------
// weight pack (written by maru_b200/convert.py)
// ------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct PackHeader {
  char magic[8];  // "MARUB200"
  uint32_t version;
  uint32_t n_tensors;
  int32_t arch[8];  // blocks, channels, attn_every, head_dim, head_channels, value_channels, 0, 0
};
struct PackEntry {
  char name[48];
  uint32_t dtype;  // 0 = f32, 1 = bf16
  uint32_t ndim;
  uint32_t dims[4];
  uint64_t offset;  // from file start
  uint64_t nbytes;
};
#pragma pack(pop)

struct Pack {
  std::vector<uint8_t> data;
  PackHeader hdr;
  std::map<std::string, PackEntry> entries;
  const PackEntry* find(const std::string& name) const {
    auto it = entries.find(name);
    return it == entries.end() ? nullptr : &it->second;
  }
};

static int load_pack(const std::string& path_in, Pack& pk) {
  std::string path = path_in;
  const std::string ext = ".b200";
  if (path.size() < ext.size() || path.compare(path.size() - ext.size(), ext.size(), ext) != 0)
    path += ext;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    set_error("cannot open weight pack '" + path + "' (run: python -m maru_b200.convert <file>.model)");
    return 1;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < (long)sizeof(PackHeader)) {
    fclose(f);
    set_error("short read on weight pack '" + path + "'");
    return 1;
  }
  pk.data.resize(sz);
  const size_t rd = fread(pk.data.data(), 1, sz, f);
  fclose(f);
  if ((long)rd != sz || sz < (long)sizeof(PackHeader)) {
    set_error("short read on weight pack '" + path + "'");
    return 1;
  }
  memcpy(&pk.hdr, pk.data.data(), sizeof(PackHeader));
  if (memcmp(pk.hdr.magic, "MARUB200", 8) != 0 || pk.hdr.version != 1) {
    set_error("'" + path + "' is not a MARUB200 v1 weight pack");
    return 1;
  }
  const size_t table = sizeof(PackHeader) + (size_t)pk.hdr.n_tensors * sizeof(PackEntry);
  if ((size_t)sz < table) {
    set_error("corrupt weight pack table");
    return 1;
  }
  for (uint32_t i = 0; i < pk.hdr.n_tensors; i++) {
    PackEntry e;
    memcpy(&e, pk.data.data() + sizeof(PackHeader) + (size_t)i * sizeof(PackEntry), sizeof(PackEntry));
    e.name[47] = 0;
    uint64_t count = 1;
    bool dims_ok = e.ndim >= 1 && e.ndim <= 4 && e.dtype <= 1;
    for (uint32_t d = 0; dims_ok && d < e.ndim; d++) {
      if (e.dims[d] == 0 || e.dims[d] > (1u << 24)) dims_ok = false;
      else count *= e.dims[d];
      if (count > (1ull << 32)) dims_ok = false;
    }
    // overflow-safe bounds, and the byte count must be what the dims say (a truncated or crafted pack must not
    // make upload_* read past the file image)
    if (!dims_ok || e.offset > (uint64_t)sz || e.nbytes > (uint64_t)sz - e.offset ||
        e.nbytes != count * (e.dtype == 1 ? 2u : 4u)) {
      set_error(std::string("corrupt weight pack entry ") + e.name);
      return 1;
    }
    pk.entries[e.name] = e;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
static int choose_nt(int cin, int cout, int taps) {
  const int cands[3] = {128, 64, 32};
  for (int nt : cands) {
    if (cout % nt != 0) continue;
    // keep in sync with MARU_CONV_CASES (conv_gemm.cu)
    const bool ok =
        (taps == 9 && ((cin == 64 && (nt == 128 || nt == 64)) || (cin == 32 && nt == 32) ||
                       (cin == 128 && nt == 64))) ||
        (taps == 1 && ((cin == 256 && (nt == 128 || nt == 32)) ||
                       (cin == 128 && (nt == 128 || nt == 64 || nt == 32)) ||
                       (cin == 64 && (nt == 128 || nt == 64 || nt == 32)) ||
                       (cin == 32 && (nt == 64 || nt == 32))));
    if (!ok) continue;
    if (taps == 9 && cin == 64 && nt == 128 && cout % 128 != 0) continue;
    return nt;
  }
  return 0;
}

int Engine::upload_conv(const Pack& pk, const std::string& name, int cin, int cout, int taps,
                        ConvLayer& L) {
  const PackEntry* w = pk.find(name + ".w");
  const PackEntry* b = pk.find(name + ".b");
  if (!w || !b) {
    set_error("weight pack lacks tensor " + name);
    return 1;
  }
  if (w->dtype != 1 || w->ndim != 3 || (int)w->dims[0] != taps || (int)w->dims[1] != cout ||
      (int)w->dims[2] != cin || b->dtype != 0 || b->ndim != 1 || (int)b->dims[0] != cout) {
    set_error("weight pack tensor " + name + " has unexpected shape");
    return 1;
  }
  L.name = name;
  L.cin = cin;
  L.cout = cout;
  L.taps = taps;
  L.nt = choose_nt(cin, cout, taps);
  if (L.nt == 0) {
    char msg[160];
    snprintf(msg, sizeof msg, "no conv_gemm instantiation for %s (cin=%d cout=%d taps=%d)",
             name.c_str(), cin, cout, taps);
    set_error(msg);
    return 1;
  }
  // [tap][cout][cin] -> UMMA K-major SWIZZLE_NONE image [tap][cin/8][cout][8]
  const uint16_t* src = reinterpret_cast<const uint16_t*>(pk.data.data() + w->offset);
  std::vector<uint16_t> img((size_t)taps * cin * cout);
  const int kch = cin / 8;
  for (int t = 0; t < taps; t++)
    for (int c8 = 0; c8 < kch; c8++)
      for (int co = 0; co < cout; co++)
        for (int e = 0; e < 8; e++)
          img[(((size_t)t * kch + c8) * cout + co) * 8 + e] =
              src[((size_t)t * cout + co) * cin + c8 * 8 + e];
  CK(cudaMalloc(&L.w, img.size() * 2));
  CK(cudaMemcpy(L.w, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&L.b, (size_t)cout * 4));
  CK(cudaMemcpy(L.b, pk.data.data() + b->offset, (size_t)cout * 4, cudaMemcpyHostToDevice));
  allocs.push_back(L.w);
  allocs.push_back(L.b);

  // flow blob: per channel group [TAPS][KCH][nt][8] weights followed by the bias operand of the bias
  // MMA ([2 k-chunks][nt rows][8]: (hi, lo, 0...) in chunk 0) -> one bulk copy per CTA and layer
  const int cands[4] = {L.nt, 128, 64, 32};
  for (int nt : cands) {
    if (cout % nt == 0 && conv_flow_supports(cin, nt, taps)) {
      L.nt_flow = nt;
      break;
    }
  }
  L.nt_blob = L.nt_flow > 0 ? L.nt_flow : (conv_ksplit_supports(cin, L.nt, taps) ? L.nt : 0);
  const float* bias = reinterpret_cast<const float*>(pk.data.data() + b->offset);
  auto bf16_rn = [](float f) -> uint16_t {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);     // inf / nan
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
  };
  auto bf16_to_f = [](uint16_t h) -> float {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  // per group of nt output channels: [TAPS][KCH][nt][8] weights | bias operand image ([2 k-chunks][nt][8]: (hi, lo, 0...))
  auto make_blob = [&](int nt, uint8_t** dst) -> int {
    const int groups = cout / nt;
    const size_t wb = (size_t)taps * cin * nt * 2, bb = (size_t)2 * nt * 16;
    std::vector<uint8_t> blob((wb + bb) * groups, 0);
    for (int g = 0; g < groups; g++) {
      uint16_t* wdst = reinterpret_cast<uint16_t*>(blob.data() + (wb + bb) * g);
      for (int t = 0; t < taps; t++)
        for (int c8 = 0; c8 < kch; c8++)
          for (int r = 0; r < nt; r++)
            for (int e = 0; e < 8; e++)
              wdst[(((size_t)t * kch + c8) * nt + r) * 8 + e] =
                  src[((size_t)t * cout + g * nt + r) * cin + c8 * 8 + e];
      uint16_t* bdst = reinterpret_cast<uint16_t*>(blob.data() + (wb + bb) * g + wb);
      for (int r = 0; r < nt; r++) {
        const float bv = bias[g * nt + r];
        const uint16_t hi = bf16_rn(bv);
        const uint16_t lo = bf16_rn(bv - bf16_to_f(hi));
        bdst[(size_t)r * 8 + 0] = hi;
        bdst[(size_t)r * 8 + 1] = lo;
      }
    }
    CK(cudaMalloc(dst, blob.size()));
    CK(cudaMemcpy(*dst, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    allocs.push_back(*dst);
    return 0;
  };
  if (L.nt_blob > 0 && make_blob(L.nt_blob, &L.blob)) return 1;
  if (conv_wide_supports(cin, cout, taps) && make_blob(256, &L.blob_wide)) return 1;
  if (conv_pair_supports(cin, cout, taps)) {
    if (L.nt_blob == conv_pair_blob_nt(cout)) L.blob_pair = L.blob;
    else if (make_blob(conv_pair_blob_nt(cout), &L.blob_pair)) return 1;
  }
  return 0;
}

int Engine::upload_f32(const Pack& pk, const std::string& name, size_t count, float** dst) {
  const PackEntry* e = pk.find(name);
  if (!e || e->dtype != 0 || e->nbytes != count * 4) {
    set_error("weight pack lacks f32 tensor " + name);
    return 1;
  }
  CK(cudaMalloc(dst, count * 4));
  CK(cudaMemcpy(*dst, pk.data.data() + e->offset, count * 4, cudaMemcpyHostToDevice));
  allocs.push_back(*dst);
  return 0;
}

int Engine::alloc_act(int channels, bool blocked, Act* a) {
  // both layouts need (C/8) * 16 bytes per row; rows_alloc covers guard rows + whole tiles + slack
  const size_t bytes = (size_t)(channels / 8) * rows_alloc * 16;
  CK(cudaMalloc(&a->p, bytes));
  CK(cudaMemset(a->p, 0, bytes));
  allocs.push_back(a->p);
  a->channels = channels;
  a->blocked = blocked;
  return 0;
}

int Engine::init(const char* weights_path, int gpu, int max_batch_, unsigned flags_) {
  if (gpu < 0) {
    set_error("maru_b200 has no CPU path: gpu_id must be >= 0");
    return 1;
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (gpu >= ndev) {
    set_error("Specified GPU device is not available.");  // same text as Model.cpp:30
    return 1;
  }
  gpu_id = gpu;
#ifdef MARU_EXPERIMENTS
  if (const char* pe = getenv("MARU_B200_PAIR64")) {      // "0": never, "N": from batch N up
    pair64_on = pe[0] != '0' || pe[1] != 0;
    if (pair64_on) pair64_min_batch = atoi(pe);
  }
#endif
  if (const char* ff = getenv("MARU_B200_FLOW")) {
    // "1": the persistent flow kernels at every batch size, "0": never (default: for 249..354 positions on 148 SMs)
    if (ff[0] == '1') dbg_desc |= 128;
    if (ff[0] == '0') dbg_desc |= 32;
  }
  if (const char* ft = getenv("MARU_B200_FLOW_TUNE")) flow_tune_on = ft[0] != '0';
  device_ref(+1);
  dev_counted = true;
  max_batch = max_batch_ > 0 ? max_batch_ : 256;
  flags = flags_;
  CK(cudaSetDevice(gpu_id));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, gpu_id));
  if (prop.major != 10) {
    set_error(std::string("maru_b200 kernels are sm_100a only; device is ") + prop.name);
    return 1;
  }
  sm_count = prop.multiProcessorCount;
  CK(configure_conv_gemm());
  CK(configure_conv_flow());
  if (conv_flow_ctas_per_sm() < 1) flow_broken = true;   // the persistent kernels cannot be resident here: per-layer launches
  CK(configure_conv_ksplit());
  CK(configure_conv_wide());
  CK(configure_conv_pair());
  CK(configure_attention());
  CK(configure_misc_kernels());
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));

  Pack pk;
  if (load_pack(weights_path, pk)) return 1;
  blocks = pk.hdr.arch[0];
  C = pk.hdr.arch[1];
  attn_every = pk.hdr.arch[2];
  head_dim = pk.hdr.arch[3];
  Ch = pk.hdr.arch[4];
  Cv = pk.hdr.arch[5];
  if (blocks < 1 || blocks > 64 || C < 64 || C % 64 != 0 || C > 256 || !head_dim_supported(head_dim) || Ch != 32 ||
      Cv < 1 || Cv > 128 || attn_every < 0 || attn_every > blocks) {
    set_error("unsupported architecture in weight pack");
    return 1;
  }
  n_attn = attn_every > 0 ? blocks / attn_every : 0;
  rows_alloc = rows_alloc_for(max_batch);

  if (upload_conv(pk, "stem", STEM_CIN, C, 9, stem)) return 1;
  blk.resize(blocks);
  for (int i = 0; i < blocks; i++) {
    const std::string b = "blk" + std::to_string(i);
    if (upload_conv(pk, b + ".reduce", C, C / 2, 1, blk[i].reduce)) return 1;
    for (int j = 0; j < 2; j++) {
      const std::string in = b + ".in" + std::to_string(j);
      if (upload_conv(pk, in + ".c1", C / 2, C / 2, 9, blk[i].c[2 * j])) return 1;
      if (upload_conv(pk, in + ".c2", C / 2, C / 2, 9, blk[i].c[2 * j + 1])) return 1;
    }
    if (upload_conv(pk, b + ".expand", C / 2, C, 1, blk[i].expand)) return 1;
  }
  att.resize(n_attn);
  for (int i = 0; i < n_attn; i++) {
    const std::string a = "att" + std::to_string(i);
    if (upload_conv(pk, a + ".qkv", C, 3 * C, 1, att[i].qkv)) return 1;
    if (upload_conv(pk, a + ".proj", C, C, 1, att[i].proj)) return 1;
  }
  if (upload_conv(pk, "head.g", C, Ch, 1, headg)) return 1;
  if (upload_f32(pk, "head.planes.w", 6 * Ch, &w_planes)) return 1;
  if (upload_f32(pk, "head.planes.b", 6, &b_planes)) return 1;
  if (upload_f32(pk, "head.fc1.w", (size_t)Cv * 2 * Ch, &w_fc1)) return 1;
  if (upload_f32(pk, "head.fc1.b", Cv, &b_fc1)) return 1;
  if (upload_f32(pk, "head.fc2.w", 3 * Cv, &w_fc2)) return 1;
  if (upload_f32(pk, "head.fc2.b", 3, &b_fc2)) return 1;

  // h and o are only consumed tile-locally (1x1 convs, residuals) -> tile-blocked; x0, r, s feed 3x3
  // convs (halo) and qkv / g are read per position -> chunk-planar
  if (alloc_act(STEM_CIN, false, &x0)) return 1;
  if (alloc_act(C, true, &h)) return 1;
  if (alloc_act(C / 2, false, &r)) return 1;
  if (alloc_act(C / 2, false, &s)) return 1;
  if (n_attn > 0) {
    if (alloc_act(3 * C, false, &qkv)) return 1;
    if (alloc_act(C, true, &o)) return 1;
  }
  if (alloc_act(Ch, false, &g)) return 1;
  {
    n_tiles_max = (total_rows(max_batch) + TILE_M - 1) / TILE_M;
    Act* acts[7] = {&x0, &h, &r, &s, &qkv, &o, &g};
    flags_bytes = (size_t)7 * n_tiles_max * sizeof(unsigned);
    CK(cudaMalloc(&d_flags, flags_bytes));
    CK(cudaMemset(d_flags, 0, flags_bytes));
    allocs.push_back(d_flags);
    CK(cudaMalloc(&d_sync, 2 * 256 * sizeof(unsigned)));
    CK(cudaMemset(d_sync, 0, 2 * 256 * sizeof(unsigned)));
    allocs.push_back(d_sync);
    CK(cudaMalloc(&d_stamps, 512 * sizeof(long long)));
    CK(cudaMemset(d_stamps, 0, 512 * sizeof(long long)));
    allocs.push_back(d_stamps);
    for (int i = 0; i < 7; i++) acts[i]->flags = d_flags + (size_t)i * n_tiles_max;
  }
  CK(cudaMalloc(&mask, (size_t)rows_alloc));
  CK(cudaMemset(mask, 0, (size_t)rows_alloc));
  allocs.push_back(mask);
  CK(cudaMalloc(&mask_c, (size_t)rows_alloc));
  CK(cudaMemset(mask_c, 0, (size_t)rows_alloc));
  allocs.push_back(mask_c);
  if (const char* cc = getenv("MARU_B200_COMPACT")) compact_on = cc[0] != '0';
  CK(cudaMalloc(&d_trace, (size_t)256 * TRACE_TILES * TRACE_SLOTS * 8));
  CK(cudaMemset(d_trace, 0, (size_t)256 * TRACE_TILES * TRACE_SLOTS * 8));
  allocs.push_back(d_trace);
  CK(cudaMalloc(&d_timeline, 64 * 2 * 8 * 8));
  CK(cudaMemset(d_timeline, 0, 64 * 2 * 8 * 8));
  allocs.push_back(d_timeline);
  CK(cudaMalloc(&d_n, 64));
  allocs.push_back(d_n);
  CK(cudaMalloc(&d_logits, (size_t)max_batch * 2 * CELLS * 4));
  allocs.push_back(d_logits);
  CK(cudaMallocHost(&h_n, 64));
  for (int i = 0; i < 3; i++) {
    Slot& sl = slot[i];
    CK(cudaMalloc(&sl.d_rows, (size_t)max_batch * IN_ROW * 4));
    CK(cudaMalloc(&sl.d_states, (size_t)max_batch * sizeof(maru_state)));
    CK(cudaMalloc(&sl.d_out, (size_t)max_batch * OUT_ROW * 4));
    CK(cudaMallocHost(&sl.h_rows, (size_t)max_batch * IN_ROW * 4));
    CK(cudaMallocHost(&sl.h_states, (size_t)max_batch * sizeof(maru_state)));
    CK(cudaMallocHost(&sl.h_out, (size_t)max_batch * OUT_ROW * 4));
    CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    CK(cudaEventCreate(&sl.t0));
    CK(cudaEventCreate(&sl.t1));
  }

  // FLOP accounting: algorithmic at T = 361 tokens, executed at 400 padded rows
  {
    const double T = CELLS, Tp = POS_ROWS;
    const double c = C;
    flops_alg = 2 * T * IN_PLANES * c * 9 + 2.0 * IN_INFOS * c;
    flops_alg += blocks * (2 * (2 * T * c * (c / 2)) + 4 * (2 * T * (c / 2) * (c / 2) * 9));
    flops_alg += n_attn * (2 * T * c * 3 * c + 4 * T * T * c + 2 * T * c * c);
    flops_alg += 2 * T * c * Ch + 2 * T * Ch * 6 + 2.0 * (2 * Ch) * Cv + 2.0 * Cv * 3;
    flops_exec = 2 * Tp * STEM_CIN * c * 9;
    flops_exec += blocks * (2 * (2 * Tp * c * (c / 2)) + 4 * (2 * Tp * (c / 2) * (c / 2) * 9));
    flops_exec += n_attn * (2 * Tp * c * 3 * c + 4 * 384.0 * Tp * c + 2 * Tp * c * c);   // 3 query tiles x 400 keys
    flops_exec += 2 * Tp * c * Ch + 2 * T * Ch * 6 + 2.0 * (2 * Ch) * Cv + 2.0 * Cv * 3;
  }
  launches_per_forward = 1 + 1 + blocks * 6 + n_attn * 3 + 2;
  CK(cudaDeviceSynchronize());
  return 0;
}

Engine::~Engine() {
  if (dev_counted) device_ref(-1);
  if (gpu_id >= 0) cudaSetDevice(gpu_id);
  if (stream) cudaStreamSynchronize(stream);
  for (auto& kv : graphs)
    if (kv.second) cudaGraphExecDestroy(kv.second);
  for (void* p : allocs) cudaFree(p);
  for (int i = 0; i < 3; i++) {
    Slot& sl = slot[i];
    if (sl.d_rows) cudaFree(sl.d_rows);
    if (sl.d_states) cudaFree(sl.d_states);
    if (sl.d_out) cudaFree(sl.d_out);
    if (sl.h_rows) cudaFreeHost(sl.h_rows);
    if (sl.h_states) cudaFreeHost(sl.h_states);
    if (sl.h_out) cudaFreeHost(sl.h_out);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.t0) cudaEventDestroy(sl.t0);
    if (sl.t1) cudaEventDestroy(sl.t1);
  }
  if (h_n) cudaFreeHost(h_n);
  for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (stream) cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------
// the launch sequence of one forward (trunk input planes + mask already written)
// ------------------------------------------------------------------------------------------
int Engine::conv(const ConvLayer& L, Act& in, Act* res, Act& out, int relu, int n, const int* n_dev,
                 cudaStream_t st) {
  ConvParams p;
  memset(&p, 0, sizeof p);
  if (in.blocked && L.taps != 1) {
    set_error("internal: tile-blocked tensor fed to a 3x3 convolution");
    return 1;
  }
  p.in = in.p;
  p.in_blocked = in.blocked;
  p.in_planes = in.channels / 8;
  p.res_blocked = res ? res->blocked : 0;
  p.res_planes = res ? res->channels / 8 : 0;
  p.out_blocked = out.blocked;
  p.out_planes = out.channels / 8;
  p.in_chunk0 = 0;
  p.w = L.w;
  p.bias = L.b;
  p.blob = (L.nt_blob == L.nt) ? L.blob : nullptr;      // packed for this channel-group width
  p.res = res ? res->p : nullptr;
  p.res_chunk0 = 0;
  p.out = out.p;
  p.out_chunk0 = 0;
  p.mask = cur_mask();
  p.rows_alloc = rows_alloc;
  p.pitch = geom.pitch;
  p.pos_rows = geom.pos_rows;
  p.m_rows = total_rows_g(n, geom.pos_rows);
  p.cout_total = L.cout;
  p.relu = relu;
  p.n_dev = n_dev;
  p.stages = 0;
  p.pdl = (dbg_desc & 1) ? 0 : 1;
  // Per-tile cross-layer flags are an EXPERIMENT that lost ("what did not work"): the forward
  // is ~20 % slower with them than with plain grid dependencies, so they are off unless debug bit 3
  // is set; never while profiling (repeated launches of one layer).
#ifdef MARU_EXPERIMENTS
  const bool use_flags = p.pdl && prof == nullptr && (dbg_desc & 8);
#else
  const bool use_flags = false;             // per-tile cross-layer flags: experiment builds only
#endif
  p.in_flags = (use_flags && in.flagged) ? in.flags : nullptr;
  p.in_expect = in.version;
  p.out_flags = use_flags ? out.flags : nullptr;
  p.trace = (dbg_trace_launch >= 0 && launches_done == dbg_trace_launch && !(dbg_desc & 64)) ? d_trace : nullptr;
  p.timeline = (dbg_timeline && launches_done < 64) ? d_timeline : nullptr;
  p.seq = launches_done;
  if (dbg_stop_after >= 0 && launches_done >= dbg_stop_after) return 0;
  launches_done++;
  // Flow mode (default): consecutive conv layers are collected and launched as ONE persistent kernel by
  // flush_chain() (conv_flow.cu). Layers whose weights or slabs do not fit next to the chain's (e.g. the
  // 147 KB stem) run as one PDL launch each, and so does everything when a debug knob asks for per-layer
  // behaviour (stop-after, traces, timeline, tile flags) or another engine shares the device.
  const bool flowable = flow_enabled() && !use_flags && p.trace == nullptr && p.timeline == nullptr && L.nt_flow > 0 &&
                        ((p.m_rows + TILE_M - 1) / TILE_M + sm_count - 1) / sm_count <= FLOW_MAX_TILES;
  if (flowable) {
    const int groups = L.cout / L.nt_flow;
    for (int attempt = 0; attempt < 2; attempt++) {
      const int blob = (conv_flow_blob_bytes(L.cin, L.nt_flow, L.taps) + 1023) & ~1023;
      const int w_new = blob > pending_w ? blob : pending_w;
      bool fits = pending.n_layers + groups <= FLOW_MAX && conv_flow_fits(L.cin, L.nt_flow, L.taps, w_new);
      for (int k = 0; fits && k < pending.n_layers; k++) {
        const FlowLayer& Q = pending.layer[k];
        fits = conv_flow_fits(Q.cin, Q.nt, Q.taps, w_new);
      }
      if (fits) {
        for (int grp = 0; grp < groups; grp++) {
          FlowLayer& Q = pending.layer[pending.n_layers + grp];
          Q.in = p.in;
          Q.res = p.res;
          Q.out = p.out;
          Q.blob = L.blob + (size_t)grp * conv_flow_blob_bytes(L.cin, L.nt_flow, L.taps);
          Q.in_blocked = p.in_blocked;
          Q.in_planes = p.in_planes;
          Q.res_blocked = p.res_blocked;
          Q.res_planes = p.res_planes;
          Q.out_blocked = p.out_blocked;
          Q.out_planes = p.out_planes;
          Q.out_chunk0 = grp * (L.nt_flow / 8);
          Q.cin = L.cin;
          Q.nt = L.nt_flow;
          Q.taps = L.taps;
          Q.relu = relu;
          // every channel group reads the output of the conv before this one (none for the first conv of a chain)
          Q.dep_back = (pending.n_layers == 0) ? 0 : 1 + grp;
        }
        if (pending.n_layers == 0) {
          pending_single = p;
          pending_first = &L;
          pending.m_rows = p.m_rows;
          pending.pitch = p.pitch;
          pending.pos_rows = p.pos_rows;
          pending.mask = p.mask;
          pending.n_dev = n_dev;
          pending.pdl = p.pdl;
          pending_flops = pending_bytes = 0.0;
          pending_convs = 0;
        }
        pending.n_layers += groups;
        pending_convs++;
        pending_w = w_new;
        {
          const double T = (double)CELLS * n;
          const double cin_alg = (L.name == "stem") ? IN_PLANES : L.cin;     // 33 planes (+ 7 infos) are algorithmic
          pending_flops += 2.0 * T * cin_alg * L.cout * L.taps + ((L.name == "stem") ? 2.0 * IN_INFOS * L.cout * n : 0.0);
          pending_bytes += T * (cin_alg + L.cout + (res ? L.cout : 0)) * 2.0 + (double)L.taps * L.cin * L.cout * 2.0;
        }
        out.flagged = false;
        return 0;
      }
      if (pending.n_layers == 0) break;       // does not fit even alone: one-layer launch below
      if (flush_chain(st)) return 1;
    }
  }
  if (flush_chain(st)) return 1;
  {
    // algorithmic figures at T = 361 on-board tokens per position (stem: 33 planes + 7 infos)
    const double T = (double)CELLS * n;
    const double cin_alg = (L.name == "stem") ? IN_PLANES : L.cin;
    double fl = 2.0 * T * cin_alg * L.cout * L.taps;
    if (L.name == "stem") fl += 2.0 * IN_INFOS * L.cout * n;
    const double by = T * (cin_alg + L.cout + (res ? L.cout : 0)) * 2.0 + (double)L.taps * L.cin * L.cout * 2.0;
    prof_begin(L.name.c_str(), L.taps == 9 ? 1 : 2, L.cin, L.cout, L.taps, fl, by, st);
  }
  n_launched++;
  // 128 input channels, 3x3 (the C = 256 nets): the K-split kernel pipelines half-slabs where conv_gemm has room for one
  // stage only (debug bit 4: conv_gemm, A/B)
  const bool plain = !(dbg_desc & 16) && p.trace == nullptr && p.timeline == nullptr && !use_flags;
  const bool ksplit = plain && conv_ksplit_supports(L.cin, L.nt, L.taps) && p.blob != nullptr && !in.blocked;
  // 128 -> 128 3x3 on CTA pairs (cta_group::2): N = 128 per SM in 64 cycles per K step instead of two N = 64 MMAs of 48
  const bool pair = plain && !(dbg_desc & (1 << 20)) && conv_pair_supports(L.cin, L.cout, L.taps) && L.blob_pair != nullptr &&
                    !in.blocked && (res == nullptr || res->blocked == out.blocked) && pair_pays(L.cin, n);
  ConvParams pp = p;
  pp.blob = L.blob_pair;
  // 256 -> k * 256 channels, 1x1 (qkv and the output projection of the C = 256 nets): 256 output channels per CTA, half the
  // redundant input-tile loads of the 128-channel groups
  const bool wide = plain && conv_wide_supports(L.cin, L.cout, L.taps) && L.blob_wide != nullptr && in.blocked &&
                    (res == nullptr || res->blocked == out.blocked);
  ConvParams pw = p;
  pw.blob = L.blob_wide;
  for (int rep = 0; rep < reps_now(); rep++) {
    if (pair) CK(launch_conv_pair(pp, L.cin, sm_count, st));
    else if (ksplit) CK(launch_conv_ksplit(p, sm_count, st));
    else if (wide) CK(launch_conv_wide(pw, sm_count, st));
    else CK(launch_conv_gemm(p, L.cin, L.nt, L.taps, sm_count, st));
  }
  prof_end(st);
  if (use_flags) {
    out.version += 8u * (unsigned)(L.cout / L.nt);   // 8 epilogue warps per CTA, one CTA per channel group
    out.flagged = true;
  } else {
    out.flagged = false;
  }
  return 0;
}

void Engine::prof_begin(const char* name, int kind, int cin, int cout, int taps, double flops,
                        double bytes, cudaStream_t st) {
  if (prof == nullptr) return;
  maru_launch_time t;
  memset(&t, 0, sizeof t);
  snprintf(t.name, sizeof t.name, "%s", name);
  t.kind = kind;
  t.cin = cin;
  t.cout = cout;
  t.taps = taps;
  t.flops = flops;
  t.bytes = bytes;
  prof->push_back(t);
  while ((int)prof_ev.size() < 2 * (int)prof->size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    prof_ev.push_back(e);
  }
  cudaEventRecord(prof_ev[2 * (prof->size() - 1)], st);
}
// how often the launch being enqueued is repeated: prof_reps for the launch a profiling pass is timing, else once
int Engine::reps_now() const {
  if (prof == nullptr) return 1;
  return (prof_only < 0 || (int)prof->size() - 1 == prof_only) ? prof_reps : 1;
}
void Engine::prof_end(cudaStream_t st) {
  if (prof == nullptr) return;
  cudaEventRecord(prof_ev[2 * (prof->size() - 1) + 1], st);
}

__global__ void set_live_n_kernel(int* d_n, int n) { *d_n = n; }

// Holds the stream busy so that the CPU can enqueue every launch + event of the profiled forward
// ahead of the GPU: the event deltas then measure GPU time, not API submission time.
__global__ void spin_kernel(long long ns) {
  long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
}

int Engine::profile_states(const maru_state* d_states, float* d_out, int n, maru_launch_time* out,
                           int max, int* count) {
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(gpu_id));
  if (n <= 0 || n > max_batch) {
    set_error("maru_eval_profile_states: n out of range");
    return 1;
  }
  fwd = (compact_on && !(dbg_desc & (1 << 21))) ? Geom::board(dev_w, dev_h) : Geom();
  flow_live = flow_choice(n);
  std::vector<maru_launch_time> rec;
  rec.reserve(256);
  while ((int)prof_ev.size() < 512) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    prof_ev.push_back(e);
  }
  // One CLEAN forward per launch of the sequence: in pass i only launch i is repeated prof_reps times back to back
  // inside its CUDA-event pair, everything before it ran once, on the real activations. (Repeating EVERY launch in one
  // pass, as round 1 did, piles the in-place residual layers up tenfold per layer: the later kernels were then timed on
  // garbage — harmless for the convolutions, but attention takes a different path when the scores explode.)
  if (launch_all(KIND_STATES, d_states, d_out, n, nullptr, stream)) return 1;
  prof_reps = 10;
  int n_launches = 1;
  for (int i = 0; i < n_launches; i++) {
    std::vector<maru_launch_time> pass;
    pass.reserve(256);
    spin_kernel<<<1, 1, 0, stream>>>(1000000);   // 1 ms head start for the CPU: the event deltas measure GPU time
    prof = &pass;
    prof_only = i;
    const int rc = launch_all(KIND_STATES, d_states, d_out, n, nullptr, stream);
    prof = nullptr;
    prof_only = -1;
    if (rc) return 1;
    CK(cudaStreamSynchronize(stream));
    if (i == 0) n_launches = (int)pass.size();
    if (i >= (int)pass.size()) break;
    cudaEventElapsedTime(&pass[i].ms, prof_ev[2 * i], prof_ev[2 * i + 1]);
    pass[i].ms /= (float)prof_reps;              // average duration of one launch
    rec.push_back(pass[i]);
  }
  // the last pass repeated the last launch only: the buffers hold a clean forward
  *count = (int)rec.size();
  for (int i = 0; i < (int)rec.size() && i < max; i++) out[i] = rec[i];
  return 0;
}

int Engine::flush_chain(cudaStream_t st) {
  if (pending.n_layers == 0) return 0;
  n_launched++;
  const int nl = pending.n_layers;
  pending.n_layers = 0;
  pending_w = 0;
  if (pending_convs == 1 && nl == 1) {
    // a chain of one layer: the one-layer kernel (its launch boundary overlaps the neighbours' through PDL)
    prof_begin(pending_first->name.c_str(), pending_first->taps == 9 ? 1 : 2, pending_first->cin, pending_first->cout,
               pending_first->taps, pending_flops, pending_bytes, st);
    for (int rep = 0; rep < reps_now(); rep++)
      CK(launch_conv_gemm(pending_single, pending_first->cin, pending_first->nt, pending_first->taps, sm_count, st));
    prof_end(st);
    return 0;
  }
  pending.n_layers = nl;
  pending.rows_alloc = rows_alloc;
  pending.flags = d_sync;
  pending.stamps = (dbg_desc & 64) && stamps_used + 4 * nl + 4 <= 512 ? d_stamps + stamps_used : nullptr;
  stamps_used += 4 * nl + 4;
  // bring-up: debug key 2 (trace launch) = -(1 + cta) traces that CTA of the FIRST flow kernel of the forward
  pending.trace = (dbg_flow_trace_cta >= 0 && n_flow_launches == dbg_flow_trace_chain) ? d_trace : nullptr;
  pending.trace_cta = dbg_flow_trace_cta;
  pending.dbg = (dbg_desc >> 8) & 1023;
  char name[48];
  snprintf(name, sizeof name, "flow%d[%s..%d layers]", n_flow_launches++, pending_first->name.c_str(), pending_convs);
  prof_begin(name, 5, 0, 0, 0, pending_flops, pending_bytes, st);
  cudaError_t e = cudaSuccess;
  for (int rep = 0; rep < reps_now() && e == cudaSuccess; rep++) {
    pending.flag_base = flag_next;      // flags only grow: every launch of a forward gets its own value range
    flag_next += (unsigned)nl;
    e = launch_conv_flow(pending, sm_count, st);
  }
  prof_end(st);
  pending.n_layers = 0;
  CK(e);
  return 0;
}

// The flow kernel spins on flags written by the other CTAs of its grid: every CTA must become resident,
// which holds when the engine has the device to itself (one stream, kernels of a forward in order).
static std::mutex g_dev_mu;
static int g_engines_on_device[64] = {0};
void Engine::device_ref(int delta) {
  std::lock_guard<std::mutex> lock(g_dev_mu);
  if (gpu_id >= 0 && gpu_id < 64) g_engines_on_device[gpu_id] += delta;
}
// The persistent flow kernels win in a window around 5-7 tiles per CTA. Below it a layer inside the flow kernel is one
// serial store -> publish -> load -> MMA chain per tile and the three big launches lose the prologue overlap that PDL
// gives 34 small ones; above it the per-layer kernels amortise their launch boundary over 10+ tiles and their lighter
// epilogue (residual by identity MMA instead of L2 loads in the epilogue, ~7 % at batch 1024) wins.
// Measured (b4c128, forward in us, flow / per-layer): batch 1: 186 / 159, 64: 246 / 222, 128: 323 / 305, 192: 415 / 401,
// 224: 444 / 433, 256: 448-458 / 466, 320: 544 / 562, 384: 666 / 659, 512: 838 / 824, 1024: 1710 / 1617.
bool Engine::flow_pays(int n) const {
  const int n_tiles = (total_rows_g(n, fwd.pos_rows) + TILE_M - 1) / TILE_M;
  // C = 256 nets: only the 128 -> 256 expansions fit a chain (two-layer launches, measured 0.7 % slower at batch 256)
  return C <= 128 && 4 * n_tiles >= 21 * sm_count && 2 * n_tiles < 15 * sm_count;     // 5.25 <= tiles per CTA < 7.5: batch 249..354
}
// CTA pairs for the 64 -> 64 layers of the C = 128 nets: decided by measurement
bool Engine::pair_pays(int cin, int n) const {
  if (cin == 128) return true;
  return pair64_on && n >= pair64_min_batch;
}
bool Engine::flow_choice(int n) const {
  if (!flow_pays(n)) return false;
  const auto it = flow_tuned.find(fwd.w * 32 + fwd.h);
  return it == flow_tuned.end() || it->second != 0;
}
bool Engine::flow_enabled() {
  if (dbg_desc & 128) return flow_device_ok();      // debug: force the flow kernels at any batch size
  return flow_live && flow_device_ok();
}
bool Engine::flow_device_ok() {
  if (flow_broken) return false;
  if ((dbg_desc & 32) || dbg_stop_after >= 0 || dbg_timeline || dbg_trace_launch >= 0) return false;
  std::lock_guard<std::mutex> lock(g_dev_mu);
  return gpu_id >= 0 && gpu_id < 64 && g_engines_on_device[gpu_id] == 1;
}

void Engine::clear_graphs() {
  for (auto& kv : graphs)
    if (kv.second) cudaGraphExecDestroy(kv.second);
  graphs.clear();
  graph_launches.clear();
}

int Engine::launch_trunk(int kind, const void* d_in, void* d_out, int n, const int* n_dev, cudaStream_t st) {
  launches_done = 0;
  {
    // per-tile completion counters start from zero in every forward (also under graph replay)
    CK(cudaMemsetAsync(d_flags, 0, flags_bytes, st));
    CK(cudaMemsetAsync(d_sync, 0, 2 * 256 * sizeof(unsigned), st));
    flag_next = 0;
    n_flow_launches = 0;
    stamps_used = 0;
    pending.n_layers = 0;
    pending_w = 0;
    Act* acts[7] = {&x0, &h, &r, &s, &qkv, &o, &g};
    for (Act* a : acts) {
      a->version = 0;
      a->flagged = false;
    }
  }
  // The stem always runs on the reference's 19x19 frame (its input planes are laid out there, ko quirk included). A batch
  // of small boards then moves to the compact frame: the stem writes into `o` (free until the first attention layer)
  // and the crop kernel moves the board's rows to `h`; every later layer sees (w + 1)(h + 1) rows per position.
  geom = Geom();
  if (fwd.compact()) {
    if (conv(stem, x0, nullptr, o, 1, n, n_dev, st)) return 1;
    if (flush_chain(st)) return 1;
    if (!(dbg_stop_after >= 0 && launches_done >= dbg_stop_after)) {
      launches_done++;
      n_launched++;
      prof_begin("crop", 0, C, C, 0, 0.0, (double)fwd.w * fwd.h * C * 2.0 * 2.0 * n, st);
      for (int rep = 0; rep < reps_now(); rep++)
        CK(launch_crop(o.p, h.p, mask, mask_c, C, o.blocked ? 1 : 0, h.blocked ? 1 : 0, rows_alloc, fwd.w, fwd.h, n, n_dev,
                       (dbg_desc & 1) ? 0 : 1, st));
      prof_end(st);
    }
    geom = fwd;
  } else {
    if (conv(stem, x0, nullptr, h, 1, n, n_dev, st)) return 1;
  }
  last = fwd;
  int ai = 0;
  for (int i = 0; i < blocks; i++) {
    Block& B = blk[i];
    if (conv(B.reduce, h, nullptr, r, 1, n, n_dev, st)) return 1;
    for (int j = 0; j < 2; j++) {
      if (conv(B.c[2 * j], r, nullptr, s, 1, n, n_dev, st)) return 1;
      if (conv(B.c[2 * j + 1], s, &r, r, 1, n, n_dev, st)) return 1;  // r = relu(r + conv(s)), in place
    }
    if (conv(B.expand, r, &h, h, 1, n, n_dev, st)) return 1;         // h = relu(h + conv(r)), in place
    if (attn_every > 0 && (i + 1) % attn_every == 0 && ai < n_attn) {
      Attn& A = att[ai++];
      if (conv(A.qkv, h, nullptr, qkv, 0, n, n_dev, st)) return 1;
      AttnParams ap;
      ap.qkv = qkv.p;
      ap.out = o.p;
      ap.out_blocked = o.blocked;
      ap.mask = cur_mask();
      ap.rows_alloc = rows_alloc;
      ap.pitch = geom.pitch;
      ap.pos_rows = geom.pos_rows;
      ap.n = n;
      ap.channels = C;
      ap.head_dim = head_dim;
      ap.n_dev = n_dev;
      ap.pdl = (dbg_desc & 1) ? 0 : 1;
      ap.dbg = ((dbg_desc >> 1) & 3) | ((dbg_desc >> 21) & 4);     // bit23 -> force the shifted softmax
      ap.trace = (dbg_att_trace_cta >= 0 && ai == 1) ? d_trace : nullptr;
      ap.trace_cta = dbg_att_trace_cta;
      if (flush_chain(st)) return 1;
      if (!(dbg_stop_after >= 0 && launches_done >= dbg_stop_after)) {
        launches_done++;
        n_launched++;
        const double T = CELLS;
        prof_begin(("att" + std::to_string(ai - 1) + ".core").c_str(), 3, C, C, 0,
                   4.0 * T * T * C * n, T * 4.0 * C * 2.0 * n, st);
        for (int rep = 0; rep < reps_now(); rep++) CK(launch_attention(ap, st));
        prof_end(st);
        o.flagged = false;   // written by a kernel without tile flags: consumers use the grid dependency
      }
      if (conv(A.proj, o, &h, h, 0, n, n_dev, st)) return 1;          // h = h + proj(o), in place
    }
  }
  if (conv(headg, h, nullptr, g, 1, n, n_dev, st)) return 1;
  HeadParams hp;
  hp.g = g.p;
  hp.mask = cur_mask();
  hp.pitch = geom.pitch;
  hp.pos_rows = geom.pos_rows;
  hp.board_w = geom.w;
  hp.board_h = geom.h;
  hp.w_planes = w_planes;
  hp.b_planes = b_planes;
  hp.w_fc1 = w_fc1;
  hp.b_fc1 = b_fc1;
  hp.w_fc2 = w_fc2;
  hp.b_fc2 = b_fc2;
  hp.out = (kind == KIND_LEAVES) ? nullptr : static_cast<float*>(d_out);
  hp.leaf = (kind == KIND_LEAVES) ? static_cast<maru_leaf*>(d_out) : nullptr;
  hp.states = (kind == KIND_LEAVES) ? static_cast<const maru_state*>(d_in) : nullptr;
  hp.logits = d_logits;
  hp.rows_alloc = rows_alloc;
  hp.n = n;
  hp.ch = Ch;
  hp.cv = Cv;
  hp.n_dev = n_dev;
  hp.pdl = (dbg_desc & 1) ? 0 : 1;
  if (flush_chain(st)) return 1;
  if (dbg_stop_after >= 0) return 0;
  n_launched++;
  prof_begin("heads", 4, Ch, 6, 0, (2.0 * CELLS * Ch * 6 + 2.0 * 2 * Ch * Cv + 2.0 * Cv * 3) * n,
             ((double)CELLS * Ch * 2 + (double)out_row_bytes(kind)) * n, st);
  for (int rep = 0; rep < reps_now(); rep++) CK(launch_heads(hp, st));
  prof_end(st);
  return 0;
}

int Engine::launch_all(int kind, const void* d_in, void* d_out, int n, const int* n_dev,
                       cudaStream_t st) {
  n_launched = 1;   // the encode / ingest kernel
  NvtxRange whole("maru.forward.enqueue");
  {
  NvtxRange enc("maru.encode");
  if (kind != KIND_PLANES) {
    prof_begin("encode", 0, 0, IN_PLANES, 0, 0.0,
               ((double)sizeof(maru_state) + (double)CELLS * IN_PLANES * 2.0) * n, st);
    for (int rep = 0; rep < reps_now(); rep++)
      CK(launch_encode_trunk(static_cast<const maru_state*>(d_in), x0.p, mask, rows_alloc, n, n_dev, st));
    prof_end(st);
  } else {
    prof_begin("ingest", 0, 0, IN_PLANES, 0, 0.0, ((double)IN_ROW * 4.0 + (double)CELLS * IN_PLANES * 2.0) * n,
               st);
    CK(launch_ingest_rows(static_cast<const float*>(d_in), x0.p, mask, rows_alloc, n, n_dev, st));
    prof_end(st);
  }
  }
  NvtxRange trunk("maru.trunk+heads");
  const int rc = launch_trunk(kind, d_in, d_out, n, n_dev, st);
  if (rc == 0 && prof == nullptr && dbg_stop_after < 0) launches_per_forward = n_launched;
  return rc;
}

static int bucket_for(int n, int max_batch) {
  int b = 1;
  while (b < n) b <<= 1;
  return b > max_batch ? max_batch : b;
}

// Enqueue one forward of n <= max_batch positions on `stream` (no host sync).
void Engine::uniform_board(const maru_state* s, int n, int* bw, int* bh) {
  *bw = *bh = 0;
  if (s == nullptr || n <= 0) return;
  const int w = s[0].width, h = s[0].height;
  for (int i = 1; i < n; i++)
    if (s[i].width != w || s[i].height != h) return;     // mixed sizes: the 19x19 frame serves them all
  *bw = w;
  *bh = h;
}

// Plane 32 of a reference row is 1 on the w x h board centred in the 19x19 frame (Board.cpp:496-497, 596-601): the middle
// frame row always crosses the board, so w = its number of ones and h = (all ones) / w.
void Engine::uniform_board_rows(const float* rows, int n, int* bw, int* bh) {
  *bw = *bh = 0;
  int w0 = 0, h0 = 0;
  for (int i = 0; i < n; i++) {
    const float* m = rows + (size_t)i * IN_ROW + (size_t)(IN_PLANES - 1) * CELLS;
    int w = 0, total = 0;
    for (int x = 0; x < BOARD; x++) w += m[(BOARD / 2) * BOARD + x] != 0.0f;
    for (int c = 0; c < CELLS; c++) total += m[c] != 0.0f;
    if (w == 0 || total % w != 0) return;                 // not a board mask: the 19x19 frame is always right
    const int h = total / w;
    if (i == 0) {
      w0 = w;
      h0 = h;
    } else if (w != w0 || h != h0) {
      return;
    }
  }
  *bw = w0;
  *bh = h0;
}

// One forward on the stream with the path flow_live says (graph replay, or plain launches with MARU_FLAG_NO_GRAPH).
// *ok = false: the launch failed (error text set), the stream is intact; return 1: a CUDA call failed.
int Engine::enqueue_forward(int kind, const void* d_in, void* d_out, int n, bool* ok) {
  const bool with_flow = flow_enabled();
  *ok = true;
  if (flags & MARU_FLAG_NO_GRAPH) {
    *ok = launch_all(kind, d_in, d_out, n, nullptr, stream) == 0;
  } else {
    // one graph per (kind, bucket, input, output); the live n is read from d_n by every kernel
    const int bucket = bucket_for(n, max_batch);
    GraphKey key{kind, bucket, d_in, d_out, with_flow, fwd.w, fwd.h};
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      // The cache is keyed by the caller's buffers: a device-API caller that rotates buffers must not grow it
      // without bound. Graphs still in flight keep running (a cudaGraphExec_t may be destroyed once launched work
      // has been submitted only after it completes, hence the sync).
      if (graphs.size() >= 64) {
        CK(cudaStreamSynchronize(stream));
        clear_graphs();
      }
      cudaGraph_t gr = nullptr;
      cudaGraphExec_t ex = nullptr;
      CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      const int rc = launch_all(kind, d_in, d_out, bucket, d_n, stream);
      cudaError_t ce = cudaStreamEndCapture(stream, &gr);
      if (rc == 0 && ce == cudaSuccess) ce = cudaGraphInstantiate(&ex, gr, 0);
      if (gr != nullptr) cudaGraphDestroy(gr);
      if (rc != 0 || ce != cudaSuccess) {
        if (rc == 0) set_error(std::string("forward graph: ") + cudaGetErrorString(ce));
        *ok = false;
      } else {
        it = graphs.emplace(key, ex).first;
        graph_launches[key] = n_launched;
      }
    }
    if (*ok) {
      // The live batch size travels BY VALUE as a kernel parameter (fixed at enqueue time): any number of forwards
      // with different n may be in flight on the stream. (A pinned staging int read by an async copy at execution
      // time was overwritten by later calls once more than a handful of forwards were outstanding.)
      set_live_n_kernel<<<1, 1, 0, stream>>>(d_n, n);
      CK(cudaGetLastError());
      const cudaError_t le = cudaGraphLaunch(it->second, stream);
      if (le == cudaSuccess) launches_per_forward = graph_launches[key];
      if (le != cudaSuccess) {
        set_error(std::string("cudaGraphLaunch: ") + cudaGetErrorString(le));
        cudaGraphExecDestroy(it->second);
        graphs.erase(it);
        *ok = false;
      }
    }
  }
  return 0;
}

// Times both paths on this forward's own inputs (a forward recomputes everything from its input, so repeating it is
// harmless) and records the faster for the board geometry. ~8 forwards and one synchronisation, once.
int Engine::tune_flow(int kind, const void* d_in, void* d_out, int n) {
  const int key = fwd.w * 32 + fwd.h;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms[2] = {0.f, 0.f};
  int rc = 0;
  for (int v = 1; v >= 0 && rc == 0; v--) {        // 1: flow kernels, 0: one launch per layer
    flow_tuned[key] = v;
    flow_live = flow_choice(n);
    bool ok = true;
    for (int r = 0; r < 2 && rc == 0 && ok; r++) rc = enqueue_forward(kind, d_in, d_out, n, &ok);     // capture + warm
    if (rc == 0 && ok) rc = cudaEventRecord(e0, stream) != cudaSuccess;
    for (int r = 0; r < 3 && rc == 0 && ok; r++) rc = enqueue_forward(kind, d_in, d_out, n, &ok);
    if (rc == 0 && ok) rc = cudaEventRecord(e1, stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess;
    if (rc == 0 && ok) cudaEventElapsedTime(&ms[v], e0, e1);
    if (rc == 0 && !ok) {
      // the flow kernels cannot be launched here: forward_device's fallback handles (and reports) it
      cudaGetLastError();
      ms[v] = 1e30f;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc) {
    flow_tuned.erase(key);
    return 1;
  }
  flow_tuned[key] = ms[1] <= ms[0] ? 1 : 0;
  if (getenv("MARU_B200_VERBOSE"))
    fprintf(stderr, "maru_b200: GPU %d, %dx%d frame, %d positions: flow kernels %.1f us, one launch per layer %.1f us\n", gpu_id,
            fwd.w, fwd.h, n, ms[1] * 1e3f / 3, ms[0] * 1e3f / 3);
  return 0;
}

int Engine::forward_device(int kind, const void* d_in, void* d_out, int n, int bw, int bh) {
  if (n <= 0) return 0;
  // a batch whose positions all hold the same small board runs on that board's compact frame after the stem
  fwd = (compact_on && !(dbg_desc & (1 << 21))) ? Geom::board(bw, bh) : Geom();
  if (n > max_batch) {
    set_error("batch larger than max_batch");
    return 1;
  }
  if (flow_tune_on && !(flags & MARU_FLAG_NO_GRAPH) && !(dbg_desc & (128 | 32)) && flow_pays(n) && flow_device_ok() &&
      flow_tuned.find(fwd.w * 32 + fwd.h) == flow_tuned.end()) {
    if (tune_flow(kind, d_in, d_out, n)) return 1;
  }
  // Two attempts: if the persistent flow kernels cannot be launched (the cooperative launch does not fit: the
  // device is shared), the engine gives them up for good and the forward is enqueued again with one launch per layer.
  for (int attempt = 0; attempt < 2; attempt++) {
    flow_live = flow_choice(n);             // by the LIVE batch size: a bucket has a flow graph and a per-layer graph
    const bool with_flow = flow_enabled();
    bool ok = true;
    if (enqueue_forward(kind, d_in, d_out, n, &ok)) return 1;
    if (ok) break;
    if (!with_flow || attempt == 1) return 1;
    cudaGetLastError();                     // launch-configuration errors are not sticky
    flow_broken = true;                     // flow_device_ok() is false from here on
    fprintf(stderr, "maru_b200: persistent conv kernels unavailable on GPU %d (%s); using one launch per layer\n",
            gpu_id, g_last_error.c_str());
  }
  n_forward++;
  n_positions += n;
  return 0;
}

// Blocking forward on host buffers: H2D, kernels, D2H (Model::forward, Model.cpp:88-100).
// `in` / `out` may be pageable (the reference passes stack arrays, Evaluator.cpp:48-49).
// The compact transports are small enough to cross PCIe INSIDE the kernels that use them: 448-byte records are read by
// the encode kernel (and the heads kernel's mask lookup) straight from pinned host memory, 1456-byte results are written
// by the heads kernel straight into it — no copy operation, no launch gap before the first and after the last kernel of a
// forward (a batch of 256: 115 KB up, 373 KB down; the two copies and their gaps were ~20 us of a 496 us blocking call).
// Only for page-locked memory (cudaMallocHost / cudaHostRegister: the caller's, or the queue's staging); pageable buffers
// (the reference passes stack arrays, Evaluator.cpp:48-49) and the large fp32 transports keep the staged copies.
static bool zero_copy_on() {
  static const bool on = [] {
    const char* v = getenv("MARU_B200_ZEROCOPY");
    return v == nullptr || v[0] != '0';
  }();
  return on;
}
static void* mapped_host(const void* p) {
  if (!zero_copy_on()) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int Engine::forward_host(int kind, const void* in, void* out, int n) {
  NvtxRange range("maru.forward_host");
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(gpu_id));
  const size_t in_row = in_row_bytes(kind), out_row = out_row_bytes(kind);
  uint8_t* const zin = kind != KIND_PLANES ? static_cast<uint8_t*>(mapped_host(in)) : nullptr;
  uint8_t* const zout = kind == KIND_LEAVES ? static_cast<uint8_t*>(mapped_host(out)) : nullptr;
  for (int off = 0; off < n; off += max_batch) {
    const int m = (n - off < max_batch) ? n - off : max_batch;
    Slot& sl = slot[2];
    void* d_in = kind != KIND_PLANES ? (void*)sl.d_states : (void*)sl.d_rows;
    if (zin != nullptr) {
      d_in = zin + (size_t)off * in_row;
    } else {
      CK(cudaMemcpyAsync(d_in, static_cast<const uint8_t*>(in) + (size_t)off * in_row, (size_t)m * in_row,
                         cudaMemcpyHostToDevice, stream));
    }
    void* const d_res = zout != nullptr ? (void*)(zout + (size_t)off * out_row) : (void*)sl.d_out;
    CK(cudaEventRecord(sl.t0, stream));
    int bw = 0, bh = 0;
    if (kind != KIND_PLANES) uniform_board(static_cast<const maru_state*>(in) + off, m, &bw, &bh);
    else uniform_board_rows(static_cast<const float*>(in) + (size_t)off * IN_ROW, m, &bw, &bh);
    if (forward_device(kind, d_in, d_res, m, bw, bh)) return 1;
    CK(cudaEventRecord(sl.t1, stream));
    if (zout == nullptr)
      CK(cudaMemcpyAsync(static_cast<uint8_t*>(out) + (size_t)off * out_row, sl.d_out, (size_t)m * out_row,
                         cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, sl.t0, sl.t1);
    device_ms_last = ms;
  }
  return 0;
}

int Engine::enqueue_slot(int kind, int slot_idx, int n) {
  NvtxRange range("maru.queue.batch.enqueue");
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(gpu_id));
  Slot& sl = slot[slot_idx];
  const size_t in_row = in_row_bytes(kind);
  void* d_in = kind != KIND_PLANES ? (void*)sl.d_states : (void*)sl.d_rows;
  const void* h_in = kind != KIND_PLANES ? (const void*)sl.h_states : (const void*)sl.h_rows;
  // the staging is page-locked: records in and compact results out cross PCIe inside the kernels (see forward_host)
  const bool zin = kind != KIND_PLANES && zero_copy_on(), zout = kind == KIND_LEAVES && zero_copy_on();
  if (zin) d_in = sl.h_states;
  else CK(cudaMemcpyAsync(d_in, h_in, (size_t)n * in_row, cudaMemcpyHostToDevice, stream));
  CK(cudaEventRecord(sl.t0, stream));
  int bw = 0, bh = 0;
  if (kind != KIND_PLANES) uniform_board(sl.h_states, n, &bw, &bh);
  else uniform_board_rows(sl.h_rows, n, &bw, &bh);
  if (forward_device(kind, d_in, zout ? (void*)sl.h_out : (void*)sl.d_out, n, bw, bh)) return 1;
  CK(cudaEventRecord(sl.t1, stream));
  if (!zout) CK(cudaMemcpyAsync(sl.h_out, sl.d_out, (size_t)n * out_row_bytes(kind), cudaMemcpyDeviceToHost, stream));
  CK(cudaEventRecord(sl.done, stream));
  return 0;
}

int Engine::wait_slot(int slot_idx, float* device_ms) {
  NvtxRange range("maru.queue.batch.wait");
  Slot& sl = slot[slot_idx];
  CK(cudaEventSynchronize(sl.done));
  if (device_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, sl.t0, sl.t1);
    *device_ms = ms;
  }
  return 0;
}

// ---- evaluator lanes -------------------------------------------------------------------------------------------
// The device side of a lane is the evaluator's own staging (slot 2): every batch is one H2D -> forward -> D2H sequence
// enqueued atomically under the evaluator's mutex on its one stream, so the D2H of batch i has drained the device
// buffers before the H2D of batch i+1 overwrites them, whichever lane or blocking call i+1 comes from.
Lane* Engine::lane_create(int rows) {
  if (rows < 1 || rows > max_batch) {
    set_error("maru_eval_lane_create: rows outside 1..max_batch");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (cudaSetDevice(gpu_id) != cudaSuccess) return nullptr;
  Lane* l = new (std::nothrow) Lane();
  if (l == nullptr) return nullptr;
  l->e = this;
  l->rows = rows;
  cudaError_t er = cudaMallocHost(&l->h_states, (size_t)rows * sizeof(maru_state));
  if (er == cudaSuccess) er = cudaMallocHost(&l->h_leaves, (size_t)rows * sizeof(maru_leaf));
  if (er == cudaSuccess) er = cudaEventCreateWithFlags(&l->done, cudaEventBlockingSync | cudaEventDisableTiming);
  if (er == cudaSuccess) er = cudaEventCreate(&l->t0);
  if (er == cudaSuccess) er = cudaEventCreate(&l->t1);
  if (er != cudaSuccess) {
    set_error(std::string("maru_eval_lane_create: ") + cudaGetErrorString(er));
    if (l->h_states) cudaFreeHost(l->h_states);
    if (l->h_leaves) cudaFreeHost(l->h_leaves);
    if (l->done) cudaEventDestroy(l->done);
    if (l->t0) cudaEventDestroy(l->t0);
    if (l->t1) cudaEventDestroy(l->t1);
    delete l;
    return nullptr;
  }
  return l;
}

void Engine::lane_destroy(Lane* l) {
  if (l == nullptr) return;
  std::lock_guard<std::mutex> lock(mu);
  cudaSetDevice(gpu_id);
  if (l->inflight) cudaEventSynchronize(l->done);
  cudaFreeHost(l->h_states);
  cudaFreeHost(l->h_leaves);
  cudaEventDestroy(l->done);
  cudaEventDestroy(l->t0);
  cudaEventDestroy(l->t1);
  delete l;
}

int Engine::lane_submit(Lane* l, int n) {
  if (n < 1 || n > l->rows || l->inflight) {
    set_error("maru_lane_submit: n outside 1..max_rows, or the lane's previous batch has not been waited for");
    return 1;
  }
  NvtxRange range("maru.lane.submit");
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(gpu_id));
  Slot& sl = slot[2];
  // (staged copies, not the zero-copy transport of forward_host: a forward graph is cached per buffer pair, and dozens of
  // lanes would each want their own)
  CK(cudaMemcpyAsync(sl.d_states, l->h_states, (size_t)n * sizeof(maru_state), cudaMemcpyHostToDevice, stream));
  CK(cudaEventRecord(l->t0, stream));
  int bw = 0, bh = 0;
  uniform_board(l->h_states, n, &bw, &bh);
  if (forward_device(KIND_LEAVES, sl.d_states, sl.d_out, n, bw, bh)) return 1;
  CK(cudaEventRecord(l->t1, stream));
  CK(cudaMemcpyAsync(l->h_leaves, sl.d_out, (size_t)n * sizeof(maru_leaf), cudaMemcpyDeviceToHost, stream));
  CK(cudaEventRecord(l->done, stream));
  l->n = n;
  l->inflight = true;
  return 0;
}

int Engine::lane_wait(Lane* l, float* device_ms) {
  if (!l->inflight) return 0;
  CK(cudaEventSynchronize(l->done));
  l->inflight = false;
  if (device_ms) {
    float ms = 0;
    cudaEventElapsedTime(&ms, l->t0, l->t1);
    *device_ms = ms;
  }
  return 0;
}

int Engine::lane_poll(Lane* l, int* done) {
  *done = 1;
  if (!l->inflight) return 0;
  const cudaError_t er = cudaEventQuery(l->done);
  if (er == cudaErrorNotReady) {
    *done = 0;
    return 0;
  }
  CK(er);
  return 0;
}

// K1 alone on host buffers: records -> reference rows (bit-exact Board::getInputs)
int Engine::encode_rows_host(const maru_state* st, float* out_rows, int n) {
  std::lock_guard<std::mutex> lock(mu);
  CK(cudaSetDevice(gpu_id));
  for (int off = 0; off < n; off += max_batch) {
    const int m = (n - off < max_batch) ? n - off : max_batch;
    Slot& sl = slot[2];
    CK(cudaMemcpyAsync(sl.d_states, st + off, (size_t)m * sizeof(maru_state), cudaMemcpyHostToDevice,
                       stream));
    CK(launch_encode_rows(sl.d_states, sl.d_rows, m, nullptr, stream));
    CK(cudaMemcpyAsync(out_rows + (size_t)off * IN_ROW, sl.d_rows, (size_t)m * IN_ROW * 4,
                       cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  return 0;
}

}  // namespace maru
