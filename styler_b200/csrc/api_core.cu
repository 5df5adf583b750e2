// C-ABI plumbing: version, thread-local error string, launch counter, TMA descriptor cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>
#include <unordered_map>

#include "common.cuh"
#include "tmap.cuh"

namespace sb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

int num_sms() {
  static std::atomic<int> cache[64];   // zero-initialised: 0 = not queried yet
  const int dev = current_device();
  if (dev >= 0 && dev < 64) {
    const int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

// ---- debug timeline -------------------------------------------------------------------------------------------
namespace {
struct TraceRec { const char* name; int a, b, c, d; cudaStream_t stream; cudaEvent_t e0, e1; };
std::atomic<int> g_trace_on{0};
std::vector<TraceRec> g_trace;
std::mutex g_trace_mu;
}  // namespace

TraceScope::TraceScope(const char* name, void* stream, int a, int b, int c, int d) : idx(-1) {
  if (g_trace_on.load(std::memory_order_relaxed) == 0) return;
  TraceRec r{name, a, b, c, d, static_cast<cudaStream_t>(stream), nullptr, nullptr};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, r.stream);
  std::lock_guard<std::mutex> lk(g_trace_mu);
  idx = static_cast<int>(g_trace.size());
  g_trace.push_back(r);
}
TraceScope::~TraceScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  cudaEventRecord(g_trace[idx].e1, g_trace[idx].stream);
}

namespace {
struct TuneDef { const char* name; int dflt, lo, hi; };
// TC_2CTA: CTA pairs for the big bf16 convs (0 off | 1 where it pays | 2 wherever legal: tests);  TC_PERSIST: persistent conv
// form (0 | 1 two CTAs/SM | 2 also one CTA/SM);  CONV_WIN: window kernel for <= 64-channel convs;  TC_BN / TC_SMEM_KB: tile and
// pipeline-depth overrides;  PDL: programmatic dependent launch (off: measured 8.71 ms/step with it vs 8.07 without);
// ATTN_PERSIST: persistent attention CTAs (0 | 1);  TC_WIDE: full-width N tile (two MMAs per k-step) for 256 < N <= 512;
// LSTM_MULTI: four utterances per BiLSTM CTA for batches >= 16 (0 | 1);  ATTN_POLY: eighths of the softmax exponentials computed
// by an FMA-pipe polynomial instead of MUFU.EX2 (0 | 2 | 3 | 4; 16-bit operands);  LSTM_MMA: BiLSTM recurrence on mma.sync,
// sixteen utterances per CTA (16-bit activations, batches >= 8);  STFT_OCC: STFT shape with three CTAs per SM (16 frames per CTA,
// magnitudes aliased onto the FFT exchange buffer: 1) or two CTAs of twelve warps (24 frames per CTA: 2) instead of two CTAs of
// eight warps (0); 3 = shape 1 with the samples read straight from global memory by the frame's own warp (no staged block
// window, one barrier per item); bitwise-equal results
const TuneDef kTune[TUNE_COUNT] = {{"TC_2CTA", 1, 0, 2}, {"TC_PERSIST", 1, 0, 2}, {"CONV_WIN", 1, 0, 1}, {"TC_BN", 0, 0, 256},
                                   {"TC_SMEM_KB", 113, 64, 220}, {"PDL", 0, 0, 1}, {"ATTN_PERSIST", 1, 0, 1}, {"TC_WIDE", 1, 0, 1},
                                   {"LSTM_MULTI", 0, 0, 1}, {"ATTN_POLY", 2, 0, 4},
                                   {"LSTM_MMA", 1, 0, 1}, {"STFT_OCC", 3, 0, 3}};
std::atomic<int> g_tune[TUNE_COUNT];          // 0 = not resolved yet, else value + 1
}  // namespace

int tuning(Tuning t) {
  int v = g_tune[t].load(std::memory_order_relaxed);
  if (v == 0) {
    const std::string env = std::string("STYLER_") + kTune[t].name;
    const char* e = getenv(env.c_str());
    int x = e != nullptr ? atoi(e) : kTune[t].dflt;
    if (x < kTune[t].lo || x > kTune[t].hi) x = kTune[t].dflt;
    v = x + 1;
    g_tune[t].store(v, std::memory_order_relaxed);
  }
  return v - 1;
}

bool pdl_enabled() { return tuning(TUNE_PDL) == 1; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- cuTensorMapEncodeTiled via the runtime's driver-entry-point lookup ------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  uint64_t v[16];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return static_cast<size_t>(h);
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

int make_tmap(CUtensorMap* out, const void* base, int elem, int rank, const uint64_t* dims, const uint64_t* strides,
              const uint32_t* box) {
  SB_REQUIRE(rank >= 2 && rank <= 4, "make_tmap: rank %d unsupported", rank);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(elem) << 8) | static_cast<uint64_t>(rank);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = i + 1 < rank ? strides[i] : 0;
    key.v[10 + i] = box[i];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn enc = get_encode_fn();
  SB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver / GPU?)");
  const int es = elem == 1 ? 2 : 4;
  SB_REQUIRE((reinterpret_cast<uint64_t>(base) & 15) == 0, "TMA base pointer %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    SB_REQUIRE(strides[i] % 16 == 0, "TMA stride[%d]=%llu bytes not a multiple of 16", i,
               (unsigned long long)strides[i]);
  SB_REQUIRE(box[0] * es == 128, "TMA inner box must span 128 bytes (got %u)", box[0] * es);
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], est[4];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; est[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUtensorMapDataType dt = elem == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                         : elem == 0 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32
                                     : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u)",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1]);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() > 8192) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return 0;
}

}  // namespace sb

extern "C" {
int styler_set_tuning(const char* name, int32_t value) {
  using namespace sb;
  SB_REQUIRE(name != nullptr, "set_tuning: null name");
  for (int i = 0; i < TUNE_COUNT; ++i) {
    if (strcmp(name, kTune[i].name) == 0) {
      if (value < 0) { g_tune[i].store(0); return 0; }          // back to the environment / default
      SB_REQUIRE(value >= kTune[i].lo && value <= kTune[i].hi, "set_tuning: %s=%d outside [%d, %d]", name, value, kTune[i].lo, kTune[i].hi);
      g_tune[i].store(value + 1);
      return 0;
    }
  }
  set_error("set_tuning: unknown switch %s", name);
  return -1;
}
// Debug timeline: styler_debug_trace(1) clears and starts recording, (0) stops.  styler_debug_trace_dump synchronises the
// device and writes one text line per leaf call -- "name a b c d stream start_ms end_ms" (times relative to the first record)
// -- into buf; returns the number of bytes needed (call again with a larger buffer if it exceeds cap).
int styler_debug_trace(int32_t on) {
  using namespace sb;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (on) {
    for (auto& r : g_trace) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_trace.clear();
  }
  g_trace_on.store(on ? 1 : 0);
  return 0;
}
int64_t styler_debug_trace_dump(char* buf, int64_t cap) {
  using namespace sb;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_trace_mu);
  std::string out;
  char line[256];
  for (size_t i = 0; i < g_trace.size(); ++i) {
    const TraceRec& r = g_trace[i];
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, g_trace[0].e0, r.e0) != cudaSuccess) t0 = -1.f;
    if (cudaEventElapsedTime(&t1, g_trace[0].e0, r.e1) != cudaSuccess) t1 = -1.f;
    snprintf(line, sizeof(line), "%s %d %d %d %d %llu %.4f %.4f\n", r.name, r.a, r.b, r.c, r.d,
             static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(r.stream)), t0, t1);
    out += line;
  }
  cudaGetLastError();
  if (buf != nullptr && cap > 0) {
    const size_t n = out.size() < static_cast<size_t>(cap - 1) ? out.size() : static_cast<size_t>(cap - 1);
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(out.size()) + 1;
}
int styler_version(void) { return 200; }
const char* styler_last_error(void) { return sb::g_err; }
int64_t styler_launch_count(void) { return sb::g_launch_count.load(); }
}
