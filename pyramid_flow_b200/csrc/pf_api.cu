// pf_api.cu — error plumbing, driver entry points and device queries for libpf_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/pf_b200.h"
#include "pf_common.cuh"

namespace pf {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  PF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver / no GPU): libpf_b200 has no CPU fallback");
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %u dims [%llu,%llu,%llu] stride0 %llu box [%u,%u] base %p",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
              box[0], rank > 1 ? box[1] : 0, base);
    return -3;
  }
  return 0;
}

int ensure_dyn_smem(const void* kernel, int bytes, const char* what) {
  struct Entry { const void* k; unsigned long long devmask; };
  static std::mutex mu;
  static Entry table[128];
  static int n_entries = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("%s: no CUDA device", what);
    return -2;
  }
  const unsigned long long bit = 1ull << (dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  Entry* e = nullptr;
  for (int i = 0; i < n_entries; ++i)
    if (table[i].k == kernel) { e = &table[i]; break; }
  if (e != nullptr && (e->devmask & bit)) return 0;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaFuncSetAttribute(%s, dynamic smem %d): %s", what, bytes, cudaGetErrorString(err));
    return -2;
  }
  if (e == nullptr && n_entries < 128) { e = &table[n_entries++]; e->k = kernel; e->devmask = 0; }
  if (e != nullptr) e->devmask |= bit;
  return 0;
}

// Library options (pf_set_option): new data paths stay opt-in until a hardware run has validated them; the defaults below are
// the validated choices.
static std::atomic<int> g_options[PF_OPT_COUNT] = {};
static std::once_flag g_options_once;
static void options_init() {
  g_options[PF_OPT_GEMM_STAGED_RESID].store(PF_OPT_DEFAULT_GEMM_STAGED_RESID);
  g_options[PF_OPT_GEMM_WAVE_TILING].store(PF_OPT_DEFAULT_GEMM_WAVE_TILING);
  g_options[PF_OPT_ATTN_PAIR_KERNEL].store(PF_OPT_DEFAULT_ATTN_PAIR_KERNEL);
  g_options[PF_OPT_ATTN_TILE_PHASE].store(PF_OPT_DEFAULT_ATTN_TILE_PHASE);
  g_options[PF_OPT_ATTN_TRIPLE_KERNEL].store(PF_OPT_DEFAULT_ATTN_TRIPLE_KERNEL);
}
int get_option(int key) {
  std::call_once(g_options_once, options_init);
  return (key >= 0 && key < PF_OPT_COUNT) ? g_options[key].load(std::memory_order_relaxed) : 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    n = prop.multiProcessorCount;
  }
  return n;
}

int warmup_gemm();
int warmup_conv();
int warmup_attn();

}  // namespace pf

extern "C" {

// ---- pf_ctx: a recorded launch sequence (one DiT step, one VAE chunk ...) owned by the library ---------------------------
struct pf_ctx {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t rec_stream = nullptr;
  bool recording = false;
  size_t nodes = 0;
};

int pf_ctx_create(pf_ctx** out) {
  if (out == nullptr) {
    pf::set_error("pf_ctx_create: null");
    return -1;
  }
  *out = new pf_ctx();
  return 0;
}

static void ctx_drop(pf_ctx* c) {
  if (c->exec) cudaGraphExecDestroy(c->exec);
  if (c->graph) cudaGraphDestroy(c->graph);
  c->exec = nullptr;
  c->graph = nullptr;
  c->nodes = 0;
}

int pf_ctx_destroy(pf_ctx* c) {
  if (c == nullptr) return 0;
  if (c->recording) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(c->rec_stream, &g);
    if (g) cudaGraphDestroy(g);
  }
  ctx_drop(c);
  (void)cudaGetLastError();
  delete c;
  return 0;
}

int pf_ctx_record_begin(pf_ctx* c, void* stream) {
  using namespace pf;
  PF_REQUIRE(c != nullptr && !c->recording, "pf_ctx_record_begin: null context or already recording");
  int rc = pf_warmup();          // nothing may initialise host-side while the stream is capturing
  if (rc) return rc;
  ctx_drop(c);
  c->rec_stream = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaStreamBeginCapture(c->rec_stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("pf_ctx_record_begin: cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    return -2;
  }
  c->recording = true;
  return 0;
}

int pf_ctx_record_end(pf_ctx* c) {
  using namespace pf;
  PF_REQUIRE(c != nullptr && c->recording, "pf_ctx_record_end: not recording");
  c->recording = false;
  cudaError_t e = cudaStreamEndCapture(c->rec_stream, &c->graph);
  if (e == cudaSuccess) e = cudaGraphGetNodes(c->graph, nullptr, &c->nodes);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&c->exec, c->graph, 0);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ctx_drop(c);
    set_error("pf_ctx_record_end: %s", cudaGetErrorString(e));
    return -2;
  }
  return static_cast<int>(c->nodes);
}

int pf_ctx_replay(pf_ctx* c, void* stream) {
  using namespace pf;
  PF_REQUIRE(c != nullptr && c->exec != nullptr, "pf_ctx_replay: nothing recorded");
  cudaError_t e = cudaGraphLaunch(c->exec, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("pf_ctx_replay: cudaGraphLaunch: %s", cudaGetErrorString(e));
    return -2;
  }
  g_launches.fetch_add(static_cast<int64_t>(c->nodes), std::memory_order_relaxed);
  return 0;
}

int pf_dit_step_flux(pf_ctx* c, void* stream) { return pf_ctx_replay(c, stream); }
int pf_dit_step_mmdit(pf_ctx* c, void* stream) { return pf_ctx_replay(c, stream); }
int pf_vae_decode_chunk(pf_ctx* c, void* stream) { return pf_ctx_replay(c, stream); }

int pf_set_option(int key, int value) {
  std::call_once(pf::g_options_once, pf::options_init);
  if (key < 0 || key >= PF_OPT_COUNT) {
    pf::set_error("pf_set_option: unknown key %d", key);
    return -1;
  }
  pf::g_options[key].store(value);
  return 0;
}
int pf_get_option(int key) { return pf::get_option(key); }

int pf_warmup(void) {
  int rc = pf::warmup_gemm();
  if (!rc) rc = pf::warmup_conv();
  if (!rc) rc = pf::warmup_attn();
  return rc;
}

const char* pf_last_error(void) { return pf::g_err; }
int pf_version(void) { return 100; }
int64_t pf_launch_count(void) { return pf::g_launches.load(); }

int pf_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    pf::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return -1;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    pf::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return -1;
  }
  if (prop.major != 10) {
    pf::set_error("libpf_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    return -1;
  }
  if (pf::get_encode_fn() == nullptr) {
    pf::set_error("driver does not expose cuTensorMapEncodeTiled");
    return -1;
  }
  return 0;
}

}  // extern "C"
