// psk_jit.cpp -- see psk_jit.h.  Host-only C++; binds libnvrtc / libcuda with dlopen.
#include "psk_jit.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "psk_ops.h"

// kJitHeaderNames / kJitHeaderSources: the kernel headers of this directory, embedded by
// pyskip_b200/build.py so that NVRTC sees exactly the sources nvcc compiled.
#include "_jit_embed.inc"

#define PSK_STR2(x) #x
#define PSK_STR(x) PSK_STR2(x)

namespace psk {
namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
typedef int (*nvrtcCreateProgram_t)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                                    const char* const*);
typedef int (*nvrtcCompileProgram_t)(nvrtcProgram, int, const char* const*);
typedef int (*nvrtcGetSize_t)(nvrtcProgram, size_t*);
typedef int (*nvrtcGetData_t)(nvrtcProgram, char*);
typedef int (*nvrtcDestroyProgram_t)(nvrtcProgram*);
typedef const char* (*nvrtcGetErrorString_t)(int);

typedef int (*cuModuleLoadData_t)(void**, const void*);
typedef int (*cuModuleGetFunction_t)(void**, void*, const char*);
typedef int (*cuFuncSetAttribute_t)(void*, int, int);
typedef int (*cuLaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                void*, void**, void**);
typedef int (*cuGetErrorString_t)(int, const char**);

struct Api {
  bool tried = false, ok = false;
  std::string why;
  nvrtcCreateProgram_t nvrtcCreateProgram = nullptr;
  nvrtcCompileProgram_t nvrtcCompileProgram = nullptr;
  nvrtcGetSize_t nvrtcGetCUBINSize = nullptr, nvrtcGetProgramLogSize = nullptr;
  nvrtcGetData_t nvrtcGetCUBIN = nullptr, nvrtcGetProgramLog = nullptr;
  nvrtcDestroyProgram_t nvrtcDestroyProgram = nullptr;
  nvrtcGetErrorString_t nvrtcGetErrorString = nullptr;
  cuModuleLoadData_t cuModuleLoadData = nullptr;
  cuModuleGetFunction_t cuModuleGetFunction = nullptr;
  cuFuncSetAttribute_t cuFuncSetAttribute = nullptr;
  cuLaunchKernel_t cuLaunchKernel = nullptr;
  cuGetErrorString_t cuGetErrorString = nullptr;
};

Api g_api;
std::mutex g_mu;
std::unordered_map<std::string, JitKernel*> g_cache;  // value nullptr = known failure
int g_compiled = 0;

void* open_first(const char* const* names) {
  for (; *names; ++names)
    if (void* h = dlopen(*names, RTLD_NOW | RTLD_LOCAL)) return h;
  return nullptr;
}

template <typename F>
bool sym(void* h, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(h, name));
  return *out != nullptr;
}

bool load_nvrtc() {
  static bool tried = false, ok = false;
  if (tried) return ok;
  tried = true;
  // The toolkit's own NVRTC first (the one matching the nvcc that built this library): a
  // bare "libnvrtc.so.12" resolves to whatever the process loaded before -- e.g. the older
  // NVRTC bundled with PyTorch, whose code for this kernel measured 22 % slower.
  static const char* const rtc_names[] = {
#ifdef PSK_NVRTC_PATH
      PSK_NVRTC_PATH,
#endif
      "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so", nullptr};
  void* rtc = open_first(rtc_names);
  if (!rtc) {
    g_api.why = "libnvrtc not found";
    return false;
  }
  ok = sym(rtc, "nvrtcCreateProgram", &g_api.nvrtcCreateProgram) &&
       sym(rtc, "nvrtcCompileProgram", &g_api.nvrtcCompileProgram) &&
       sym(rtc, "nvrtcGetCUBINSize", &g_api.nvrtcGetCUBINSize) &&
       sym(rtc, "nvrtcGetCUBIN", &g_api.nvrtcGetCUBIN) &&
       sym(rtc, "nvrtcGetProgramLogSize", &g_api.nvrtcGetProgramLogSize) &&
       sym(rtc, "nvrtcGetProgramLog", &g_api.nvrtcGetProgramLog) &&
       sym(rtc, "nvrtcDestroyProgram", &g_api.nvrtcDestroyProgram) &&
       sym(rtc, "nvrtcGetErrorString", &g_api.nvrtcGetErrorString);
  if (!ok) g_api.why = "missing symbol in libnvrtc";
  return ok;
}

bool load_api() {
  if (g_api.tried) return g_api.ok;
  g_api.tried = true;
  const char* dis = getenv("PSKB200_JIT");
  if (dis && dis[0] == '0') {
    g_api.why = "disabled by PSKB200_JIT=0";
    return false;
  }
  if (!load_nvrtc()) return false;
  static const char* const cu_names[] = {"libcuda.so.1", "libcuda.so", nullptr};
  void* cu = open_first(cu_names);
  if (!cu) {
    g_api.why = "libcuda not found";
    return false;
  }
  void* rtc = cu;  // nvrtc symbols are already bound
  bool ok = true &&
            sym(cu, "cuModuleLoadData", &g_api.cuModuleLoadData) &&
            sym(cu, "cuModuleGetFunction", &g_api.cuModuleGetFunction) &&
            sym(cu, "cuFuncSetAttribute", &g_api.cuFuncSetAttribute) &&
            sym(cu, "cuLaunchKernel", &g_api.cuLaunchKernel) &&
            sym(cu, "cuGetErrorString", &g_api.cuGetErrorString);
  (void)rtc;
  if (!ok) g_api.why = "missing symbol in libcuda";
  g_api.ok = ok;
  return ok;
}

// postfix program -> straight-line SSA over apply_unary / apply_binary with literal op-codes
// (the switch inside them folds away at compile time)
std::string gen_source(int K, bool typed_float, bool views, int nprog, const psk_node* prog) {
  std::string body;
  std::vector<int> st;
  int next = 0;
  char buf[1024];
  for (int i = 0; i < nprog; ++i) {
    const int op = prog[i].op;
    if (op == PSK_OP_SRC) {
      snprintf(buf, sizeof buf, "    const uint32_t t%d = v[%u];\n", next, prog[i].arg);
      st.push_back(next++);
    } else if (op == PSK_OP_CONST) {
      snprintf(buf, sizeof buf, "    const uint32_t t%d = c%d;\n", next, i);
      st.push_back(next++);
    } else if (op == PSK_OP_SELECT) {
      int c = st.back(); st.pop_back();
      int b = st.back(); st.pop_back();
      int a = st.back(); st.pop_back();
      snprintf(buf, sizeof buf, "    const uint32_t t%d = t%d ? t%d : t%d;\n", next, a, b, c);
      st.push_back(next++);
    } else if (op_arity(op) == 1) {
      int a = st.back(); st.pop_back();
      snprintf(buf, sizeof buf, "    const uint32_t t%d = apply_unary(%d, t%d);\n", next, op, a);
      st.push_back(next++);
    } else {
      int b = st.back(); st.pop_back();
      int a = st.back(); st.pop_back();
      snprintf(buf, sizeof buf, "    const uint32_t t%d = apply_binary(%d, t%d, t%d);\n", next, op, a, b);
      st.push_back(next++);
    }
    body += buf;
  }
  snprintf(buf, sizeof buf, "    return t%d;\n", st.back());
  body += buf;

  // immediates are run-time data (one module serves every set of constants): loaded from
  // the plan into registers once per thread
  std::string members, inits;
  for (int i = 0; i < nprog; ++i) {
    if (prog[i].op != PSK_OP_CONST) continue;
    snprintf(buf, sizeof buf, "  uint32_t c%d;\n", i);
    members += buf;
    snprintf(buf, sizeof buf, "  ev.c%d = p->prog[%d].arg;\n", i, i);
    inits += buf;
  }
  std::string src = "#include \"psk_eval.cuh\"\nnamespace psk {\nstruct JitEval {\n" + members +
                    "  template <typename ValArray>\n"
                    "  __device__ __forceinline__ uint32_t operator()(const ValArray& v) const {\n";
  src += body;
  src += "  }\n};\n}  // namespace psk\n";
  snprintf(buf, sizeof buf,
           "  psk::merge_eval_tile_loop<%d, %s, %d>(p, ev);\n}\n", K, typed_float ? "true" : "false", views ? 1 : 0);
  src += "extern \"C\" __global__ void __launch_bounds__(psk::kTileThreads, psk::kCtasPerSm) psk_jit_merge_eval(psk::DPlan* p) {\n"
         "  psk::JitEval ev;\n";
  src += inits;
  src += buf;
  return src;
}

std::string cache_key(int K, bool typed_float, bool views, int T, int CAP, int regions, int nprog, const psk_node* prog) {
  std::string key;
  key.reserve(nprog * 3 + 24);
  key += std::to_string(T) + "/" + std::to_string(CAP) + "/" + std::to_string(regions) + "/";
  key += (char)K;
  key += typed_float ? 'T' : 'B';
  key += views ? 'V' : 'F';
  for (int i = 0; i < nprog; ++i) {
    key += (char)prog[i].op;
    key += (char)(prog[i].op == PSK_OP_SRC ? prog[i].arg : 0);
  }
  return key;
}

bool compile_cubin(int K, bool typed_float, bool views, int T, int CAP, int regions, int nprog, const psk_node* prog,
                   std::vector<char>* cubin, std::string* why) {
  const std::string src = gen_source(K, typed_float, views, nprog, prog);
  nvrtcProgram np = nullptr;
  int r = g_api.nvrtcCreateProgram(&np, src.c_str(), "psk_jit_merge_eval.cu", kJitNumHeaders, kJitHeaderSources,
                                   kJitHeaderNames);
  if (r != 0) {
    *why = std::string("nvrtcCreateProgram: ") + g_api.nvrtcGetErrorString(r);
    return false;
  }
  const std::string opt_t = "-DPSK_TILE_THREADS=" + std::to_string(T);
  const std::string opt_c = "-DPSK_TILE_CAP=" + std::to_string(CAP);
  const std::string opt_r = "-DPSK_PARK_DEPTH=" + std::to_string(regions);
  // tuning aid: extra definitions for the kernel source, e.g. PSKB200_JIT_DEFS="-DPSK_COARSE_GROUP=2"
  const char* extra_env = getenv("PSKB200_JIT_DEFS");
  const std::string opt_x = (extra_env && strncmp(extra_env, "-D", 2) == 0 && !strchr(extra_env, ' ')) ? extra_env : "-DPSK_JIT";
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", opt_x.c_str(),
#ifdef PSK_PHASE_TIMING
                        "-DPSK_PHASE_TIMING",
#endif
                        opt_t.c_str(), opt_c.c_str(), opt_r.c_str(),
  };
  r = g_api.nvrtcCompileProgram(np, (int)(sizeof(opts) / sizeof(opts[0])), opts);
  if (r != 0) {
    size_t n = 0;
    g_api.nvrtcGetProgramLogSize(np, &n);
    std::string log(n, '\0');
    if (n) g_api.nvrtcGetProgramLog(np, &log[0]);
    *why = std::string("nvrtcCompileProgram: ") + g_api.nvrtcGetErrorString(r) + "\n" + log;
    g_api.nvrtcDestroyProgram(&np);
    return false;
  }
  size_t n = 0;
  g_api.nvrtcGetCUBINSize(np, &n);
  cubin->resize(n);
  g_api.nvrtcGetCUBIN(np, cubin->data());
  g_api.nvrtcDestroyProgram(&np);
  if (const char* dump = getenv("PSKB200_JIT_DUMP")) {  // for cuobjdump -sass / -res-usage
    static int serial = 0;
    std::string base = std::string(dump) + "/psk_jit_" + std::to_string(serial++);
    if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) {
      fwrite(cubin->data(), 1, cubin->size(), f);
      fclose(f);
    }
    if (FILE* f = fopen((base + ".cu").c_str(), "w")) {
      fputs(src.c_str(), f);
      fclose(f);
    }
  }
  return true;
}

JitKernel* compile(int K, bool typed_float, bool views, int T, int CAP, int regions, int nprog, const psk_node* prog,
                   std::string* why) {
  std::vector<char> cubin;
  if (!compile_cubin(K, typed_float, views, T, CAP, regions, nprog, prog, &cubin, why)) return nullptr;
  void* mod = nullptr;
  int r = g_api.cuModuleLoadData(&mod, cubin.data());
  if (r != 0) {
    const char* s = nullptr;
    g_api.cuGetErrorString(r, &s);
    *why = std::string("cuModuleLoadData: ") + (s ? s : "?");
    return nullptr;
  }
  void* fn = nullptr;
  r = g_api.cuModuleGetFunction(&fn, mod, "psk_jit_merge_eval");
  if (r != 0) {
    *why = "cuModuleGetFunction failed";
    return nullptr;
  }
  return new JitKernel{fn};
}

}  // namespace

const JitKernel* jit_get_merge_eval(int K, bool typed_float, bool views, int T, int CAP, int regions, int nprog,
                                    const psk_node* prog, const char** why) {
  std::lock_guard<std::mutex> l(g_mu);
  static std::string s_why;
  if (!load_api()) {
    if (why) *why = g_api.why.c_str();
    return nullptr;
  }
  const std::string key = cache_key(K, typed_float, views, T, CAP, regions, nprog, prog);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    if (!it->second && why) *why = "previous compilation of this program failed";
    return it->second;
  }
  JitKernel* k = compile(K, typed_float, views, T, CAP, regions, nprog, prog, &s_why);
  if (k) {
    ++g_compiled;
  } else {
    if (getenv("PSKB200_JIT_VERBOSE")) fprintf(stderr, "pskb200 jit: %s\n", s_why.c_str());
    if (why) *why = s_why.c_str();
  }
  g_cache.emplace(key, k);
  return k;
}

int jit_launch(const JitKernel* k, unsigned grid, unsigned block, size_t smem, void* stream, void* plan_ptr) {
  // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8
  static std::unordered_map<void*, size_t> s_attr;
  {
    std::lock_guard<std::mutex> l(g_mu);
    auto it = s_attr.find(k->function);
    if (it == s_attr.end() || it->second < smem) {
      int r = g_api.cuFuncSetAttribute(k->function, 8, (int)smem);
      if (r != 0) return r;
      s_attr[k->function] = smem;
    }
  }
  void* args[] = {&plan_ptr};
  return g_api.cuLaunchKernel(k->function, grid, 1, 1, block, 1, 1, (unsigned)smem, stream, args, nullptr);
}

int jit_compile_check(int K, bool typed_float, int T, int CAP, int regions, int nprog, const psk_node* prog, char* log, int loglen) {
  std::lock_guard<std::mutex> l(g_mu);
  std::string why;
  std::vector<char> cubin;
  bool ok = load_nvrtc() && compile_cubin(K, typed_float, false, T, CAP, regions, nprog, prog, &cubin, &why) &&
            compile_cubin(K, typed_float, true, T, CAP, regions, nprog, prog, &cubin, &why);  // both staging variants
  if (!ok && why.empty()) why = g_api.why;
  if (log && loglen > 0) {
    std::string msg = ok ? ("ok cubin bytes=" + std::to_string(cubin.size())) : why;
    snprintf(log, (size_t)loglen, "%s", msg.c_str());
  }
  return ok ? 0 : 1;
}

int jit_compiled_count() {
  std::lock_guard<std::mutex> l(g_mu);
  return g_compiled;
}

}  // namespace psk
