// modules.inl — plug-in models: dynamics functors supplied as CUDA source at run time
// (the reference's extension mechanism is subclassing DiscreteDynamics,
// altro/problem/dynamics.hpp:148-187 there — a virtual call cannot run on the device, a functor
// compiled into the kernels can).
//
// The kernels of this library are templates on a model functor M (device.cuh documents the concept:
// static constexpr n, m, kDiscrete, kStage3RepeatsStage2; static __device__ eval / jac).  A plug-in is
// the source text of such a struct.  When a solver is created for it, the kernel templates are
// instantiated on it with NVRTC (sm_100a cubin), loaded with the driver API and launched through the
// same host code as the built-in instantiations (KernelTable in altro_b200.cu).  Compiled modules are
// cached on disk next to the library (modules/<hash>.cubin; ALTRO_B200_MODULE_CACHE overrides), keyed
// by the plug-in source, the kernel headers and the compiler version, so a model is compiled once —
// __graft_entry__.build() does it for the models shipped under plugins/.
//
// libnvrtc and libcuda are opened lazily with dlopen: the library itself links against neither, so it
// loads (and its CPU-side entry points work) on a machine without a driver.
// Included by altro_b200.cu, inside its anonymous namespace.

// (system headers are included at the top of altro_b200.cu)

struct PluginModel {
  std::string name;    // struct name inside namespace altro_b200
  std::string source;  // text of the struct
  int n = 0, m = 0, nparams = 0;
};

std::mutex g_plugin_mutex;
std::vector<PluginModel> g_plugins;  // model id = kFirstPluginId + index
constexpr int kFirstPluginId = 100;

// ---- lazily bound driver / NVRTC entry points -------------------------------------------------------
struct DriverApi {
  void* lib = nullptr;
  int (*cuModuleLoadData)(void**, const void*) = nullptr;
  int (*cuModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  int (*cuFuncSetAttribute)(void*, int, int) = nullptr;
  int (*cuGetErrorString)(int, const char**) = nullptr;
  bool ok = false;
};
DriverApi& driver() {
  static DriverApi d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!d.lib) return;
    auto sym = [&](const char* nm) { return dlsym(d.lib, nm); };
    d.cuModuleLoadData = reinterpret_cast<decltype(d.cuModuleLoadData)>(sym("cuModuleLoadData"));
    d.cuModuleGetFunction = reinterpret_cast<decltype(d.cuModuleGetFunction)>(sym("cuModuleGetFunction"));
    d.cuLaunchKernel = reinterpret_cast<decltype(d.cuLaunchKernel)>(sym("cuLaunchKernel"));
    d.cuFuncSetAttribute = reinterpret_cast<decltype(d.cuFuncSetAttribute)>(sym("cuFuncSetAttribute"));
    d.cuGetErrorString = reinterpret_cast<decltype(d.cuGetErrorString)>(sym("cuGetErrorString"));
    d.ok = d.cuModuleLoadData && d.cuModuleGetFunction && d.cuLaunchKernel && d.cuFuncSetAttribute;
  });
  return d;
}

struct NvrtcApi {
  void* lib = nullptr;
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
  int (*AddNameExpression)(void*, const char*) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  bool ok = false;
};
NvrtcApi& nvrtc() {
  static NvrtcApi r;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* nm : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      r.lib = dlopen(nm, RTLD_NOW);
      if (r.lib) break;
    }
    if (!r.lib) return;
    auto sym = [&](const char* nm) { return dlsym(r.lib, nm); };
    r.CreateProgram = reinterpret_cast<decltype(r.CreateProgram)>(sym("nvrtcCreateProgram"));
    r.DestroyProgram = reinterpret_cast<decltype(r.DestroyProgram)>(sym("nvrtcDestroyProgram"));
    r.AddNameExpression = reinterpret_cast<decltype(r.AddNameExpression)>(sym("nvrtcAddNameExpression"));
    r.CompileProgram = reinterpret_cast<decltype(r.CompileProgram)>(sym("nvrtcCompileProgram"));
    r.GetProgramLogSize = reinterpret_cast<decltype(r.GetProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
    r.GetProgramLog = reinterpret_cast<decltype(r.GetProgramLog)>(sym("nvrtcGetProgramLog"));
    r.GetCUBINSize = reinterpret_cast<decltype(r.GetCUBINSize)>(sym("nvrtcGetCUBINSize"));
    r.GetCUBIN = reinterpret_cast<decltype(r.GetCUBIN)>(sym("nvrtcGetCUBIN"));
    r.GetLoweredName = reinterpret_cast<decltype(r.GetLoweredName)>(sym("nvrtcGetLoweredName"));
    r.Version = reinterpret_cast<decltype(r.Version)>(sym("nvrtcVersion"));
    r.ok = r.CreateProgram && r.DestroyProgram && r.AddNameExpression && r.CompileProgram && r.GetProgramLogSize &&
           r.GetProgramLog && r.GetCUBINSize && r.GetCUBIN && r.GetLoweredName;
  });
  return r;
}

cudaError_t driver_launch(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
  DriverApi& d = driver();
  if (!d.ok) return cudaErrorInitializationError;
  void* f = const_cast<void*>(fn);
  if (smem > 48 * 1024) {
    constexpr int kMaxDynamicSharedSizeBytes = 8;  // CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES
    if (d.cuFuncSetAttribute(f, kMaxDynamicSharedSizeBytes, static_cast<int>(smem)) != 0) return cudaErrorInvalidValue;
  }
  const int rc = d.cuLaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, static_cast<unsigned>(smem), st, args,
                                  nullptr);
  return rc == 0 ? cudaSuccess : cudaErrorLaunchFailure;
}

// ---- where the kernel headers and the module cache live ---------------------------------------------
std::string library_dir() {
  Dl_info info;
  if (dladdr(reinterpret_cast<const void*>(&driver_launch), &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t slash = p.rfind('/');
    return slash == std::string::npos ? "." : p.substr(0, slash);
  }
  return ".";
}
std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}
const char* const kModuleHeaders[] = {"common.cuh", "device.cuh", "kernels.cuh", "phased.cuh", "outer.cuh", "backward_coop.cuh"};

// kernel name expressions of one model, in KernelId order ("" = not part of a module)
std::vector<std::string> module_kernel_names(const PluginModel& pm) {
  const std::string M = "altro_b200::" + pm.name;
  const bool coop = use_coop(pm.n, pm.m, kPhasedTile);
  std::vector<std::string> k(K_NUM);
  auto t = [&](const char* kernel, const char* rest) { return std::string("altro_b200::") + kernel + "<" + M + ", 8" + rest + ">"; };
  k[K_SOLVE] = t("k_solve", "");
  k[K_PHASE] = t("k_phase", "");
  k[K_EXP] = t("k_update_expansions", ", false");
  k[K_EXP_PHASED] = t("k_update_expansions", ", true");
  k[K_CON_VALUES] = t("k_constraint_values", "");
  k[K_BP_CTG] = t("k_backward_mat", ", 4, true, false");
  k[K_BP_STREAM] = t("k_backward_mat", ", 4, false, false");
  k[K_BP_PHASED] = t("k_backward_mat", ", 4, false, true");
  if (coop) {
    k[K_COOP_CTG] = "altro_b200::k_backward_coop<" + M + ", 3, true, false>";
    k[K_COOP_STREAM] = "altro_b200::k_backward_coop<" + M + ", 3, false, false>";
    k[K_COOP_PHASED] = "altro_b200::k_backward_coop<" + M + ", 3, false, true>";
  }
  k[K_ROLL_WIDE] = t("k_roll_wide", "");
  k[K_COST_WIDE] = t("k_cost_wide", "");
  k[K_ACC_WIDE] = t("k_acc_wide", "");
  k[K_ROLL_DEEP] = t("k_roll_deep", "");
  k[K_COST_DEEP] = t("k_cost_deep", "");
  k[K_ACC_DEEP] = t("k_acc_deep", "");
  k[K_LS_WIDE] = t("k_ls_wide", "");
  k[K_LS_DEEP1] = t("k_ls_deep", ", 1");
  k[K_LS_DEEP2] = t("k_ls_deep", ", 2");
  k[K_OUTER_REGEN] = t("k_outer_rollout", ", true");
  k[K_OUTER_ROLLOUT] = t("k_outer_rollout", ", false");
  k[K_OUTER_DUALS] = t("k_outer_duals", "");
  k[K_OUTER_COST] = t("k_outer_cost", "");
  return k;
}

struct CompiledModule {
  std::string cubin;
  std::vector<std::string> lowered;  // mangled kernel names, KernelId order
};

// cache file: "ALTROMOD1\n" + K_NUM lines of lowered names + "\n" + cubin bytes
bool load_cached(const std::string& path, CompiledModule* out) {
  const std::string raw = read_file(path);
  if (raw.compare(0, 10, "ALTROMOD1\n") != 0) return false;
  size_t pos = 10;
  out->lowered.assign(K_NUM, "");
  for (int i = 0; i < K_NUM; ++i) {
    const size_t nl = raw.find('\n', pos);
    if (nl == std::string::npos) return false;
    out->lowered[i] = raw.substr(pos, nl - pos);
    pos = nl + 1;
  }
  out->cubin = raw.substr(pos);
  return !out->cubin.empty();
}

int compile_module(const PluginModel& pm, CompiledModule* out, std::string* cache_path_out) {
  const std::string dir = library_dir();
  const std::string csrc = dir + "/csrc";
  std::string key = pm.name + "\n" + pm.source + "\n";
  for (const char* h : kModuleHeaders) {
    const std::string text = read_file(csrc + "/" + h);
    if (text.empty()) return fail(ALTRO_B200_ERR_STATE, "plug-in models need the kernel headers next to the library: " + csrc + "/" + h + " not found");
    key += text;
  }
  NvrtcApi& r = nvrtc();
  int major = 0, minor = 0;
  if (r.ok && r.Version) r.Version(&major, &minor);
  key += "nvrtc " + std::to_string(major) + "." + std::to_string(minor) + " sm_100a v1";
  char hex[32];
  std::snprintf(hex, sizeof(hex), "%016llx", static_cast<unsigned long long>(fnv1a(key)));
  const char* env = std::getenv("ALTRO_B200_MODULE_CACHE");
  const std::string cache_dir = env ? env : dir + "/modules";
  const std::string path = cache_dir + "/" + pm.name + "_" + hex + ".cubin";
  if (cache_path_out) *cache_path_out = path;
  if (load_cached(path, out)) return 0;
  if (!r.ok) return fail(ALTRO_B200_ERR_UNSUPPORTED, "plug-in model '" + pm.name + "' is not in the module cache and libnvrtc could not be loaded");

  const std::string src = "#include \"phased.cuh\"\n#include \"outer.cuh\"\n#include \"backward_coop.cuh\"\nnamespace altro_b200 {\n" +
                          pm.source + "\n}\n";
  void* prog = nullptr;
  if (r.CreateProgram(&prog, src.c_str(), (pm.name + ".cu").c_str(), 0, nullptr, nullptr) != 0)
    return fail(ALTRO_B200_ERR_CUDA, "nvrtcCreateProgram failed");
  const std::vector<std::string> names = module_kernel_names(pm);
  for (const std::string& nm : names)
    if (!nm.empty() && r.AddNameExpression(prog, nm.c_str()) != 0) {
      r.DestroyProgram(&prog);
      return fail(ALTRO_B200_ERR_CUDA, "nvrtcAddNameExpression failed for " + nm);
    }
  const std::string inc = "-I" + csrc;
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "-split-compile=0", inc.c_str()};
  const int rc = r.CompileProgram(prog, 6, opts);
  if (rc != 0) {
    size_t n = 0;
    r.GetProgramLogSize(prog, &n);
    std::string log(n, ' ');
    if (n) r.GetProgramLog(prog, &log[0]);
    r.DestroyProgram(&prog);
    return fail(ALTRO_B200_ERR_ARG, "plug-in model '" + pm.name + "' does not compile:\n" + log.substr(0, 4000));
  }
  out->lowered.assign(K_NUM, "");
  for (int i = 0; i < K_NUM; ++i) {
    if (names[i].empty()) continue;
    const char* low = nullptr;
    if (r.GetLoweredName(prog, names[i].c_str(), &low) != 0 || !low) {
      r.DestroyProgram(&prog);
      return fail(ALTRO_B200_ERR_CUDA, "nvrtcGetLoweredName failed for " + names[i]);
    }
    out->lowered[i] = low;
  }
  size_t sz = 0;
  r.GetCUBINSize(prog, &sz);
  out->cubin.assign(sz, '\0');
  r.GetCUBIN(prog, &out->cubin[0]);
  r.DestroyProgram(&prog);
  // best-effort cache write (atomic rename)
  mkdir(cache_dir.c_str(), 0755);
  const std::string tmp = path + ".tmp" + std::to_string(static_cast<long>(getpid()));
  {
    std::ofstream f(tmp, std::ios::binary);
    f << "ALTROMOD1\n";
    for (const std::string& l : out->lowered) f << l << "\n";
    f.write(out->cubin.data(), static_cast<std::streamsize>(out->cubin.size()));
  }
  std::rename(tmp.c_str(), path.c_str());
  return 0;
}

// loaded modules per (model id, device)
std::map<std::pair<int, int>, KernelTable> g_loaded_modules;

int plugin_ops(int model, int device, Ops* out) {
  PluginModel pm;
  {
    std::lock_guard<std::mutex> lock(g_plugin_mutex);
    const int idx = model - kFirstPluginId;
    if (idx < 0 || idx >= static_cast<int>(g_plugins.size())) return fail(ALTRO_B200_ERR_ARG, "unknown plug-in model id " + std::to_string(model));
    pm = g_plugins[idx];
    auto it = g_loaded_modules.find({model, device});
    if (it != g_loaded_modules.end()) {
      if (out) *out = make_ops_from(it->second);
      return 0;
    }
  }
  CompiledModule cm;
  int rc = compile_module(pm, &cm, nullptr);
  if (rc) return rc;
  DriverApi& d = driver();
  if (!d.ok) return fail(ALTRO_B200_ERR_CUDA, "libcuda.so.1 could not be loaded (no usable CUDA device; there is no CPU fallback)");
  cudaFree(nullptr);  // make sure the runtime's primary context is current on this thread
  void* mod = nullptr;
  int drc = d.cuModuleLoadData(&mod, cm.cubin.data());
  if (drc != 0) {
    const char* msg = nullptr;
    if (d.cuGetErrorString) d.cuGetErrorString(drc, &msg);
    return fail(ALTRO_B200_ERR_CUDA, std::string("cuModuleLoadData failed for plug-in model '") + pm.name + "': " + (msg ? msg : "?"));
  }
  KernelTable kt;
  kt.driver = true;
  kt.n = pm.n; kt.m = pm.m; kt.W = kPhasedTile;
  kt.coop = use_coop(pm.n, pm.m, kPhasedTile);
  for (int i = 0; i < K_NUM; ++i) {
    if (cm.lowered[i].empty()) continue;
    void* fn = nullptr;
    if (d.cuModuleGetFunction(&fn, mod, cm.lowered[i].c_str()) != 0)
      return fail(ALTRO_B200_ERR_CUDA, "kernel " + cm.lowered[i] + " is missing from the module of '" + pm.name + "'");
    kt.f[i] = fn;
  }
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  g_loaded_modules[{model, device}] = kt;
  if (out) *out = make_ops_from(kt);
  return 0;
}
