// User-supplied ODE right-hand sides, compiled at run time with NVRTC for sm_100a.
//
// The reference accepts any traced Python callable as `ode_fun` (src/rodeo/solve.py:219).  Here a user model is a
// CUDA source string defining `struct UserModel` with the functor interface of rodeo_models.cuh (the Python side,
// rodeo_b200.models.CudaOde, wraps a one-line right-hand side into that struct; the block-diagonal Jacobian then
// comes from dual numbers, i.e. what jax.jacfwd would yield).  The SAME kernel templates as the ahead-of-time
// instantiations are compiled: the three device headers are embedded in this library as strings and handed to NVRTC
// as in-memory includes.  One cubin per (model, op, interrogation, structure) is built on first use and cached.
//
// libnvrtc and libcuda are opened lazily with dlopen so that the library still loads on a machine that has neither.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "rodeo_host.h"

namespace rodeo {
namespace host {

extern const char* const kEmbeddedCore;      // rodeo_core.cuh
extern const char* const kEmbeddedModels;    // rodeo_models.cuh
extern const char* const kEmbeddedKernels;   // rodeo_kernels.cuh

namespace {

struct Api {
  void *nvrtc = nullptr, *cuda = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  CUresult (*ModuleLoadData)(CUmodule*, const void*);
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**);
  bool ok = false;
};

template <typename F>
bool sym(void* lib, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  return out != nullptr;
}

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* n : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"})
      if ((a.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    for (const char* n : {"libcuda.so.1", "libcuda.so"})
      if ((a.cuda = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!a.nvrtc || !a.cuda) return;
    a.ok = sym(a.nvrtc, "nvrtcCreateProgram", a.CreateProgram) && sym(a.nvrtc, "nvrtcCompileProgram", a.CompileProgram) &&
           sym(a.nvrtc, "nvrtcAddNameExpression", a.AddNameExpression) &&
           sym(a.nvrtc, "nvrtcGetLoweredName", a.GetLoweredName) && sym(a.nvrtc, "nvrtcGetCUBINSize", a.GetCUBINSize) &&
           sym(a.nvrtc, "nvrtcGetCUBIN", a.GetCUBIN) && sym(a.nvrtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) &&
           sym(a.nvrtc, "nvrtcGetProgramLog", a.GetProgramLog) && sym(a.nvrtc, "nvrtcDestroyProgram", a.DestroyProgram) &&
           sym(a.cuda, "cuModuleLoadData", a.ModuleLoadData) && sym(a.cuda, "cuModuleGetFunction", a.ModuleGetFunction) &&
           sym(a.cuda, "cuFuncSetAttribute", a.FuncSetAttribute) && sym(a.cuda, "cuLaunchKernel", a.LaunchKernel);
  });
  return a;
}

struct UserModel {
  std::string name, src;
  int nb, p, m, ntheta;
  std::map<std::string, CUfunction> fns;   // key: kernel instantiation expression
};

std::mutex g_mu;
std::vector<UserModel> g_models;

}  // namespace

// Compile (or fetch) `expr`, e.g. "rodeo::dalton_kernel<double, UserModel, 0, 1, 1>", for user model `id`.
int user_kernel(int id, const std::string& expr, CUfunction* out) {
  Api& a = api();
  if (!a.ok) { set_error("NVRTC path unavailable: could not dlopen libnvrtc.so.12 / libcuda.so.1"); return RODEO_ERR_NVRTC; }
  std::lock_guard<std::mutex> lk(g_mu);
  const int k = id - RODEO_MODEL_USER_BASE;
  if (k < 0 || k >= (int)g_models.size()) { set_error("unknown user model id %d", id); return RODEO_ERR_INVALID; }
  UserModel& um = g_models[k];
  auto it = um.fns.find(expr);
  if (it != um.fns.end()) { *out = it->second; return RODEO_OK; }
  RODEO_CUDA_OK(cudaFree(0));   // make sure the primary context exists and is current
  const std::string tu = std::string("#include \"rodeo_kernels.cuh\"\n") + um.src + "\n";
  const char* hdr_src[3] = {kEmbeddedCore, kEmbeddedModels, kEmbeddedKernels};
  const char* hdr_name[3] = {"rodeo_core.cuh", "rodeo_models.cuh", "rodeo_kernels.cuh"};
  nvrtcProgram prog;
  if (a.CreateProgram(&prog, tu.c_str(), (um.name + ".cu").c_str(), 3, hdr_src, hdr_name) != NVRTC_SUCCESS) {
    set_error("nvrtcCreateProgram failed"); return RODEO_ERR_NVRTC;
  }
  a.AddNameExpression(prog, expr.c_str());
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const nvrtcResult rc = a.CompileProgram(prog, 3, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0; a.GetProgramLogSize(prog, &n);
    std::string log(n, '\0'); a.GetProgramLog(prog, &log[0]);
    if (log.size() > 400) log = log.substr(0, 400);
    set_error("NVRTC compilation of model '%s' failed: %s", um.name.c_str(), log.c_str());
    a.DestroyProgram(&prog);
    return RODEO_ERR_NVRTC;
  }
  const char* lowered = nullptr;
  a.GetLoweredName(prog, expr.c_str(), &lowered);
  size_t nbin = 0; a.GetCUBINSize(prog, &nbin);
  std::vector<char> cubin(nbin); a.GetCUBIN(prog, cubin.data());
  CUmodule mod; CUfunction fn;
  if (!lowered || a.ModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS ||
      a.ModuleGetFunction(&fn, mod, lowered) != CUDA_SUCCESS) {
    set_error("loading the NVRTC cubin of model '%s' failed", um.name.c_str());
    a.DestroyProgram(&prog);
    return RODEO_ERR_NVRTC;
  }
  a.DestroyProgram(&prog);
  um.fns[expr] = fn;
  *out = fn;
  return RODEO_OK;
}

int user_dims(int id, int* nb, int* p, int* m, int* ntheta) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int k = id - RODEO_MODEL_USER_BASE;
  if (k < 0 || k >= (int)g_models.size()) { set_error("unknown user model id %d", id); return RODEO_ERR_INVALID; }
  *nb = g_models[k].nb; *p = g_models[k].p; *m = g_models[k].m; *ntheta = g_models[k].ntheta;
  return RODEO_OK;
}

// Launch a user-model kernel: `args` are pointers to the kernel parameters AFTER the leading FilterConsts, which is
// packed here (Q[nb][p][p], R[nb][p(p+1)/2], W[nb][m][p] doubles: exactly the layout of FilterConsts<double,..>).
int user_launch(const RodeoProblem& p, const char* kernel, const char* extra_targs, const double* W, const double* Q,
                const double* R, int wcol, long long threads, int smem, std::vector<void*> args, cudaStream_t s) {
  int nb, ps, m, nth;
  if (int rc = user_dims(p.model_id, &nb, &ps, &m, &nth)) return rc;
  if (p.n_block != nb || p.n_bstate != ps || p.n_bmeas != m || p.n_theta != nth) {
    set_error("user model expects (n_block,n_bstate,n_bmeas,n_theta)=(%d,%d,%d,%d), got (%d,%d,%d,%d)", nb, ps, m, nth,
              p.n_block, p.n_bstate, p.n_bmeas, p.n_theta);
    return RODEO_ERR_INVALID;
  }
  const int qk = detect_structure<double>(Q, W, nb, ps, m, wcol);
  const int ns = ps * (ps + 1) / 2;
  std::vector<double> consts((size_t)nb * (ps * ps + ns + m * ps));
  size_t o = 0;
  for (int i = 0; i < nb * ps * ps; ++i) consts[o++] = Q[i];
  for (int b = 0; b < nb; ++b)
    for (int i = 0; i < ps; ++i)
      for (int j = i; j < ps; ++j) consts[o++] = R[(b * ps + i) * ps + j];
  for (int i = 0; i < nb * m * ps; ++i) consts[o++] = W[i];
  char expr[256];
  snprintf(expr, sizeof(expr), "rodeo::%s<double, UserModel, %d, %d%s>", kernel, p.interrogate, qk, extra_targs);
  CUfunction fn;
  if (int rc = user_kernel(p.model_id, expr, &fn)) return rc;
  Api& a = api();
  if (smem > 0) a.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
  args.insert(args.begin(), consts.data());
  if (threads <= 0) return RODEO_OK;
  const CUresult rc = a.LaunchKernel(fn, grid_for(threads, 32), 1, 1, 32, 1, 1, (unsigned)smem, (CUstream)s, args.data(), nullptr);
  g_launches++;
  if (rc != CUDA_SUCCESS) { set_error("cuLaunchKernel failed for %s (CUresult %d)", expr, (int)rc); return RODEO_ERR_CUDA; }
  return RODEO_OK;
}

// launch an arbitrary instantiation (no FilterConsts parameter), e.g. ode_init_pad_kernel<double, UserModel>
int user_launch_raw(int model_id, const char* expr, long long threads, int block, std::vector<void*> args,
                    cudaStream_t s) {
  CUfunction fn;
  if (int rc = user_kernel(model_id, expr, &fn)) return rc;
  if (threads <= 0) return RODEO_OK;
  const CUresult rc = api().LaunchKernel(fn, grid_for(threads, block), 1, 1, block, 1, 1, 0, (CUstream)s, args.data(), nullptr);
  g_launches++;
  if (rc != CUDA_SUCCESS) { set_error("cuLaunchKernel failed for %s (CUresult %d)", expr, (int)rc); return RODEO_ERR_CUDA; }
  return RODEO_OK;
}

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" int rodeo_b200_register_model_nvrtc(const char* name, const char* src, int n_block, int n_bstate,
                                               int n_bmeas, int n_theta, int* model_id) {
  if (!name || !src || !model_id) { set_error("NULL argument"); return RODEO_ERR_INVALID; }
  if (n_block < 1 || n_bstate < 1 || n_bmeas != 1 || n_theta < 0) {
    set_error("user models need n_block >= 1, n_bstate >= 1, n_bmeas == 1 (got %d, %d, %d)", n_block, n_bstate, n_bmeas);
    return RODEO_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  UserModel um;
  um.name = name;
  char guard[512];
  snprintf(guard, sizeof(guard),
           "\nstatic_assert(UserModel::NB == %d && UserModel::P == %d && UserModel::M == %d && UserModel::NTHETA == %d, "
           "\"UserModel dimensions differ from the registered ones\");\n", n_block, n_bstate, n_bmeas, n_theta);
  um.src = std::string(src) + guard;
  um.nb = n_block; um.p = n_bstate; um.m = n_bmeas; um.ntheta = n_theta;
  g_models.push_back(um);
  *model_id = RODEO_MODEL_USER_BASE + (int)g_models.size() - 1;
  return RODEO_OK;
}
