// altro_b200.cu — C ABI (include/altro_b200.h) over the kernels in kernels.cuh.
//
// Host side only marshals: it flattens the problem description into one blob, owns the device
// arrays, converts between the instance-major API layout and the tile-major device layout and
// launches kernels.  No numerical work of the hot path runs on the host and there is no CPU
// fallback — without a usable CUDA device every compute entry point fails.
#include "../../include/altro_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "backward_coop.cuh"
#include "outer.cuh"
#include "phased.cuh"

// large-state path, compiled in its own translation unit (altro_b200_large.cu) without FMA
// contraction so that its ill-conditioned LLT decisions match the CPU oracle bit for bit
cudaError_t altro_b200_launch_solve_large_32_8(const altro_b200::SolverParams& P, int mode, cudaStream_t st);
// the same path on the fp64 tensor instruction, one instance per warp (altro_b200_large_mma.cu)
cudaError_t altro_b200_launch_solve_large_mma_32_8(const altro_b200::SolverParams& P, int mode, int sm_count, cudaStream_t st);

using namespace altro_b200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ALTRO_B200_ERR_CUDA,                                                    \
                  std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                    \
  } while (0)

struct HostCost {
  std::vector<double> data;  // Q, R, H, q, r, c
};

}  // namespace

// ------------------------------------------------------------------------------------------
// Problem (host description)
// ------------------------------------------------------------------------------------------
struct altro_b200_problem {
  int n, m, N;
  int model = kUnicycle;
  std::vector<double> params;
  std::vector<float> h, t;
  std::vector<int> cost_id;            // per knot, -1 = unset
  std::vector<HostCost> costs;
  std::vector<std::vector<ConBlock>> eq, ineq;  // per knot, insertion order
  std::vector<double> x0;
  bool step_set = false, model_set = false;
};

namespace {

// Flatten to the device blob.  use_constraints = false drops every constraint (plain iLQR).
int build_blob(const altro_b200_problem& p, bool use_constraints, std::vector<char>* out,
               int* pmax_out) {
  const int n = p.n, m = p.m, N = p.N;
  for (int k = 0; k <= N; ++k)
    if (p.cost_id[k] < 0)
      return fail(ALTRO_B200_ERR_STATE, "cost function of knot " + std::to_string(k) + " is not set");
  if (!p.step_set) return fail(ALTRO_B200_ERR_STATE, "time step is not set");
  if (!p.model_set) return fail(ALTRO_B200_ERR_STATE, "dynamics model is not set");
  // constraint sets: dedupe identical knots
  std::vector<ConSet> sets;
  std::vector<int> conset_id(N + 1, 0);
  int pmax = 0;
  for (int k = 0; k <= N; ++k) {
    ConSet cs;
    std::memset(&cs, 0, sizeof(cs));
    if (use_constraints) {
      int row = 0;
      for (const auto* vec : {&p.eq[k], &p.ineq[k]})
        for (const ConBlock& b : *vec) {
          if (cs.nblocks >= kMaxBlocks)
            return fail(ALTRO_B200_ERR_UNSUPPORTED, "more than 4 constraints at one knot");
          ConBlock c = b;
          c.row0 = row;
          row += c.p;
          cs.blk[cs.nblocks++] = c;
        }
      cs.p_total = row;
      pmax = std::max(pmax, row);
      for (int bi = 0; bi < cs.nblocks; ++bi) {
        const ConBlock& c = cs.blk[bi];
        int shape = kShapeGeneric;
        if (c.kind == kControlBound && c.nl == m && c.nu == m && c.p == 2 * m) {
          bool canonical = true;
          for (int i = 0; i < 2 * m; ++i) canonical = canonical && c.idx[i] == i % m;
          if (canonical) shape = kShapeBoundsFull;
        } else if (c.kind == kCircle && c.p <= 4) {
          shape = kShapeCirclesOneChunk;
        }
        cs.hdr[bi] = make_int4(c.kind | (c.equality << 8) | (shape << 12) | (c.p << 16), c.row0, c.nl,
                               c.xi | (c.yi << 8));
      }
    }
    int id = -1;
    for (size_t i = 0; i < sets.size(); ++i)
      if (std::memcmp(&sets[i], &cs, sizeof(cs)) == 0) id = static_cast<int>(i);
    if (id < 0) {
      sets.push_back(cs);
      id = static_cast<int>(sets.size()) - 1;
    }
    conset_id[k] = id;
  }
  const int cost_stride = n * n + m * m + n * m + n + m + 1 + 1;  // + the CostShape slot
  auto align = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
  BlobHeader h;
  std::memset(&h, 0, sizeof(h));
  h.n = n; h.m = m; h.N = N; h.model = p.model;
  h.ncost = static_cast<int>(p.costs.size());
  h.nconset = static_cast<int>(sets.size());
  h.pmax = pmax;
  h.nparams = static_cast<int>(p.params.size());
  h.cost_stride = cost_stride;
  size_t off = align(sizeof(BlobHeader), 16);
  h.off_cost_id = static_cast<int>(off); off = align(off + sizeof(int) * (N + 1), 16);
  h.off_conset_id = static_cast<int>(off); off = align(off + sizeof(int) * (N + 1), 16);
  h.off_h = static_cast<int>(off); off = align(off + sizeof(float) * (N + 1), 16);
  h.off_t = static_cast<int>(off); off = align(off + sizeof(float) * (N + 1), 16);
  h.off_params = static_cast<int>(off); off = align(off + sizeof(double) * std::max<size_t>(1, p.params.size()), 16);
  h.off_cost = static_cast<int>(off); off = align(off + sizeof(double) * cost_stride * p.costs.size(), 16);
  h.off_conset = static_cast<int>(off); off = align(off + sizeof(ConSet) * sets.size(), 16);
  h.bytes = static_cast<int>(off);
  out->assign(off, 0);
  char* b = out->data();
  std::memcpy(b, &h, sizeof(h));
  {
    std::vector<int> cost_off(N + 1), conset_off(N + 1);
    for (int k = 0; k <= N; ++k) {
      cost_off[k] = h.off_cost + static_cast<int>(sizeof(double)) * cost_stride * p.cost_id[k];
      conset_off[k] = h.off_conset + static_cast<int>(sizeof(ConSet)) * conset_id[k];
    }
    std::memcpy(b + h.off_cost_id, cost_off.data(), sizeof(int) * (N + 1));
    std::memcpy(b + h.off_conset_id, conset_off.data(), sizeof(int) * (N + 1));
  }
  std::memcpy(b + h.off_h, p.h.data(), sizeof(float) * (N + 1));
  std::memcpy(b + h.off_t, p.t.data(), sizeof(float) * (N + 1));
  if (!p.params.empty())
    std::memcpy(b + h.off_params, p.params.data(), sizeof(double) * p.params.size());
  for (size_t i = 0; i < p.costs.size(); ++i) {
    const std::vector<double>& c = p.costs[i].data;  // Q, R, H, q, r, c
    double* dst = reinterpret_cast<double*>(b + h.off_cost + sizeof(double) * cost_stride * i);
    std::memcpy(dst, c.data(), sizeof(double) * (cost_stride - 1));
    bool diag = true;
    for (int col = 0; col < n; ++col)
      for (int row = 0; row < n; ++row)
        if (row != col && c[row + col * n] != 0.0) diag = false;
    for (int col = 0; col < m; ++col)
      for (int row = 0; row < m; ++row)
        if (row != col && c[n * n + row + col * m] != 0.0) diag = false;
    for (int q = 0; q < n * m; ++q)
      if (c[n * n + m * m + q] != 0.0) diag = false;
    if (std::getenv("ALTRO_B200_DENSE_COST")) diag = false;  // test knob: the dense path gives the same bits
    dst[cost_stride - 1] = diag ? kCostDiagonal : kCostDense;
  }
  std::memcpy(b + h.off_conset, sets.data(), sizeof(ConSet) * sets.size());
  *pmax_out = pmax;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Kernel dispatch per device-capable model
// ------------------------------------------------------------------------------------------
// The kernels of one (model, tile width): either instantiations compiled into this library
// (handles = host stubs, launched through the runtime) or functions of a module compiled at run
// time from a plug-in model's source (handles = CUfunction, launched through the driver; modules.inl).
enum KernelId : int {
  K_SOLVE = 0, K_PHASE, K_EXP, K_EXP_PHASED, K_CON_VALUES, K_BP_CTG, K_BP_STREAM, K_BP_PHASED,
  K_COOP_CTG, K_COOP_STREAM, K_COOP_PHASED, K_ROLL_WIDE, K_COST_WIDE, K_ACC_WIDE, K_ROLL_DEEP, K_COST_DEEP,
  K_ACC_DEEP, K_LS_WIDE, K_LS_DEEP1, K_LS_DEEP2, K_OUTER_REGEN, K_OUTER_ROLLOUT, K_OUTER_DUALS, K_OUTER_COST,
  K_MICROBENCH, K_NUM
};
struct KernelTable {
  bool driver = false;  // handles are CUfunction (run-time compiled module)
  int n = 0, m = 0, W = 0;
  bool coop = false;    // the two-lanes-per-instance backward pass exists for this model
  const void* f[K_NUM] = {};
};

struct Ops {
  std::function<cudaError_t(const SolverParams&, int mode, int budget, int parts, cudaStream_t)> solve;
  std::function<cudaError_t(const SolverParams&, int phase, cudaStream_t)> phase;
  std::function<cudaError_t(const SolverParams&, cudaStream_t)> expansions;
  std::function<cudaError_t(const SolverParams&, bool store_ctg, cudaStream_t)> backward_mat;
  std::function<cudaError_t(const SolverParams&, int k, double* out, cudaStream_t)> con_values;
  // phased engine (phased.cuh); empty when the tile width has no instantiation
  std::function<cudaError_t(const SolverParams&, cudaStream_t)> expansions_phased;
  std::function<cudaError_t(const SolverParams&, cudaStream_t)> backward_phased;
  std::function<cudaError_t(const SolverParams&, int mode, cudaStream_t)> ls_wide;
  std::function<cudaError_t(const SolverParams&, int mode, int max_instances, cudaStream_t)> ls_deep;
  // outer-loop work of the phased engine on a dense list (outer.cuh); returns the kernels launched
  std::function<cudaError_t(const SolverParams&, int mode, int sm_count, cudaStream_t, int* nlaunch)> outer_step;
  std::function<cudaError_t(const SolverParams&, double* sink, long long* out, int reps, cudaStream_t)> microbench;
  // split line search: rollout / per-knot cost / acceptance kernels (wide: all tiles; deep: the list)
  std::function<cudaError_t(const SolverParams&, int mode, cudaStream_t)> ls_split_wide;
  std::function<cudaError_t(const SolverParams&, int mode, int max_instances, cudaStream_t)> ls_split_deep;
  bool large = false;  // one instance per CTA (large.cuh): whole solves only, W = 1 layout
};

constexpr int kBpStages = 4;
constexpr int kCoopStages = 3;  // ring depth of the two-lanes-per-instance backward pass (47 KB per warp at n = 6)
constexpr int kPhasedTile = 8;  // tile width of the phased engine's workspaces
constexpr int kCostGrid = 148 * 12;  // persistent CTAs of the per-knot cost kernels

// models whose backward pass runs two lanes per instance (backward_coop.cuh): medium n with even n and m,
// tile width 8
constexpr bool use_coop(int n, int m, int W) {
  return W == kCoopTile && n >= 5 && n <= 12 && n % kCoopLanes == 0 && m % kCoopLanes == 0;
}
template <class M, int W>
constexpr bool kUseCoop = use_coop(M::n, M::m, W);

cudaError_t driver_launch(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args);  // modules.inl

// one launch, whichever kind of handle the table holds
cudaError_t launch(const KernelTable& kt, int id, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
  if (!kt.f[id]) return cudaErrorInvalidDeviceFunction;
  if (kt.driver) return driver_launch(kt.f[id], grid, block, smem, st, args);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kt.f[id], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  return cudaLaunchKernel(kt.f[id], grid, block, args, smem, st);
}

// shared-memory sizes of the kernels from the run-time dimensions (the same formulas as the
// constexpr helpers next to the kernels: stage_doubles, p_stage_doubles, p_roll_stage_doubles, ...)
struct Dims {
  int n, m, nz, nkd, nexp;
  Dims(int n_, int m_) : n(n_), m(m_), nz(n_ + m_), nkd(m_ * n_ + m_), nexp(exp_fields(n_, m_)) {}
  int stage(int pmax, int W) const {
    const int fwd = (nz + nkd + pmax) * W, bwd = (nz + pmax) * kWarp;
    return 2 * (fwd > bwd ? fwd : bwd);
  }
  int p_stage(int pmax, int WI) const { return 2 * (nz + nkd + pmax) * WI; }
  int p_roll_stage(int WI) const { return kRollStages * (nz + nkd) * WI; }
  int p_deep_warp(int pmax, int N, int WI) const { return p_stage(pmax, WI) + WI * N; }
  int coop_smem() const {  // backward_coop_smem<M, kCoopStages>()
    const int scratch = n * n + m * n + n * m + m * m + m + m * n + n + 1;
    return kCoopStages * nexp * kCoopInst * 8 + kCoopInst * scratch * 8 + kCoopStages * 8;
  }
};

Ops make_ops_from(const KernelTable& kt) {
  Ops o;
  const Dims D(kt.n, kt.m);
  const int W = kt.W;
  auto blob16 = [](const SolverParams& P) { return ((P.blob_bytes + 15) / 16) * 16; };
  auto arg = [](const void* p) { return const_cast<void*>(p); };
  o.solve = [=](const SolverParams& P, int mode, int budget, int parts, cudaStream_t st) -> cudaError_t {
    const size_t smem = blob16(P) + kSolveWarps * D.stage(P.pmax, W) * sizeof(double);
    void* args[] = {arg(&P), &mode, &budget, &parts};
    return launch(kt, K_SOLVE, (P.T + kSolveWarps - 1) / kSolveWarps, kSolveWarps * kWarp, smem, st, args);
  };
  o.phase = [=](const SolverParams& P, int phase, cudaStream_t st) -> cudaError_t {
    const size_t smem = blob16(P) + kSolveWarps * D.stage(P.pmax, W) * sizeof(double);
    void* args[] = {arg(&P), &phase};
    return launch(kt, K_PHASE, (P.T + kSolveWarps - 1) / kSolveWarps, kSolveWarps * kWarp, smem, st, args);
  };
  o.expansions = [=](const SolverParams& P, cudaStream_t st) -> cudaError_t {
    void* args[] = {arg(&P)};
    return launch(kt, K_EXP, dim3((P.B + 127) / 128, P.N + 1), 128, 0, st, args);
  };
  o.con_values = [=](const SolverParams& P, int k, double* out, cudaStream_t st) -> cudaError_t {
    void* args[] = {arg(&P), &k, &out};
    return launch(kt, K_CON_VALUES, (P.B + 127) / 128, 128, P.blob_bytes, st, args);
  };
  auto backward = [=](const SolverParams& P, int id_mat, int id_coop, cudaStream_t st) -> cudaError_t {
    void* args[] = {arg(&P)};
    if (kt.coop && !std::getenv("ALTRO_B200_NO_COOP"))  // medium n: two lanes per instance (backward_coop.cuh)
      return launch(kt, id_coop, (P.T + kCoopInst / kCoopTile - 1) / (kCoopInst / kCoopTile), kWarp, D.coop_smem(), st, args);
    const size_t smem = kBpStages * D.nexp * kWarp * sizeof(double) + kBpStages * 8;
    return launch(kt, id_mat, (P.T + kWarp / W - 1) / (kWarp / W), kWarp, smem, st, args);
  };
  o.backward_mat = [=](const SolverParams& P, bool store_ctg, cudaStream_t st) -> cudaError_t {
    return store_ctg ? backward(P, K_BP_CTG, K_COOP_CTG, st) : backward(P, K_BP_STREAM, K_COOP_STREAM, st);
  };
  if (W != kPhasedTile) return o;
  o.microbench = [=](const SolverParams& P, double* sink, long long* out, int reps, cudaStream_t st) -> cudaError_t {
    void* args[] = {arg(&P), &sink, &out, &reps};
    return launch(kt, K_MICROBENCH, 1, kWarp, P.blob_bytes, st, args);
  };
  o.expansions_phased = [=](const SolverParams& P, cudaStream_t st) -> cudaError_t {
    void* args[] = {arg(&P)};
    return launch(kt, K_EXP_PHASED, dim3((P.B + 127) / 128, P.N + 1), 128, 0, st, args);
  };
  o.backward_phased = [=](const SolverParams& P, cudaStream_t st) -> cudaError_t {
    return backward(P, K_BP_PHASED, K_COOP_PHASED, st);
  };
  o.ls_split_wide = [=](const SolverParams& P, int mode, cudaStream_t st) -> cudaError_t {
    const size_t smem_roll = blob16(P) + kLsWarps * D.p_roll_stage(W) * sizeof(double);
    const size_t smem_acc = kLsWarps * W * P.N * sizeof(double);
    const int grid = (P.T + kLsWarps - 1) / kLsWarps;
    void* a1[] = {arg(&P)};
    cudaError_t e = launch(kt, K_ROLL_WIDE, grid, kLsWarps * kWarp, smem_roll, st, a1);
    if (e != cudaSuccess) return e;
    const long items = static_cast<long>(P.T) * (P.N + 1);
    const int cgrid = static_cast<int>(std::min<long>((items + kLsWarps - 1) / kLsWarps, kCostGrid));
    e = launch(kt, K_COST_WIDE, cgrid, kLsWarps * kWarp, P.blob_bytes, st, a1);
    if (e != cudaSuccess) return e;
    void* a2[] = {arg(&P), &mode};
    return launch(kt, K_ACC_WIDE, grid, kLsWarps * kWarp, smem_acc, st, a2);
  };
  o.ls_split_deep = [=](const SolverParams& P, int mode, int max_instances, cudaStream_t st) -> cudaError_t {
    const size_t smem_roll = blob16(P) + kLsWarps * D.p_roll_stage(1) * sizeof(double);
    const size_t smem_acc = kLsWarps * P.N * sizeof(double);
    const int grid = (max_instances + kLsWarps - 1) / kLsWarps;
    int wide_tries = kWarp / W;
    void* a1[] = {arg(&P), &wide_tries};
    cudaError_t e = launch(kt, K_ROLL_DEEP, grid, kLsWarps * kWarp, smem_roll, st, a1);
    if (e != cudaSuccess) return e;
    const long items = static_cast<long>(max_instances) * (P.N + 1);
    const int cgrid = static_cast<int>(std::min<long>((items + kLsWarps - 1) / kLsWarps, kCostGrid));
    e = launch(kt, K_COST_DEEP, cgrid, kLsWarps * kWarp, P.blob_bytes, st, a1);
    if (e != cudaSuccess) return e;
    void* a2[] = {arg(&P), &mode, &wide_tries};
    return launch(kt, K_ACC_DEEP, grid, kLsWarps * kWarp, smem_acc, st, a2);
  };
  o.ls_wide = [=](const SolverParams& P, int mode, cudaStream_t st) -> cudaError_t {
    const size_t smem = blob16(P) + kLsWarps * D.p_stage(P.pmax, W) * sizeof(double);
    void* args[] = {arg(&P), &mode};
    return launch(kt, K_LS_WIDE, (P.T + kLsWarps - 1) / kLsWarps, kLsWarps * kWarp, smem, st, args);
  };
  o.outer_step = [=](const SolverParams& P, int mode, int sm_count, cudaStream_t st, int* nlaunch) -> cudaError_t {
    // grids: enough CTAs for every instance to have outer work pending (first slot of a solve),
    // capped at a few waves; the kernels loop over the device-side list with a grid stride
    const int nk = P.N + 1;
    const int lane_grid = std::min((P.B + kOuterThreads - 1) / kOuterThreads, sm_count * 8);
    const long items = static_cast<long>(P.B) * nk;
    const int item_grid = static_cast<int>(std::min<long>((items + kOuterThreads - 1) / kOuterThreads, sm_count * 16));
    int force = std::getenv("ALTRO_B200_ALWAYS_ROLLOUT") ? 1 : 0;
    void* a1[] = {arg(&P)};
    void* a2[] = {arg(&P), &force};
    k_outer_select<<<(P.B + 255) / 256, 256, 0, st>>>(P);
    cudaError_t e = launch(kt, K_OUTER_REGEN, lane_grid, kOuterThreads, 0, st, a2);
    if (e == cudaSuccess) e = launch(kt, K_OUTER_DUALS, item_grid, kOuterThreads, 0, st, a1);
    k_outer_decide<<<lane_grid, kOuterThreads, 0, st>>>(P);
    if (e == cudaSuccess) e = launch(kt, K_OUTER_ROLLOUT, lane_grid, kOuterThreads, 0, st, a2);
    if (e == cudaSuccess) e = launch(kt, K_OUTER_COST, item_grid, kOuterThreads, 0, st, a1);
    k_outer_finish<<<lane_grid, kOuterThreads, 0, st>>>(P, mode);
    *nlaunch = 7;
    return e != cudaSuccess ? e : cudaGetLastError();
  };
  o.ls_deep = [=](const SolverParams& P, int mode, int max_instances, cudaStream_t st) -> cudaError_t {
    int wide_tries = kWarp / W;
    void* args[] = {arg(&P), &mode, &wide_tries};
    if (P.opt.line_search_max_iterations - kWarp / W <= kWarp / 2) {  // two instances per warp
      const size_t smem = blob16(P) + kLsWarps * D.p_deep_warp(P.pmax, P.N, 2) * sizeof(double);
      return launch(kt, K_LS_DEEP2, (max_instances + 2 * kLsWarps - 1) / (2 * kLsWarps), kLsWarps * kWarp, smem, st, args);
    }
    const size_t smem = blob16(P) + kLsWarps * D.p_deep_warp(P.pmax, P.N, 1) * sizeof(double);
    return launch(kt, K_LS_DEEP1, (max_instances + kLsWarps - 1) / kLsWarps, kLsWarps * kWarp, smem, st, args);
  };
  return o;
}

// kernel table of an instantiation compiled into this library
template <class M, int W>
KernelTable builtin_table() {
  KernelTable kt;
  kt.n = M::n; kt.m = M::m; kt.W = W;
  kt.f[K_SOLVE] = reinterpret_cast<const void*>(k_solve<M, W>);
  kt.f[K_PHASE] = reinterpret_cast<const void*>(k_phase<M, W>);
  kt.f[K_EXP] = reinterpret_cast<const void*>(k_update_expansions<M, W, false>);
  kt.f[K_CON_VALUES] = reinterpret_cast<const void*>(k_constraint_values<M, W>);
  kt.f[K_BP_CTG] = reinterpret_cast<const void*>(k_backward_mat<M, W, kBpStages, true, false>);
  kt.f[K_BP_STREAM] = reinterpret_cast<const void*>(k_backward_mat<M, W, kBpStages, false, false>);
  if constexpr (W == kPhasedTile) {
    kt.f[K_EXP_PHASED] = reinterpret_cast<const void*>(k_update_expansions<M, W, true>);
    kt.f[K_BP_PHASED] = reinterpret_cast<const void*>(k_backward_mat<M, W, kBpStages, false, true>);
    if constexpr (kUseCoop<M, W>) {
      kt.coop = true;
      kt.f[K_COOP_CTG] = reinterpret_cast<const void*>(k_backward_coop<M, kCoopStages, true, false>);
      kt.f[K_COOP_STREAM] = reinterpret_cast<const void*>(k_backward_coop<M, kCoopStages, false, false>);
      kt.f[K_COOP_PHASED] = reinterpret_cast<const void*>(k_backward_coop<M, kCoopStages, false, true>);
    }
    kt.f[K_ROLL_WIDE] = reinterpret_cast<const void*>(k_roll_wide<M, W>);
    kt.f[K_COST_WIDE] = reinterpret_cast<const void*>(k_cost_wide<M, W>);
    kt.f[K_ACC_WIDE] = reinterpret_cast<const void*>(k_acc_wide<M, W>);
    kt.f[K_ROLL_DEEP] = reinterpret_cast<const void*>(k_roll_deep<M, W>);
    kt.f[K_COST_DEEP] = reinterpret_cast<const void*>(k_cost_deep<M, W>);
    kt.f[K_ACC_DEEP] = reinterpret_cast<const void*>(k_acc_deep<M, W>);
    kt.f[K_LS_WIDE] = reinterpret_cast<const void*>(k_ls_wide<M, W>);
    kt.f[K_LS_DEEP1] = reinterpret_cast<const void*>(k_ls_deep<M, W, 1>);
    kt.f[K_LS_DEEP2] = reinterpret_cast<const void*>(k_ls_deep<M, W, 2>);
    kt.f[K_OUTER_REGEN] = reinterpret_cast<const void*>(k_outer_rollout<M, W, true>);
    kt.f[K_OUTER_ROLLOUT] = reinterpret_cast<const void*>(k_outer_rollout<M, W, false>);
    kt.f[K_OUTER_DUALS] = reinterpret_cast<const void*>(k_outer_duals<M, W>);
    kt.f[K_OUTER_COST] = reinterpret_cast<const void*>(k_outer_cost<M, W>);
    kt.f[K_MICROBENCH] = reinterpret_cast<const void*>(k_microbench<M, W>);
  }
  return kt;
}

template <class M, int W>
Ops make_ops() {
  return make_ops_from(builtin_table<M, W>());
}

#include "modules.inl"

int default_engine();

// Large-state path: engine "fused" = the exact-order kernel, one instance per CTA (large.cuh, bit-equal to
// the oracle); otherwise the tensor-instruction kernel, one instance per warp (large_mma.cuh, equal to rounding).
Ops make_large_ops_32_8() {
  Ops o;
  o.large = true;
  if (default_engine() == ALTRO_B200_ENGINE_FUSED) {
    o.solve = [](const SolverParams& P, int mode, int, int, cudaStream_t st) -> cudaError_t {
      return altro_b200_launch_solve_large_32_8(P, mode, st);
    };
  } else {
    o.solve = [](const SolverParams& P, int mode, int, int, cudaStream_t st) -> cudaError_t {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      return altro_b200_launch_solve_large_mma_32_8(P, mode, sms, st);
    };
  }
  return o;
}

// Tile widths with instantiations: one.  Round 1 compiled five widths x four models x ~20 kernels
// eagerly (46 MB, 3.6 - 15 minutes); the phased engine only ever used width 8 and the fused engine — kept
// as its cross-check — runs on it too.  ALTRO_DEV_BUILD (altro_cpp_b200/build.py --dev) additionally
// restricts the models to the unicycle: a kernel-tuning build.
constexpr int kWidths[] = {8};
bool width_available(int W) {
  for (int w : kWidths)
    if (w == W) return true;
  return false;
}

template <class M>
bool ops_for_width(int W, Ops* out) {
  if (!width_available(W)) return false;
  if (W == 8) { if (out) *out = make_ops<M, 8>(); return true; }
  return false;
}

// The cart-pole ships as a plug-in (plugins/cartpole.cuh): ALTRO_B200_MODEL_CARTPOLE resolves to its id.
int builtin_cartpole_id() {
  static int id = -1;
  static std::once_flag once;
  std::call_once(once, [] {
    const std::string text = read_file(library_dir() + "/plugins/cartpole.cuh");
    if (text.empty()) return;
    std::lock_guard<std::mutex> lock(g_plugin_mutex);
    PluginModel pm;
    pm.name = "Cartpole";
    pm.source = text;
    pm.n = 4; pm.m = 1; pm.nparams = 4;
    g_plugins.push_back(pm);
    id = kFirstPluginId + static_cast<int>(g_plugins.size()) - 1;
  });
  return id;
}
// Triple integrators other than the perf benchmark's dof = 2 (compiled in) are instantiated at run time from
// the TripleIntegrator<dof> template of device.cuh: the "plug-in" is one line.  dof <= 3 (n <= 9).
int builtin_triple_id(int dof) {
  static std::map<int, int> ids;
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  auto it = ids.find(dof);
  if (it != ids.end()) return it->second;
  PluginModel pm;
  pm.name = "TripleIntegratorDof" + std::to_string(dof);
  pm.source = "struct " + pm.name + " : TripleIntegrator<" + std::to_string(dof) + "> {};\n";
  pm.n = 3 * dof; pm.m = dof; pm.nparams = 0;
  g_plugins.push_back(pm);
  const int id = kFirstPluginId + static_cast<int>(g_plugins.size()) - 1;
  ids[dof] = id;
  return id;
}
int resolve_model(int model, int n = 0, int m = 0) {
  if (model == kCartpole) return builtin_cartpole_id();
  if (model == kTripleIntegrator && (m == 1 || m == 3) && n == 3 * m) return builtin_triple_id(m);
  return model;
}

// out == nullptr: is there a device path for (n, m, model) at tile width W?  (no compilation, no device)
// Otherwise fills *out; for a plug-in model that may compile or load its module on the CURRENT device.
int lookup_ops_rc(int n, int m, int model, int W, Ops* out) {
  model = resolve_model(model, n, m);
  if (model >= kFirstPluginId) {
    {
      std::lock_guard<std::mutex> lock(g_plugin_mutex);
      const int idx = model - kFirstPluginId;
      if (idx >= static_cast<int>(g_plugins.size()) || g_plugins[idx].n != n || g_plugins[idx].m != m || W != kPhasedTile)
        return ALTRO_B200_ERR_UNSUPPORTED;
    }
    if (!out) return 0;
    int device = 0;
    cudaGetDevice(&device);
    return plugin_ops(model, device, out);
  }
#ifdef ALTRO_DEV_TI2  // kernel-tuning build for the n = 6 path
  if (model == kTripleIntegrator && n == 6 && m == 2) return ops_for_width<TripleIntegrator<2>>(W, out) ? 0 : ALTRO_B200_ERR_UNSUPPORTED;
  return ALTRO_B200_ERR_UNSUPPORTED;
#endif
  bool ok = false;
  if (model == kUnicycle && n == 3 && m == 2) ok = ops_for_width<Unicycle>(W, out);
#ifndef ALTRO_DEV_BUILD
  else if (model == kTripleIntegrator && n == 6 && m == 2) ok = ops_for_width<TripleIntegrator<2>>(W, out);
  else if (model == kLinear && n == 32 && m == 8) { if (out) *out = make_large_ops_32_8(); ok = true; }
#endif
  return ok ? 0 : ALTRO_B200_ERR_UNSUPPORTED;
}
bool lookup_ops(int n, int m, int model, int W, Ops* out) { return lookup_ops_rc(n, m, model, W, out) == 0; }

// Tile width: the serial sweeps are latency-bound, so a B200 wants >= ~14 warps per SM in flight.
// Narrow tiles turn a small batch into more warps and free lane groups for the parallel line
// search; a batch that already fills the machine uses full-width tiles.
// Warps of the solve kernel that fit on one SM at its register budget (kernels.cuh:
// ALTRO_SOLVE_MINB blocks of ALTRO_SOLVE_WARPS warps; narrow tiles get the full register file).
int resident_warps_per_sm(int W) {
  return (W <= 4 ? ALTRO_SOLVE_MINB_NARROW : ALTRO_SOLVE_MINB) * ALTRO_SOLVE_WARPS;
}

// Tile width: the narrowest tile (most warps, most lane groups for the parallel line search and
// the knot-parallel backward sweep) whose warps are all resident at once — a second wave would
// double the makespan of these latency-bound sweeps.
int g_default_engine = -1;  // -1: not set through the API -> environment, then built-in default
int default_engine() {
  if (g_default_engine >= 0) return g_default_engine;
  if (const char* e = std::getenv("ALTRO_B200_ENGINE")) {
    if (std::strcmp(e, "fused") == 0) return ALTRO_B200_ENGINE_FUSED;
    if (std::strcmp(e, "phased") == 0) return ALTRO_B200_ENGINE_PHASED;
  }
  return ALTRO_B200_ENGINE_PHASED;
}

int choose_tile_width(int batch, int sm_count) {
  if (const char* e = std::getenv("ALTRO_B200_TILE")) {
    const int w = std::atoi(e);
    if (width_available(w)) return w;
  }
  int pick = kWidths[sizeof(kWidths) / sizeof(int) - 1];
  for (int w = 32; w >= 2; w /= 2) {
    if (!width_available(w)) continue;
    const long warps = (batch + w - 1) / w;
    if (warps <= static_cast<long>(sm_count) * resident_warps_per_sm(w)) pick = w;
  }
  return pick;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Solver
// ------------------------------------------------------------------------------------------
// ALTRO_B200_NO_FILL_SYNC=1 restores the round-1 behaviour (no synchronisation after the zero-fill of
// lazily allocated scratch) — only to demonstrate the race it caused on non-blocking streams.
static void sync_after_fill() {
  if (!std::getenv("ALTRO_B200_NO_FILL_SYNC")) cudaDeviceSynchronize();
}

static int env_split_max() {
  const char* e = std::getenv("ALTRO_B200_SPLIT_MAX");
  return e ? std::atoi(e) : 2048;
}

struct altro_b200_solver {
  int n, m, N, B, T, Bp, W, G, pmax, device, use_al;
  int engine = ALTRO_B200_ENGINE_FUSED;
  Ops ops;
  SolverParams P;
  char* d_blob = nullptr;
  int blob_bytes = 0;
  double* d_io = nullptr;   // staging for instance-major transfers
  size_t io_bytes = 0;
  size_t dev_bytes = 0;
  int64_t launches = 0;
  std::vector<void*> allocs;
  bool inputs_set = false;
  bool insolve_ready = false;     // measurement: phases set up for altro_b200_backward_pass_insolve
  bool expansions_valid = false;  // EXP holds the records altro_b200_update_expansions wrote for the current Z_
  int model = 0, sm_count = 148;
  std::vector<int> p_knot;  // constraint rows per knot (ALCost order)
  // secondary workspaces: unfinished instances are re-packed into them between k_solve launches
  struct Secondary {
    SolverParams P;
    Ops ops;
    bool allocated = false;
    size_t zcap[kMaxZ] = {};  // instances each trajectory buffer has room for
  } sec[2];
  int* d_list = nullptr;
  int* h_count = nullptr;  // pinned
  int hist_instances = 0, hist_rows = 0;  // per-iteration SolverStats rows recorded for the first instances

  int alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    allocs.push_back(*p);
    dev_bytes += bytes;
    // debug aid (tools/gpu_repro.py): fill every allocation with 0xFF (NaN doubles, -1 ints) so that a
    // read of memory the solve never wrote shows up in the results instead of depending on the heap
    if (std::getenv("ALTRO_B200_POISON")) {
      cudaMemset(*p, 0xFF, bytes);
      cudaDeviceSynchronize();
    }
    return 0;
  }
  // Lazily allocated arrays that must start as zeros.  cudaMemset runs on the legacy default stream,
  // which a non-blocking user stream (e.g. a torch stream) does not wait for: synchronise so that the
  // zeros cannot land after kernels queued next on the caller's stream.
  int alloc_zeroed(void** p, size_t bytes) {
    int rc = alloc(p, bytes);
    if (rc) return rc;
    cudaMemset(*p, 0, bytes);
    sync_after_fill();
    return 0;
  }
  int ensure_io(size_t bytes) {
    if (bytes <= io_bytes) return 0;
    if (d_io) { cudaFree(d_io); dev_bytes -= io_bytes; }
    cudaError_t e = cudaMalloc(&d_io, bytes);
    if (e != cudaSuccess) { d_io = nullptr; io_bytes = 0; return fail(ALTRO_B200_ERR_CUDA, "cudaMalloc(staging) failed"); }
    io_bytes = bytes;
    dev_bytes += bytes;
    if (std::getenv("ALTRO_B200_POISON")) cudaMemset(d_io, 0xFF, bytes);
    return 0;
  }
  int ensure_phased() {  // scratch of the phased engine: expansions, lists, line-search / outer-step scratch
    const size_t knots = static_cast<size_t>(T) * (N + 1) * W * sizeof(double);
    int rc;
    if (!P.EXP && (rc = alloc_zeroed(reinterpret_cast<void**>(&P.EXP), knots * exp_fields(n, m)))) return rc;
    // the split line-search kernels only run while at most split_cap instances are in flight
    P.split_cap = std::min(Bp, std::max(kWarp, env_split_max()));
    if (!P.CAND) {
      const size_t bytes = static_cast<size_t>(P.split_cap) * (N + 1) * (n + m) * kWarp * sizeof(double);
      if ((rc = alloc(reinterpret_cast<void**>(&P.CAND), bytes))) return rc;
    }
    if (!P.list && (rc = alloc(reinterpret_cast<void**>(&P.list), static_cast<size_t>(Bp) * sizeof(int)))) return rc;
    if (!P.olist && (rc = alloc(reinterpret_cast<void**>(&P.olist), static_cast<size_t>(Bp) * sizeof(int)))) return rc;
    if (!P.COSTK) {
      // rows of 32 per-knot costs: one per tile (wide kernels) or per list entry (split deep kernels);
      // the outer-step kernels use it as [entry][N+1], at most Bp entries
      const size_t rows = std::max<size_t>(std::max(T, P.split_cap), (static_cast<size_t>(Bp) + kWarp - 1) / kWarp);
      if ((rc = alloc(reinterpret_cast<void**>(&P.COSTK), static_cast<size_t>(N + 1) * rows * kWarp * sizeof(double)))) return rc;
      if ((rc = alloc(reinterpret_cast<void**>(&P.TRYST), rows * kWarp * sizeof(int)))) return rc;
    }
    return 0;
  }
  int ensure_stepwise() {  // EXP / CTG / COSTS are only needed by the step-wise API
    const size_t knots = static_cast<size_t>(T) * (N + 1) * W * sizeof(double);
    int rc;
    if (!P.EXP && (rc = alloc_zeroed(reinterpret_cast<void**>(&P.EXP), knots * exp_fields(n, m)))) return rc;
    if (!P.CTG) {
      if ((rc = alloc_zeroed(reinterpret_cast<void**>(&P.CTG), knots * (n * n + n)))) return rc;
      if ((rc = alloc_zeroed(reinterpret_cast<void**>(&P.COSTS), knots))) return rc;
    }
    return 0;
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Largest double t with sqrt(t) <= mx, so that `sqrt(s) > mx` and `s > t` are the same predicate
// for every s >= 0 (sqrt is monotone and correctly rounded on host and device alike).
double sqrt_threshold(double mx) {
  if (std::isnan(mx)) return std::numeric_limits<double>::quiet_NaN();  // sqrt(s) > NaN is never true
  if (mx < 0) return -1.0;                                              // always true for s >= 0
  if (std::isinf(mx)) return mx;
  double t = mx * mx;
  if (std::isinf(t)) return std::numeric_limits<double>::max();
  while (std::sqrt(t) > mx) t = std::nextafter(t, 0.0);
  while (std::sqrt(std::nextafter(t, HUGE_VAL)) <= mx) t = std::nextafter(t, HUGE_VAL);
  return t;
}

DevOptions to_dev(const altro_b200_options& o) {
  DevOptions d;
  std::memset(&d, 0, sizeof(d));
  d.max_iterations_total = o.max_iterations_total;
  d.max_iterations_outer = o.max_iterations_outer;
  d.max_iterations_inner = o.max_iterations_inner;
  d.bp_reg_fail_threshold = o.bp_reg_fail_threshold;
  d.check_forwardpass_bounds = o.check_forwardpass_bounds;
  d.line_search_max_iterations = o.line_search_max_iterations;
  d.reset_duals = o.reset_duals;
  d.skip_repeated_iterations = o.skip_repeated_iterations;
  d.cost_tolerance = o.cost_tolerance;
  d.gradient_tolerance = o.gradient_tolerance;
  d.bp_reg_increase_factor = o.bp_reg_increase_factor;
  d.bp_reg_initial = o.bp_reg_initial;
  d.bp_reg_max = o.bp_reg_max;
  d.bp_reg_min = o.bp_reg_min;
  d.state_max = o.state_max;
  d.control_max = o.control_max;
  d.state_max_sq = sqrt_threshold(o.state_max);
  d.control_max_sq = sqrt_threshold(o.control_max);
  d.line_search_lower_bound = o.line_search_lower_bound;
  d.line_search_upper_bound = o.line_search_upper_bound;
  d.line_search_decrease_factor = o.line_search_decrease_factor;
  d.constraint_tolerance = o.constraint_tolerance;
  d.maximum_penalty = o.maximum_penalty;
  d.initial_penalty = o.initial_penalty;
  d.penalty_scaling = o.penalty_scaling;
  return d;
}

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

int check_launch(altro_b200_solver* s, int nlaunch) {
  s->launches += nlaunch;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(ALTRO_B200_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

int fill_scalar(altro_b200_solver* s, int field, double v, cudaStream_t st) {
  k_fill_scalar<<<(s->Bp + 255) / 256, 256, 0, st>>>(s->P.sc + static_cast<size_t>(field) * s->Bp, v, s->Bp);
  return check_launch(s, 1);
}
int fill_int(altro_b200_solver* s, int field, int v, cudaStream_t st) {
  k_fill_int<<<(s->Bp + 255) / 256, 256, 0, st>>>(s->P.is + static_cast<size_t>(field) * s->Bp, v, s->Bp);
  return check_launch(s, 1);
}

// gather one tile-major array into instance-major staging, then to host
int unpack_to(altro_b200_solver* s, const double* src0, int K, int F, int f0,
              int nf, int k0, int nk, double* dst, bool dst_is_host, bool by_zsel,
              cudaStream_t st) {
  const size_t bytes = static_cast<size_t>(s->B) * nk * nf * sizeof(double);
  double* d = dst;
  if (dst_is_host) {
    int rc = s->ensure_io(bytes);
    if (rc) return rc;
    d = s->d_io;
  }
  dim3 grid((s->B + 127) / 128, nk);
  k_unpack<<<grid, 128, 0, st>>>(s->P, src0, K, F, f0, nf, k0, nk, d, by_zsel ? 1 : 0);
  int rc = check_launch(s, 1);
  if (rc) return rc;
  if (dst_is_host) {
    CU(cudaMemcpyAsync(dst, d, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return 0;
}

}  // namespace

// ==========================================================================================
extern "C" {

const char* altro_b200_last_error(void) { return g_err.c_str(); }
const char* altro_b200_version(void) { return "altro-cpp_b200 0.1.0 (sm_100a)"; }

void altro_b200_default_options(altro_b200_options* o) {  // solver_options.hpp:23-56
  std::memset(o, 0, sizeof(*o));
  o->max_iterations_total = 300;
  o->max_iterations_outer = 30;
  o->max_iterations_inner = 100;
  o->cost_tolerance = 1e-4;
  o->gradient_tolerance = 1e-2;
  o->bp_reg_increase_factor = 1.6;
  o->bp_reg_initial = 0.0;
  o->bp_reg_max = 1e8;
  o->bp_reg_min = 1e-8;
  o->bp_reg_fail_threshold = 100;
  o->check_forwardpass_bounds = 1;
  o->state_max = 1e8;
  o->control_max = 1e8;
  o->line_search_max_iterations = 20;
  o->line_search_lower_bound = 1e-8;
  o->line_search_upper_bound = 10.0;
  o->line_search_decrease_factor = 2;
  o->constraint_tolerance = 1e-4;
  o->maximum_penalty = 1e8;
  o->initial_penalty = 1.0;
  o->reset_duals = 1;
  o->penalty_scaling = 10.0;
}

int altro_b200_is_supported(int n, int m, int model) { return lookup_ops(n, m, model, kPhasedTile, nullptr) ? 1 : 0; }

// ---------------------------------------------------------------- plug-in models
int altro_b200_register_model(const char* name, const char* cuda_source, int n, int m, int nparams, int* model_id) {
  if (!name || !cuda_source || !model_id || n <= 0 || m <= 0 || n > 16 || m > 16 || nparams < 0)
    return fail(ALTRO_B200_ERR_ARG, "register_model: bad argument (1 <= n, m <= 16)");
  const std::string nm = name;
  if (nm.empty() || !(std::isalpha(static_cast<unsigned char>(nm[0])) || nm[0] == '_'))
    return fail(ALTRO_B200_ERR_ARG, "register_model: the name must be a C++ identifier");
  for (char c : nm)
    if (!(std::isalnum(static_cast<unsigned char>(c)) || c == '_')) return fail(ALTRO_B200_ERR_ARG, "register_model: the name must be a C++ identifier");
  if (std::string(cuda_source).find("struct " + nm) == std::string::npos)
    return fail(ALTRO_B200_ERR_ARG, "register_model: the source must define `struct " + nm + "`");
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  for (size_t i = 0; i < g_plugins.size(); ++i)
    if (g_plugins[i].name == nm) {
      if (g_plugins[i].source != cuda_source || g_plugins[i].n != n || g_plugins[i].m != m || g_plugins[i].nparams != nparams)
        return fail(ALTRO_B200_ERR_STATE, "register_model: a different model named '" + nm + "' is already registered");
      *model_id = kFirstPluginId + static_cast<int>(i);
      return 0;
    }
  PluginModel pm;
  pm.name = nm;
  pm.source = cuda_source;
  pm.n = n; pm.m = m; pm.nparams = nparams;
  g_plugins.push_back(pm);
  *model_id = kFirstPluginId + static_cast<int>(g_plugins.size()) - 1;
  return 0;
}
static int precompile_impl(int model, char* cache_path, int cache_path_cap);
int altro_b200_precompile_model(int model, char* cache_path, int cache_path_cap) {
  return precompile_impl(resolve_model(model), cache_path, cache_path_cap);
}
int altro_b200_precompile_builtin_model(int model, int n, int m, char* cache_path, int cache_path_cap) {
  return precompile_impl(resolve_model(model, n, m), cache_path, cache_path_cap);
}
static int precompile_impl(int model, char* cache_path, int cache_path_cap) {
  PluginModel pm;
  {
    std::lock_guard<std::mutex> lock(g_plugin_mutex);
    const int idx = model - kFirstPluginId;
    if (idx < 0 || idx >= static_cast<int>(g_plugins.size())) return fail(ALTRO_B200_ERR_ARG, "precompile_model: not a plug-in model id");
    pm = g_plugins[idx];
  }
  CompiledModule cm;
  std::string path;
  int rc = compile_module(pm, &cm, &path);
  if (rc) return rc;
  if (cache_path && cache_path_cap > 0) std::snprintf(cache_path, static_cast<size_t>(cache_path_cap), "%s", path.c_str());
  return 0;
}

// ---------------------------------------------------------------- problem
int altro_b200_problem_create(int n, int m, int N, altro_b200_problem** out) {
  if (!out || n <= 0 || m <= 0 || N <= 0 || n > kMaxDim || m > kMaxDim)
    return fail(ALTRO_B200_ERR_ARG, "problem_create: bad dimensions");
  auto* p = new altro_b200_problem();
  p->n = n; p->m = m; p->N = N;
  p->h.assign(N + 1, 0.f);
  p->t.assign(N + 1, 0.f);
  p->cost_id.assign(N + 1, -1);
  p->eq.resize(N + 1);
  p->ineq.resize(N + 1);
  p->x0.assign(n, 0.0);
  *out = p;
  return 0;
}
void altro_b200_problem_destroy(altro_b200_problem* p) { delete p; }

int altro_b200_problem_set_model(altro_b200_problem* p, int model, const double* params, int nparams) {
  if (!p || nparams < 0 || (nparams > 0 && !params)) return fail(ALTRO_B200_ERR_ARG, "set_model: bad argument");
  if (model == kCartpole && nparams != 4) return fail(ALTRO_B200_ERR_ARG, "cartpole needs params mc, mp, l, g");
  if (model >= kFirstPluginId) {
    std::lock_guard<std::mutex> lock(g_plugin_mutex);
    const int idx = model - kFirstPluginId;
    if (idx >= static_cast<int>(g_plugins.size())) return fail(ALTRO_B200_ERR_ARG, "set_model: unknown plug-in model id");
    if (g_plugins[idx].nparams != nparams) return fail(ALTRO_B200_ERR_ARG, "set_model: the plug-in model takes " + std::to_string(g_plugins[idx].nparams) + " parameters");
  }
  p->model = model;
  p->params.assign(params, params + nparams);
  p->model_set = true;
  return 0;
}
int altro_b200_problem_set_uniform_step(altro_b200_problem* p, float h) {
  if (!p) return fail(ALTRO_B200_ERR_ARG, "null problem");
  for (int k = 0; k < p->N; ++k) {  // trajectory.hpp:122-130
    p->h[k] = h;
    p->t[k] = static_cast<float>(k) * h;
  }
  p->h[p->N] = 0.0f;
  p->t[p->N] = static_cast<float>(h) * p->N;
  p->step_set = true;
  return 0;
}
int altro_b200_problem_set_steps(altro_b200_problem* p, const float* t, const float* h) {
  if (!p || !t || !h) return fail(ALTRO_B200_ERR_ARG, "set_steps: null argument");
  for (int k = 0; k < p->N; ++k)
    if (!(h[k] > 0.0f)) return fail(ALTRO_B200_ERR_ARG, "set_steps: every step before the terminal knot must be positive");
  for (int k = 0; k <= p->N; ++k) {  // Trajectory::SetStep / SetTime per knot, altro/common/trajectory.hpp:119-120
    p->h[k] = h[k];
    p->t[k] = t[k];
  }
  p->step_set = true;
  return 0;
}
int altro_b200_problem_set_cost(altro_b200_problem* p, int k0, int k1, const double* Q, const double* R,
                                const double* H, const double* q, const double* r, double c) {
  if (!p || !Q || !R || !H || !q || !r) return fail(ALTRO_B200_ERR_ARG, "set_cost: null argument");
  if (k0 < 0 || k1 > p->N + 1 || k0 >= k1) return fail(ALTRO_B200_ERR_ARG, "set_cost: bad knot range");
  const int n = p->n, m = p->m;
  HostCost hc;
  hc.data.insert(hc.data.end(), Q, Q + n * n);
  hc.data.insert(hc.data.end(), R, R + m * m);
  hc.data.insert(hc.data.end(), H, H + n * m);
  hc.data.insert(hc.data.end(), q, q + n);
  hc.data.insert(hc.data.end(), r, r + m);
  hc.data.push_back(c);
  int id = -1;
  for (size_t i = 0; i < p->costs.size(); ++i)
    if (p->costs[i].data == hc.data) id = static_cast<int>(i);
  if (id < 0) {
    p->costs.push_back(hc);
    id = static_cast<int>(p->costs.size()) - 1;
  }
  for (int k = k0; k < k1; ++k) p->cost_id[k] = id;
  return 0;
}
int altro_b200_problem_add_goal(altro_b200_problem* p, int k, const double* xf) {
  if (!p || !xf || k < 0 || k > p->N) return fail(ALTRO_B200_ERR_ARG, "add_goal: bad argument");
  ConBlock b;
  std::memset(&b, 0, sizeof(b));
  b.kind = kGoal;
  b.equality = 1;
  b.p = p->n;
  for (int i = 0; i < p->n; ++i) b.a[i] = xf[i];
  p->eq[k].push_back(b);
  return 0;
}
int altro_b200_problem_add_control_bound(altro_b200_problem* p, int k, const double* lb, const double* ub) {
  if (!p || !lb || !ub || k < 0 || k > p->N) return fail(ALTRO_B200_ERR_ARG, "add_control_bound: bad argument");
  ConBlock b;
  std::memset(&b, 0, sizeof(b));
  b.kind = kControlBound;
  b.equality = 0;
  int row = 0;
  {
    int finite = 0;
    for (int i = 0; i < p->m; ++i) finite += (std::fabs(lb[i]) < DBL_MAX) + (std::fabs(ub[i]) < DBL_MAX);
    if (finite > kMaxDim) return fail(ALTRO_B200_ERR_UNSUPPORTED, "too many bound rows");
  }
  for (int i = 0; i < p->m; ++i) {
    if (lb[i] > ub[i]) return fail(ALTRO_B200_ERR_ARG, "Lower bound isn't less than the upper bound.");
    if (std::fabs(lb[i]) < DBL_MAX) {  // basic_constraints.hpp:136-143
      b.idx[row] = i;
      b.a[row] = lb[i];
      ++row;
    }
  }
  b.nl = row;
  for (int i = 0; i < p->m; ++i)
    if (std::fabs(ub[i]) < DBL_MAX) {
      b.idx[row] = i;
      b.a[row] = ub[i];
      ++row;
    }
  b.nu = row - b.nl;
  b.p = row;
  p->ineq[k].push_back(b);
  return 0;
}
static int add_circles_impl(altro_b200_problem* p, int k, int nc, const double* cx, const double* cy,
                            const double* cr, bool squared, int xi, int yi) {
  if (!p || !cx || !cy || !cr || k < 0 || k > p->N || nc <= 0 || nc > kMaxDim || xi < 0 || yi < 0 ||
      xi >= p->n || yi >= p->n)
    return fail(ALTRO_B200_ERR_ARG, "add_circles: bad argument");
  ConBlock b;
  std::memset(&b, 0, sizeof(b));
  b.kind = kCircle;
  b.equality = 0;
  b.p = nc;
  b.xi = xi;
  b.yi = yi;
  for (int i = 0; i < nc; ++i) {
    b.a[i] = cx[i];
    b.b[i] = cy[i];
    b.c[i] = squared ? cr[i] : cr[i] * cr[i];
  }
  p->ineq[k].push_back(b);
  return 0;
}
int altro_b200_problem_add_circles(altro_b200_problem* p, int k, int nc, const double* cx, const double* cy,
                                   const double* cr, int xi, int yi) {
  return add_circles_impl(p, k, nc, cx, cy, cr, false, xi, yi);
}
int altro_b200_problem_add_circles_r2(altro_b200_problem* p, int k, int nc, const double* cx, const double* cy,
                                      const double* cr2, int xi, int yi) {
  return add_circles_impl(p, k, nc, cx, cy, cr2, true, xi, yi);
}
int altro_b200_problem_set_initial_state(altro_b200_problem* p, const double* x0) {
  if (!p || !x0) return fail(ALTRO_B200_ERR_ARG, "set_initial_state: null argument");
  p->x0.assign(x0, x0 + p->n);
  return 0;
}

// ---------------------------------------------------------------- solver
int altro_b200_solver_create(const altro_b200_problem* p, int batch, int use_constraints, int device,
                             altro_b200_solver** out) {
  if (!p || !out || batch <= 0) return fail(ALTRO_B200_ERR_ARG, "solver_create: bad argument");
  if (!lookup_ops(p->n, p->m, p->model, kPhasedTile, nullptr))
    return fail(ALTRO_B200_ERR_UNSUPPORTED, "no device instantiation for (n=" + std::to_string(p->n) + ", m=" +
                                               std::to_string(p->m) + ", model=" + std::to_string(p->model) + ")");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return fail(ALTRO_B200_ERR_CUDA, "no usable CUDA device (there is no CPU fallback)");
  std::vector<char> blob;
  int pmax = 0;
  int rc = build_blob(*p, use_constraints != 0, &blob, &pmax);
  if (rc) return rc;
  DeviceGuard guard(device);
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
  // any failure below (typically cudaMalloc running out of memory at a large batch) releases what
  // was allocated so far: the handle is destroyed through the public destructor
  std::unique_ptr<altro_b200_solver, void (*)(altro_b200_solver*)> s(new altro_b200_solver(), altro_b200_solver_destroy);
  s->device = device;
  s->n = p->n; s->m = p->m; s->N = p->N; s->B = batch;
  Ops probe;
  lookup_ops(p->n, p->m, p->model, kPhasedTile, &probe);
  if (probe.large && pmax > 0)
    return fail(ALTRO_B200_ERR_UNSUPPORTED, "the large-state path (n=32) supports unconstrained problems only");
  if (probe.large && static_cast<int>(p->params.size()) != p->n * (p->n + p->m))
    return fail(ALTRO_B200_ERR_ARG, "linear model needs params = [A (n*n), B (n*m)]");
  s->engine = default_engine();
  s->W = probe.large ? 1 : choose_tile_width(batch, sm_count);
  if (s->engine == ALTRO_B200_ENGINE_PHASED && !probe.large) {
    if (std::getenv("ALTRO_B200_TILE") && s->W != kPhasedTile) s->engine = ALTRO_B200_ENGINE_FUSED;
    else s->W = kPhasedTile;
  }
  s->G = probe.large ? 1 : kWarp / s->W;
  s->T = (batch + s->W - 1) / s->W;
  s->Bp = s->T * s->W;
  s->pmax = pmax;
  s->p_knot.assign(p->N + 1, 0);
  if (use_constraints)
    for (int k = 0; k <= p->N; ++k) {
      for (const ConBlock& b : p->eq[k]) s->p_knot[k] += b.p;
      for (const ConBlock& b : p->ineq[k]) s->p_knot[k] += b.p;
    }
  s->device = device;
  s->use_al = use_constraints != 0;
  if ((rc = lookup_ops_rc(p->n, p->m, p->model, s->W, &s->ops))) {
    if (g_err.empty() || rc == ALTRO_B200_ERR_UNSUPPORTED) return fail(rc, "no kernels for this model at tile width " + std::to_string(s->W));
    return rc;  // the message of the module compiler / loader stands
  }
  std::memset(&s->P, 0, sizeof(s->P));
  SolverParams& P = s->P;
  P.B = batch; P.T = s->T; P.N = p->N; P.W = s->W; P.Bp = s->Bp;
  P.n = p->n; P.m = p->m; P.pmax = pmax; P.use_al = s->use_al;
  const size_t knots = static_cast<size_t>(s->T) * (p->N + 1) * s->W * sizeof(double);
  const int nz = p->n + p->m, nkd = p->m * p->n + p->m;
  if ((rc = s->alloc(reinterpret_cast<void**>(&s->d_blob), blob.size()))) return rc;
  for (int zb = 0; zb < 1 + s->G; ++zb) {
    if ((rc = s->alloc(reinterpret_cast<void**>(&P.Z[zb]), knots * nz))) return rc;
    CU(cudaMemset(P.Z[zb], 0, knots * nz));
  }
  if ((rc = s->alloc(reinterpret_cast<void**>(&P.KD), knots * nkd))) return rc;
  if (pmax > 0 && (rc = s->alloc(reinterpret_cast<void**>(&P.LAM), knots * pmax))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&P.X0), static_cast<size_t>(s->Bp) * p->n * sizeof(double)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&P.sc), static_cast<size_t>(S_NUM) * s->Bp * sizeof(double)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&P.is), static_cast<size_t>(I_NUM) * s->Bp * sizeof(int)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&P.counters), 8 * sizeof(int)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&s->d_list), static_cast<size_t>(s->Bp) * sizeof(int)))) return rc;
  CU(cudaMemset(P.counters, 0, 8 * sizeof(int)));
  CU(cudaHostAlloc(reinterpret_cast<void**>(&s->h_count), 4 * sizeof(int), cudaHostAllocDefault));
  s->model = p->model;
  s->sm_count = sm_count;
  CU(cudaMemcpy(s->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  s->blob_bytes = static_cast<int>(blob.size());
  P.blob = s->d_blob;
  P.blob_bytes = s->blob_bytes;
  CU(cudaMemset(P.KD, 0, knots * nkd));  // KnotPointFunctions::Init :271-278
  if (pmax > 0) CU(cudaMemset(P.LAM, 0, knots * pmax));
  CU(cudaMemset(P.X0, 0, static_cast<size_t>(s->Bp) * p->n * sizeof(double)));
  CU(cudaMemset(P.sc, 0, static_cast<size_t>(S_NUM) * s->Bp * sizeof(double)));
  CU(cudaMemset(P.is, 0, static_cast<size_t>(I_NUM) * s->Bp * sizeof(int)));
  altro_b200_options o;
  altro_b200_default_options(&o);
  P.opt = to_dev(o);
  if ((rc = fill_scalar(s.get(), S_PENALTY, 1.0, 0))) return rc;   // constraint_values.hpp:45 penalty_.setOnes
  if ((rc = fill_scalar(s.get(), S_CSRC_ALPHA, -1.0, 0))) return rc;
  if ((rc = fill_int(s.get(), I_STATUS, kUnsolved, 0))) return rc;
  if ((rc = fill_int(s.get(), I_STATUS_AL, kUnsolved, 0))) return rc;
  CU(cudaDeviceSynchronize());
  *out = s.release();
  return 0;
}

void altro_b200_solver_destroy(altro_b200_solver* s) {
  if (!s) return;
  DeviceGuard guard(s->device);
  for (void* p : s->allocs) cudaFree(p);
  for (auto& w : s->sec)
    for (int zb = 0; zb < kMaxZ; ++zb)
      if (w.P.Z[zb]) cudaFree(w.P.Z[zb]);
  if (s->d_io) cudaFree(s->d_io);
  if (s->h_count) cudaFreeHost(s->h_count);
  delete s;
}

int altro_b200_solver_set_options(altro_b200_solver* s, const altro_b200_options* o) {
  if (!s || !o) return fail(ALTRO_B200_ERR_ARG, "set_options: null argument");
  s->P.opt = to_dev(*o);
  return 0;
}
int altro_b200_solver_batch(const altro_b200_solver* s) { return s ? s->B : 0; }

static int set_inputs_impl(altro_b200_solver* s, const double* x0_dev, const double* U0_dev,
                           const double* u_nominal, cudaStream_t st) {
  Unom un;
  std::memset(&un, 0, sizeof(un));
  if (!U0_dev) {
    if (!u_nominal) return fail(ALTRO_B200_ERR_ARG, "set_inputs: need U0 or u_nominal");
    for (int i = 0; i < s->m; ++i) un.v[i] = u_nominal[i];
  }
  dim3 grid((s->Bp + 127) / 128, s->N + 1);
  k_pack_inputs<<<grid, 128, 0, st>>>(s->P, x0_dev, U0_dev, un, 1 + s->G);
  s->inputs_set = true;
  return check_launch(s, 1);
}

int altro_b200_solver_set_inputs_host(altro_b200_solver* s, const double* x0, const double* U0,
                                      const double* u_nominal, void* stream) {
  if (!s || !x0) return fail(ALTRO_B200_ERR_ARG, "set_inputs: null argument");
  DeviceGuard guard(s->device);
  const size_t bx = static_cast<size_t>(s->B) * s->n * sizeof(double);
  const size_t bu = U0 ? static_cast<size_t>(s->B) * s->N * s->m * sizeof(double) : 0;
  int rc = s->ensure_io(bx + bu);
  if (rc) return rc;
  CU(cudaMemcpyAsync(s->d_io, x0, bx, cudaMemcpyHostToDevice, S(stream)));
  double* dU = nullptr;
  if (U0) {
    dU = s->d_io + static_cast<size_t>(s->B) * s->n;
    CU(cudaMemcpyAsync(dU, U0, bu, cudaMemcpyHostToDevice, S(stream)));
  }
  return set_inputs_impl(s, s->d_io, dU, u_nominal, S(stream));
}
int altro_b200_solver_set_inputs_dev(altro_b200_solver* s, const double* x0_dev, const double* U0_dev,
                                     const double* u_nominal, void* stream) {
  if (!s || !x0_dev) return fail(ALTRO_B200_ERR_ARG, "set_inputs: null argument");
  DeviceGuard guard(s->device);
  return set_inputs_impl(s, x0_dev, U0_dev, u_nominal, S(stream));
}
int altro_b200_solver_set_states_host(altro_b200_solver* s, const double* X, void* stream) {
  if (!s || !X) return fail(ALTRO_B200_ERR_ARG, "set_states: null argument");
  DeviceGuard guard(s->device);
  const size_t bytes = static_cast<size_t>(s->B) * (s->N + 1) * s->n * sizeof(double);
  int rc = s->ensure_io(bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(s->d_io, X, bytes, cudaMemcpyHostToDevice, S(stream)));
  dim3 grid((s->B + 127) / 128, s->N + 1);
  k_set_states<<<grid, 128, 0, S(stream)>>>(s->P, s->d_io);
  return check_launch(s, 1);
}
int altro_b200_solver_set_penalty(altro_b200_solver* s, double rho, void* stream) {
  if (!s || !(rho >= 0)) return fail(ALTRO_B200_ERR_ARG, "Penalty must be positive.");
  DeviceGuard guard(s->device);
  return fill_scalar(s, S_PENALTY, rho, S(stream));
}
int altro_b200_solver_set_initial_cost(altro_b200_solver* s, double cost, void* stream) {
  if (!s || cost != cost) return fail(ALTRO_B200_ERR_ARG, "set_initial_cost: bad argument");
  DeviceGuard guard(s->device);
  return fill_scalar(s, S_INITIAL_COST, cost, S(stream));
}
int altro_b200_solver_set_duals_host(altro_b200_solver* s, int k, const double* lambda, int p, void* stream) {
  if (!s || !lambda || k < 0 || k > s->N || p < 0 || p > s->pmax)
    return fail(ALTRO_B200_ERR_ARG, "set_duals: bad argument");
  DeviceGuard guard(s->device);
  int rc = s->ensure_io(sizeof(double) * kMaxDim * kMaxBlocks);
  if (rc) return rc;
  CU(cudaMemcpyAsync(s->d_io, lambda, sizeof(double) * p, cudaMemcpyHostToDevice, S(stream)));
  k_fill_duals<<<(s->Bp + 127) / 128, 128, 0, S(stream)>>>(s->P, k, s->d_io, p);
  return check_launch(s, 1);
}

// ---------------------------------------------------------------- solves
// Secondary workspace with room for `cap` instances at any tile width.
static int ensure_secondary(altro_b200_solver* s, int which) {
  altro_b200_solver::Secondary& w = s->sec[which];
  if (w.allocated) return 0;
  const int cap = s->Bp + kWarp;
  const size_t knots = static_cast<size_t>(cap) * (s->N + 1) * sizeof(double);
  const int nz = s->n + s->m, nkd = s->m * s->n + s->m;
  std::memset(&w.P, 0, sizeof(w.P));
  int rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&w.P.KD), knots * nkd))) return rc;
  if (s->pmax > 0 && (rc = s->alloc(reinterpret_cast<void**>(&w.P.LAM), knots * s->pmax))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&w.P.X0), static_cast<size_t>(cap) * s->n * sizeof(double)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&w.P.sc), static_cast<size_t>(S_NUM) * cap * sizeof(double)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&w.P.is), static_cast<size_t>(I_NUM) * cap * sizeof(int)))) return rc;
  if ((rc = s->alloc(reinterpret_cast<void**>(&w.P.counters), 8 * sizeof(int)))) return rc;
  // defined contents for the padding lanes of the last tile (they are staged, never used)
  const int fill = std::getenv("ALTRO_B200_POISON") ? 0xFF : 0;
  cudaMemset(w.P.KD, fill, knots * nkd);
  if (s->pmax > 0) cudaMemset(w.P.LAM, fill, knots * s->pmax);
  cudaMemset(w.P.X0, fill, static_cast<size_t>(cap) * s->n * sizeof(double));
  cudaMemset(w.P.sc, 0, static_cast<size_t>(S_NUM) * cap * sizeof(double));
  cudaMemset(w.P.is, 0, static_cast<size_t>(I_NUM) * cap * sizeof(int));
  sync_after_fill();  // the legacy-stream memsets above are not ordered with a non-blocking user stream
  w.allocated = true;
  return 0;
}

// trajectory buffers 0 .. nbuf-1 of a secondary workspace, each with room for `count` instances
// (+ one tile of padding).  Narrow tiles need many buffers but hold few instances, wide tiles the
// opposite, so the footprint stays ~(W + 32) * 2000 instance-trajectories whatever the batch.
static int ensure_secondary_buffers(altro_b200_solver* s, int which, int nbuf, int count) {
  altro_b200_solver::Secondary& w = s->sec[which];
  const size_t need = static_cast<size_t>(count) + kWarp;
  const size_t per_instance = static_cast<size_t>(s->N + 1) * (s->n + s->m) * sizeof(double);
  bool fresh = false;
  for (int zb = 0; zb < nbuf; ++zb) {
    if (w.P.Z[zb] && w.zcap[zb] >= need) continue;
    if (w.P.Z[zb]) {
      cudaFree(w.P.Z[zb]);
      s->dev_bytes -= w.zcap[zb] * per_instance;
      w.P.Z[zb] = nullptr;
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&w.P.Z[zb]), need * per_instance);
    if (e != cudaSuccess) {
      w.zcap[zb] = 0;
      return fail(ALTRO_B200_ERR_CUDA, std::string("cudaMalloc(re-pack buffer): ") + cudaGetErrorString(e));
    }
    w.zcap[zb] = need;
    s->dev_bytes += need * per_instance;
    cudaMemset(w.P.Z[zb], std::getenv("ALTRO_B200_POISON") ? 0xFF : 0, need * per_instance);
    fresh = true;
  }
  // cudaMemset runs on the legacy default stream, which a non-blocking user stream does not wait
  // for: without this the memset could land AFTER the re-pack kernel queued next on the solve's stream
  if (fresh) sync_after_fill();
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

// AugmentedLagrangianiLQR::Solve / iLQR::Solve for the whole batch.  Everything stays on the
// device; the host only relaunches the resumable solve kernel and, when most instances are
// done, re-packs the unfinished ones densely (their results are identical wherever they run).
static int solve_impl(altro_b200_solver* s, int mode, cudaStream_t st) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  DeviceGuard guard(s->device);
  // a solver that records history keeps part of its bookkeeping in registers: it solves in one launch
  const int budget = s->P.HIST ? (1 << 30) : std::max(1, env_int("ALTRO_B200_BUDGET", 16));
  const int repack_pct = s->P.HIST ? 0 : env_int("ALTRO_B200_REPACK_PCT", 100);  // 0 disables re-packing
  int rc;
  if ((rc = fill_int(s, I_PHASE, mode == 1 ? kPhAlInit : kPhSolveStart, st))) return rc;
  if ((rc = fill_int(s, I_LSFAIL, 0, st))) return rc;
  k_iota<<<(s->Bp + 255) / 256, 256, 0, st>>>(s->P.is + static_cast<size_t>(I_ORIG) * s->Bp, s->Bp);
  if ((rc = check_launch(s, 1))) return rc;
  SolverParams* cur = &s->P;
  Ops cur_ops = s->ops;
  int cur_sec = -1;  // -1: primary
  auto scatter_back = [&](SolverParams& from) -> int {
    dim3 grid((from.B + 127) / 128, s->N + 2);
    k_move_instances<<<grid, 128, 0, st>>>(from, s->P, nullptr, from.B, 1);
    return check_launch(s, 1);
  };
  for (int launch = 0; launch < 100000; ++launch) {
    cur->opt = s->P.opt;
    CU(cudaMemsetAsync(cur->counters, 0, 4 * sizeof(int), st));
    cudaError_t e = cur_ops.solve(*cur, mode, budget, 3, st);
    s->launches += 1;
    if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("k_solve launch: ") + cudaGetErrorString(e));
    CU(cudaMemcpyAsync(s->h_count, cur->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int unfinished = s->h_count[0];
    if (unfinished == 0) break;
    if (!s->ops.large && repack_pct > 0 &&
        static_cast<long>(unfinished) * 100 <= static_cast<long>(cur->B) * repack_pct) {
      const int nxt = (cur_sec == 0) ? 1 : 0;
      if ((rc = ensure_secondary(s, nxt))) return rc;
      altro_b200_solver::Secondary& w = s->sec[nxt];
      k_list_unfinished<<<(cur->B + 255) / 256, 256, 0, st>>>(*cur, s->d_list, unfinished);
      if ((rc = check_launch(s, 1))) return rc;
      SolverParams& Q = w.P;
      Q.B = unfinished;
      Q.W = choose_tile_width(unfinished, s->sm_count);
      {
        const int w2 = env_int("ALTRO_B200_TILE2", 0);  // experiment knob: tile width of re-packed workspaces
        if ((w2 == 2 || w2 == 4 || w2 == 8 || w2 == 16) && (unfinished + w2 - 1) / w2 <= s->sm_count * 16) Q.W = w2;
      }
      if ((rc = ensure_secondary_buffers(s, nxt, 1 + kWarp / Q.W, unfinished))) return rc;
      Q.T = (unfinished + Q.W - 1) / Q.W;
      Q.Bp = Q.T * Q.W;
      Q.N = s->N; Q.n = s->n; Q.m = s->m; Q.pmax = s->pmax; Q.use_al = s->use_al;
      Q.blob = s->P.blob; Q.blob_bytes = s->P.blob_bytes;
      lookup_ops(s->n, s->m, s->model, Q.W, &w.ops);
      dim3 grid((unfinished + 127) / 128, s->N + 2);
      k_move_instances<<<grid, 128, 0, st>>>(*cur, Q, s->d_list, unfinished, 0);
      if ((rc = check_launch(s, 1))) return rc;
      if (cur_sec >= 0 && (rc = scatter_back(*cur))) return rc;  // leave nothing behind in a secondary
      cur = &Q;
      cur_ops = w.ops;
      cur_sec = nxt;
    }
  }
  if (cur_sec >= 0 && (rc = scatter_back(*cur))) return rc;
  return 0;
}
// The same solve on the phased engine: every slot is [outer-step kernels on a dense list
// (outer.cuh)] -> expansions -> TMA-streamed backward pass -> wide line search -> deep line
// search, all queued on ONE stream without host round trips; every `poll` slots the host reads
// the unfinished count and, when enough instances are done, re-packs the rest densely.
static int solve_phased_impl(altro_b200_solver* s, int mode, cudaStream_t st) {
  DeviceGuard guard(s->device);
  // Outer steps cost ~0.1 ms whatever the number of instances taking one (two serial chains over
  // the horizon): they run every outer_period slots; an instance that ended an iLQR solve waits
  // for the next one (results do not depend on when it is served).
  const int outer_period = std::max(1, env_int("ALTRO_B200_OUTER_PERIOD", 2));
  const int poll = outer_period * std::max(1, env_int("ALTRO_B200_POLL", 4) / outer_period);
  const int repack_pct = env_int("ALTRO_B200_REPACK_PCT", 70);  // 0 disables re-packing
  int rc;
  if ((rc = s->ensure_phased())) return rc;
  if ((rc = fill_int(s, I_PHASE, mode == 1 ? kPhAlInit : kPhSolveStart, st))) return rc;
  if ((rc = fill_int(s, I_LSFAIL, 0, st))) return rc;
  k_iota<<<(s->Bp + 255) / 256, 256, 0, st>>>(s->P.is + static_cast<size_t>(I_ORIG) * s->Bp, s->Bp);
  if ((rc = check_launch(s, 1))) return rc;
  SolverParams* cur = &s->P;
  Ops cur_ops = s->ops;
  int cur_sec = -1;
  auto scatter_back = [&](SolverParams& from) -> int {
    dim3 grid((from.B + 127) / 128, s->N + 2);
    k_move_instances<<<grid, 128, 0, st>>>(from, s->P, nullptr, from.B, 1);
    return check_launch(s, 1);
  };
#define PH(call, what)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    s->launches += 1;                                                                      \
    if (e_ != cudaSuccess)                                                                 \
      return fail(ALTRO_B200_ERR_CUDA, std::string(what " launch: ") + cudaGetErrorString(e_)); \
  } while (0)
  // the split line search evaluates the tries in two rounds (32/W, then 32): longer searches use
  // the fused line-search kernels, which loop
  // ... and only pays when few instances are in flight (the slot is then bound by the length of
  // the serial chains, not by instruction throughput)
  const int kEarlySlots = env_int("ALTRO_B200_EARLY_SLOTS", 8);
  const int split_max = std::min(env_split_max(), s->P.split_cap);
  const bool split_ok = s->P.opt.line_search_max_iterations <= kWarp / kPhasedTile + kWarp;
  for (long slot = 0; slot < 10000000; ++slot) {
    cur->opt = s->P.opt;
    CU(cudaMemsetAsync(cur->counters, 0, 8 * sizeof(int), st));
    // The host learns how many instances are still unreported from a counter the outer step leaves behind.
    // Early in a solve it asks after every outer step and waits for the answer before queueing the inner
    // kernels (short solves — two or three iterations — end there, and a no-op slot is ~10 launches); later
    // every `poll` slots, with the answer read at the end of the slot so the queue never drains.
    const bool early = slot < kEarlySlots && slot > 0;
    const bool polling = (slot % poll) == 0 || (early && slot % outer_period == 0);
    if (slot % outer_period == 0) {
      int nl = 0;
      cudaError_t e = cur_ops.outer_step(*cur, mode, s->sm_count, st, &nl);
      s->launches += nl;
      if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("outer-step launch: ") + cudaGetErrorString(e));
      if (polling) CU(cudaMemcpyAsync(s->h_count, cur->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
      if (early) {
        CU(cudaStreamSynchronize(st));
        if (s->h_count[0] == 0) break;
      }
    }
    PH(cur_ops.expansions_phased(*cur, st), "k_update_expansions");
    PH(cur_ops.backward_phased(*cur, st), "k_backward_mat");
    if (split_ok && cur->B <= split_max) {
      PH(cur_ops.ls_split_wide(*cur, mode, st), "k_roll/cost/acc_wide");
      PH(cur_ops.ls_split_deep(*cur, mode, cur->B, st), "k_roll/cost/acc_deep");
      s->launches += 4;
    } else {
      PH(cur_ops.ls_wide(*cur, mode, st), "k_ls_wide");
      PH(cur_ops.ls_deep(*cur, mode, cur->B, st), "k_ls_deep");
    }
    if (!polling) continue;
    CU(cudaStreamSynchronize(st));
    const int unfinished = s->h_count[0];
    if (unfinished == 0) break;
    if (repack_pct > 0 && static_cast<long>(unfinished) * 100 <= static_cast<long>(cur->B) * repack_pct) {
      const int nxt = (cur_sec == 0) ? 1 : 0;
      if ((rc = ensure_secondary(s, nxt))) return rc;
      altro_b200_solver::Secondary& w = s->sec[nxt];
      CU(cudaMemsetAsync(cur->counters, 0, 8 * sizeof(int), st));
      k_list_unfinished<<<(cur->B + 255) / 256, 256, 0, st>>>(*cur, s->d_list, unfinished);
      if ((rc = check_launch(s, 1))) return rc;
      SolverParams& Q = w.P;
      Q.B = unfinished;
      Q.W = kPhasedTile;
      if ((rc = ensure_secondary_buffers(s, nxt, 1 + kWarp / Q.W, unfinished))) return rc;
      Q.T = (unfinished + Q.W - 1) / Q.W;
      Q.Bp = Q.T * Q.W;
      Q.N = s->N; Q.n = s->n; Q.m = s->m; Q.pmax = s->pmax; Q.use_al = s->use_al;
      Q.blob = s->P.blob; Q.blob_bytes = s->P.blob_bytes;
      Q.EXP = s->P.EXP; Q.CAND = s->P.CAND; Q.list = s->P.list; Q.olist = s->P.olist;  // per-slot scratch, shared
      Q.COSTK = s->P.COSTK; Q.TRYST = s->P.TRYST; Q.split_cap = s->P.split_cap;
      lookup_ops(s->n, s->m, s->model, Q.W, &w.ops);
      dim3 grid((unfinished + 127) / 128, s->N + 2);
      k_move_instances<<<grid, 128, 0, st>>>(*cur, Q, s->d_list, unfinished, 0);
      if ((rc = check_launch(s, 1))) return rc;
      if (cur_sec >= 0 && (rc = scatter_back(*cur))) return rc;
      cur = &Q;
      cur_ops = w.ops;
      cur_sec = nxt;
    }
  }
#undef PH
  if (cur_sec >= 0 && (rc = scatter_back(*cur))) return rc;
  return 0;
}

static int solve_dispatch(altro_b200_solver* s, int mode, cudaStream_t st) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  s->expansions_valid = false;  // the phased engine uses EXP as per-slot scratch
  if (s->engine == ALTRO_B200_ENGINE_PHASED && s->ops.ls_wide) return solve_phased_impl(s, mode, st);
  return solve_impl(s, mode, st);
}
int altro_b200_solve_al(altro_b200_solver* s, void* stream) { return solve_dispatch(s, 1, S(stream)); }
int altro_b200_solve_ilqr(altro_b200_solver* s, void* stream) { return solve_dispatch(s, 0, S(stream)); }
/* latency probe of the per-knot device functions (tools/gpu_microbench.py); cycles[16] */
int altro_b200_microbench(altro_b200_solver* s, long long* cycles, int reps) {
  if (!s || !cycles || !s->ops.microbench) return fail(ALTRO_B200_ERR_UNSUPPORTED, "microbench: unsupported solver");
  DeviceGuard guard(s->device);
  {
    int rc = s->ensure_phased();
    if (rc) return rc;
  }
  double* sink = nullptr;
  long long* out = nullptr;
  CU(cudaMalloc(&sink, 32 * sizeof(double)));
  CU(cudaMalloc(&out, 32 * sizeof(long long)));
  CU(cudaMemset(out, 0, 32 * sizeof(long long)));
  cudaError_t e = s->ops.microbench(s->P, sink, out, reps, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles, out, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(sink);
  cudaFree(out);
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("microbench: ") + cudaGetErrorString(e));
  return 0;
}
void altro_b200_set_default_engine(int engine) {
  g_default_engine = (engine == ALTRO_B200_ENGINE_FUSED || engine == ALTRO_B200_ENGINE_PHASED) ? engine : -1;
}
int altro_b200_solver_engine(const altro_b200_solver* s) { return s ? s->engine : -1; }

int altro_b200_solve_al_host(altro_b200_solver* s, const double* x0, const double* U0, const double* u_nominal,
                             double* X, double* U, double* cost, double* viol, int32_t* status,
                             int32_t* iters, void* stream) {
  int rc = altro_b200_solver_set_inputs_host(s, x0, U0, u_nominal, stream);
  if (rc) return rc;
  if ((rc = altro_b200_solve_al(s, stream))) return rc;
  if (X || U)
    if ((rc = altro_b200_get_trajectory_host(s, X, U, stream))) return rc;
  return altro_b200_get_results_host(s, cost, viol, status, iters, stream);
}

// ---------------------------------------------------------------- step-wise phases
static int phase_impl(altro_b200_solver* s, int phase, cudaStream_t st) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  if (s->ops.large) return fail(ALTRO_B200_ERR_UNSUPPORTED, "step-wise methods are not available on the large-state path");
  DeviceGuard guard(s->device);
  int rc = s->ensure_stepwise();
  if (rc) return rc;
  cudaError_t e = s->ops.phase(s->P, phase, st);
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("k_phase launch: ") + cudaGetErrorString(e));
  return 0;
}
int altro_b200_rollout(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseRollout, S(stream)); }
int altro_b200_cost(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseCost, S(stream)); }
int altro_b200_update_expansions(altro_b200_solver* s, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  if (s->ops.large) return fail(ALTRO_B200_ERR_UNSUPPORTED, "step-wise methods are not available on the large-state path");
  DeviceGuard guard(s->device);
  int rc = s->ensure_stepwise();
  if (rc) return rc;
  cudaError_t e = s->ops.expansions(s->P, S(stream));
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("k_update_expansions launch: ") + cudaGetErrorString(e));
  s->expansions_valid = true;
  s->insolve_ready = false;
  return 0;
}
int altro_b200_backward_pass(altro_b200_solver* s, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  if (s->ops.large) return fail(ALTRO_B200_ERR_UNSUPPORTED, "step-wise methods are not available on the large-state path");
  // the records of a whole solve's last slot are not the expansion of the current Z_ (and the phased
  // engine's scratch has no cost-to-go / costs_ arrays): require an explicit UpdateExpansions
  if (!s->expansions_valid) return fail(ALTRO_B200_ERR_STATE, "UpdateExpansions must run before BackwardPass");
  DeviceGuard guard(s->device);
  {
    int rc = s->ensure_stepwise();
    if (rc) return rc;
  }
  cudaError_t e = s->ops.backward_mat(s->P, /*store_ctg=*/true, S(stream));
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("backward_pass: ") + cudaGetErrorString(e));
  return 0;
}
int altro_b200_forward_pass(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseForward, S(stream)); }
int altro_b200_update_convergence_statistics(altro_b200_solver* s, void* stream) {
  return phase_impl(s, kPhaseStats, S(stream));
}
int altro_b200_update_duals(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseDuals, S(stream)); }
int altro_b200_al_init(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseAlInit, S(stream)); }
int altro_b200_update_penalties(altro_b200_solver* s, void* stream) {
  return phase_impl(s, kPhasePenalties, S(stream));
}

// Measurement entry points (not part of the reference surface): the bare kernels.
int altro_b200_backward_pass_stream_only(altro_b200_solver* s, void* stream) {
  if (!s || !s->expansions_valid) return fail(ALTRO_B200_ERR_STATE, "UpdateExpansions must run before BackwardPass");
  DeviceGuard guard(s->device);
  cudaError_t e = s->ops.backward_mat(s->P, /*store_ctg=*/false, S(stream));
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("backward_pass: ") + cudaGetErrorString(e));
  return 0;
}
// The backward-pass kernel exactly as the phased engine launches it inside a solve
// (k_backward_mat<..., kPhased = true>: per-instance phase mask, regularisation hand-off, list of
// stalled instances), on the records of the last UpdateExpansions with EVERY instance marked as
// being in an inner iteration.  Scrambles the solve state: set the inputs again afterwards.
int altro_b200_backward_pass_insolve(altro_b200_solver* s, void* stream) {
  if (!s || !s->expansions_valid) return fail(ALTRO_B200_ERR_STATE, "UpdateExpansions must run before BackwardPass");
  if (!s->ops.backward_phased) return fail(ALTRO_B200_ERR_UNSUPPORTED, "no phased engine for this solver");
  DeviceGuard guard(s->device);
  int rc;
  if ((rc = s->ensure_phased())) return rc;
  if (!s->insolve_ready) {  // once per UpdateExpansions: the timed launches are the kernel alone
    if ((rc = fill_int(s, I_PHASE, kPhInner, S(stream)))) return rc;
    if ((rc = fill_int(s, I_LSFAIL, 0, S(stream)))) return rc;
    s->insolve_ready = true;
  }
  cudaError_t e = s->ops.backward_phased(s->P, S(stream));
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("backward_pass: ") + cudaGetErrorString(e));
  return 0;
}
int altro_b200_backward_pass_fused(altro_b200_solver* s, void* stream) {
  return phase_impl(s, kPhaseBackwardFused, S(stream));
}
int altro_b200_solve_setup(altro_b200_solver* s, void* stream) { return phase_impl(s, kPhaseSolveSetup, S(stream)); }

// ---------------------------------------------------------------- outputs
int altro_b200_get_trajectory_host(altro_b200_solver* s, double* X, double* U, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  DeviceGuard guard(s->device);
  const int nz = s->n + s->m;
  int rc;
  if (X && (rc = unpack_to(s, s->P.Z[0], s->N + 1, nz, 0, s->n, 0, s->N + 1, X, true, true, S(stream)))) return rc;
  if (U && (rc = unpack_to(s, s->P.Z[0], s->N + 1, nz, s->n, s->m, 0, s->N, U, true, true, S(stream)))) return rc;
  return 0;
}
int altro_b200_get_trajectory_dev(altro_b200_solver* s, double* X, double* U, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  DeviceGuard guard(s->device);
  const int nz = s->n + s->m;
  int rc;
  if (X && (rc = unpack_to(s, s->P.Z[0], s->N + 1, nz, 0, s->n, 0, s->N + 1, X, false, true, S(stream)))) return rc;
  if (U && (rc = unpack_to(s, s->P.Z[0], s->N + 1, nz, s->n, s->m, 0, s->N, U, false, true, S(stream)))) return rc;
  return 0;
}
int altro_b200_get_gains_host(altro_b200_solver* s, double* K, double* d, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  DeviceGuard guard(s->device);
  const int nkd = s->m * s->n + s->m;
  int rc;
  if (K && (rc = unpack_to(s, s->P.KD, s->N, nkd, 0, s->m * s->n, 0, s->N, K, true, false, S(stream)))) return rc;
  if (d && (rc = unpack_to(s, s->P.KD, s->N, nkd, s->m * s->n, s->m, 0, s->N, d, true, false, S(stream)))) return rc;
  return 0;
}
int altro_b200_get_ctg_host(altro_b200_solver* s, int k, double* Pm, double* p, void* stream) {
  if (!s || k < 0 || k > s->N) return fail(ALTRO_B200_ERR_ARG, "get_ctg: bad argument");
  if (!s->P.CTG) return fail(ALTRO_B200_ERR_STATE, "BackwardPass has not run");
  DeviceGuard guard(s->device);
  const int F = s->n * s->n + s->n;
  int rc;
  if (Pm && (rc = unpack_to(s, s->P.CTG, s->N + 1, F, 0, s->n * s->n, k, 1, Pm, true, false, S(stream)))) return rc;
  if (p && (rc = unpack_to(s, s->P.CTG, s->N + 1, F, s->n * s->n, s->n, k, 1, p, true, false, S(stream)))) return rc;
  return 0;
}
int altro_b200_get_costs_host(altro_b200_solver* s, double* costs, void* stream) {
  if (!s || !costs) return fail(ALTRO_B200_ERR_ARG, "get_costs: bad argument");
  if (!s->P.COSTS) return fail(ALTRO_B200_ERR_STATE, "Cost / UpdateExpansions has not run");
  DeviceGuard guard(s->device);
  return unpack_to(s, s->P.COSTS, s->N + 1, 1, 0, 1, 0, s->N + 1, costs, true, false, S(stream));
}
int altro_b200_get_expansion_host(altro_b200_solver* s, int k, double* A, double* Bm, double* lxx, double* lxu,
                                  double* luu, double* lx, double* lu, void* stream) {
  if (!s || k < 0 || k > s->N) return fail(ALTRO_B200_ERR_ARG, "get_expansion: bad argument");
  if (!s->P.EXP) return fail(ALTRO_B200_ERR_STATE, "UpdateExpansions has not run");
  DeviceGuard guard(s->device);
  const int n = s->n, m = s->m, F = exp_fields(n, m);
  double* outs[7] = {A, Bm, lxx, lxu, luu, lx, lu};
  const int sizes[7] = {n * n, n * m, n * n, n * m, m * m, n, m};
  int f0 = 0;
  for (int i = 0; i < 7; ++i) {
    if (outs[i]) {
      int rc = unpack_to(s, s->P.EXP, s->N + 1, F, f0, sizes[i], k, 1, outs[i], true, false, S(stream));
      if (rc) return rc;
    }
    f0 += sizes[i];
  }
  return 0;
}
int altro_b200_get_duals_host(altro_b200_solver* s, int k, double* lambda, int* p_out, void* stream) {
  if (!s || k < 0 || k > s->N) return fail(ALTRO_B200_ERR_ARG, "get_duals: bad argument");
  DeviceGuard guard(s->device);
  if (p_out) *p_out = s->pmax;
  if (lambda && s->pmax > 0)
    return unpack_to(s, s->P.LAM, s->N + 1, s->pmax, 0, s->pmax, k, 1, lambda, true, false, S(stream));
  return 0;
}
int altro_b200_get_constraint_values_host(altro_b200_solver* s, int k, double* c, int* p_out, void* stream) {
  if (!s || k < 0 || k > s->N) return fail(ALTRO_B200_ERR_ARG, "get_constraint_values: bad argument");
  if (!s->inputs_set) return fail(ALTRO_B200_ERR_STATE, "Initial state must be set before solving.");
  if (p_out) *p_out = s->p_knot[k];
  if (!c || s->pmax == 0) return 0;
  if (!s->ops.con_values)
    return fail(ALTRO_B200_ERR_UNSUPPORTED, "constraint values are not available on the large-state path");
  DeviceGuard guard(s->device);
  const size_t bytes = static_cast<size_t>(s->B) * s->pmax * sizeof(double);
  int rc = s->ensure_io(bytes);
  if (rc) return rc;
  cudaError_t e = s->ops.con_values(s->P, k, s->d_io, S(stream));
  s->launches += 1;
  if (e != cudaSuccess) return fail(ALTRO_B200_ERR_CUDA, std::string("k_constraint_values launch: ") + cudaGetErrorString(e));
  CU(cudaMemcpyAsync(c, s->d_io, bytes, cudaMemcpyDeviceToHost, S(stream)));
  CU(cudaStreamSynchronize(S(stream)));
  return 0;
}
int altro_b200_get_results_host(altro_b200_solver* s, double* cost, double* viol, int32_t* status,
                                int32_t* iters, void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  DeviceGuard guard(s->device);
  const size_t Bp = s->Bp, B = s->B;
  cudaStream_t st = S(stream);
  if (cost) CU(cudaMemcpyAsync(cost, s->P.sc + S_COST * Bp, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (viol) CU(cudaMemcpyAsync(viol, s->P.sc + S_VIOL * Bp, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status)
    CU(cudaMemcpyAsync(status, s->P.is + (s->use_al ? I_STATUS_AL : I_STATUS) * Bp, B * sizeof(int),
                       cudaMemcpyDeviceToHost, st));
  std::vector<int> tmp;
  if (iters) {
    tmp.resize(3 * B);
    for (int f = 0; f < 3; ++f)
      CU(cudaMemcpyAsync(tmp.data() + f * B, s->P.is + (I_ITERS_INNER + f) * Bp, B * sizeof(int),
                         cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  if (iters)
    for (size_t b = 0; b < B; ++b)
      for (int f = 0; f < 3; ++f) iters[3 * b + f] = tmp[f * B + b];
  return 0;
}
int altro_b200_get_ilqr_status_host(altro_b200_solver* s, int32_t* status, void* stream) {
  if (!s || !status) return fail(ALTRO_B200_ERR_ARG, "null argument");
  DeviceGuard guard(s->device);
  CU(cudaMemcpyAsync(status, s->P.is + static_cast<size_t>(I_STATUS) * s->Bp, s->B * sizeof(int),
                     cudaMemcpyDeviceToHost, S(stream)));
  CU(cudaStreamSynchronize(S(stream)));
  return 0;
}
int altro_b200_get_scalars_host(altro_b200_solver* s, double* reg, double* dV0, double* dV1, double* alpha,
                                double* z, double* dJ, double* grad, double* penalty, double* initial_cost,
                                void* stream) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  DeviceGuard guard(s->device);
  double* outs[9] = {reg, dV0, dV1, alpha, z, dJ, grad, penalty, initial_cost};
  const int fields[9] = {S_REG, S_DV0, S_DV1, S_ALPHA, S_ZRATIO, S_DJ, S_GRAD, S_PENALTY, S_INITIAL_COST};
  for (int i = 0; i < 9; ++i)
    if (outs[i])
      CU(cudaMemcpyAsync(outs[i], s->P.sc + static_cast<size_t>(fields[i]) * s->Bp, s->B * sizeof(double),
                         cudaMemcpyDeviceToHost, S(stream)));
  CU(cudaStreamSynchronize(S(stream)));
  return 0;
}

// ---------------------------------------------------------------- batch sharded over several GPUs
// Instances are independent: contiguous slices of the batch go to one solver per
// device; a host thread per device drives its slice (H2D of its inputs, the solve, D2H of its
// results) on its own stream, so the devices work concurrently.  No collective anywhere.
}  // extern "C"

#include <thread>

struct altro_b200_multi {
  int n = 0, m = 0, N = 0, B = 0;
  std::vector<int> devices, lo, hi;
  std::vector<altro_b200_solver*> parts;
  std::vector<cudaStream_t> streams;
  std::vector<double> t_scatter, t_solve, t_gather;  // ms of the last solve, per device
};

extern "C" {

int altro_b200_multi_create(const altro_b200_problem* p, int batch, int use_constraints, const int* devices,
                            int ndev, altro_b200_multi** out) {
  if (!p || !out || !devices || ndev <= 0 || batch < ndev)
    return fail(ALTRO_B200_ERR_ARG, "multi_create: bad argument (need at least one instance per device)");
  std::unique_ptr<altro_b200_multi, void (*)(altro_b200_multi*)> mm(new altro_b200_multi(), altro_b200_multi_destroy);
  mm->n = p->n; mm->m = p->m; mm->N = p->N; mm->B = batch;
  const int base = batch / ndev, rem = batch % ndev;
  int lo = 0;
  for (int g = 0; g < ndev; ++g) {  // contiguous slices, sizes differing by at most one
    const int hi = lo + base + (g < rem ? 1 : 0);
    mm->devices.push_back(devices[g]);
    mm->lo.push_back(lo);
    mm->hi.push_back(hi);
    lo = hi;
  }
  mm->parts.assign(ndev, nullptr);
  mm->streams.assign(ndev, nullptr);
  mm->t_scatter.assign(ndev, 0.0);
  mm->t_solve.assign(ndev, 0.0);
  mm->t_gather.assign(ndev, 0.0);
  for (int g = 0; g < ndev; ++g) {
    int rc = altro_b200_solver_create(p, mm->hi[g] - mm->lo[g], use_constraints, devices[g], &mm->parts[g]);
    if (rc) return rc;
    DeviceGuard guard(devices[g]);
    CU(cudaStreamCreateWithFlags(&mm->streams[g], cudaStreamNonBlocking));
  }
  *out = mm.release();
  return 0;
}

void altro_b200_multi_destroy(altro_b200_multi* mm) {
  if (!mm) return;
  for (size_t g = 0; g < mm->parts.size(); ++g) {
    if (mm->streams[g]) {
      DeviceGuard guard(mm->devices[g]);
      cudaStreamDestroy(mm->streams[g]);
    }
    altro_b200_solver_destroy(mm->parts[g]);
  }
  delete mm;
}

int altro_b200_multi_num_devices(const altro_b200_multi* mm) { return mm ? static_cast<int>(mm->parts.size()) : 0; }

int altro_b200_multi_set_options(altro_b200_multi* mm, const altro_b200_options* o) {
  if (!mm || !o) return fail(ALTRO_B200_ERR_ARG, "multi_set_options: null argument");
  for (altro_b200_solver* s : mm->parts) {
    int rc = altro_b200_solver_set_options(s, o);
    if (rc) return rc;
  }
  return 0;
}

int altro_b200_multi_solve_al_host(altro_b200_multi* mm, const double* x0, const double* U0, const double* u_nominal,
                                   double* X, double* U, double* cost, double* viol, int32_t* status, int32_t* iters) {
  if (!mm || !x0) return fail(ALTRO_B200_ERR_ARG, "multi_solve_al_host: null argument");
  const int G = static_cast<int>(mm->parts.size());
  const size_t n = mm->n, m = mm->m, N = mm->N;
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  auto work = [&](int g) {
    cudaSetDevice(mm->devices[g]);
    altro_b200_solver* s = mm->parts[g];
    cudaStream_t st = mm->streams[g];
    const size_t lo = mm->lo[g];
    cudaEvent_t e[4];
    for (cudaEvent_t& ev : e) cudaEventCreate(&ev);
    int rc = 0;
    cudaEventRecord(e[0], st);
    rc = altro_b200_solver_set_inputs_host(s, x0 + lo * n, U0 ? U0 + lo * N * m : nullptr, u_nominal, st);
    cudaEventRecord(e[1], st);
    if (!rc) rc = altro_b200_solve_al(s, st);
    cudaEventRecord(e[2], st);
    if (!rc && (X || U)) rc = altro_b200_get_trajectory_host(s, X ? X + lo * (N + 1) * n : nullptr, U ? U + lo * N * m : nullptr, st);
    if (!rc) rc = altro_b200_get_results_host(s, cost ? cost + lo : nullptr, viol ? viol + lo : nullptr,
                                              status ? status + lo : nullptr, iters ? iters + lo * 3 : nullptr, st);
    cudaEventRecord(e[3], st);
    cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e[0], e[1]); mm->t_scatter[g] = ms;
    cudaEventElapsedTime(&ms, e[1], e[2]); mm->t_solve[g] = ms;
    cudaEventElapsedTime(&ms, e[2], e[3]); mm->t_gather[g] = ms;
    for (cudaEvent_t& ev : e) cudaEventDestroy(ev);
    rcs[g] = rc;
    if (rc) errs[g] = altro_b200_last_error();  // thread-local: carried back to the caller's thread
  };
  std::vector<std::thread> threads;
  for (int g = 1; g < G; ++g) threads.emplace_back(work, g);
  int prev = -1;
  cudaGetDevice(&prev);
  work(0);
  if (prev >= 0) cudaSetDevice(prev);
  for (std::thread& t : threads) t.join();
  for (int g = 0; g < G; ++g)
    if (rcs[g]) return fail(rcs[g], "device " + std::to_string(mm->devices[g]) + ": " + errs[g]);
  return 0;
}

int altro_b200_multi_last_timings(const altro_b200_multi* mm, double* scatter_ms, double* solve_ms, double* gather_ms) {
  if (!mm) return fail(ALTRO_B200_ERR_ARG, "null handle");
  for (size_t g = 0; g < mm->parts.size(); ++g) {
    if (scatter_ms) scatter_ms[g] = mm->t_scatter[g];
    if (solve_ms) solve_ms[g] = mm->t_solve[g];
    if (gather_ms) gather_ms[g] = mm->t_gather[g];
  }
  return 0;
}

size_t altro_b200_backward_pass_bytes(const altro_b200_solver* s) {
  if (!s) return 0;
  const size_t n = s->n, m = s->m, N = s->N;
  const size_t per = 8 * (N * (n * (n + m) + n * n + n * m + m * m + n + m + m * n + m) + n * n + n);
  return per * s->B;
}
int64_t altro_b200_kernel_launches(const altro_b200_solver* s) { return s ? s->launches : 0; }

int altro_b200_solver_enable_history(altro_b200_solver* s, int instances, int rows) {
  if (!s) return fail(ALTRO_B200_ERR_ARG, "null solver");
  if (s->ops.large) return fail(ALTRO_B200_ERR_UNSUPPORTED, "history is not recorded on the large-state path");
  if (instances <= 0 || rows <= 0) return fail(ALTRO_B200_ERR_ARG, "enable_history: instances and rows must be positive");
  if (s->P.HIST) return fail(ALTRO_B200_ERR_STATE, "enable_history: already enabled");
  DeviceGuard guard(s->device);
  instances = std::min(instances, s->B);
  void* p = nullptr;
  int rc = s->alloc_zeroed(&p, static_cast<size_t>(instances) * rows * kHistCols * sizeof(double));
  if (rc) return rc;
  s->P.HIST = static_cast<double*>(p);
  s->P.hist_instances = s->hist_instances = instances;
  s->P.hist_rows = s->hist_rows = rows;
  s->engine = ALTRO_B200_ENGINE_FUSED;  // the rows are written by k_solve
  return 0;
}

int altro_b200_get_history_host(altro_b200_solver* s, int instance, double* rows_out, int max_rows, int* nrows,
                                void* stream) {
  if (!s || !rows_out || !nrows) return fail(ALTRO_B200_ERR_ARG, "get_history: null argument");
  if (!s->P.HIST) return fail(ALTRO_B200_ERR_STATE, "get_history: altro_b200_solver_enable_history was not called");
  if (instance < 0 || instance >= s->hist_instances) return fail(ALTRO_B200_ERR_ARG, "get_history: instance not recorded");
  DeviceGuard guard(s->device);
  int it_total = 0;
  CU(cudaMemcpyAsync(&it_total, s->P.is + static_cast<size_t>(I_ITERS_TOTAL) * s->Bp + instance, sizeof(int),
                     cudaMemcpyDeviceToHost, S(stream)));
  CU(cudaStreamSynchronize(S(stream)));
  const int n = std::max(0, std::min(std::min(it_total, s->hist_rows), max_rows));
  if (n > 0)
    CU(cudaMemcpyAsync(rows_out, s->P.HIST + static_cast<size_t>(instance) * s->hist_rows * kHistCols,
                       static_cast<size_t>(n) * kHistCols * sizeof(double), cudaMemcpyDeviceToHost, S(stream)));
  CU(cudaStreamSynchronize(S(stream)));
  *nrows = n;
  return 0;
}
size_t altro_b200_device_bytes(const altro_b200_solver* s) { return s ? s->dev_bytes : 0; }

}  // extern "C"
