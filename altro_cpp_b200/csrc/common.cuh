// common.cuh — problem descriptor, batch layout and per-instance state shared by the
// kernels (kernels.cuh) and the C ABI (altro_b200.cu).
//
// Batch layout ("tile-major"): the batch is cut into tiles of W instances (W = 2 .. 32,
// chosen per solver from the batch size).  Every per-knot quantity is stored as
//
//     arr[tile][knot][field][i]          (i = instance within the tile, fastest; W*8 B per row)
//
// so an access to one field of one knot of a tile is one contiguous, sector-aligned row and
// the whole record of a (tile, knot) is one contiguous block (F*W*8 B) that a single
// cp.async.bulk (TMA 1-D bulk copy) can stage into shared memory.
//
// A warp always owns ONE tile.  Its 32 lanes are G = 32/W groups of W lanes: lane = a*W + i.
// Lane group a = 0 carries the instances; groups a > 0 are the extra hands used where one
// instance has parallel work of its own — the line search evaluates G step lengths of every
// instance at once.  Small batches (B200: 148 SMs want >= ~2000 warps) use W = 8 so that the
// latency-bound serial sweeps are spread over 4x more warps; huge batches use W = 32.
#pragma once

#ifdef __CUDACC_RTC__
// run-time compilation of a model module (NVRTC has no host headers): the few fixed-width names used below
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace altro_b200 {

constexpr int kWarp = 32;
constexpr int kMaxZ = 17;       // trajectory buffers: Z_ plus up to G = 16 line-search candidates
constexpr int kMaxDim = 32;     // max rows of one constraint block (and max n)
constexpr int kMaxBlocks = 4;   // constraint blocks per knot (ALCost eq_ + ineq_ entries)
constexpr unsigned kFull = 0xffffffffu;

enum ModelKind : int { kUnicycle = 0, kTripleIntegrator = 1, kCartpole = 2, kLinear = 3 };
enum ConKind : int { kGoal = 0, kControlBound = 1, kCircle = 2 };

// altro/common/solver_stats.hpp:20-31
enum Status : int {
  kSolved = 0,
  kUnsolved = 1,
  kStateLimit = 2,
  kControlLimit = 3,
  kCostIncrease = 4,
  kMaxIterations = 5,
  kMaxOuterIterations = 6,
  kMaxInnerIterations = 7,
  kMaxPenalty = 8,
  kBackwardPassRegularizationFailed = 9,
};

// Shapes of a constraint block the device code has straight-line paths for (same arithmetic as the
// generic row loop, fewer instructions); detected when the blob is built.
enum BlockShape : int {
  kShapeGeneric = 0,
  kShapeBoundsFull = 1,        // control bound with finite lower and upper bounds on every control
  kShapeCirclesOneChunk = 2,   // circle block with at most 4 circles
};

// Shapes of a QuadraticCost the device code has shorter paths for (same sums, the exact-zero terms
// dropped); detected when the blob is built and stored in the last slot of the cost record.
enum CostShape : int {
  kCostDense = 0,
  kCostDiagonal = 1,  // Q and R diagonal, H = 0 (what QuadraticCost::LQRCost builds from diagonal weights)
};

// One ConstraintValues<n,m,ConType> (altro/constraints/constraint_values.hpp:24) of a knot.
struct ConBlock {
  int kind;      // ConKind
  int equality;  // 1: ZeroCone (dual cone = identity); 0: NegativeOrthant
  int p;         // OutputDimension()
  int row0;      // first dual row of this block inside the knot's dual column
  int nl, nu;    // control bound: number of lower / upper rows (basic_constraints.hpp:94-96)
  int xi, yi;    // circle: state indices of the position (obstacle_constraints.hpp:89-94)
  int idx[kMaxDim];   // control bound: control index of row i
  double a[kMaxDim];  // goal: xf | bound: bound value of row i | circle: cx
  double b[kMaxDim];  // circle: cy
  double c[kMaxDim];  // circle: r^2, rounded once on the host like the reference's pow(r, 2)
                      // (obstacle_constraints.hpp:37 there)
};

// The constraints of one knot in ALCost order: equalities then inequalities
// (altro/augmented_lagrangian/al_cost.hpp:264-273).
struct ConSet {
  int nblocks;
  int p_total;
  int _pad[2];
  // packed copy of the scalar fields of blk[i], one 16-byte load per block:
  // x = kind | equality << 8 | shape << 12 | p << 16, y = row0, z = nl, w = xi | yi << 8
  int4 hdr[kMaxBlocks];
  ConBlock blk[kMaxBlocks];
};

// Header of the problem blob.  All offsets are bytes from the start of the blob; the blob is
// copied verbatim into shared memory by every CTA.
struct BlobHeader {
  int n, m, N, model;
  int ncost, nconset, pmax, nparams;
  int cost_stride;     // doubles per QuadraticCost: n*n + m*m + n*m + n + m + 1, + 1 slot holding the CostShape
  int off_cost_id;     // int[N+1]: byte offset (from the blob base) of knot k's QuadraticCost
  int off_conset_id;   // int[N+1]: byte offset of knot k's ConSet
  int off_h;           // float[N+1]
  int off_t;           // float[N+1]
  int off_params;      // double[nparams]
  int off_cost;        // double[ncost*cost_stride]  (Q, R, H col-major, q, r, c)
  int off_conset;      // ConSet[nconset]
  int bytes;
  int _pad;
};

// Device view of a blob.  The table offsets are read once; every per-knot lookup is then a single
// shared-memory load (the tables hold byte offsets, not ids).
struct Desc {
  const char* base;
  const int* cost_off;
  const int* conset_off;
  const float* hs;
  const double* prm;
  __device__ __forceinline__ explicit Desc(const char* b) : base(b) {
    if (b) {
      const BlobHeader& h = *reinterpret_cast<const BlobHeader*>(b);
      cost_off = reinterpret_cast<const int*>(b + h.off_cost_id);
      conset_off = reinterpret_cast<const int*>(b + h.off_conset_id);
      hs = reinterpret_cast<const float*>(b + h.off_h);
      prm = reinterpret_cast<const double*>(b + h.off_params);
    } else {
      cost_off = conset_off = nullptr;
      hs = nullptr;
      prm = nullptr;
    }
  }
  __device__ __forceinline__ const BlobHeader& hdr() const {
    return *reinterpret_cast<const BlobHeader*>(base);
  }
  __device__ __forceinline__ const double* cost(int k) const {
    return reinterpret_cast<const double*>(base + cost_off[k]);
  }
  __device__ __forceinline__ const ConSet& conset(int k) const {
    return *reinterpret_cast<const ConSet*>(base + conset_off[k]);
  }
  __device__ __forceinline__ float h(int k) const { return hs[k]; }
  __device__ __forceinline__ float t(int k) const {
    return reinterpret_cast<const float*>(base + hdr().off_t)[k];
  }
  __device__ __forceinline__ const double* params() const { return prm; }
};

// altro/common/solver_options.hpp:19-65 (numeric fields)
struct DevOptions {
  int max_iterations_total, max_iterations_outer, max_iterations_inner;
  int bp_reg_fail_threshold, check_forwardpass_bounds, line_search_max_iterations, reset_duals;
  int skip_repeated_iterations;  // extension, default 0 (see finish_inner in kernels.cuh)
  double cost_tolerance, gradient_tolerance;
  double bp_reg_increase_factor, bp_reg_initial, bp_reg_max, bp_reg_min;
  double state_max, control_max;
  // largest t with sqrt(t) <= state_max / control_max: `sqrt(s) > max` is `s > t` (ilqr.hpp:484-495)
  double state_max_sq, control_max_sq;
  double line_search_lower_bound, line_search_upper_bound, line_search_decrease_factor;
  double constraint_tolerance, maximum_penalty, initial_penalty, penalty_scaling;
};

// Per-instance scalar state, SoA: sc[field][Bp], is[field][Bp].
enum ScalarField : int {
  S_REG = 0,       // rho_   (ilqr.hpp:802)
  S_DREG,          // drho_  (ilqr.hpp:803)
  S_DV0, S_DV1,    // deltaV_ (ilqr.hpp:804)
  S_PENALTY,       // penalty rho of every constraint (uniform, constraint_values.hpp:202-207)
  S_VIOL,          // GetMaxViolation() of the stored constraint values
  S_COST,          // last Cost()/accepted J of the current trajectory  (costs_.sum())
  S_INITIAL_COST,  // stats.initial_cost
  S_COST_CUR,      // stats.cost.back()          (carry-forward, solver_stats.cpp:54-66)
  S_COST_PREV,     // stats.cost.rbegin()[1]
  S_DJ, S_GRAD, S_ALPHA, S_ZRATIO,   // stats.{cost_decrease,gradient,alpha,improvement_ratio}.back()
  S_CSRC_ALPHA,    // Q8: < 0 -> stored constraint values come from Z; else alpha of the last
                   //        evaluated (rejected) candidate
  S_J0,            // costs_.sum() of the current Z_ (carried between launches of k_solve)
  S_VTMP,          // phased engine scratch: bit pattern of a running max violation (atomicMax on the
                   // unsigned view of non-negative doubles: exact and order-independent)
  S_REG_IN, S_DREG_IN,  // regularisation at the entry of the current inner iteration (phased engine hand-off)
  S_CAND_ALPHA,    // phased engine: step length of the rejected candidate that trajectory buffer zsel + 1 still
                   // holds (the last try of a failed search), or < 0; saves regenerating it for Q8
  S_NUM
};
enum IntField : int {
  I_STATUS = 0,    // iLQR status_
  I_STATUS_AL,     // AugmentedLagrangianiLQR status_
  I_ITERS_INNER, I_ITERS_OUTER, I_ITERS_TOTAL,
  I_ZSEL,          // which trajectory buffer currently is Z_ (the others hold line-search candidates)
  I_PHASE,         // SolvePhase: where the instance is in its solve (k_solve is resumable)
  I_ORIG,          // index of the instance in the solver's primary workspace (compaction)
  I_LSFAIL,        // 1: the instance's last line search failed completely (scheduling hint only)
  I_ROLLED,        // 1: the states of Z_ are the rollout of its controls from x0 (set when a solve has rolled it
                   // out, cleared when inputs or states are set): phased engine skips re-rollouts that change nothing
  I_NUM
};

// k_solve is a resumable state machine: every instance advances by whole inner iterations
// ("slots"); between launches the host may re-pack the unfinished instances densely.
enum SolvePhase : int {
  kPhAlInit = 0,      // AugmentedLagrangianiLQR::Init() pending
  kPhSolveStart = 1,  // iLQR::Solve(): SolveSetup + Rollout + initial Cost pending
  kPhInner = 2,       // next: UpdateExpansions/BackwardPass/ForwardPass of one inner iteration
  kPhOuter = 3,       // iLQR solve ended: UpdateDuals / IsDone / UpdatePenalties pending
  kPhDone = 4,        // terminated; final Cost() not reported yet
  kPhReported = 5,    // everything written
  kPhMoved = 6,       // continued in another workspace
};

struct SolverParams {
  int B, T, N;    // instances, tiles, segments
  int W, Bp;      // tile width, padded batch T*W
  int n, m, pmax, use_al;
  const char* blob;
  int blob_bytes;
  double* Z[kMaxZ];  // [T][N+1][n+m][W]    trajectory buffers (1 + 32/W of them are allocated)
  double* KD;     // [T][N][m*n+m][W]       K (col-major m x n) then d
  double* LAM;    // [T][N+1][pmax][W]      duals (nullptr when pmax == 0)
  double* X0;     // [T][n][W]              initial states
  double* EXP;    // [T][N+1][fexp][W]      materialised expansions (step-wise API, phased engine)
  double* CTG;    // [T][N+1][n*n+n][W]     cost-to-go P, p (step-wise API only)
  double* COSTS;  // [T][N+1][W]            costs_ vector (step-wise API only)
  double* sc;     // [S_NUM][Bp]
  int* is;        // [I_NUM][Bp]
  int* counters;  // [0] = instances not yet reported after the last k_solve launch / outer step; [1],[2]
                  // re-pack cursors; [3] = entries of `list`; [4] = entries of `olist` (phased engine); 8 ints
  int* list;      // [Bp] phased engine: instances whose line search continues in k_ls_deep
  int* olist;     // [Bp] phased engine: instances with outer-loop work pending (outer.cuh)
  double* CAND;   // [cap][N+1][n+m][32]  phased engine, split line search: candidate trajectories of
                  //                k_roll_deep (scratch), cap = split_cap list entries
  double* COSTK;  // [rows][N+1][32]  phased engine: per-knot costs of line-search candidates (scratch),
                  //                row = tile (wide kernels) or list entry (deep kernels), lane fastest;
                  //                also [entry][N+1] per-knot costs of the outer-step kernels
  int* TRYST;     // [rows][32]     phased engine: status left by each candidate rollout
  int split_cap;  // list entries CAND / COSTK / TRYST have room for in the split deep kernels
  double* HIST;   // [hist_instances][hist_rows][kHistCols] per-iteration SolverStats rows (nullptr: not recorded)
  int hist_instances, hist_rows;
  DevOptions opt;
};

// Columns of one history row = the double-valued SolverStats vectors (altro/common/solver_stats.hpp:54-61).
enum HistCol : int { kHistCost = 0, kHistAlpha, kHistZ, kHistGradient, kHistCostDecrease, kHistRegularization,
                     kHistViolations, kHistMaxPenalty, kHistCols };

__host__ __device__ inline int exp_fields(int n, int m) {
  return n * (n + m) + n * n + n * m + m * m + n + m;
}

}  // namespace altro_b200
