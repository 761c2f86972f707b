// large_mma.cuh — the large-state path on the fp64 tensor instruction (BASELINE config C5: random
// LQR, n = 32, m = 8).  One problem instance per WARP, eight instances in flight per CTA, one CTA per
// SM; no block-wide barrier after the prologue, so the eight warps drift apart and the tensor phase
// of one overlaps the factorisation / substitution / rollout latency of the others.
//
// Every dense product of the Riccati step (CalcActionValueExpansion, knot_point_function_type.hpp:149-164;
// CalcCostToGo, :180-195) is issued as mma.sync.m8n8k4.f64 on fragments read straight from shared
// memory:
//     T = A'P, BtP = B'P                      (160 DMMA)      P dead afterwards
//     Qux = H' + BtP A, Quu = R + BtP B       ( 40 DMMA)
//     LLT(Quu + reg I), K = -Quu^-1 Qux, d    (registers; lane c owns column c of K)
//     M = Qux + Quu K                         (  8 DMMA)
//     P+ = Q + T A + K'M + Qxu K              (192 DMMA)      accumulated in 64 registers per lane
// = 400 DMMA per knot point, against 7 800 dependent multiply-add pairs per thread-owned element in
// the exact-order kernel (large.cuh).  Measured issue rate on a B200: 0.25 DMMA / clk / SM
// (profiles/r02_dmma_rate.txt), the same 37 TFLOP/s as the DFMA pipe — what the tensor instruction buys
// is 16x fewer shared-memory operand reads per flop and 256x fewer issue slots, which is what bounds
// the exact-order kernel.
//
// Shared-memory matrices are padded to a leading dimension congruent to 4 mod 16 doubles (36 for the
// 32-wide ones, 12 for the 8-wide ones): an A-operand fragment (row = lane / 4, k = lane % 4) read
// row-major and a B-operand fragment (k = lane % 4, col = lane / 4) read column-major then touch 16
// distinct 8-byte banks per half-warp — conflict-free in both orientations, so A serves as A' (left
// operand of A'P) and as A (right operand of T A) from one copy.
//
// Results: the summation order differs from the reference's (tiles of 4 along k, fused multiply-add,
// one warp reduction per rollout for the cost), so this kernel agrees with the oracle to rounding
// (tests: <= 1e-9 after a whole solve), not bit for bit; the exact-order kernel stays selectable
// (engine "fused").
#pragma once

#include "device.cuh"

namespace altro_b200 {

constexpr int kMmaWarps = 8;
constexpr int kMmaThreads = kMmaWarps * 32;
constexpr int kLdW = 36;  // leading dimension of the 32-wide matrices
constexpr int kLdT = 12;  // leading dimension of the 8-wide matrices

// offsets in doubles (n = 32, m = 8)
struct MmaLayout {
  static constexpr int n = 32, m = 8;
  static constexpr int cost_len = n * n + m * m + n * m + n + m + 1;  // [Q | R | H | q | r | c]
  static constexpr int A = 0, B = A + n * kLdW, C = B + m * kLdW,      // shared by the CTA
                       warp0 = (C + cost_len + 1) & ~1;
  // per warp
  static constexpr int P = 0;                 // cost-to-go Hessian, column-major ld 36
  static constexpr int K = P;                 // (P is dead after T, BtP) gains, column-major ld 12
  static constexpr int M = P + n * kLdT;      // Qux + Quu K, column-major ld 12
  static constexpr int T = P + n * kLdW;      // A'P, row-major ld 36
  static constexpr int BtP = T + n * kLdW;    // B'P, row-major ld 36 (8 rows)
  static constexpr int Qux = BtP + m * kLdW;  // column-major ld 12 (= Qxu row-major)
  static constexpr int Quu = Qux + n * kLdT;  // row-major ld 12
  static constexpr int vec = Quu + m * kLdT;
  static constexpr int vx = vec, vu = vx + n, vp = vu + m, vQx = vp + n, vQu = vQx + n, vdx = vQu + m,
                       vend = vdx + n;
  static constexpr int per_warp = vend;
  static constexpr int total = warp0 + kMmaWarps * per_warp;
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// One whole AL-iLQR / iLQR solve per warp (mode as in k_solve).
template <int n, int m>
__global__ void __launch_bounds__(kMmaThreads, 1) k_solve_large_mma(SolverParams P, int mode) {
  static_assert(n == 32 && m == 8, "fragment schedule written for n = 32, m = 8");
  using Lay = MmaLayout;
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  constexpr int nz = n + m, nkd = m * n + m;
  constexpr int NT = n / 8, KS = n / 4;  // 8x8 tiles per side, k-steps of 4
  const Desc D(P.blob);
  const DevOptions& o = P.opt;
  const int N = P.N;
  double* As = sm + Lay::A;
  double* Bs = sm + Lay::B;
  {
    const double* mp = D.params();
    for (int e = threadIdx.x; e < n * n; e += kMmaThreads) As[e % n + (e / n) * kLdW] = mp[e];
    for (int e = threadIdx.x; e < n * m; e += kMmaThreads) Bs[e % n + (e / n) * kLdW] = mp[n * n + e];
  }
  // When every stage knot uses the same cost record (the usual LQR / tracking set-up) it is staged in
  // shared memory once; otherwise, and for the terminal knot, the record is read through L1/L2.
  __shared__ int s_cost_varies, s_cost_dense;
  if (threadIdx.x == 0) s_cost_varies = s_cost_dense = 0;
  __syncthreads();
  const int off0 = D.cost_off[0];
  for (int k = threadIdx.x; k < N; k += kMmaThreads)
    if (D.cost_off[k] != off0) s_cost_varies = 1;
  for (int e = threadIdx.x; e < Lay::cost_len; e += kMmaThreads) {
    const double v = D.cost(0)[e];
    sm[Lay::C + e] = v;
    // diagonal Q and R, no cross term (QuadraticCost::LQRCost, examples/quadratic_cost.cpp:31-45): the exact
    // zeros are skipped below — same bits as the dense evaluation, a fraction of the instructions
    const bool structural_zero = e < n * n ? (e % n != e / n)
                               : e < n * n + m * m ? ((e - n * n) % m != (e - n * n) / m)
                               : e < n * n + m * m + n * m;
    if (structural_zero && v != 0.0) s_cost_dense = 1;
  }
  __syncthreads();
  const bool staged = s_cost_varies == 0;
  const bool diag = staged && s_cost_dense == 0;
  auto cost_rec = [&](int k) -> const double* {
    return (staged && k < N) ? static_cast<const double*>(sm + Lay::C) : D.cost(k);
  };
  double* W = sm + Lay::warp0 + warp * Lay::per_warp;
  double *Pw = W + Lay::P, *Kw = W + Lay::K, *Mw = W + Lay::M, *Tw = W + Lay::T, *BtPw = W + Lay::BtP;
  double *Quxw = W + Lay::Qux, *Quuw = W + Lay::Quu;
  double *vx = W + Lay::vx, *vu = W + Lay::vu, *vp = W + Lay::vp, *vQx = W + Lay::vQx, *vQu = W + Lay::vQu;
  double* vdx = W + Lay::vdx;
  // fragment offsets: A operand row-major (row g, k q), B operand column-major (k q, column g)
  const int fa36 = g * kLdW + q, fa12 = g * kLdT + q;
  const int fb36 = q + g * kLdW, fb12 = q + g * kLdT;

  // instance b = blockIdx.x + gridDim.x * (warp + kMmaWarps * round): a ragged last round spreads
  // over all SMs with fewer warps each instead of leaving SMs empty
  for (int slot = warp;; slot += kMmaWarps) {
    const long long bl = blockIdx.x + static_cast<long long>(gridDim.x) * slot;
    if (bl >= P.B) break;
    const int b = static_cast<int>(bl);
    const double* X0 = P.X0 + static_cast<size_t>(b) * n;
    auto zptr = [&](int sel, int k) { return P.Z[sel] + (static_cast<size_t>(b) * (N + 1) + k) * nz; };
    auto kdptr = [&](int k) { return P.KD + (static_cast<size_t>(b) * N + k) * nkd; };

    // ---- per-lane share of one knot's cost (QuadraticCost::Evaluate); x, u already in vx, vu
    auto knot_cost_lane = [&](int k, double x, double u) -> double {
      const double* C = cost_rec(k);
      const double *Q = C, *R = C + n * n, *H = R + m * m, *qv = H + n * m, *r = qv + n;
      double c;
      if (diag && k < N) {
        c = x * (0.5 * (Q[lane + lane * n] * x) + qv[lane]);
        if (lane < m) c += u * (0.5 * (R[lane + lane * m] * u) + r[lane]);
      } else {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
        for (int j = 0; j < n; j += 2) {
          a0 = fma(Q[lane + j * n], vx[j], a0);
          a1 = fma(Q[lane + (j + 1) * n], vx[j + 1], a1);
        }
        double h2 = 0.0;
#pragma unroll
        for (int j = 0; j < m; ++j) h2 = fma(H[lane + j * n], vu[j], h2);
        c = x * (0.5 * (a0 + a1) + h2 + qv[lane]);
        if (lane < m) {
          double ru = 0.0;
#pragma unroll
          for (int j = 0; j < m; ++j) ru = fma(R[lane + j * m], vu[j], ru);
          c += u * (0.5 * ru + r[lane]);
        }
      }
      if (lane == 0) c += r[m];
      return c;
    };

    // ---- forward sweep (closed: RolloutClosedLoop(alpha) into zout, else Rollout() in place)
    auto forward = [&](bool closed, int zsel, int zout, double alpha, double& J_out, double& g_out,
                       int& status) -> bool {
      double x = X0[lane];
      double Jl = 0.0, gs = 0.0;
      bool ok = true;
      const double* Zc = zptr(zsel, 0);
      double* Zn = zptr(closed ? zout : zsel, 0);
      double arow[n], brow[m];
#pragma unroll
      for (int j = 0; j < n; ++j) arow[j] = As[lane + j * kLdW];
#pragma unroll
      for (int j = 0; j < m; ++j) brow[j] = Bs[lane + j * kLdW];
      const double* KDb = kdptr(0);
      // the next knot's reference point and gains are requested one knot ahead (HBM latency off the chain)
      double zx_n = Zc[lane], zu_n = lane < m ? Zc[n + lane] : 0.0, kq_n[8], dq_n = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) kq_n[t] = closed ? KDb[lane + 32 * t] : 0.0;
      if (closed && lane < m) dq_n = KDb[m * n + lane];
      for (int k = 0; k <= N; ++k) {
        double* zn = Zn + static_cast<size_t>(k) * nz;
        const double zx = zx_n, zu = zu_n, dq = dq_n;
        double kq[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) kq[t] = kq_n[t];
        if (k < N) {
          const double* zc1 = Zc + static_cast<size_t>(k + 1) * nz;
          zx_n = zc1[lane];
          if (lane < m) zu_n = zc1[n + lane];
          if (closed && k + 1 < N) {
            const double* pk1 = KDb + static_cast<size_t>(k + 1) * nkd;
#pragma unroll
            for (int t = 0; t < 8; ++t) kq_n[t] = pk1[lane + 32 * t];
            if (lane < m) dq_n = pk1[m * n + lane];
          }
        }
        double u = 0.0;
        if (k < N && closed) {
          vdx[lane] = x - zx;
          __syncwarp();
          double acc = 0.0, acc1 = 0.0;  // lane l: row l % 8 of K, columns l / 8 + 4 t
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            acc = fma(kq[t], vdx[(lane >> 3) + 4 * t], acc);
            acc1 = fma(kq[t + 1], vdx[(lane >> 3) + 4 * t + 4], acc1);
          }
          acc += acc1;
          acc += __shfl_xor_sync(0xffffffffu, acc, 8);
          acc += __shfl_xor_sync(0xffffffffu, acc, 16);
          double ratio = 0.0;
          if (lane < m) {
            u = zu + acc + dq * alpha;
            ratio = fabs(dq) / (fabs(u) + 1);
          }
#pragma unroll
          for (int s = 1; s < 8; s <<= 1) ratio = fmax(ratio, __shfl_xor_sync(0xffffffffu, ratio, s));
          gs += __shfl_sync(0xffffffffu, ratio, 0);
        } else if (lane < m) {
          u = (k < N || !closed) ? zu : 0.0;
        }
        vx[lane] = x;
        if (lane < m) vu[lane] = u;
        __syncwarp();
        zn[lane] = x;
        if (closed && lane < m) zn[n + lane] = u;
        Jl += knot_cost_lane(k, x, u);
        if (k < N) {  // x+ = A x + B u
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // row `lane` of A and B sits in registers for the whole sweep
#pragma unroll
          for (int j = 0; j < n; j += 4) {
            a0 = fma(arow[j], vx[j], a0);
            a1 = fma(arow[j + 1], vx[j + 1], a1);
            a2 = fma(arow[j + 2], vx[j + 2], a2);
            a3 = fma(arow[j + 3], vx[j + 3], a3);
          }
          double b0 = 0.0, b1 = 0.0;
#pragma unroll
          for (int j = 0; j < m; j += 2) {
            b0 = fma(brow[j], vu[j], b0);
            b1 = fma(brow[j + 1], vu[j + 1], b1);
          }
          x = ((a0 + a1) + (a2 + a3)) + (b0 + b1);
          if (closed && o.check_forwardpass_bounds) {  // ilqr.hpp:484-495 (sqrt(s) > max <=> s > max_sq)
            // |x|^2 > t needs some x_i^2 > t / n (n + 1 below: slack for the rounding of the sum): one vote rules
            // the norms out on almost every knot
            double sx = x * x, su = lane < m ? u * u : 0.0;
            if (__any_sync(0xffffffffu, sx * (n + 1) > o.state_max_sq || su * (m + 1) > o.control_max_sq)) {
#pragma unroll
              for (int s = 16; s > 0; s >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, s);
                su += __shfl_xor_sync(0xffffffffu, su, s);
              }
              if (sx > o.state_max_sq) {
                status = kStateLimit;
                ok = false;
              } else if (su > o.control_max_sq) {
                status = kControlLimit;
                ok = false;
              }
            }
          }
        }
        __syncwarp();
        if (!ok) break;
      }
      if (ok && closed) status = kUnsolved;
      J_out = warp_sum(Jl);
      g_out = gs;
      return ok;
    };

    // ---- backward sweep with the regularisation restart loop (ilqr.hpp:385-445)
    auto backward = [&](int zsel, double& reg, double& dreg, double& dV0, double& dV1, int& status, double& gsum) {
      int max_reg_count = 0;
      bool repeat = true;
      dV0 = 0.0;  // not reset by a restart (ilqr.hpp:390-392: set once, before the repeat loop)
      dV1 = 0.0;
      while (repeat) {
        {  // terminal cost-to-go: P = Qf, p = Qf x + q + H u
          const double* C = cost_rec(N);
          const double* zc = zptr(zsel, N);
          const double *Q = C, *H = C + n * n + m * m, *qv = H + n * m;
          vx[lane] = zc[lane];
          if (lane < m) vu[lane] = zc[n + lane];
          for (int j = 0; j < n; ++j) Pw[lane + j * kLdW] = Q[lane + j * n];
          __syncwarp();
          double a = 0.0, h2 = 0.0;
          for (int j = 0; j < n; ++j) a = fma(Q[lane + j * n], vx[j], a);
          for (int j = 0; j < m; ++j) h2 = fma(H[lane + j * n], vu[j], h2);
          vp[lane] = a + qv[lane] + h2;
          __syncwarp();
        }
        double gs = 0.0;
        bool failed = false;
        double zx_n, zu_n = 0.0;
        {
          const double* zc1 = zptr(zsel, N - 1);
          zx_n = zc1[lane];
          if (lane < m) zu_n = zc1[n + lane];
        }
        for (int k = N - 1; k >= 0; --k) {
          const double* C = cost_rec(k);
          const double *Q = C, *R = C + n * n, *H = R + m * m, *qv = H + n * m, *r = qv + n;
          vx[lane] = zx_n;
          if (lane < m) vu[lane] = zu_n;
          if (k > 0) {  // next knot's point, one knot ahead
            const double* zc1 = zptr(zsel, k - 1);
            zx_n = zc1[lane];
            if (lane < m) zu_n = zc1[n + lane];
          }
          __syncwarp();
          // lx (lane r) = Q x + q + H u;  lu (lanes 4 i, i < m) = R u + r + H'x with H'x summed by the four
          // lanes of a quad (inner index skewed by the quad: conflict-free).  A'p and B'p ride along with
          // T = A'P below as one more operand column.
          double lxr, lui;
          if (diag) {
            lxr = Q[lane + lane * n] * vx[lane] + qv[lane];
            lui = R[g + g * m] * vu[g] + r[g];
          } else {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
            for (int j = 0; j < n; j += 2) {
              a0 = fma(Q[lane + j * n], vx[j], a0);
              a1 = fma(Q[lane + (j + 1) * n], vx[j + 1], a1);
            }
            double h2 = 0.0;
#pragma unroll
            for (int j = 0; j < m; ++j) h2 = fma(H[lane + j * n], vu[j], h2);
            lxr = (a0 + a1) + qv[lane] + h2;
            double ru = 0.0, hx = 0.0;
#pragma unroll
            for (int j = 0; j < m; ++j) ru = fma(R[g + j * m], vu[j], ru);
#pragma unroll
            for (int t = 0; t < n / 4; ++t) {
              const int l = (4 * t + q + 4 * g) & (n - 1);
              hx = fma(H[l + g * n], vx[l], hx);
            }
            hx += __shfl_xor_sync(0xffffffffu, hx, 1);
            hx += __shfl_xor_sync(0xffffffffu, hx, 2);
            lui = ru + r[g] + hx;
          }
          {  // T = A'P (row-major), BtP = B'P (row-major)
            double t[NT][NT][2], bt[NT][2], ap[NT][2], bp[2];
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
              for (int j = 0; j < NT; ++j) t[i][j][0] = t[i][j][1] = 0.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) bt[j][0] = bt[j][1] = ap[j][0] = ap[j][1] = 0.0;
            bp[0] = bp[1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              double a[NT], bf[NT];
#pragma unroll
              for (int i = 0; i < NT; ++i) a[i] = As[fa36 + 8 * i * kLdW + 4 * ks];
              const double ab = Bs[fa36 + 4 * ks];
#pragma unroll
              for (int j = 0; j < NT; ++j) bf[j] = Pw[fb36 + 8 * j * kLdW + 4 * ks];
#pragma unroll
              for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(t[i][j], a[i], bf[j]);
#pragma unroll
              for (int j = 0; j < NT; ++j) dmma884(bt[j], ab, bf[j]);
              const double pf = g == 0 ? vp[4 * ks + q] : 0.0;  // [p 0 ... 0] as a B operand: column 0 of the product
#pragma unroll
              for (int i = 0; i < NT; ++i) dmma884(ap[i], a[i], pf);
              dmma884(bp, ab, pf);
            }
            if (q == 0) {  // column 0 of a C fragment lives in the q == 0 lanes
#pragma unroll
              for (int i = 0; i < NT; ++i) vdx[8 * i + g] = ap[i][0];
              vQu[g] = lui + bp[0];
            }
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
              for (int j = 0; j < NT; ++j)
                *reinterpret_cast<double2*>(Tw + (8 * i + g) * kLdW + 8 * j + 2 * q) = make_double2(t[i][j][0], t[i][j][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j)
              *reinterpret_cast<double2*>(BtPw + g * kLdW + 8 * j + 2 * q) = make_double2(bt[j][0], bt[j][1]);
          }
          __syncwarp();
          double qux[NT][2];  // C fragment of Qux: row g, columns 8 j + 2 q + {0, 1}
          {                   // Qux = H' + BtP A, Quu = R + BtP B
            double quu[2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              qux[j][0] = diag ? 0.0 : H[(8 * j + 2 * q) + g * n];
              qux[j][1] = diag ? 0.0 : H[(8 * j + 2 * q + 1) + g * n];
            }
            quu[0] = (!diag || g == 2 * q) ? R[g + (2 * q) * m] : 0.0;
            quu[1] = (!diag || g == 2 * q + 1) ? R[g + (2 * q + 1) * m] : 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const double ab = BtPw[fa36 + 4 * ks];
              const double bb = Bs[fb36 + 4 * ks];
#pragma unroll
              for (int j = 0; j < NT; ++j) dmma884(qux[j], ab, As[fb36 + 8 * j * kLdW + 4 * ks]);
              dmma884(quu, ab, bb);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              Quxw[g + (8 * j + 2 * q) * kLdT] = qux[j][0];
              Quxw[g + (8 * j + 2 * q + 1) * kLdT] = qux[j][1];
            }
            *reinterpret_cast<double2*>(Quuw + g * kLdT + 2 * q) = make_double2(quu[0], quu[1]);
          }
          __syncwarp();
          // RegularizeActionValue + LLT (lower; pivot <= 0 fails, Q19), every lane redundantly in registers
          double L[m][m], rinv[m];
          bool okc = true;
#pragma unroll
          for (int j = 0; j < m; ++j)
#pragma unroll
            for (int i = j; i < m; ++i) L[i][j] = Quuw[i * kLdT + j] + (i == j ? reg : 0.0);
#pragma unroll
          for (int kk = 0; kk < m; ++kk) {
            double piv = L[kk][kk];
            if (kk > 0) {
              double sq = 0.0;
#pragma unroll
              for (int j = 0; j < kk; ++j) sq = fma(L[kk][j], L[kk][j], sq);
              piv -= sq;
            }
            okc = okc && !(piv <= 0.0);
            rinv[kk] = rsqrt(piv);
            L[kk][kk] = piv * rinv[kk];
#pragma unroll
            for (int i = kk + 1; i < m; ++i) {
              double a = L[i][kk];
              if (kk > 0) {
                double dot = 0.0;
#pragma unroll
                for (int j = 0; j < kk; ++j) dot = fma(L[i][j], L[kk][j], dot);
                a -= dot;
              }
              L[i][kk] = a * rinv[kk];
            }
          }
          if (!okc) {  // ilqr.hpp:409-427 (uniform over the warp: every lane factorised the same matrix)
            increase_reg(o, reg, dreg);
            if (reg >= o.bp_reg_max) max_reg_count++;
            if (max_reg_count >= o.bp_reg_fail_threshold) {
              status = kBackwardPassRegularizationFailed;
              repeat = false;
            }
            failed = true;
            break;
          }
          // CalcGains: lane c solves column c of K; every lane solves d
          double kc[m], quc[m], dv[m];
          {
            auto llt_solve = [&](double(&bv)[m]) {
#pragma unroll
              for (int i = 0; i < m; ++i) {
                double s = bv[i];
#pragma unroll
                for (int j = 0; j < i; ++j) s = fma(-L[i][j], bv[j], s);
                bv[i] = s * rinv[i];
              }
#pragma unroll
              for (int i = m - 1; i >= 0; --i) {
                double s = bv[i];
#pragma unroll
                for (int j = i + 1; j < m; ++j) s = fma(-L[j][i], bv[j], s);
                bv[i] = s * rinv[i];
              }
            };
#pragma unroll
            for (int i = 0; i < m; i += 2) {
              const double2 v = *reinterpret_cast<const double2*>(Quxw + i + lane * kLdT);
              quc[i] = v.x;
              quc[i + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < m; ++i) {
              kc[i] = quc[i];
              dv[i] = vQu[i];
            }
            llt_solve(kc);
            llt_solve(dv);
#pragma unroll
            for (int i = 0; i < m; ++i) {
              kc[i] = -kc[i];
              dv[i] = -dv[i];
            }
          }
          double* pk = kdptr(k);
#pragma unroll
          for (int i = 0; i < m; i += 2) {
            *reinterpret_cast<double2*>(Kw + i + lane * kLdT) = make_double2(kc[i], kc[i + 1]);
            *reinterpret_cast<double2*>(pk + i + lane * m) = make_double2(kc[i], kc[i + 1]);
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < m; i += 2) *reinterpret_cast<double2*>(pk + m * n + i) = make_double2(dv[i], dv[i + 1]);
          }
          // qd = Quu d (unregularised, Q3), w = qd + Qu; dV, gradient statistic: lane l works on control l % 8
          double pn;
          {
            const int i8 = lane & 7;
            double di = dv[0];
#pragma unroll
            for (int j = 1; j < m; ++j) di = (i8 == j) ? dv[j] : di;
            double qd = 0.0;
#pragma unroll
            for (int l = 0; l < m; ++l) qd = fma(Quuw[i8 * kLdT + l], dv[l], qd);
            const double qu = vQu[i8];
            if (lane < m) vQx[lane] = qd + qu;  // w, for every lane below
            double dqu = di * qu, dqd = di * qd, gk = fabs(di) / (fabs(vu[i8]) + 1);
#pragma unroll
            for (int sft = 1; sft < 8; sft <<= 1) {
              dqu += __shfl_xor_sync(0xffffffffu, dqu, sft);
              dqd += __shfl_xor_sync(0xffffffffu, dqd, sft);
              gk = fmax(gk, __shfl_xor_sync(0xffffffffu, gk, sft));
            }
            dV0 += dqu;
            dV1 += 0.5 * dqd;
            gs += gk;
            __syncwarp();
            // p+ = Qx + K'(Quu d + Qu) + Qxu d   (lane c)
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int i = 0; i < m; ++i) {
              s1 = fma(kc[i], vQx[i], s1);
              s2 = fma(quc[i], dv[i], s2);
            }
            pn = (lxr + vdx[lane]) + s1 + s2;
          }
          __syncwarp();
          {  // M = Qux + Quu K (column-major ld 12)
#pragma unroll
            for (int ks = 0; ks < m / 4; ++ks) {
              const double a = Quuw[fa12 + 4 * ks];
#pragma unroll
              for (int j = 0; j < NT; ++j) dmma884(qux[j], a, Kw[fb12 + 8 * j * kLdT + 4 * ks]);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              Mw[g + (8 * j + 2 * q) * kLdT] = qux[j][0];
              Mw[g + (8 * j + 2 * q + 1) * kLdT] = qux[j][1];
            }
          }
          __syncwarp();
          {  // P+ = Q + T A + K'M + Qxu K
            double pp[NT][NT][2];
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                pp[i][j][0] = (!diag || (i == j && g == 2 * q)) ? Q[(8 * i + g) + (8 * j + 2 * q) * n] : 0.0;
                pp[i][j][1] = (!diag || (i == j && g == 2 * q + 1)) ? Q[(8 * i + g) + (8 * j + 2 * q + 1) * n] : 0.0;
              }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              double a[NT], bf[NT];
#pragma unroll
              for (int i = 0; i < NT; ++i) a[i] = Tw[fa36 + 8 * i * kLdW + 4 * ks];
#pragma unroll
              for (int j = 0; j < NT; ++j) bf[j] = As[fb36 + 8 * j * kLdW + 4 * ks];
#pragma unroll
              for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(pp[i][j], a[i], bf[j]);
            }
#pragma unroll
            for (int ks = 0; ks < m / 4; ++ks) {
              double a1[NT], a2[NT], b1[NT], b2[NT];
#pragma unroll
              for (int i = 0; i < NT; ++i) {
                a1[i] = Kw[fa12 + 8 * i * kLdT + 4 * ks];    // K'[r][k] = K[k + r * 12]
                a2[i] = Quxw[fa12 + 8 * i * kLdT + 4 * ks];  // Qxu[r][k] = Qux[k + r * 12]
              }
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                b1[j] = Mw[fb12 + 8 * j * kLdT + 4 * ks];
                b2[j] = Kw[fb12 + 8 * j * kLdT + 4 * ks];
              }
#pragma unroll
              for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                  dmma884(pp[i][j], a1[i], b1[j]);
                  dmma884(pp[i][j], a2[i], b2[j]);
                }
            }
            __syncwarp();  // K, M live in P's storage: every lane is done reading them
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                Pw[(8 * i + g) + (8 * j + 2 * q) * kLdW] = pp[i][j][0];
                Pw[(8 * i + g) + (8 * j + 2 * q + 1) * kLdW] = pp[i][j][1];
              }
          }
          vp[lane] = pn;
          __syncwarp();
          if (k == 0) repeat = false;
        }
        (void)failed;
        gsum = gs;
      }
      decrease_reg(o, reg, dreg);
    };

    // ---- the solve (same control flow as k_solve / k_solve_large; the warp runs in lock-step) ----
    auto SC = [&](int f) -> double& { return P.sc[static_cast<size_t>(f) * P.Bp + b]; };
    auto IS = [&](int f) -> int& { return P.is[static_cast<size_t>(f) * P.Bp + b]; };
    double penalty = SC(S_PENALTY);
    double reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0, cost_cur = SC(S_COST_CUR), cost_prev = SC(S_COST_PREV);
    double initial_cost = 0.0, dJ = 0.0, grad = 0.0, alpha_stat = 0.0, z_stat = 0.0, J0 = 0.0, viol = 0.0;
    int zsel = IS(I_ZSEL), it_inner = 0, it_outer = IS(I_ITERS_OUTER), it_total = IS(I_ITERS_TOTAL);
    int st = kUnsolved, st_al = kUnsolved;
    if (mode == 1) {  // Init()
      if (o.initial_penalty > 0) penalty = o.initial_penalty;
      it_outer = 0;
      it_total = 0;
      cost_cur = 0.0;
      cost_prev = 0.0;
    }
    {
      it_inner = 0;
      st = kUnsolved;
      reg = o.bp_reg_initial;
      dreg = 0.0;
      double gtmp;
      forward(false, zsel, zsel, 0.0, J0, gtmp, st);
      initial_cost = J0;
      bool run = o.max_iterations_inner > 0;
      while (run) {
        double gs_bwd = 0.0;
        backward(zsel, reg, dreg, dV0, dV1, st, gs_bwd);
        double alpha = 1.0, z = -1.0, gs_acc = 0.0;
        bool success = false;
        for (int t = 0; t < o.line_search_max_iterations; ++t) {  // ForwardPass, ilqr.hpp:512-558
          double J, gs;
          if (forward(true, zsel, zsel ^ 1, alpha, J, gs, st)) {
            const double expected = -alpha * (dV0 + alpha * dV1);
            z = (expected > 0.0) ? (J0 - J) / expected : -1.0;
            if (o.line_search_lower_bound <= z && z <= o.line_search_upper_bound && J < J0) {
              success = true;
              cost_cur = J;
              alpha_stat = alpha;
              z_stat = z;
              J0 = J;
              gs_acc = gs;
              break;
            }
          }
          alpha /= o.line_search_decrease_factor;
        }
        if (success) {
          zsel ^= 1;
          grad = gs_acc / static_cast<double>(N);
        } else {
          increase_reg(o, reg, dreg);
          grad = gs_bwd / static_cast<double>(N);
        }
        dJ = (it_inner == 0) ? (initial_cost - cost_cur) : (cost_prev - cost_cur);
        it_inner++;
        it_total++;
        cost_prev = cost_cur;
        if (dJ < o.cost_tolerance && grad < o.gradient_tolerance) {
          st = kSolved;
          run = false;
        } else if (it_inner >= o.max_iterations_inner) {
          st = kMaxInnerIterations;
          run = false;
        } else if (it_total >= o.max_iterations_total) {
          st = kMaxIterations;
          run = false;
        } else if (st != kUnsolved) {
          run = false;
        }
      }
      if (mode != 0) {  // no constraints: UpdateDuals is a no-op, violation and max penalty are 0
        it_outer++;
        viol = 0.0;
        st_al = (st != kSolved) ? st : kSolved;
      }
    }
    double Jf;
    {  // final Cost()
      double Jl = 0.0;
      for (int k = 0; k <= N; ++k) {
        const double* zc = zptr(zsel, k);
        const double x = zc[lane], u = lane < m ? zc[n + lane] : 0.0;
        vx[lane] = x;
        if (lane < m) vu[lane] = u;
        __syncwarp();
        Jl += knot_cost_lane(k, x, u);
        __syncwarp();
      }
      Jf = warp_sum(Jl);
    }
    if (lane == 0) {
      SC(S_COST) = Jf;
      SC(S_REG) = reg; SC(S_DREG) = dreg; SC(S_DV0) = dV0; SC(S_DV1) = dV1; SC(S_PENALTY) = penalty;
      SC(S_VIOL) = viol; SC(S_INITIAL_COST) = initial_cost; SC(S_COST_CUR) = cost_cur;
      SC(S_COST_PREV) = cost_prev; SC(S_DJ) = dJ; SC(S_GRAD) = grad; SC(S_ALPHA) = alpha_stat;
      SC(S_ZRATIO) = z_stat; SC(S_CSRC_ALPHA) = -1.0; SC(S_J0) = J0;
      IS(I_STATUS) = st; IS(I_STATUS_AL) = st_al; IS(I_ITERS_INNER) = it_inner;
      IS(I_ITERS_OUTER) = it_outer; IS(I_ITERS_TOTAL) = it_total; IS(I_ZSEL) = zsel;
      IS(I_PHASE) = kPhReported;
    }
    __syncwarp();
  }
}

}  // namespace altro_b200
