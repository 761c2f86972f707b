// backward_coop.cuh — BackwardPass over materialised expansions for medium state dimensions
// (e.g. the triple integrator of BASELINE config C3: n = 6, m = 2), TWO lanes per instance.
//
// k_backward_mat (kernels.cuh) gives an instance one lane with every matrix in registers.  For
// n = 6 that is ~250 live doubles: 255 registers plus spills, a 125 KB staging ring per warp, one warp
// per SM and two waves — the serial chain of the Riccati recursion
// (knot_point_function_type.hpp:149-230 there) then runs at a quarter of the HBM roofline.  Here the
// lanes of a pair split the COLUMNS of the n-column matrices (lane 0: state columns 0 .. n/2-1 and
// control columns 0 .. m/2-1, lane 1 the rest); what the other lane needs goes through a small
// per-instance scratch in shared memory.  A warp is 16 instances = two tiles of eight and owns one TMA
// ring: 47 KB of shared memory, four warps per SM, the whole C3 slice (512 warps) in one wave, ~100
// live doubles per lane.  (Eight lanes per instance were tried first: every lane then re-reads all of
// A and A'P from shared memory and the kernel is bound by shared-memory bandwidth — slower than one
// lane per instance.)
//
// Every scalar of the recursion is still produced by ONE lane as the same sequence of operations
// as in riccati_step (device.cuh) — only WHICH lane computes it changes — so K, d, P, p and deltaV
// are bit-identical to the one-lane kernel and to the fused engine (tests/test_gpu_parity.py
// test_engines_are_bit_identical, c3).
#pragma once

#include "kernels.cuh"

namespace altro_b200 {

constexpr int kCoopLanes = 2;   // lanes per instance
constexpr int kCoopTile = 8;    // tile width of the workspaces it runs on
constexpr int kCoopInst = kWarp / kCoopLanes;  // instances per warp (two tiles)

template <class M>
struct CoopScratch {  // per instance, in shared memory
  static constexpr int n = M::n, m = M::m;
  double AtP[n * n];   // A'P, column-major
  double BtP[m * n];   // B'P (m x n)
  double Qxu[n * m];
  double Quu[m * m];
  double Qu[m];
  double K[m * n];
  double p[n];         // cost-to-go gradient of knot k+1 (both lanes read all of it)
  double pad;          // odd number of doubles: the instances of a warp start in different banks
};

template <class M, int kStages, bool kStoreCtg, bool kPhased>
__global__ void __launch_bounds__(kWarp) k_backward_coop(SolverParams P) {
  constexpr int n = M::n, m = M::m, W = kCoopTile, nexp = Lane<M, W>::nexp, LPI = kCoopLanes;
  constexpr int CN = n / LPI, CM = m / LPI;  // state / control columns per lane
  static_assert(n % LPI == 0 && m % LPI == 0, "the lanes of a pair split the columns evenly");
  constexpr int TPW = kCoopInst / W;  // tiles per warp
  constexpr uint32_t kRecBytes = nexp * W * sizeof(double);
  constexpr int kSlotDoubles = nexp * W * TPW;
  constexpr int oA = 0, oB = n * n, oLxx = n * (n + m), oLxu = oLxx + n * n, oLuu = oLxu + n * m, oLx = oLuu + m * m,
                oLu = oLx + n;
  extern __shared__ __align__(128) char smem[];
  double* ring = reinterpret_cast<double*>(smem);
  CoopScratch<M>* scratch = reinterpret_cast<CoopScratch<M>*>(smem + static_cast<size_t>(kStages) * kSlotDoubles * sizeof(double));
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + kCoopInst);
  const int lane = threadIdx.x, iw = lane / LPI, j = lane % LPI;  // instance within the warp, column role
  const unsigned gmask = ((1u << LPI) - 1u) << (iw * LPI);        // the lanes of this instance
  const int tile0 = blockIdx.x * TPW;
  const int ntiles = (P.T - tile0 < TPW) ? (P.T - tile0) : TPW;
  const Lane<M, W> L(P, nullptr, tile0 + iw / W, iw % W);
  const DevOptions& o = P.opt;
  const int N = P.N;
  const bool valid = L.valid && (!kPhased || L.is(I_PHASE) == kPhInner);
  if (kPhased && !__any_sync(kFull, valid)) return;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const size_t tile_stride = static_cast<size_t>(N + 1) * nexp * W;
  const double* rec0 = P.EXP + static_cast<size_t>(tile0) * tile_stride;
  auto load_slot = [&](int slot, int k) {  // lane 0 only
    mbar_expect_tx(&bars[slot], kRecBytes * ntiles);
    for (int t = 0; t < ntiles; ++t)
      tma_load_1d(ring + static_cast<size_t>(slot) * kSlotDoubles + t * nexp * W,
                  rec0 + t * tile_stride + static_cast<size_t>(k) * nexp * W, kRecBytes, &bars[slot]);
  };
  CoopScratch<M>& S = scratch[iw];

  double reg = 0.0, dreg = 0.0, dV0 = 0.0, dV1 = 0.0;
  int st = kUnsolved;
  if (valid) {
    reg = L.sc(S_REG);
    dreg = L.sc(S_DREG);
    st = L.is(I_STATUS);
  }
  const double reg_in = reg, dreg_in = dreg;
  int max_reg_count = 0;
  bool repeat = valid;
  uint32_t issued = 0, consumed = 0;
  while (__any_sync(kFull, repeat)) {
    // terminal cost-to-go: this lane's columns of P = lxx(N); p = lx(N)
    double Pc[CN][n];
    {
      const double* e = L.exp(N);
      ALTRO_UNROLL
      for (int c = 0; c < CN; ++c) {
        const int col = j * CN + c;
        ALTRO_UNROLL
        for (int r = 0; r < n; ++r) Pc[c][r] = valid ? e[(oLxx + r + col * n) * W] : 0.0;
        S.p[col] = valid ? e[(oLx + col) * W] : 0.0;
        if (kStoreCtg && repeat) {
          double* ct = L.ctg(N);
          ALTRO_UNROLL
          for (int r = 0; r < n; ++r) ct[(r + col * n) * W] = Pc[c][r];
          ct[(n * n + col) * W] = S.p[col];
        }
      }
    }
    __syncwarp();
    int next_k = N - 1;
    for (int s = 0; s < kStages && next_k >= 0; ++s, --next_k, ++issued)
      if (lane == 0) load_slot(issued % kStages, next_k);
    bool live = repeat;
    for (int k = N - 1; k >= 0; --k, ++consumed) {
      const int slot = consumed % kStages;
      mbar_wait(&bars[slot], (consumed / kStages) & 1);
      // field f of this instance: e[f * W]
      const double* e = ring + static_cast<size_t>(slot) * kSlotDoubles + (iw / W) * nexp * W + iw % W;
      // ---- phase 1: A'P, B'P for this lane's columns of P; A'p, B'p rows (CalcActionValueExpansion :149-164)
      double pv[n];
      ALTRO_UNROLL
      for (int l = 0; l < n; ++l) pv[l] = S.p[l];
      double Qx[CN], Qu_own[CM > 0 ? CM : 1];
      ALTRO_UNROLL
      for (int c = 0; c < CN; ++c) {
        const int col = j * CN + c;
        ALTRO_UNROLL
        for (int r = 0; r < n; ++r) {  // (A'P)(r, col) = sum_l A(l, r) P(l, col)
          double acc = e[(oA + r * n) * W] * Pc[c][0];
          ALTRO_UNROLL
          for (int l = 1; l < n; ++l) acc += e[(oA + l + r * n) * W] * Pc[c][l];
          S.AtP[r + col * n] = acc;
        }
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) {  // (B'P)(r, col) = sum_l B(l, r) P(l, col)
          double acc = e[(oB + r * n) * W] * Pc[c][0];
          ALTRO_UNROLL
          for (int l = 1; l < n; ++l) acc += e[(oB + l + r * n) * W] * Pc[c][l];
          S.BtP[r + col * m] = acc;
        }
        double v = e[(oA + col * n) * W] * pv[0];  // (A'p)(col)
        ALTRO_UNROLL
        for (int l = 1; l < n; ++l) v += e[(oA + l + col * n) * W] * pv[l];
        Qx[c] = e[(oLx + col) * W] + v;
      }
      ALTRO_UNROLL
      for (int c = 0; c < CM; ++c) {
        const int col = j * CM + c;
        double w = e[(oB + col * n) * W] * pv[0];  // (B'p)(col)
        ALTRO_UNROLL
        for (int l = 1; l < n; ++l) w += e[(oB + l + col * n) * W] * pv[l];
        Qu_own[c] = e[(oLu + col) * W] + w;
        S.Qu[col] = Qu_own[c];
      }
      __syncwarp(gmask);
      // ---- phase 2: this lane's columns of Qxx, Qxu and Quu -------------------------------------------------
      double Qc[CN][n];
      ALTRO_UNROLL
      for (int c = 0; c < CN; ++c) {
        const int col = j * CN + c;
        ALTRO_UNROLL
        for (int r = 0; r < n; ++r) {  // ((A'P) A)(r, col)
          double acc = S.AtP[r] * e[(oA + col * n) * W];
          ALTRO_UNROLL
          for (int l = 1; l < n; ++l) acc += S.AtP[r + l * n] * e[(oA + l + col * n) * W];
          Qc[c][r] = e[(oLxx + r + col * n) * W] + acc;
        }
      }
      ALTRO_UNROLL
      for (int c = 0; c < CM; ++c) {
        const int col = j * CM + c;
        ALTRO_UNROLL
        for (int r = 0; r < n; ++r) {  // ((A'P) B)(r, col)
          double acc = S.AtP[r] * e[(oB + col * n) * W];
          ALTRO_UNROLL
          for (int l = 1; l < n; ++l) acc += S.AtP[r + l * n] * e[(oB + l + col * n) * W];
          S.Qxu[r + col * n] = e[(oLxu + r + col * n) * W] + acc;
        }
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) {  // ((B'P) B)(r, col)
          double acc = S.BtP[r] * e[(oB + col * n) * W];
          ALTRO_UNROLL
          for (int l = 1; l < n; ++l) acc += S.BtP[r + l * m] * e[(oB + l + col * n) * W];
          S.Quu[r + col * m] = e[(oLuu + r + col * m) * W] + acc;
        }
      }
      // every lane of the warp is done with the ring slot (its values went straight into the products above, so
      // the loads have completed); the proxy fence orders those generic-proxy reads before the async-proxy
      // refill (see k_backward_mat)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (next_k >= 0) {
        if (lane == 0) load_slot(issued % kStages, next_k);
        --next_k;
        ++issued;
      }
      // ---- phase 3: LLT of Quu + reg I, on both lanes (RegularizeActionValue :175-186, CalcGains :197-211) ---
      double Lm[m * m], rL[m], Quu[m * m], Qxu[n * m], Qu[m];
      ALTRO_UNROLL
      for (int q = 0; q < m * m; ++q) Quu[q] = S.Quu[q];
      ALTRO_UNROLL
      for (int q = 0; q < n * m; ++q) Qxu[q] = S.Qxu[q];
      ALTRO_UNROLL
      for (int q = 0; q < m; ++q) Qu[q] = S.Qu[q];
      ALTRO_UNROLL
      for (int c = 0; c < m; ++c)
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) Lm[r + c * m] = Quu[r + c * m] + (r == c ? 1.0 : 0.0) * reg;
      bool ok = true;
      ALTRO_UNROLL
      for (int kk = 0; kk < m; ++kk) {
        double xk = Lm[kk + kk * m];
        if (kk > 0) {
          double sq = 0.0;
          ALTRO_UNROLL
          for (int q = 0; q < kk; ++q) sq += Lm[kk + q * m] * Lm[kk + q * m];
          xk -= sq;
        }
        if (xk <= 0.0) ok = false;
        xk = sqrt(xk);
        Lm[kk + kk * m] = xk;
        rL[kk] = 1.0 / xk;
        ALTRO_UNROLL
        for (int r = kk + 1; r < m; ++r) {
          double a = Lm[r + kk * m];
          if (kk > 0) {
            double dot = 0.0;
            ALTRO_UNROLL
            for (int q = 0; q < kk; ++q) dot += Lm[r + q * m] * Lm[kk + q * m];
            a -= dot;
          }
          Lm[r + kk * m] = div_by(a, xk, rL[kk]);
        }
      }
      // -(L L')^-1 b on a register array (kept by value: a pointer argument would put the arrays in local memory)
      struct Vm { double v[m]; };
      auto solve = [&](Vm bv) -> Vm {
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) {
          double s = bv.v[r];
          ALTRO_UNROLL
          for (int q = 0; q < r; ++q) s -= Lm[r + q * m] * bv.v[q];
          bv.v[r] = div_by(s, Lm[r + r * m], rL[r]);
        }
        ALTRO_UNROLL
        for (int r = m - 1; r >= 0; --r) {
          double s = bv.v[r];
          ALTRO_UNROLL
          for (int q = r + 1; q < m; ++q) s -= Lm[q + r * m] * bv.v[q];
          bv.v[r] = div_by(s, Lm[r + r * m], rL[r]);
        }
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) bv.v[r] = bv.v[r] * -1;
        return bv;
      };
      // ---- phase 4: gains — this lane's state columns of K; d on both lanes ---------------------------------
      // (whatever is indexed by this lane's own column number — a run-time value — is read from shared
      // memory or kept in per-column registers; the full-size register copies are indexed statically)
      double d[m], Kown[CN][m], Qxu_own[CN][m];
      if (live && ok) {
        ALTRO_UNROLL
        for (int c = 0; c < CN; ++c) {
          const int col = j * CN + c;
          Vm b;
          ALTRO_UNROLL
          for (int r = 0; r < m; ++r) {
            Qxu_own[c][r] = S.Qxu[col + r * n];
            b.v[r] = Qxu_own[c][r];
          }
          b = solve(b);
          ALTRO_UNROLL
          for (int r = 0; r < m; ++r) {
            Kown[c][r] = b.v[r];
            S.K[r + col * m] = b.v[r];
          }
        }
        Vm b;
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) b.v[r] = Qu[r];
        b = solve(b);
        ALTRO_UNROLL
        for (int r = 0; r < m; ++r) d[r] = b.v[r];
      }
      __syncwarp(gmask);
      // ---- phase 5: cost-to-go of knot k (CalcCostToGo :220-230, unregularised Q) ---------------------------
      double pnew[CN];
      if (live && ok) {
        double Kall[m * n];
        ALTRO_UNROLL
        for (int q = 0; q < m * n; ++q) Kall[q] = S.K[q];
        double KtQuu[n * m];  // K'Quu, every row (the P update needs all of them)
        ALTRO_UNROLL
        for (int r = 0; r < n; ++r)
          ALTRO_UNROLL
          for (int c = 0; c < m; ++c) {
            double acc = Kall[r * m] * Quu[c * m];
            ALTRO_UNROLL
            for (int l = 1; l < m; ++l) acc += Kall[l + r * m] * Quu[l + c * m];
            KtQuu[r + c * n] = acc;
          }
        ALTRO_UNROLL
        for (int c = 0; c < CN; ++c) {
          const int col = j * CN + c;
          double kq[m];  // row col of K'Quu from this lane's own column of K
          ALTRO_UNROLL
          for (int cc = 0; cc < m; ++cc) {
            double acc = Kown[c][0] * Quu[cc * m];
            ALTRO_UNROLL
            for (int l = 1; l < m; ++l) acc += Kown[c][l] * Quu[l + cc * m];
            kq[cc] = acc;
          }
          double v1 = kq[0] * d[0], v2 = Kown[c][0] * Qu[0], v3 = Qxu_own[c][0] * d[0];
          ALTRO_UNROLL
          for (int l = 1; l < m; ++l) {
            v1 += kq[l] * d[l];
            v2 += Kown[c][l] * Qu[l];
            v3 += Qxu_own[c][l] * d[l];
          }
          pnew[c] = Qx[c] + v1 + v2 + v3;
          // column col of P: P(r, col) = Qxx(r, col) + (K'Quu K)(r, col) + (K'Qux)(r, col) + (Qxu K)(r, col)
          ALTRO_UNROLL
          for (int r = 0; r < n; ++r) {
            double t1 = KtQuu[r] * Kown[c][0], t2 = Kall[r * m] * Qxu_own[c][0], t3 = Qxu[r] * Kown[c][0];
            ALTRO_UNROLL
            for (int l = 1; l < m; ++l) {
              t1 += KtQuu[r + l * n] * Kown[c][l];
              t2 += Kall[l + r * m] * Qxu_own[c][l];
              t3 += Qxu[r + l * n] * Kown[c][l];
            }
            Pc[c][r] = Qc[c][r] + t1 + t2 + t3;
          }
          double* pk = L.kd(k);
          ALTRO_UNROLL
          for (int r = 0; r < m; ++r) pk[(r + col * m) * W] = Kown[c][r];
          if (kStoreCtg) {
            double* ct = L.ctg(k);
            ALTRO_UNROLL
            for (int r = 0; r < n; ++r) ct[(r + col * n) * W] = Pc[c][r];
            ct[(n * n + col) * W] = pnew[c];
          }
        }
        if (j == 0) {
          double* pk = L.kd(k);
          ALTRO_UNROLL
          for (int r = 0; r < m; ++r) pk[(m * n + r) * W] = d[r];
          double a = d[0] * Qu[0];
          ALTRO_UNROLL
          for (int r = 1; r < m; ++r) a += d[r] * Qu[r];
          double Quud[m];
          ALTRO_UNROLL
          for (int r = 0; r < m; ++r) {
            double acc = Quu[r] * d[0];
            ALTRO_UNROLL
            for (int l = 1; l < m; ++l) acc += Quu[r + l * m] * d[l];
            Quud[r] = acc;
          }
          double bq = d[0] * Quud[0];
          ALTRO_UNROLL
          for (int r = 1; r < m; ++r) bq += d[r] * Quud[r];
          dV0 += a;
          dV1 += 0.5 * bq;
        }
        ALTRO_UNROLL
        for (int c = 0; c < CN; ++c) S.p[j * CN + c] = pnew[c];  // read again only after the pair barrier below
      }
      if (live && !ok) {  // ilqr.hpp:409-427: raise the regularisation, restart from k = N-1
        increase_reg(o, reg, dreg);
        if (reg >= o.bp_reg_max) max_reg_count++;
        if (max_reg_count >= o.bp_reg_fail_threshold) {
          st = kBackwardPassRegularizationFailed;
          repeat = false;
        }
        live = false;
      } else if (live && k == 0) {
        repeat = false;
      }
      __syncwarp(gmask);  // p of knot k is complete before the next phase 1 reads it
    }
  }
  if (valid && j == 0) {  // the lane that accumulated deltaV writes the instance's scalars
    decrease_reg(o, reg, dreg);
    L.sc(S_REG) = reg;
    L.sc(S_DREG) = dreg;
    L.sc(S_DV0) = dV0;
    L.sc(S_DV1) = dV1;
    L.is(I_STATUS) = st;
    if (kPhased) {
      L.sc(S_REG_IN) = reg_in;
      L.sc(S_DREG_IN) = dreg_in;
    }
  }
}

template <class M, int kStages>
constexpr int backward_coop_smem() {
  return kStages * Lane<M, kCoopTile>::nexp * kCoopInst * static_cast<int>(sizeof(double)) +
         kCoopInst * static_cast<int>(sizeof(CoopScratch<M>)) + kStages * 8;
}

}  // namespace altro_b200
