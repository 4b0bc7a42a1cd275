// rodeo_b200: covariance schedules -- solve_sim under the state-independent interrogations.
//
// interrogate_chkrebtii, interrogate_schober and interrogate_rodeo return wgt_meas == 0 and a var_meas that depends on
// the predicted VARIANCE only (reference src/rodeo/interrogate.py:13-60, 87-115), so in _solve_filter
// (src/rodeo/solve.py:59-88) W_meas = ode_weight and the whole covariance recursion -- predicted and filtered
// variances, Kalman gains, the factor interrogate_chkrebtii draws with, and in the backward pass the smoothing gains
// and the factors of the conditional variances of smooth_sim (src/rodeo/kalmantv/standard.py:220-255) -- is a function
// of (Q, R, W, n_steps) alone: the same for every theta, every particle and every key.  Only the MEANS see the ODE.
// (interrogate_kramer puts the Jacobian into W_meas; its covariances depend on the state and it keeps the full kernels.)
//
// So the covariance recursion is run ONCE per (prior, n_steps) into a small table in device memory ("schedule",
// ~300 B per block-step, L2-resident; cached by the library across calls -- a pseudo-marginal MCMC re-uses it in every
// iteration) by the SAME __device__ functions the full kernels inline, and the per-theta kernel carries the block means
// only: per block-step one mean predict, the draw, the right-hand side and a gain application forward, one gain
// application and one draw backward.  The history the backward sweep needs shrinks from means + packed covariances
// (NB (P + P(P+1)/2) values per theta*step) to the means (NB P values).
//
// Results: the table entries are bitwise the values every theta of the full kernel computes for itself (same functions,
// same inputs, same order), and the mean arithmetic below repeats the full kernels' operation order, so without a
// per-theta prior scale the draws are BITWISE those of solve_sim_kernel / solve_sim_bl_kernel
// (tests/test_gpu_parity.py::test_solve_sim_schedule_*).  A per-theta prior scale R(theta, b) = rs R[b] (an IBM prior
// whose sigma is part of theta, the reference's pseudo-marginal walk-through, docs/examples/parameter.md:218-236) scales
// every covariance of block b by rs exactly -- S_0 = 0, S_p = Q S Q^T + rs R, V = W S_p W^T or 0 -- so gains are
// unchanged and factors scale by sqrt(rs): applied to the tabulated unit-scale factors per theta (agreement with the
// full kernel to rounding, ~1e-15).  interrogate_schober with a per-theta scale stays on the full kernels: its filtered
// variance is exactly singular, and the sign of a rounding-noise pivot is not scale invariant.
//
// No host / toolkit includes (kept NVRTC-clean like the other device headers).
#pragma once
#include "rodeo_kernels.cuh"

#ifndef RODEO_SCHED_MINB
#define RODEO_SCHED_MINB 14
#endif
#ifndef RODEO_SCHED_SMEM
#define RODEO_SCHED_SMEM 12000     // staging bytes per warp for the draws (x_out != nullptr)
#endif
#ifndef RODEO_SCHED_PF
#define RODEO_SCHED_PF 6           // L2 prefetch distance of the backward sweep's history loads, in steps
#endif

namespace rodeo {

template <typename T, class Model, int INTERR, int QK>
struct Sched {
  static constexpr int NB = Model::NB, P = Model::P, M = Model::M, JC = Model::JCOLS, WK = Model::WCOL;
  static constexpr int NS = P * (P + 1) / 2, MS = M * (M + 1) / 2, ROW = NB * P;
  static constexpr bool UNITW = (QK == QK_UNIT_UPPER) && (M == 1);
  static constexpr bool DRAW = INTERR == INTERR_CHKREBTII;
  static constexpr bool HAS_V = INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII;
  // forward row of (step n, block b): [leading JC rows of the factor of S_p (chkrebtii)] [gain]
  //   gain = S_p e_WK (P values) and 1 / S_z           (structured instantiation, update_unit_row)
  //        = K^T (M x P)                               (dense instantiation, update)
  static constexpr int NFA = DRAW ? JC * (JC + 1) / 2 : 0;
  static constexpr int NGAIN = UNITW ? P + 1 : M * P;
  static constexpr int FWD = (NFA + NGAIN + 1) & ~1;          // even => rows stay aligned for 2-element vector loads
  // backward row of (row n, block b), n = 1 .. N: smoothing gain G (P x P), lower factor A of the conditional variance
  // (packed, lidx); row N carries the factor of S_f[N] (terminal draw, solve.py:182-186)
  static constexpr int BWD = (P * P + NS + 1) & ~1;
  __host__ __device__ static constexpr i64 bwd_off(int N) { return (i64)N * NB * FWD; }
  __host__ __device__ static constexpr i64 sf_off(int N) { return bwd_off(N) + (i64)N * NB * BWD; }
  __host__ __device__ static constexpr i64 total(int N) { return sf_off(N) + (i64)N * NB * NS; }
  // staged time rows of the draws per warp
  static constexpr int KOUT_RAW = RODEO_SCHED_SMEM / (ROW * SEG_PITCH * (int)sizeof(T));
  static constexpr int KOUT = KOUT_RAW < 1 ? 1 : (KOUT_RAW > 16 ? 16 : KOUT_RAW);
  static constexpr int SMEM = KOUT * ROW * SEG_PITCH * (int)sizeof(T);
};

// CNT table values at a warp-uniform, 2-element aligned address
template <typename T, int CNT>
RD_DEV void sched_load(const T* __restrict__ p, T (&r)[CNT]) {
  static_assert(CNT % 2 == 0, "rows are padded to an even length");
  if constexpr (sizeof(T) == 8) {
    RD_UNROLL for (int k = 0; k < CNT / 2; ++k) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(p) + k);
      r[2 * k] = v.x; r[2 * k + 1] = v.y;
    }
  } else {
    RD_UNROLL for (int k = 0; k < CNT / 2; ++k) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p) + k);
      r[2 * k] = v.x; r[2 * k + 1] = v.y;
    }
  }
}

// mean part of predict<T, P, QK> (same operation order)
template <typename T, int P, int QK, typename MT>
RD_DEV void sched_predict_mean(const T (&Q)[P][P], const MT (&mu)[P], MT (&mup)[P]) {
  RD_UNROLL for (int i = 0; i < P; ++i) {
    if (QK == QK_UNIT_UPPER) {
      MT m = mu[i];
      RD_UNROLL for (int j = i + 1; j < P; ++j) m = rd_fma((MT)Q[i][j], mu[j], m);
      mup[i] = m;
    } else {
      MT m = (MT)Q[i][0] * mu[0];
      RD_UNROLL for (int j = 1; j < P; ++j) m = rd_fma((MT)Q[i][j], mu[j], m);
      mup[i] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Schedule, forward part: one thread per block walks the covariance recursion of _solve_filter (unit prior scale)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK>
__global__ void __launch_bounds__(32)
sched_forward_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C, int N,
                     T* __restrict__ tab) {
  typedef Sched<T, Model, INTERR, QK> SC;
  typedef typename MeanOf<T>::type MT;
  constexpr int NB = SC::NB, P = SC::P, M = SC::M, JC = SC::JC, WK = SC::WK, NS = SC::NS, MS = SC::MS;
  const int b = threadIdx.x;
  if (b >= NB || blockIdx.x != 0) return;
  T S[NS];
  MT mu0[P], mp[P];
  RD_UNROLL for (int k = 0; k < NS; ++k) S[k] = T(0);
  RD_UNROLL for (int i = 0; i < P; ++i) mu0[i] = MT(0);
  T* sf = tab + SC::sf_off(N);
  for (int n = 0; n < N; ++n) {
    T Sp[NS];
    predict<T, P, QK>(C.Q[b], C.R[b], T(1), mu0, S, mp, Sp);
    T* row = tab + ((i64)n * NB + b) * SC::FWD;
    if constexpr (SC::DRAW) {
      T A[P][P];
      psd_factor<T, P>(Sp, A);
      int k = 0;
      RD_UNROLL for (int j = 0; j < JC; ++j)
        RD_UNROLL for (int kk = 0; kk <= j; ++kk) row[k++] = A[j][kk];
    }
    if constexpr (SC::UNITW) {
      // update_unit_row without a Jacobian: v = S_p e_WK, S_z = V + v_WK with V = S_p[WK][WK] (chkrebtii, rodeo) or 0
      T v[P];
      RD_UNROLL for (int i = 0; i < P; ++i) v[i] = Sp[sym<P>(i, WK)];
      const T V = Sp[sidx<P>(WK, WK)];
      const T Sm = SC::HAS_V ? V + v[WK] : v[WK];
      const T rS = rcp(Sm);
      RD_UNROLL for (int i = 0; i < P; ++i) row[SC::NFA + i] = v[i];
      row[SC::NFA + P] = rS;
      RD_UNROLL for (int i = 0; i < P; ++i) {
        const T k = v[i] * rS;
        RD_UNROLL for (int j = i; j < P; ++j) Sp[sidx<P>(i, j)] = rd_fma(-k, v[j], Sp[sidx<P>(i, j)]);
      }
    } else {
      // Fwd::interrogate's var_meas = W S_p W^T and update<T, P, M>'s covariance part
      T V[MS];
      if constexpr (SC::HAS_V) {
        T u[M][P];
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int i = 0; i < P; ++i) {
            T a = Sp[sym<P>(i, 0)] * C.W[b][r][0];
            RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(Sp[sym<P>(i, j)], C.W[b][r][j], a);
            u[r][i] = a;
          }
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int s = r; s < M; ++s) {
            T a = C.W[b][r][0] * u[s][0];
            RD_UNROLL for (int i = 1; i < P; ++i) a = rd_fma(C.W[b][r][i], u[s][i], a);
            V[sidx<M>(r, s)] = a;
          }
      } else {
        RD_UNROLL for (int k = 0; k < MS; ++k) V[k] = T(0);
      }
      T v[M][P], Kt[M][P], Sm[MS];
      RD_UNROLL for (int r = 0; r < M; ++r)
        RD_UNROLL for (int i = 0; i < P; ++i) {
          T a = Sp[sym<P>(i, 0)] * C.W[b][r][0];
          RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(Sp[sym<P>(i, j)], C.W[b][r][j], a);
          v[r][i] = a; Kt[r][i] = a;
        }
      RD_UNROLL for (int r = 0; r < M; ++r)
        RD_UNROLL for (int s = r; s < M; ++s) {
          T a = V[sidx<M>(r, s)];
          RD_UNROLL for (int i = 0; i < P; ++i) a = rd_fma(C.W[b][r][i], v[s][i], a);
          Sm[sidx<M>(r, s)] = a;
        }
      solve_small<T, M, P>(Sm, Kt);
      RD_UNROLL for (int r = 0; r < M; ++r)
        RD_UNROLL for (int i = 0; i < P; ++i) row[SC::NFA + r * P + i] = Kt[r][i];
      RD_UNROLL for (int i = 0; i < P; ++i)
        RD_UNROLL for (int j = i; j < P; ++j) {
          T s = Sp[sidx<P>(i, j)];
          RD_UNROLL for (int r = 0; r < M; ++r) s = rd_fma(-Kt[r][i], v[r][j], s);
          Sp[sidx<P>(i, j)] = s;
        }
    }
    RD_UNROLL for (int k = 0; k < NS; ++k) { S[k] = Sp[k]; sf[((i64)n * NB + b) * NS + k] = Sp[k]; }   // S_f[n + 1]
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Schedule, backward part: rows are independent given S_f[n] -- one thread per (row n = 1 .. N, block)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK>
__global__ void __launch_bounds__(128)
sched_backward_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C, int N,
                      T* __restrict__ tab) {
  typedef Sched<T, Model, INTERR, QK> SC;
  typedef typename MeanOf<T>::type MT;
  constexpr int NB = SC::NB, P = SC::P, NS = SC::NS;
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)N * NB) return;
  const int n = (int)(t / NB) + 1, b = (int)(t % NB);
  const T* sf = tab + SC::sf_off(N) + ((i64)(n - 1) * NB + b) * NS;
  T* row = tab + SC::bwd_off(N) + ((i64)(n - 1) * NB + b) * SC::BWD;
  T Sf[NS], A[P][P], G[P][P];
  RD_UNROLL for (int k = 0; k < NS; ++k) Sf[k] = sf[k];
  // FilterConsts indexed by a run-time block: copy the block's Q, R out of the constant bank
  T Q[P][P], R[NS];
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) Q[i][j] = C.Q[b][i][j];
  RD_UNROLL for (int k = 0; k < NS; ++k) R[k] = C.R[b][k];
  if (n < N) {
    MT mu0[P], mp[P];
    RD_UNROLL for (int i = 0; i < P; ++i) mu0[i] = MT(0);
    T Sp[NS], Ct[P][P], Cv[NS];
    predict<T, P, QK>(Q, R, T(1), mu0, Sf, mp, Sp);            // pred[n+1]
    smooth_gain<T, P, QK>(Q, Sf, Sp, G, Ct);
    cond_var<T, P>(Sf, G, Ct, Cv);
    psd_factor<T, P>(Cv, A);
  } else {
    psd_factor<T, P>(Sf, A);
    RD_UNROLL for (int i = 0; i < P; ++i)
      RD_UNROLL for (int j = 0; j < P; ++j) G[i][j] = T(0);
  }
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) row[i * P + j] = G[i][j];
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j <= i; ++j) row[P * P + lidx(i, j)] = A[i][j];
}

// ------------------------------------------------------------------------------------------------------------------
// solve_sim over a schedule: one lane per theta, block means only
// ------------------------------------------------------------------------------------------------------------------
// Table rows are warp-uniform.  Read straight from global memory they put an L2 round trip (~700 cycles on B200) on
// the dependency chain of every step, which is all this kernel has left (C5: 1.7 warps per SM sub-partition), so each
// warp streams them through shared memory: chunk c+1 travels (cp.async) while chunk c is consumed (LDS broadcast).

// per-warp staging of `rows` table rows of ROWSZ elements starting at row r0 (16-byte / 8-byte granules)
template <typename T, int ROWSZ>
RD_DEV void sched_stage(const T* __restrict__ tab, int r0, int rows, T* smem, int lane) {
  const T* g = tab + (i64)r0 * ROWSZ;
  for (int e = 2 * lane; e < rows * ROWSZ; e += 64) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem + e);
    if (sizeof(T) == 8) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g + e));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g + e));
  }
  cp_async_commit();
}
RD_DEV void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// CNT values of a staged row (shared memory, warp-uniform address: one broadcast LDS.128 per two values)
template <typename T, int CNT>
RD_DEV void sched_lds(const T* p, T (&r)[CNT]) {
  static_assert(CNT % 2 == 0, "rows are padded to an even length");
  if constexpr (sizeof(T) == 8) {
    RD_UNROLL for (int k = 0; k < CNT / 2; ++k) {
      const double2 v = reinterpret_cast<const double2*>(p)[k];
      r[2 * k] = v.x; r[2 * k + 1] = v.y;
    }
  } else {
    RD_UNROLL for (int k = 0; k < CNT / 2; ++k) {
      const float2 v = reinterpret_cast<const float2*>(p)[k];
      r[2 * k] = v.x; r[2 * k + 1] = v.y;
    }
  }
}

// History of filtered means: entry e (= mu_f[e + 1]) holds ROW values per theta, stored as 2-element vectors with theta
// innermost -- [entry][pair][theta][2] (+ [entry][theta] for an odd last value) -- so a lane moves two values per
// instruction and a warp access is one contiguous run (512 B of float64 pairs).
template <typename T> struct Vec2Of { typedef double2 type; };
template <> struct Vec2Of<float> { typedef float2 type; };
template <typename T, int ROW>
struct SchedHist {
  static constexpr int NP = ROW / 2;
  typedef typename Vec2Of<T>::type V2;
  T* base; i64 ldb, idx;
  RD_DEV T* entry(int e) const { return base + (i64)e * ROW * ldb; }
  template <typename MT>
  RD_DEV void store(int e, const MT (&v)[ROW]) const {
    T* h = entry(e);
    RD_UNROLL for (int j = 0; j < NP; ++j) {
      V2 w; w.x = (T)v[2 * j]; w.y = (T)v[2 * j + 1];
      reinterpret_cast<V2*>(h)[j * ldb + idx] = w;
    }
    if (ROW & 1) h[(i64)NP * 2 * ldb + idx] = (T)v[ROW - 1];
  }
  RD_DEV void load(int e, T (&v)[ROW]) const {
    const T* h = entry(e);
    RD_UNROLL for (int j = 0; j < NP; ++j) {
      const V2 w = reinterpret_cast<const V2*>(h)[j * ldb + idx];
      v[2 * j] = w.x; v[2 * j + 1] = w.y;
    }
    if (ROW & 1) v[ROW - 1] = h[(i64)NP * 2 * ldb + idx];
  }
  RD_DEV void prefetch(int e) const {
    const T* h = entry(e);
    RD_UNROLL for (int j = 0; j < NP; ++j)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const V2*>(h) + j * ldb + idx));
    if (ROW & 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(h + (i64)NP * 2 * ldb + idx));
  }
};

template <typename T, class Model, int INTERR, int QK, bool XOUT>
struct SchedSim {
  typedef Sched<T, Model, INTERR, QK> SC;
  // backward segment = staged output rows (XOUT) = table chunk
  static constexpr int K = XOUT ? SC::KOUT : 16;
  static constexpr int OUT_ELEMS = XOUT ? ((K * SC::ROW * SEG_PITCH + 1) & ~1) : 0;   // even: the table chunks stay 16-byte aligned
  static constexpr int BCH_ELEMS = K * SC::NB * SC::BWD;                 // one backward table chunk
  static constexpr int TAB_ELEMS = (XOUT ? 1 : 2) * BCH_ELEMS;             // XOUT: the next chunk travels during the copy-out
  // forward chunk: what fits the same region, at most 32 steps
  static constexpr int FROW = SC::NB * SC::FWD;
  static constexpr int CHF_RAW = (OUT_ELEMS + TAB_ELEMS) / (2 * FROW);
  static constexpr int CHF = CHF_RAW > 32 ? 32 : CHF_RAW;
  static constexpr int SMEM = (OUT_ELEMS + TAB_ELEMS) * (int)sizeof(T);
  static_assert(CHF >= 1, "forward chunk");
};

template <typename T, class Model, int INTERR, int QK, bool XOUT>
__global__ void __launch_bounds__(32, RODEO_SCHED_MINB)
solve_sim_sched_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                       const CommonArgs<T> a, const T* __restrict__ tab, const T* __restrict__ z_smooth,
                       T* __restrict__ stash, i64 ldb, T* __restrict__ x_out, const SimLoglik<T> sl) {
  typedef Sched<T, Model, INTERR, QK> SC;
  typedef SchedSim<T, Model, INTERR, QK, XOUT> SS;
  typedef typename MeanOf<T>::type MT;
  typedef typename Model::template Par<MT> Par;
  constexpr int NB = SC::NB, P = SC::P, M = SC::M, JC = SC::JC, WK = SC::WK, ROW = SC::ROW, K = SS::K;
  const int lane = threadIdx.x;
  const i64 theta0 = (i64)blockIdx.x * 32;
  i64 idx = theta0 + lane;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;                 // the whole warp takes part in the cooperative copy-out
  const Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  T* smem = reinterpret_cast<T*>(rodeo_dyn_smem);
  const T* x0 = a.ode_init + idx * ROW;

  // factors scale with the square root of the per-theta prior scale; a multiplication by exactly 1 otherwise
  MT sq[NB];
  RD_UNROLL for (int b = 0; b < NB; ++b) sq[b] = a.r_scale != nullptr ? sqrt((MT)a.r_scale[idx * NB + b]) : MT(1);

  MT mu[NB][P];
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = (MT)x0[b * P + i];
  const SchedHist<T, ROW> hist{stash, ldb, idx};

  // ---- forward: mu_f[n] -> mu_f[n+1]  (solve.py:59-88 with the variances read from the schedule) ----
  auto interr_normals = [&](int n, T (&zc)[NB][JC]) {
    if constexpr (SC::DRAW) {
      if (a.z_interr != nullptr) {
        const T* z = a.z_interr + (idx * a.n_steps + n) * (NB * P);
        RD_UNROLL for (int b = 0; b < NB; ++b)
          RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = z[b * P + j];
      } else {
        T z[NB * JC];
        philox_normals<T, NB * JC>(a.key0, a.key1, a.particle_offset + idx, n, TAG_INTERR_A, z);
        RD_UNROLL for (int b = 0; b < NB; ++b)
          RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = z[b * JC + j];
      }
    } else {
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = T(0);
    }
  };
  {
    constexpr int CHF = SS::CHF, FROW = SS::FROW;
    T zc[NB][JC];
    interr_normals(0, zc);
    MT t_next = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, 0, N) : MT(0);
    const int nch = (N + CHF - 1) / CHF;
    sched_stage<T, FROW>(tab, 0, N < CHF ? N : CHF, smem, lane);
    for (int c = 0; c < nch; ++c) {
      const int n_lo = c * CHF, n_hi = (n_lo + CHF) < N ? (n_lo + CHF) : N;
      if (c + 1 < nch) {
        sched_stage<T, FROW>(tab, n_hi, (n_hi + CHF) < N ? CHF : (N - n_hi), smem + ((c + 1) & 1) * CHF * FROW, lane);
        cp_async_wait_1();
      } else {
        cp_async_wait_all();
      }
      __syncwarp();
      const T* frow = smem + (c & 1) * CHF * FROW;
      _Pragma("unroll 1") for (int n = n_lo; n < n_hi; ++n) {
        // the next step's normals do not depend on the state: issued ahead of this step's dependency chain
        T zn[NB][JC];
        interr_normals(n + 1 < N ? n + 1 : N - 1, zn);            // (the last one is generated twice and dropped)
        const MT t = t_next;
        if (Model::USES_TIME) t_next = step_time<MT>(a.t_min, a.t_max, n + 1, N);
        MT mp[NB][P], x[NB][JC], f[NB][M];
        T r[NB][SC::FWD];
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          sched_lds<T, SC::FWD>(frow + b * SC::FWD, r[b]);
          sched_predict_mean<T, P, QK, MT>(C.Q[b], mu[b], mp[b]);
          RD_UNROLL for (int j = 0; j < JC; ++j) {
            MT acc = mp[b][j];
            if constexpr (SC::DRAW) {
              RD_UNROLL for (int k = 0; k <= j; ++k)
                acc = rd_fma((MT)r[b][j * (j + 1) / 2 + k], (MT)zc[b][k] * sq[b], acc);
            }
            x[b][j] = acc;
          }
        }
        frow += FROW;
        Model::template rhs<MT, MT>(q, t, x, f);
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          if constexpr (SC::UNITW) {
            const MT res = sub_exact(f[b][0], mp[b][WK]);
            const MT g = res * (MT)r[b][SC::NFA + P];
            RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = rd_fma((MT)r[b][SC::NFA + i], g, mp[b][i]);
          } else {
            MT res[M];
            RD_UNROLL for (int rr = 0; rr < M; ++rr) {
              MT acc = f[b][rr];
              RD_UNROLL for (int j = 0; j < P; ++j) acc = rd_fma(-(MT)C.W[b][rr][j], mp[b][j], acc);
              res[rr] = acc;
            }
            RD_UNROLL for (int i = 0; i < P; ++i) {
              MT m = mp[b][i];
              RD_UNROLL for (int rr = 0; rr < M; ++rr) m = rd_fma((MT)r[b][SC::NFA + rr * P + i], res[rr], m);
              mu[b][i] = m;
            }
          }
        }
        if (live && n + 1 < N) {                                  // history entry n = mu_f[n+1]
          MT flat[ROW];
          RD_UNROLL for (int b = 0; b < NB; ++b)
            RD_UNROLL for (int i = 0; i < P; ++i) flat[b * P + i] = mu[b][i];
          hist.store(n, flat);
        }
        RD_UNROLL for (int b = 0; b < NB; ++b)
          RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = zn[b][j];
      }
      __syncwarp();                                               // chunk c is free before chunk c+2 overwrites it
    }
  }

  // ---- backward: the sampling smoother over the tabulated gains and factors ----
  auto normals = [&](int n, T (&z)[NB * P]) {
    if (z_smooth != nullptr) {
      const T* zp = z_smooth + (idx * (i64)(N + 1) + n) * ROW;
      RD_UNROLL for (int k = 0; k < ROW; ++k) z[k] = zp[k];
    } else {
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        T zb[P];
        philox_normals<T, P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_SMOOTH, zb, b * ((P + 3) / 4));
        RD_UNROLL for (int k = 0; k < P; ++k) z[b * P + k] = zb[k];
      }
    }
  };
  constexpr int BROW = NB * SC::BWD;
  const T* btab = tab + SC::bwd_off(N);                          // row n (1 .. N) at table index n - 1
  T* tbuf = smem + SS::OUT_ELEMS;                                // two chunks of K table rows
  T* buf = smem;                                                 // staged output rows (XOUT)
  // segment j covers rows n0 = jK .. n0 + cnt - 1 (<= N - 1); its table rows are indices max(n0, 1) - 1 .. n0 + cnt - 2
  auto stage_seg = [&](int j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    const int lo = (n0 < 1 ? 1 : n0) - 1, hi = n0 + cnt - 2;     // inclusive
    if (hi >= lo) sched_stage<T, BROW>(btab, lo, hi - lo + 1, tbuf + (XOUT ? 0 : (j & 1) * SS::BCH_ELEMS), lane);
    else cp_async_commit();
  };
  const int jtop = (N - 1) / K;
  stage_seg(jtop);
  MT x[NB][P];
  {                                                             // terminal draw from N(mu_f[N], S_f[N])
    T z[NB * P];
    normals(N, z);
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T r[SC::BWD];
      sched_load<T, SC::BWD>(btab + ((i64)(N - 1) * NB + b) * SC::BWD, r);
      RD_UNROLL for (int i = 0; i < P; ++i) {
        MT acc = mu[b][i];
        RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)r[P * P + lidx(i, k)], (MT)z[b * P + k] * sq[b], acc);
        x[b][i] = acc;
      }
    }
    if (live && x_out != nullptr) store_mean_row<T, NB, P>(x_out + (idx * (i64)(N + 1) + N) * ROW, x);
  }
  SimLoglikAcc<T, MT> la;
  la.init(sl, N + 1);
  auto x_of = [&](int bb) {
    MT v = x[0][0];
    RD_UNROLL for (int c = 1; c < NB; ++c) v = bb == c ? x[c][0] : v;
    return v;
  };
  la.row(sl, N, NB, 0, NB, x_of);

  // mu_f[n] is loaded one row ahead into registers and requested from HBM RODEO_SCHED_PF rows ahead
  T nmu[ROW];
  // (row indices below 1 are clamped to 1: a redundant load instead of a branch; N == 1 has no history at all)
  auto hload = [&](int n) { hist.load((n < 1 ? 1 : n) - 1, nmu); };
  auto hprefetch = [&](int n) { hist.prefetch((n < 1 ? 1 : n) - 1); };
  T z[NB * P];
  if (N > 1) {
    for (int n = N - 2; n > N - 2 - RODEO_SCHED_PF; --n) hprefetch(n);
    hload(N - 1);
    normals(N - 1, z);
  } else {
    RD_UNROLL for (int k = 0; k < ROW; ++k) { z[k] = T(0); nmu[k] = T(0); }
  }
  for (int j = jtop; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    if (!XOUT && j > 0) { stage_seg(j - 1); cp_async_wait_1(); }
    else cp_async_wait_all();
    __syncwarp();
    const T* trow0 = tbuf + (XOUT ? 0 : (j & 1) * SS::BCH_ELEMS) - (i64)((n0 < 1 ? 1 : n0) - 1) * BROW;   // row index n-1 -> trow0 + (n-1) BROW
    _Pragma("unroll 1") for (int s = cnt - 1; s >= 0; --s) {
      const int n = n0 + s;
      if (n == 0) {                                             // row 0 = ode_init: x0 is known, not sampled
        if constexpr (XOUT)
          RD_UNROLL for (int k = 0; k < ROW; ++k) buf[(s * ROW + k) * SEG_PITCH + lane] = x0[k];
        la.row(sl, 0, NB, 0, NB, [&](int bb) { return (MT)x0[bb * P]; });
        break;
      }
      MT mf[NB][P];
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int i = 0; i < P; ++i) mf[b][i] = (MT)nmu[b * P + i];
      hload(n - 1);
      hprefetch(n - 1 - RODEO_SCHED_PF);
      T zn[NB * P];
      normals(n > 1 ? n - 1 : 1, zn);
      const T* brow = trow0 + (i64)(n - 1) * BROW;
      MT xn[NB][P];
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        T r[SC::BWD];
        sched_lds<T, SC::BWD>(brow + b * SC::BWD, r);
        MT mp[P], zs[P];
        RD_UNROLL for (int k = 0; k < P; ++k) zs[k] = (MT)z[b * P + k] * sq[b];
        sched_predict_mean<T, P, QK, MT>(C.Q[b], mf[b], mp);
        // m = mu_f + G (x' - mu_p) ;  x = m + A z      (standard.py:251-254, solve.py:179)
        RD_UNROLL for (int i = 0; i < P; ++i) {
          MT acc = mf[b][i];
          RD_UNROLL for (int jj = 0; jj < P; ++jj) acc = rd_fma((MT)r[i * P + jj], x[b][jj] - mp[jj], acc);
          RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)r[P * P + lidx(i, k)], zs[k], acc);
          xn[b][i] = acc;
        }
      }
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int i = 0; i < P; ++i) {
          x[b][i] = xn[b][i];
          if constexpr (XOUT) buf[(s * ROW + b * P + i) * SEG_PITCH + lane] = (T)xn[b][i];
        }
      RD_UNROLL for (int k = 0; k < ROW; ++k) z[k] = zn[k];
      la.row(sl, n, NB, 0, NB, x_of);
    }
    __syncwarp();                                               // table chunk j is free; staged rows are complete
    if constexpr (XOUT) {
      if (j > 0) stage_seg(j - 1);                              // lands while the rows of segment j are copied out
      // rows n0 .. n0+cnt-1: per theta one contiguous run of cnt*ROW elements, consecutive lanes store consecutive elements
      constexpr int NIT = (K * ROW + 31) / 32;
      const int run = cnt * ROW;
      const i64 stride = (i64)(N + 1) * ROW;
      T* dst = x_out + (theta0 * (i64)(N + 1) + n0) * ROW + lane;
      const int nth = (a.B - theta0) < 32 ? (int)(a.B - theta0) : 32;
      RD_UNROLL4 for (int th = 0; th < nth; ++th) {
        RD_UNROLL for (int it = 0; it < NIT; ++it) {
          const int rr = lane + 32 * it;
          if (rr < run) dst[32 * it] = buf[rr * SEG_PITCH + th];
        }
        dst += stride;
      }
      __syncwarp();
    }
  }
  if (sl.out != nullptr && live) sl.out[idx] = (T)la.ll;
}


// ------------------------------------------------------------------------------------------------------------------
// solve_sim over a schedule with one lane per (theta, block)
// ------------------------------------------------------------------------------------------------------------------
// solve_sim_sched_kernel is latency-bound when the batch does not fill the GPU: most of a step is the Philox /
// Box-Muller work of its normals, on one dependency chain.  As in solve_sim_bl_kernel the blocks of a theta are spread
// over lanes (block-major, lane = b * TW + tl, TW = 32 / n_block thetas per warp): the block recursions only meet in the
// right-hand side, through one shuffle per visible column, so a lane applies the normals of its own block only and the
// launch has n_block times more warps of a shorter chain.  It pays for small batches only (FitzHugh-Nagumo, N = 800:
// 8,192 particles 0.90 -> 0.74 ms, 16,384: 0.91 -> 0.82, 32,768: 1.06 -> 1.12): the forward step's single Philox call and
// the right-hand side are repeated by every lane of a theta, so the host uses it while its grid fills at most half of the
// resident slots (abi_solve_sim.cu).  Same values in the same order per block: the draws are bitwise those of
// solve_sim_sched_kernel; the fused log-likelihood sums per block first (as solve_sim_bl_kernel does).  Requires
// 32 % n_block == 0 (the mean history is laid out per lane).
template <typename T, class Model, int INTERR, int QK, bool XOUT>
struct SchedSimBl {
  typedef Sched<T, Model, INTERR, QK> SC;
  static constexpr bool OK = SC::NB >= 2 && 32 % SC::NB == 0;
  static constexpr int TW = 32 / SC::NB, PITCH = TW + 1;
  static constexpr int KOUT_RAW = RODEO_SCHED_SMEM / (SC::ROW * PITCH * (int)sizeof(T));
  static constexpr int K = XOUT ? (KOUT_RAW < 1 ? 1 : (KOUT_RAW > 16 ? 16 : KOUT_RAW)) : 16;
  static constexpr int OUT_ELEMS = XOUT ? ((K * SC::ROW * PITCH + 1) & ~1) : 0;
  static constexpr int BCH_ELEMS = K * SC::NB * SC::BWD;
  static constexpr int TAB_ELEMS = (XOUT ? 1 : 2) * BCH_ELEMS;
  static constexpr int FROW = SC::NB * SC::FWD;
  static constexpr int CHF_RAW = (OUT_ELEMS + TAB_ELEMS) / (2 * FROW);
  static constexpr int CHF = CHF_RAW > 32 ? 32 : CHF_RAW;
  static constexpr int SMEM = (OUT_ELEMS + TAB_ELEMS) * (int)sizeof(T);
  static_assert(CHF >= 1, "forward chunk");
};

template <typename T, class Model, int INTERR, int QK, bool XOUT>
__global__ void __launch_bounds__(32, RODEO_SCHED_MINB)
solve_sim_sched_bl_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                          const CommonArgs<T> a, const T* __restrict__ tab, const T* __restrict__ z_smooth,
                          T* __restrict__ stash, i64 ldb, T* __restrict__ x_out, const SimLoglik<T> sl) {
  typedef Sched<T, Model, INTERR, QK> SC;
  typedef SchedSimBl<T, Model, INTERR, QK, XOUT> SS;
  typedef typename MeanOf<T>::type MT;
  typedef typename Model::template Par<MT> Par;
  constexpr int NB = SC::NB, P = SC::P, M = SC::M, JC = SC::JC, WK = SC::WK, ROW = SC::ROW, K = SS::K;
  constexpr int TW = SS::TW, PITCH = SS::PITCH;
  static_assert(SS::OK, "block lanes need 32 % n_block == 0");
  const int lane = threadIdx.x;
  const int b = lane / TW, tl = lane - b * TW;              // block-major lanes: every lane is a (theta, block)
  const i64 theta0 = (i64)blockIdx.x * TW;
  i64 idx = theta0 + tl;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  T* smem = reinterpret_cast<T*>(rodeo_dyn_smem);
  const T* x0 = a.ode_init + idx * ROW + b * P;

  // this lane's block of the shared constants, in registers (the constant bank cannot be indexed per lane)
  T Q[P][P], W[M][P];
  RD_UNROLL for (int c = 0; c < NB; ++c)
    if (c == b) {
      RD_UNROLL for (int i = 0; i < P; ++i)
        RD_UNROLL for (int j = 0; j < P; ++j) Q[i][j] = C.Q[c][i][j];
      RD_UNROLL for (int r = 0; r < M; ++r)
        RD_UNROLL for (int j = 0; j < P; ++j) W[r][j] = C.W[c][r][j];
    }
  const MT sq = a.r_scale != nullptr ? sqrt((MT)a.r_scale[idx * NB + b]) : MT(1);

  MT mu[P];
  RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = (MT)x0[i];
  // mean history per lane: P values per entry, lane-global index (32 lanes per warp: NB * ldb lanes in all)
  const SchedHist<T, P> hist{stash, (i64)NB * ldb, (i64)blockIdx.x * 32 + lane};

  // ---- forward ----
  auto interr_normals = [&](int n, T (&zc)[JC]) {
    if constexpr (SC::DRAW) {
      if (a.z_interr != nullptr) {
        const T* z = a.z_interr + (idx * a.n_steps + n) * (NB * P) + b * P;
        RD_UNROLL for (int j = 0; j < JC; ++j) zc[j] = z[j];
      } else {
        // the same stream layout as the one-lane-per-theta kernels: normal b*JC + j of the step's vector
        philox_normal_range<T, JC>(a.key0, a.key1, a.particle_offset + idx, n, TAG_INTERR_A, b * JC, zc);
      }
    } else {
      RD_UNROLL for (int j = 0; j < JC; ++j) zc[j] = T(0);
    }
  };
  {
    constexpr int CHF = SS::CHF, FROW = SS::FROW;
    T zc[JC];
    interr_normals(0, zc);
    MT t_next = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, 0, N) : MT(0);
    const int nch = (N + CHF - 1) / CHF;
    sched_stage<T, FROW>(tab, 0, N < CHF ? N : CHF, smem, lane);
    for (int c = 0; c < nch; ++c) {
      const int n_lo = c * CHF, n_hi = (n_lo + CHF) < N ? (n_lo + CHF) : N;
      if (c + 1 < nch) {
        sched_stage<T, FROW>(tab, n_hi, (n_hi + CHF) < N ? CHF : (N - n_hi), smem + ((c + 1) & 1) * CHF * FROW, lane);
        cp_async_wait_1();
      } else {
        cp_async_wait_all();
      }
      __syncwarp();
      const T* frow = smem + (c & 1) * CHF * FROW + b * SC::FWD;
      _Pragma("unroll 1") for (int n = n_lo; n < n_hi; ++n) {
        T zn[JC];
        interr_normals(n + 1 < N ? n + 1 : N - 1, zn);
        const MT t = t_next;
        if (Model::USES_TIME) t_next = step_time<MT>(a.t_min, a.t_max, n + 1, N);
        MT mp[P], xo[JC], x[NB][JC], f[NB][M];
        T r[SC::FWD];
        sched_lds<T, SC::FWD>(frow, r);
        frow += FROW;
        sched_predict_mean<T, P, QK, MT>(Q, mu, mp);
        RD_UNROLL for (int j = 0; j < JC; ++j) {
          MT acc = mp[j];
          if constexpr (SC::DRAW) {
            RD_UNROLL for (int k = 0; k <= j; ++k) acc = rd_fma((MT)r[j * (j + 1) / 2 + k], (MT)zc[k] * sq, acc);
          }
          xo[j] = acc;
        }
        RD_UNROLL for (int c2 = 0; c2 < NB; ++c2)
          RD_UNROLL for (int j = 0; j < JC; ++j) x[c2][j] = __shfl_sync(0xffffffffu, xo[j], c2 * TW + tl);
        Model::template rhs<MT, MT>(q, t, x, f);
        MT fo[M];
        RD_UNROLL for (int rr = 0; rr < M; ++rr) {
          fo[rr] = f[0][rr];
          RD_UNROLL for (int c2 = 1; c2 < NB; ++c2) fo[rr] = (b == c2) ? f[c2][rr] : fo[rr];
        }
        if constexpr (SC::UNITW) {
          const MT res = sub_exact(fo[0], mp[WK]);
          const MT g = res * (MT)r[SC::NFA + P];
          RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = rd_fma((MT)r[SC::NFA + i], g, mp[i]);
        } else {
          MT res[M];
          RD_UNROLL for (int rr = 0; rr < M; ++rr) {
            MT acc = fo[rr];
            RD_UNROLL for (int j = 0; j < P; ++j) acc = rd_fma(-(MT)W[rr][j], mp[j], acc);
            res[rr] = acc;
          }
          RD_UNROLL for (int i = 0; i < P; ++i) {
            MT m = mp[i];
            RD_UNROLL for (int rr = 0; rr < M; ++rr) m = rd_fma((MT)r[SC::NFA + rr * P + i], res[rr], m);
            mu[i] = m;
          }
        }
        if (live && n + 1 < N) hist.store(n, mu);                 // history entry n = mu_f[n+1]
        RD_UNROLL for (int j = 0; j < JC; ++j) zc[j] = zn[j];
      }
      __syncwarp();
    }
  }

  // ---- backward ----
  auto normals = [&](int n, T (&z)[P]) {
    if (z_smooth != nullptr) {
      const T* zp = z_smooth + (idx * (i64)(N + 1) + n) * ROW + b * P;
      RD_UNROLL for (int k = 0; k < P; ++k) z[k] = zp[k];
    } else {
      philox_normals<T, P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_SMOOTH, z, b * ((P + 3) / 4));
    }
  };
  constexpr int BROW = NB * SC::BWD;
  const T* btab = tab + SC::bwd_off(N);
  T* tbuf = smem + SS::OUT_ELEMS;
  T* buf = smem;
  auto stage_seg = [&](int j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    const int lo = (n0 < 1 ? 1 : n0) - 1, hi = n0 + cnt - 2;
    if (hi >= lo) sched_stage<T, BROW>(btab, lo, hi - lo + 1, tbuf + (XOUT ? 0 : (j & 1) * SS::BCH_ELEMS), lane);
    else cp_async_commit();
  };
  const int jtop = (N - 1) / K;
  stage_seg(jtop);
  MT x[P];
  {                                                             // terminal draw from N(mu_f[N], S_f[N])
    T z[P], r[SC::BWD];
    normals(N, z);
    sched_load<T, SC::BWD>(btab + ((i64)(N - 1) * NB + b) * SC::BWD, r);
    RD_UNROLL for (int i = 0; i < P; ++i) {
      MT acc = mu[i];
      RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)r[P * P + lidx(i, k)], (MT)z[k] * sq, acc);
      x[i] = acc;
    }
    if (live && x_out != nullptr) {
      T* dst = x_out + (idx * (i64)(N + 1) + N) * ROW + b * P;
      RD_UNROLL for (int i = 0; i < P; ++i) dst[i] = (T)x[i];
    }
  }
  SimLoglikAcc<T, MT> la;
  la.init(sl, N + 1);
  la.row(sl, N, NB, b, b + 1, [&](int) { return x[0]; });

  T nmu[P];
  auto hload = [&](int n) { hist.load((n < 1 ? 1 : n) - 1, nmu); };
  auto hprefetch = [&](int n) { hist.prefetch((n < 1 ? 1 : n) - 1); };
  T z[P];
  if (N > 1) {
    for (int n = N - 2; n > N - 2 - RODEO_SCHED_PF; --n) hprefetch(n);
    hload(N - 1);
    normals(N - 1, z);
  } else {
    RD_UNROLL for (int k = 0; k < P; ++k) { z[k] = T(0); nmu[k] = T(0); }
  }
  for (int j = jtop; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    if (!XOUT && j > 0) { stage_seg(j - 1); cp_async_wait_1(); }
    else cp_async_wait_all();
    __syncwarp();
    const T* trow0 = tbuf + (XOUT ? 0 : (j & 1) * SS::BCH_ELEMS) - (i64)((n0 < 1 ? 1 : n0) - 1) * BROW + b * SC::BWD;
    _Pragma("unroll 1") for (int s = cnt - 1; s >= 0; --s) {
      const int n = n0 + s;
      if (n == 0) {                                             // row 0 = ode_init: known, not sampled
        if constexpr (XOUT)
          RD_UNROLL for (int k = 0; k < P; ++k) buf[(s * ROW + b * P + k) * PITCH + tl] = x0[k];
        la.row(sl, 0, NB, b, b + 1, [&](int) { return (MT)x0[0]; });
        break;
      }
      MT mf[P];
      RD_UNROLL for (int i = 0; i < P; ++i) mf[i] = (MT)nmu[i];
      hload(n - 1);
      hprefetch(n - 1 - RODEO_SCHED_PF);
      T zn[P];
      normals(n > 1 ? n - 1 : 1, zn);
      T r[SC::BWD];
      sched_lds<T, SC::BWD>(trow0 + (i64)(n - 1) * BROW, r);
      MT mp[P], zs[P], xn[P];
      RD_UNROLL for (int k = 0; k < P; ++k) zs[k] = (MT)z[k] * sq;
      sched_predict_mean<T, P, QK, MT>(Q, mf, mp);
      RD_UNROLL for (int i = 0; i < P; ++i) {
        MT acc = mf[i];
        RD_UNROLL for (int jj = 0; jj < P; ++jj) acc = rd_fma((MT)r[i * P + jj], x[jj] - mp[jj], acc);
        RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)r[P * P + lidx(i, k)], zs[k], acc);
        xn[i] = acc;
      }
      RD_UNROLL for (int i = 0; i < P; ++i) {
        x[i] = xn[i];
        if constexpr (XOUT) buf[(s * ROW + b * P + i) * PITCH + tl] = (T)xn[i];
      }
      RD_UNROLL for (int k = 0; k < P; ++k) z[k] = zn[k];
      la.row(sl, n, NB, b, b + 1, [&](int) { return x[0]; });
    }
    __syncwarp();
    if constexpr (XOUT) {
      if (j > 0) stage_seg(j - 1);
      constexpr int NIT = (K * ROW + 31) / 32;
      const int run = cnt * ROW;
      const i64 stride = (i64)(N + 1) * ROW;
      T* dst = x_out + (theta0 * (i64)(N + 1) + n0) * ROW + lane;
      const int nth = (a.B - theta0) < TW ? (int)(a.B - theta0) : TW;
      RD_UNROLL4 for (int th = 0; th < nth; ++th) {
        RD_UNROLL for (int it = 0; it < NIT; ++it) {
          const int rr = lane + 32 * it;
          if (rr < run) dst[32 * it] = buf[rr * PITCH + th];
        }
        dst += stride;
      }
      __syncwarp();
    }
  }
  if (sl.out != nullptr) {
    // sum the blocks of a theta in block order (lane c * TW + tl holds block c), as solve_sim_bl_kernel does
    MT tot = MT(0);
    RD_UNROLL for (int c = 0; c < NB; ++c) tot += __shfl_sync(0xffffffffu, la.ll, c * TW + tl);
    if (live && b == 0) sl.out[idx] = (T)tot;
  }
}

}  // namespace rodeo
