// rodeo_b200 kernels: the probabilistic-ODE filtering hot path, one thread per theta.
//
// Reference call stacks being replaced:
//   _solve_filter   src/rodeo/solve.py:31-122      forward lax.scan: vmap(predict) -> interrogate -> vmap(update)
//   solve_mv        src/rodeo/solve.py:208-302     + reverse scan of smooth_mv
//   solve_sim       src/rodeo/solve.py:125-205     + reverse scan of smooth_sim and Gaussian draws
//   dalton          src/rodeo/inference/dalton.py:39-235   two forward filters with per-step Gaussian log-pdfs
//   fenrir          src/rodeo/inference/fenrir.py:86-328   forward filter + reverse filter on the backward chain
//
// Execution model: a thread owns one theta (all n_block blocks, because the ODE right-hand side couples the
// blocks' means) and keeps the block means / packed covariances in registers across the whole time loop, which
// runs inside the kernel.  Q, R, W arrive in the kernel-parameter constant bank.  Anything that must outlive the
// forward sweep (the filtered moments the smoothers need) goes to a theta-innermost (coalesced) stash in HBM.
//
// No host / toolkit includes: this header is also fed to NVRTC.
#pragma once
#include "rodeo_core.cuh"
#include "rodeo_models.cuh"


namespace rodeo {

typedef long long i64;

// Launch-invariant arguments common to all ops.
template <typename T>
struct CommonArgs {
  i64 B;                  // thetas in this launch
  i64 particle_offset;    // global index of theta 0 (keeps Philox streams independent of sharding)
  int n_steps;
  T t_min, t_max;
  const T* theta;         // (B, NTHETA)
  const T* ode_init;      // (B, NB, P)
  unsigned key0, key1;    // PRNG key (the reference's uint32[2] jax key)
  const T* z_interr;      // optional injected normals for interrogate_chkrebtii, (B, n_steps, NSTREAM, NB, P)
};

template <typename T>
struct ObsArgs {
  int n_obs;
  const int* obs_ind;     // (n_obs) searchsorted(linspace(t_min,t_max,N+1), obs_times), computed on the host
  const T* obs_data;      // (n_obs, NB, NOBS)
  const T* obs_weight;    // (n_obs, NB, NOBS, P)
  const T* obs_var;       // (n_obs, NB, NOBS, NOBS)
};

// reference src/rodeo/solve.py:74: t = t_min + (t_max - t_min) * (n + 1) / n_steps
template <typename T>
RD_DEV T step_time(T t_min, T t_max, int n, int n_steps) {
  return t_min + (t_max - t_min) * (T)(n + 1) / (T)n_steps;
}

// stream tags for the counter-based RNG
enum : unsigned { TAG_INTERR_A = 0x100u, TAG_INTERR_B = 0x200u, TAG_SMOOTH = 0x300u };

template <typename T, int COUNT>
RD_DEV void philox_normals(unsigned key0, unsigned key1, i64 particle, int step, unsigned tag, T (&z)[COUNT]) {
  Philox ph{key0, key1};
  RD_UNROLL for (int k = 0; k < COUNT; k += 2) {
    unsigned r[4];
    ph((unsigned)particle, (unsigned)((unsigned long long)particle >> 32), (unsigned)step, tag + (unsigned)(k >> 1), r);
    T a, b;
    normal_pair(r, a, b);
    z[k] = a;
    if (k + 1 < COUNT) z[k + 1] = b;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// One theta's filter state and the forward-step pieces
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK>
struct Fwd {
  static constexpr int NB = Model::NB, P = Model::P, M = Model::M, JC = Model::JCOLS;
  static constexpr int NS = P * (P + 1) / 2, MS = M * (M + 1) / 2;
  // QK_UNIT_UPPER is the "structured" instantiation: Q unit upper triangular (every IBM prior) AND, for scalar
  // measurements, W = e_WCOL in every block (what first_order_pad / the docs' higher-order example build).  The host
  // checks both exactly and otherwise dispatches the dense instantiation (rodeo_host.h: detect_structure).
  static constexpr bool UNITW = (QK == QK_UNIT_UPPER) && (M == 1);
  static constexpr int WK = Model::WCOL;
  static constexpr bool HAS_J = (INTERR == INTERR_KRAMER);
  typedef FilterConsts<T, NB, P, M> Consts;
  typedef typename Model::template Par<T> Par;

  T mu[NB][P];
  T S[NB][NS];

  RD_DEV void init(const T* x0) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = x0[b * P + i];
      RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = T(0);
    }
  }

  // (mu, S) filtered at n  ->  predicted at n+1   (reference standard.predict, standard.py:57-59)
  RD_DEV void predict_all(const Consts& C) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T mp[P], Sp[NS];
      predict<T, P, QK>(C.Q[b], C.R[b], mu[b], S[b], mp, Sp);
      RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = mp[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = Sp[k];
    }
  }

  RD_DEV static T Wc(const Consts& C, int b, int r, int j) {
    if (UNITW) return j == WK ? T(1) : T(0);
    return C.W[b][r][j];
  }

  // Interrogation at the predicted moments (reference src/rodeo/interrogate.py:13-115) combined with
  // W_meas = ode_weight + wgt_meas (src/rodeo/solve.py:79).  zc[b][j] are the standard normals consumed by
  // interrogate_chkrebtii; only the JC columns the right-hand side can see are ever needed.
  //
  // Output: jl = the block-diagonal Jacobian (kramer; zero otherwise) so that the measurement rows are
  // wm = W - [jl, 0..]; their noise V; and the update residual res = x_meas - (wm mu_p + mean_meas), x_meas == 0.
  // For every interrogation this is  f - W mu_p  with f evaluated at mu_p (kramer, schober, rodeo) or at the draw
  // (chkrebtii): in interrogate_kramer, mean_meas = -f + J mu_p and wm = W - J, so the two J mu_p terms cancel
  // identically.  Forming the cancelled expression directly drops an O(|J mu_p|) rounding term from a residual of
  // size sqrt(S) ~ 1e-3 and is strictly more accurate than evaluating both terms.
  RD_DEV void interrogate(const Consts& C, const Par& q, T t, const T (&zc)[NB][JC],
                          T (&jl)[NB][M][JC], T (&res)[NB][M], T (&V)[NB][MS]) const {
    T x[NB][JC];
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      if constexpr (INTERR == INTERR_CHKREBTII) {
        // x_b ~ N(mu_p, S_p) by Cholesky (jax.random.multivariate_normal default, interrogate.py:30-34);
        // rows >= JC of the draw are never read by f, so only the leading JC rows of the factor are formed.
        T A[P][P];
        psd_factor<T, P>(S[b], A);
        RD_UNROLL for (int j = 0; j < JC; ++j) {
          T a = mu[b][j];
          RD_UNROLL for (int k = 0; k <= j; ++k) a = rd_fma(A[j][k], zc[b][k], a);
          x[b][j] = a;
        }
      } else {
        RD_UNROLL for (int j = 0; j < JC; ++j) x[b][j] = mu[b][j];
      }
    }
    T f[NB][M];
    if constexpr (HAS_J) {
      eval_f_jac<Model, T>(q, t, x, f, jl);
    } else {
      Model::template rhs<T, T>(q, t, x, f);
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int j = 0; j < JC; ++j) jl[b][r][j] = T(0);
    }
    RD_UNROLL for (int b = 0; b < NB; ++b)
      RD_UNROLL for (int r = 0; r < M; ++r) {
        if (UNITW) {
          res[b][r] = f[b][r] - mu[b][WK];
        } else {
          T a = f[b][r];
          RD_UNROLL for (int j = 0; j < P; ++j) a = rd_fma(-C.W[b][r][j], mu[b][j], a);
          res[b][r] = a;
        }
      }
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      if constexpr (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII) {
        // var_meas = W S_p W^T  (interrogate.py:25-29, 109-112)
        if (UNITW) {
          V[b][0] = S[b][sidx<P>(WK, WK)];
        } else {
          T u[M][P];
          RD_UNROLL for (int r = 0; r < M; ++r)
            RD_UNROLL for (int i = 0; i < P; ++i) {
              T a = S[b][sym<P>(i, 0)] * C.W[b][r][0];
              RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(S[b][sym<P>(i, j)], C.W[b][r][j], a);
              u[r][i] = a;
            }
          RD_UNROLL for (int r = 0; r < M; ++r)
            RD_UNROLL for (int s = r; s < M; ++s) {
              T a = C.W[b][r][0] * u[s][0];
              RD_UNROLL for (int i = 1; i < P; ++i) a = rd_fma(C.W[b][r][i], u[s][i], a);
              V[b][sidx<M>(r, s)] = a;
            }
        }
      } else {
        RD_UNROLL for (int k = 0; k < MS; ++k) V[b][k] = T(0);
      }
    }
  }

  // measurement rows wm = W - [jl, 0..] of block b
  RD_DEV static void rows(const Consts& C, int b, const T (&jl)[M][JC], T (&wm)[M][P]) {
    RD_UNROLL for (int r = 0; r < M; ++r)
      RD_UNROLL for (int j = 0; j < P; ++j) {
        T w = Wc(C, b, r, j);
        wm[r][j] = (HAS_J && j < JC) ? w - jl[r][j] : w;
      }
  }

  // normals for this step's chkrebtii interrogation: injected array or Philox
  template <int NSTREAM>
  RD_DEV void interr_normals(const CommonArgs<T>& a, i64 idx, int n, int stream, T (&zc)[NB][JC]) const {
    if constexpr (INTERR == INTERR_CHKREBTII) {
      if (a.z_interr != nullptr) {
        const T* z = a.z_interr + ((idx * a.n_steps + n) * NSTREAM + stream) * (NB * P);
        RD_UNROLL for (int b = 0; b < NB; ++b)
          RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = z[b * P + j];
      } else {
        T z[NB * JC];
        philox_normals<T, NB * JC>(a.key0, a.key1, a.particle_offset + idx, n,
                                   stream == 0 ? TAG_INTERR_A : TAG_INTERR_B, z);
        RD_UNROLL for (int b = 0; b < NB; ++b)
          RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = z[b * JC + j];
      }
    } else {
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int j = 0; j < JC; ++j) zc[b][j] = T(0);
    }
  }

  // plain ODE-measurement update of every block, x_meas == 0 (solve.py:51, 81-88)
  template <bool WITH_LOGPDF>
  RD_DEV void update_z(const Consts& C, const T (&jl)[NB][M][JC], const T (&res)[NB][M], const T (&V)[NB][MS],
                       LogPdfAcc<T>& acc) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      if constexpr (UNITW) {
        update_unit_row<T, P, JC, WK, WITH_LOGPDF, HAS_J>(mu[b], S[b], jl[b][0], res[b][0], V[b][0], acc);
      } else {
        T wm[M][P];
        rows(C, b, jl[b], wm);
        update<T, P, M, WITH_LOGPDF>(mu[b], S[b], wm, res[b], V[b], acc);
      }
    }
  }

  // observation-augmented update (dalton zy_update, dalton.py:136-149): rows [W~; D_i], offsets [d; 0],
  // noise blockdiag(V, Omega_i), observed value [0; y_i]  ->  residual [res; y_i - D_i mu_p]
  template <int NOBS, bool WITH_LOGPDF>
  RD_DEV void update_zy(const Consts& C, const T (&jl)[NB][M][JC], const T (&res)[NB][M], const T (&V)[NB][MS],
                        const ObsArgs<T>& o, int i, LogPdfAcc<T>& acc) {
    constexpr int MA = M + NOBS, MAS = MA * (MA + 1) / 2;
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T wa[MA][P], ra[MA], Va[MAS], wm[M][P];
      rows(C, b, jl[b], wm);
      RD_UNROLL for (int k = 0; k < MAS; ++k) Va[k] = T(0);
      RD_UNROLL for (int r = 0; r < M; ++r) {
        RD_UNROLL for (int j = 0; j < P; ++j) wa[r][j] = wm[r][j];
        ra[r] = res[b][r];
        RD_UNROLL for (int s = r; s < M; ++s) Va[sidx<MA>(r, s)] = V[b][sidx<M>(r, s)];
      }
      RD_UNROLL for (int r = 0; r < NOBS; ++r) {
        T a = __ldg(o.obs_data + (i * NB + b) * NOBS + r);
        RD_UNROLL for (int j = 0; j < P; ++j) {
          wa[M + r][j] = __ldg(o.obs_weight + ((i * NB + b) * NOBS + r) * P + j);
          a = rd_fma(-wa[M + r][j], mu[b][j], a);
        }
        ra[M + r] = a;
        RD_UNROLL for (int s = r; s < NOBS; ++s)
          Va[sidx<MA>(M + r, M + s)] = __ldg(o.obs_var + ((i * NB + b) * NOBS + r) * NOBS + s);
      }
      update<T, P, MA, WITH_LOGPDF>(mu[b], S[b], wa, ra, Va, acc);
    }
  }

  // pure observation update (fenrir backward pass, fenrir.py:160-179): rows D_i, offset 0, noise Omega_i
  template <int NOBS>
  RD_DEV void update_y(const ObsArgs<T>& o, int i, LogPdfAcc<T>& acc) {
    constexpr int OS = NOBS * (NOBS + 1) / 2;
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T wa[NOBS][P], ra[NOBS], Va[OS];
      RD_UNROLL for (int r = 0; r < NOBS; ++r) {
        T a = __ldg(o.obs_data + (i * NB + b) * NOBS + r);
        RD_UNROLL for (int j = 0; j < P; ++j) {
          wa[r][j] = __ldg(o.obs_weight + ((i * NB + b) * NOBS + r) * P + j);
          a = rd_fma(-wa[r][j], mu[b][j], a);
        }
        ra[r] = a;
        RD_UNROLL for (int s = r; s < NOBS; ++s)
          Va[sidx<NOBS>(r, s)] = __ldg(o.obs_var + ((i * NB + b) * NOBS + r) * NOBS + s);
      }
      update<T, P, NOBS, true>(mu[b], S[b], wa, ra, Va, acc);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------
// dalton: log p(Y | Z) = log p(Z, Y) - log p(Z), two forward filters, scalar output, no history
// ------------------------------------------------------------------------------------------------------------------
// Thread mapping: one thread per (theta, filter); lanes 2k / 2k+1 of a warp run the joint (Z,Y) and the marginal (Z)
// filter of the same theta and meet in a single shuffle at the end.  The two filters execute identical code except
// on the n_obs observation steps, where the joint lanes take the augmented update.  Why not one thread per theta:
// the kernel is FP64-pipe bound per SM sub-partition, so its time is ceil(warps per sub-partition) x (cost of one
// warp); 65,536 thetas are 2,048 two-filter warps = 3.46 per sub-partition (rounds up to 4, 13.5% idle) but 4,096
// one-filter warps of half the cost = 6.92 (rounds up to 7, 1.2% idle).  It also halves the live state per thread,
// so the kernel fits 128 registers (4 resident warps per sub-partition) without spilling.
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32, 16)
dalton_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
              const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ loglik) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, M = F::M, JC = F::JC, MS = F::MS;
  const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool joint = (tid & 1) == 0;
  i64 idx = tid >> 1;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;                 // keep the whole warp in the loop for the final shuffle
  const typename F::Par q = Model::template load<T>(a.theta + idx * Model::NTHETA);
  F f;
  f.init(a.ode_init + idx * NB * P);
  LogPdfAcc<T> acc;
  acc.init();

  // log p(Y_0 | X_0) when the first observation sits on t_min (dalton.py:207-215); joint filter only
  int i = 0;
  if (__ldg(o.obs_ind) == 0) {
    if (joint) {
      constexpr int OS = NOBS * (NOBS + 1) / 2;
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        T res[NOBS], Om[OS];
        RD_UNROLL for (int r = 0; r < NOBS; ++r) {
          T m = T(0);
          RD_UNROLL for (int j = 0; j < P; ++j) m = rd_fma(__ldg(o.obs_weight + (b * NOBS + r) * P + j), f.mu[b][j], m);
          res[r] = __ldg(o.obs_data + b * NOBS + r) - m;
          RD_UNROLL for (int s = r; s < NOBS; ++s) Om[sidx<NOBS>(r, s)] = __ldg(o.obs_var + (b * NOBS + r) * NOBS + s);
        }
        logpdf_terms<T, NOBS>(Om, res, acc);
      }
    }
    i = 1;
  }
  // traced out-of-range gathers clamp: obs_ind[min(i, n_obs-1)]
  int next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));

  for (int n = 0; n < a.n_steps; ++n) {
    const T t = Model::USES_TIME ? step_time<T>(a.t_min, a.t_max, n, a.n_steps) : T(0);
    T jl[NB][M][JC], res[NB][M], V[NB][MS], zc[NB][JC];
    // each filter is linearised at its own prediction (dalton.py:116-132, 168-184)
    f.predict_all(C);
    f.template interr_normals<2>(a, idx, n, joint ? 0 : 1, zc);
    f.interrogate(C, q, t, zc, jl, res, V);
    if (n + 1 == next_obs) {
      const int ic = i < o.n_obs ? i : o.n_obs - 1;
      if (joint) f.template update_zy<NOBS, true>(C, jl, res, V, o, ic, acc);
      else f.template update_z<true>(C, jl, res, V, acc);
      ++i;
      next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));
    } else {
      f.template update_z<true>(C, jl, res, V, acc);
    }
    if ((n & 7) == 7) acc.ld.renorm();
  }
  const T mine = acc.value();
  const T other = __shfl_xor_sync(0xffffffffu, mine, 1);
  if (joint && live) loglik[idx] = mine - other;          // logdens_joint - logdens_marg (dalton.py:235)
}

// ------------------------------------------------------------------------------------------------------------------
// forward filter + stash of the filtered moments, shared by solve_mv / solve_sim / fenrir
// ------------------------------------------------------------------------------------------------------------------
// Stash layout: stash[((n-1) * NSTATE + k) * ldb + idx], n = 1..N-1, k over [mu (NB*P) | S packed (NB*NS)];
// theta innermost, so a warp writes/reads 32 consecutive elements per (n, k).  pred[n+1] is not stored: it is
// recomputed from filt[n] in the backward sweep (one predict), which halves the history traffic.
template <typename T, class F>
RD_DEV void stash_store(T* __restrict__ stash, i64 ldb, i64 idx, int n, const F& f) {
  constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  T* s = stash + (i64)(n - 1) * NSTATE * ldb + idx;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) s[(i64)(b * P + i) * ldb] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) s[(i64)(NB * P + b * NS + k) * ldb] = f.S[b][k];
  }
}
template <typename T, class F>
RD_DEV void stash_load(const T* __restrict__ stash, i64 ldb, i64 idx, int n, F& f) {
  constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  const T* s = stash + (i64)(n - 1) * NSTATE * ldb + idx;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) f.mu[b][i] = s[(i64)(b * P + i) * ldb];
    RD_UNROLL for (int k = 0; k < NS; ++k) f.S[b][k] = s[(i64)(NB * P + b * NS + k) * ldb];
  }
}

template <typename T, class Model, int INTERR, int QK>
RD_DEV void forward_and_stash(const FilterConsts<T, Model::NB, Model::P, Model::M>& C, const CommonArgs<T>& a,
                              const typename Model::template Par<T>& q, i64 idx,
                              Fwd<T, Model, INTERR, QK>& f, T* __restrict__ stash, i64 ldb) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, M = F::M, JC = F::JC, MS = F::MS;
  LogPdfAcc<T> dummy;
  dummy.init();
  for (int n = 0; n < a.n_steps; ++n) {
    const T t = Model::USES_TIME ? step_time<T>(a.t_min, a.t_max, n, a.n_steps) : T(0);
    T jl[NB][M][JC], res[NB][M], V[NB][MS], zc[NB][JC];
    f.predict_all(C);
    f.template interr_normals<1>(a, idx, n, 0, zc);
    f.interrogate(C, q, t, zc, jl, res, V);
    f.template update_z<false>(C, jl, res, V, dummy);
    if (n + 1 < a.n_steps) stash_store<T, F>(stash, ldb, idx, n + 1, f);
  }
}

// full-matrix / vector stores of one time row of the outputs (reference layouts (N+1, nb, p) and (N+1, nb, p, p))
template <typename T, int NB, int P>
RD_DEV void store_mean_row(T* __restrict__ out, const T (&mu)[NB][P]) {
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i) out[b * P + i] = mu[b][i];
}
template <typename T, int NB, int P>
RD_DEV void store_var_row(T* __restrict__ out, const T (&S)[NB][P * (P + 1) / 2]) {
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i)
      RD_UNROLL for (int j = 0; j < P; ++j) out[(b * P + i) * P + j] = S[b][sym<P>(i, j)];
}

// ------------------------------------------------------------------------------------------------------------------
// solve_mv: forward filter, then the mean/variance smoother  (reference src/rodeo/solve.py:208-302)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK>
__global__ void __launch_bounds__(32)
solve_mv_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                const CommonArgs<T> a, T* __restrict__ stash, i64 ldb,
                T* __restrict__ mean_out, T* __restrict__ var_out) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, NS = F::NS;
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B) return;
  const typename F::Par q = Model::template load<T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  F f;
  f.init(a.ode_init + idx * NB * P);
  T* mrow = mean_out + idx * (i64)(N + 1) * (NB * P);
  T* vrow = var_out + idx * (i64)(N + 1) * (NB * P * P);
  // row 0 is (ode_init, 0) verbatim: x0 is known and never smoothed (solve.py:295-301)
  store_mean_row<T, NB, P>(mrow, f.mu);
  store_var_row<T, NB, P>(vrow, f.S);

  forward_and_stash<T, Model, INTERR, QK>(C, a, q, idx, f, stash, ldb);

  // smoothed[N] = filt[N]
  T ms[NB][P], Ss[NB][NS];
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) ms[b][i] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) Ss[b][k] = f.S[b][k];
  }
  store_mean_row<T, NB, P>(mrow + (i64)N * (NB * P), ms);
  store_var_row<T, NB, P>(vrow + (i64)N * (NB * P * P), Ss);

  for (int n = N - 1; n >= 1; --n) {
    stash_load<T, F>(stash, ldb, idx, n, f);     // filt[n]
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T mp[P], Sp[NS], G[P][P], Ct[P][P];
      predict<T, P, QK>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Sp);          // pred[n+1]
      smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
      // mu_s = mu_f + G (mu_s' - mu_p) ;  S_s = S_f + G (S_s' - S_p) G^T    (standard.py:213-216)
      T dm[P], D[NS];
      RD_UNROLL for (int i = 0; i < P; ++i) dm[i] = ms[b][i] - mp[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) D[k] = Ss[b][k] - Sp[k];
      RD_UNROLL for (int i = 0; i < P; ++i) {
        T m = f.mu[b][i];
        RD_UNROLL for (int j = 0; j < P; ++j) m = rd_fma(G[i][j], dm[j], m);
        ms[b][i] = m;
      }
      RD_UNROLL for (int k = 0; k < NS; ++k) Ss[b][k] = f.S[b][k];
      add_GDGt<T, P>(G, D, Ss[b]);
    }
    store_mean_row<T, NB, P>(mrow + (i64)n * (NB * P), ms);
    store_var_row<T, NB, P>(vrow + (i64)n * (NB * P * P), Ss);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// solve_sim: forward filter, then the sampling smoother  (reference src/rodeo/solve.py:125-205)
// ------------------------------------------------------------------------------------------------------------------
// z_smooth: optional injected normals (B, N+1, NB, P); row N feeds the terminal draw, rows 1..N-1 the backward
// draws.  Without it the draws come from Philox keyed by (key, particle, step).
template <typename T, class Model, int INTERR, int QK>
__global__ void __launch_bounds__(32)
solve_sim_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                 const CommonArgs<T> a, const T* __restrict__ z_smooth, T* __restrict__ stash, i64 ldb,
                 T* __restrict__ x_out) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, NS = F::NS;
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B) return;
  const typename F::Par q = Model::template load<T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  F f;
  f.init(a.ode_init + idx * NB * P);
  T* xrow = x_out + idx * (i64)(N + 1) * (NB * P);
  store_mean_row<T, NB, P>(xrow, f.mu);

  forward_and_stash<T, Model, INTERR, QK>(C, a, q, idx, f, stash, ldb);

  T x[NB][P];
  for (int n = N; n >= 1; --n) {
    T z[NB * P];
    if (z_smooth != nullptr) {
      const T* zp = z_smooth + (idx * (i64)(N + 1) + n) * (NB * P);
      RD_UNROLL for (int k = 0; k < NB * P; ++k) z[k] = zp[k];
    } else {
      philox_normals<T, NB * P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_SMOOTH, z);
    }
    if (n < N) stash_load<T, F>(stash, ldb, idx, n, f);     // filt[n]
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T m[P], Cv[NS];
      if (n == N) {
        // terminal draw from N(mu_f[N], S_f[N])  (solve.py:182-186)
        RD_UNROLL for (int i = 0; i < P; ++i) m[i] = f.mu[b][i];
        RD_UNROLL for (int k = 0; k < NS; ++k) Cv[k] = f.S[b][k];
      } else {
        T mp[P], Sp[NS], G[P][P], Ct[P][P];
        predict<T, P, QK>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Sp);        // pred[n+1]
        smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
        // m = mu_f + G (x' - mu_p) ;  C = S_f - G (S_f Q^T)^T      (standard.py:251-254)
        RD_UNROLL for (int i = 0; i < P; ++i) {
          T acc = f.mu[b][i];
          RD_UNROLL for (int j = 0; j < P; ++j) acc = rd_fma(G[i][j], x[b][j] - mp[j], acc);
          m[i] = acc;
        }
        cond_var<T, P>(f.S[b], G, Ct, Cv);
      }
      T A[P][P];
      psd_factor<T, P>(Cv, A);
      RD_UNROLL for (int i = 0; i < P; ++i) {
        T acc = m[i];
        RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma(A[i][k], z[b * P + k], acc);
        x[b][i] = acc;
      }
    }
    store_mean_row<T, NB, P>(xrow + (i64)n * (NB * P), x);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fenrir: forward filter, then a Kalman filter on the backward Markov chain with the Gaussian observations
// (reference src/rodeo/inference/fenrir.py:86-328)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32)
fenrir_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
              const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ stash, i64 ldb,
              T* __restrict__ loglik) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, NS = F::NS;
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B) return;
  const typename F::Par q = Model::template load<T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  F f;
  f.init(a.ode_init + idx * NB * P);
  forward_and_stash<T, Model, INTERR, QK>(C, a, q, idx, f, stash, ldb);

  // backward-filter state starts at filt[N]
  F bk;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) bk.mu[b][i] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) bk.S[b][k] = f.S[b][k];
  }
  LogPdfAcc<T> acc;
  acc.init();
  int i = o.n_obs - 1;
  // negative traced indices wrap NumPy-style (fenrir.py:178)
  auto obs_at = [&](int k) { return __ldg(o.obs_ind + (k < 0 ? k + o.n_obs : k)); };
  if (obs_at(i) >= N) {                                   // terminal point update (fenrir.py:196-220)
    bk.template update_y<NOBS>(o, i, acc);
    --i;
  }
  int next_obs = obs_at(i);
  int cnt = 0;
  for (int t = N - 1; t >= 0; --t) {
    if (t >= 1) stash_load<T, F>(stash, ldb, idx, t, f);  // filt[t]
    else f.init(a.ode_init + idx * NB * P);               // filt[0] = (ode_init, 0)
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T mp[P], Sp[NS], G[P][P], Ct[P][P], Cv[NS];
      predict<T, P, QK>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Sp);          // pred[t+1]
      smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
      cond_var<T, P>(f.S[b], G, Ct, Cv);
      // backward chain X_t = A X_{t+1} + bvec + N(0, Cv), A = G, bvec = mu_f - G mu_p   (standard.py:366-370)
      // predict the backward filter through it (fenrir.py:151-157)
      T nm[P];
      RD_UNROLL for (int r = 0; r < P; ++r) {
        T acc2 = f.mu[b][r];
        RD_UNROLL for (int j = 0; j < P; ++j) acc2 = rd_fma(G[r][j], bk.mu[b][j] - mp[j], acc2);
        nm[r] = acc2;
      }
      add_GDGt<T, P>(G, bk.S[b], Cv);
      RD_UNROLL for (int r = 0; r < P; ++r) bk.mu[b][r] = nm[r];
      RD_UNROLL for (int k = 0; k < NS; ++k) bk.S[b][k] = Cv[k];
    }
    if (next_obs == t) {
      bk.template update_y<NOBS>(o, i < 0 ? i + o.n_obs : i, acc);
      --i;
      next_obs = obs_at(i);
    }
    if ((++cnt & 7) == 0) acc.ld.renorm();
  }
  loglik[idx] = acc.value();
}

// ------------------------------------------------------------------------------------------------------------------
// first_order_pad initial value: X0 = [x0, f(x0, t, theta), 0, ...]   (reference src/rodeo/utils.py:94-96)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model>
__global__ void ode_init_pad_kernel(i64 B, T t, const T* __restrict__ theta, const T* __restrict__ x0,
                                    T* __restrict__ X0) {
  constexpr int NB = Model::NB, P = Model::P, M = Model::M, JC = Model::JCOLS;
  const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B) return;
  const typename Model::template Par<T> q = Model::template load<T>(theta + idx * Model::NTHETA);
  T x[NB][JC], f[NB][M];
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int j = 0; j < JC; ++j) x[b][j] = (j == 0) ? x0[idx * NB + b] : T(0);
  Model::template rhs<T, T>(q, t, x, f);
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i)
      X0[(idx * NB + b) * P + i] = (i == 0) ? x[b][0] : (i == 1 ? f[b][0] : T(0));
}

}  // namespace rodeo
