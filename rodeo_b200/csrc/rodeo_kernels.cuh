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


#ifndef RODEO_SIM_BL_MINB
#define RODEO_SIM_BL_MINB 14
#endif

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
  const T* r_scale;       // optional per-theta scale of the prior variance, (B, NB): R(theta, b) = r_scale * R[b]
  int dalton_geometry;    // dalton_kernel only: how the joint and marginal filters of a theta are laid out (see there)
  const T* q_batch;       // QK_DENSE_BATCH only: per-theta prior weight / variance, (B, NB, P, P) each -- the reference
  const T* r_batch;       //   takes any prior_pars per theta under vmap (docs/examples/parameter.md:218-236)
};

// The filter constants a thread works with: the kernel parameter itself, or (QK_DENSE_BATCH) a register copy whose Q, R
// are this theta's own rows of the batched prior.
template <typename T, int NB, int P, int M, int QK>
struct PriorConsts {
  const FilterConsts<T, NB, P, M>& C;
  RD_DEV PriorConsts(const FilterConsts<T, NB, P, M>& Cp, const CommonArgs<T>&, i64) : C(Cp) {}
};
template <typename T, int NB, int P, int M>
struct PriorConsts<T, NB, P, M, QK_DENSE_BATCH> {
  FilterConsts<T, NB, P, M> C;
  RD_DEV PriorConsts(const FilterConsts<T, NB, P, M>& Cp, const CommonArgs<T>& a, i64 idx) {
    const T* q = a.q_batch + idx * (NB * P * P);
    const T* r = a.r_batch + idx * (NB * P * P);
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i)
        RD_UNROLL for (int j = 0; j < P; ++j) {
          C.Q[b][i][j] = q[(b * P + i) * P + j];
          if (j >= i) C.R[b][sidx<P>(i, j)] = r[(b * P + i) * P + j];
        }
      RD_UNROLL for (int m = 0; m < M; ++m)
        RD_UNROLL for (int j = 0; j < P; ++j) C.W[b][m][j] = Cp.W[b][m][j];
    }
  }
};

template <typename T>
struct ObsArgs {
  int n_obs;
  const int* obs_ind;     // (n_obs) searchsorted(linspace(t_min,t_max,N+1), obs_times), computed on the host
  const T* obs_data;      // (n_obs, NB, NOBS)
  const T* obs_weight;    // (n_obs, NB, NOBS, P)
  const T* obs_var;       // (n_obs, NB, NOBS, NOBS)
};

// Data-adaptive solvers (reference src/rodeo/inference/dalton.py:242-545): the forward filter of solve_mv / solve_sim
// additionally conditions on the Gaussian observations, i.e. step n uses the augmented update of dalton.zy_update
// when step_obs[n] >= 0 (the observation index the reference's pointer walk would use at that step, precomputed by
// obs_step_map_kernel because the backward sweeps re-run forward steps out of order).
template <typename T>
struct ObsHook {
  ObsArgs<T> o;
  const int* step_obs;   // (n_steps) or nullptr
};

// Fused Gaussian observation log-likelihood of a solve_sim draw (the inner call of a pseudo-marginal MCMC step,
// reference docs/examples/parameter.md:333-354: solve_sim, then sum_i,k log N(obs_data[i, k]; Xt[obs_ind[i], k, 0], sd^2)):
// accumulated while the backward sweep produces the draw, so that Xt need not be written at all (x_out == nullptr).
// obs_ind must be non-decreasing (searchsorted of sorted observation times); out-of-range indices clamp like the
// reference's traced gather.
template <typename T>
struct SimLoglik {
  const int* obs_ind;     // (n_obs)
  const T* obs_data;      // (n_obs, NB)
  int n_obs;
  T noise_sd;
  T* out;                 // (B) or nullptr: no log-likelihood
};
// walks the observations backwards in time alongside the sweep
template <typename T, typename MT>
struct SimLoglikAcc {
  MT ll, rvar, cst;
  int oi, next_t, n_rows;
  RD_DEV void init(const SimLoglik<T>& sl, int n_rows_) {
    ll = MT(0); n_rows = n_rows_; oi = sl.out != nullptr ? sl.n_obs - 1 : -1;
    rvar = MT(1) / ((MT)sl.noise_sd * (MT)sl.noise_sd);
    cst = MT(-0.5) * MT(1.8378770664093454836) - log((MT)sl.noise_sd);
    next_t = oi >= 0 ? clampi(__ldg(sl.obs_ind + oi)) : -1;
  }
  RD_DEV int clampi(int r) const { return r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r); }
  // row t of the draw is known: xv(b) returns x_t[b][0]; blocks b0 .. b1-1 are this thread's
  template <class XV>
  RD_DEV void row(const SimLoglik<T>& sl, int t, int nb, int b0, int b1, XV xv) {
    while (next_t == t) {
      for (int b = b0; b < b1; ++b) {
        const MT d = (MT)__ldg(sl.obs_data + oi * nb + b) - xv(b);
        ll += cst - MT(0.5) * d * d * rvar;
      }
      --oi;
      next_t = oi >= 0 ? clampi(__ldg(sl.obs_ind + oi)) : -1;
    }
  }
};

// step_obs[n] = i if the reference's scan (dalton.py:313-350) applies observation i at step n (t+1 == obs_ind[i],
// i clamped, starting at 1 when obs_ind[0] == 0), else -1
template <int UNUSED = 0>
__global__ void obs_step_map_kernel(int n_steps, int n_obs, const int* __restrict__ obs_ind, int* __restrict__ map) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int i = (obs_ind[0] == 0) ? 1 : 0;
  for (int n = 0; n < n_steps; ++n) {
    const int ic = i < n_obs ? i : n_obs - 1;
    if (n + 1 == obs_ind[ic]) { map[n] = ic; ++i; } else map[n] = -1;
  }
}

// the model parameters in the mean type (double), from a theta row stored as T
template <class Model, typename T>
RD_DEV typename Model::template Par<typename MeanOf<T>::type> load_par(const T* th) {
  typedef typename MeanOf<T>::type MT;
  MT tmp[Model::NTHETA > 0 ? Model::NTHETA : 1];
  RD_UNROLL for (int k = 0; k < Model::NTHETA; ++k) tmp[k] = (MT)th[k];
  return Model::template load<MT>(tmp);
}

// reference src/rodeo/solve.py:74: t = t_min + (t_max - t_min) * (n + 1) / n_steps
template <typename T>
RD_DEV T step_time(T t_min, T t_max, int n, int n_steps) {
  return t_min + (t_max - t_min) * (T)(n + 1) / (T)n_steps;
}

// stream tags for the counter-based RNG
enum : unsigned { TAG_INTERR_A = 0x100u, TAG_INTERR_B = 0x200u, TAG_SMOOTH = 0x300u };

// Random-number layout: a stream is (key, particle, step, tag); Philox call c of a stream yields its normals 4c .. 4c+3.
//   interrogation streams (TAG_INTERR_*): normal b * JC + j belongs to block b, visible column j;
//   smoothing stream (TAG_SMOOTH):        block b owns calls b * CPB .. (b+1) * CPB - 1, CPB = ceil(P / 4), so that a
//                                         lane that holds one block generates exactly its own calls.
// Both lane mappings (one lane per theta, one lane per (theta, block)) read this layout, so a draw does not depend on
// which kernel produced it.
RD_DEV void philox_call(unsigned key0, unsigned key1, i64 particle, int step, unsigned tag, int call, unsigned (&r)[4]) {
  Philox ph{key0, key1};
  ph((unsigned)particle, (unsigned)((unsigned long long)particle >> 32), (unsigned)step, tag + (unsigned)call, r);
}

// normals 4 * call0 .. 4 * call0 + COUNT - 1 of a stream
template <typename T, int COUNT>
RD_DEV void philox_normals(unsigned key0, unsigned key1, i64 particle, int step, unsigned tag, T (&z)[COUNT],
                           int call0 = 0) {
  RD_UNROLL for (int c = 0; c < (COUNT + 3) / 4; ++c) {
    unsigned r[4];
    philox_call(key0, key1, particle, step, tag, call0 + c, r);
    T q[4];
    normal_quad(r, q);
    RD_UNROLL for (int i = 0; i < 4; ++i)
      if (4 * c + i < COUNT) z[4 * c + i] = q[i];
  }
}

// normals first .. first+COUNT-1 of a stream, `first` a run-time value (selects instead of dynamic indexing)
template <typename T, int COUNT>
RD_DEV void philox_normal_range(unsigned key0, unsigned key1, i64 particle, int step, unsigned tag, int first,
                                T (&z)[COUNT]) {
  const int c0 = first >> 2, o = first & 3;
  if constexpr (COUNT == 1) {
    // one normal: pick its pair's two words first, so that a single Box-Muller pair is evaluated
    unsigned r[4];
    philox_call(key0, key1, particle, step, tag, c0, r);
    const unsigned ru = (o & 2) ? r[2] : r[0], ra = (o & 2) ? r[3] : r[1];
    T a, b;
    normal_pair(ru, ra, a, b);
    z[0] = (o & 1) ? b : a;
  } else {
    constexpr int NCALL = (COUNT + 2) / 4 + 1;
    T pr[4 * NCALL];
    RD_UNROLL for (int k = 0; k < NCALL; ++k) {
      unsigned r[4];
      philox_call(key0, key1, particle, step, tag, c0 + k, r);
      T q[4];
      normal_quad(r, q);
      RD_UNROLL for (int i = 0; i < 4; ++i) pr[4 * k + i] = q[i];
    }
    RD_UNROLL for (int k = 0; k < COUNT; ++k)
      z[k] = o == 0 ? pr[k] : (o == 1 ? pr[k + 1] : (o == 2 ? pr[k + 2] : pr[k + 3]));
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
  typedef typename MeanOf<T>::type MT;      // means, ODE evaluation and residuals: always double (rodeo_core.cuh)
  typedef FilterConsts<T, NB, P, M> Consts;
  typedef typename Model::template Par<MT> Par;

  MT mu[NB][P];
  T S[NB][NS];
  T rs[NB];               // per-theta prior-variance scale (1 unless the prior is batched)

  RD_DEV void load_scale(const CommonArgs<T>& a, i64 idx) {
    RD_UNROLL for (int b = 0; b < NB; ++b) rs[b] = a.r_scale != nullptr ? a.r_scale[idx * NB + b] : T(1);
  }

  RD_DEV void init(const T* x0) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = (MT)x0[b * P + i];
      RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = T(0);
    }
  }

  // (mu, S) filtered at n  ->  predicted at n+1   (reference standard.predict, standard.py:57-59)
  RD_DEV void predict_all(const Consts& C) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      MT mp[P];
      T Sp[NS];
      predict<T, P, QK>(C.Q[b], C.R[b], rs[b], mu[b], S[b], mp, Sp);
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
  RD_DEV void interrogate(const Consts& C, const Par& q, MT t, const T (&zc)[NB][JC],
                          T (&jl)[NB][M][JC], MT (&res)[NB][M], T (&V)[NB][MS], const MT* frc = nullptr) const {
    MT x[NB][JC];
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      if constexpr (INTERR == INTERR_CHKREBTII) {
        // x_b ~ N(mu_p, S_p) by Cholesky (jax.random.multivariate_normal default, interrogate.py:30-34);
        // rows >= JC of the draw are never read by f, so only the leading JC rows of the factor are formed.
        T A[P][P];
        psd_factor<T, P>(S[b], A);
        RD_UNROLL for (int j = 0; j < JC; ++j) {
          MT a = mu[b][j];
          RD_UNROLL for (int k = 0; k <= j; ++k) a = rd_fma((MT)A[j][k], (MT)zc[b][k], a);
          x[b][j] = a;
        }
      } else {
        RD_UNROLL for (int j = 0; j < JC; ++j) x[b][j] = mu[b][j];
      }
    }
    MT f[NB][M];
    if constexpr (HAS_J) {
      MT jd[NB][M][JC];
      eval_f_jac<Model, MT>(q, t, x, f, jd, frc);
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int j = 0; j < JC; ++j) jl[b][r][j] = (T)jd[b][r][j];
    } else {
      eval_f<Model, MT>(q, t, x, f, frc);
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int j = 0; j < JC; ++j) jl[b][r][j] = T(0);
    }
    RD_UNROLL for (int b = 0; b < NB; ++b)
      RD_UNROLL for (int r = 0; r < M; ++r) {
        if (UNITW) {
          res[b][r] = sub_exact(f[b][r], mu[b][WK]);
        } else {
          MT a = f[b][r];
          RD_UNROLL for (int j = 0; j < P; ++j) a = rd_fma(-(MT)C.W[b][r][j], mu[b][j], a);
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
  // `acc` is one accumulator for all blocks (LogPdfAcc) or per-block partial sums (LogPdfParts): see acc_at
  template <bool WITH_LOGPDF, class ACCS>
  RD_DEV void update_z(const Consts& C, const T (&jl)[NB][M][JC], const MT (&res)[NB][M], const T (&V)[NB][MS],
                       ACCS& acc) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      auto&& ab = acc_at(acc, b);
      if constexpr (UNITW) {
        update_unit_row<T, P, JC, WK, WITH_LOGPDF, HAS_J, (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII)>(
            mu[b], S[b], jl[b][0], res[b][0], V[b][0], ab);
      } else {
        T wm[M][P];
        rows(C, b, jl[b], wm);
        update<T, P, M, WITH_LOGPDF>(mu[b], S[b], wm, res[b], V[b], ab);
      }
    }
  }

  // observation-augmented update (dalton zy_update, dalton.py:136-149): rows [W~; D_i], offsets [d; 0],
  // noise blockdiag(V, Omega_i), observed value [0; y_i]  ->  residual [res; y_i - D_i mu_p]
  template <int NOBS, bool WITH_LOGPDF, class ACCS>
  RD_DEV void update_zy(const Consts& C, const T (&jl)[NB][M][JC], const MT (&res)[NB][M], const T (&V)[NB][MS],
                        const ObsArgs<T>& o, int i, ACCS& acc) {
    constexpr int MA = M + NOBS, MAS = MA * (MA + 1) / 2;
    constexpr bool HAS_V = (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII);
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      auto&& ab = acc_at(acc, b);
      if constexpr (M == 1 && NOBS == 1 && sizeof(T) == 8) {
        // one ODE row + one observation row with uncorrelated noises: two scalar updates (rodeo_core.cuh).  float64
        // only: in float32 the Schur complement formed through the downdated covariance costs accuracy the float32
        // instantiation does not have to spare (dalton 1.2e-5 against 7.5e-6 with the stacked update; gate 1e-5)
        T D[P];
        const MT ry = (MT)__ldg(o.obs_data + (i * NB + b));
        RD_UNROLL for (int j = 0; j < P; ++j) D[j] = __ldg(o.obs_weight + (i * NB + b) * P + j);
        const T Om = __ldg(o.obs_var + (i * NB + b));
        const DenseRow<T, P> row2{D};
        if constexpr (UNITW) {
          const UnitRow<T, P, JC, WK, HAS_J> row1{jl[b][0]};
          update_two_rows<T, P, WITH_LOGPDF, HAS_V>(mu[b], S[b], row1, res[b][0], V[b][0], row2, ry, Om, ab);
        } else {
          T wm[M][P];
          rows(C, b, jl[b], wm);
          const DenseRow<T, P> row1{wm[0]};
          update_two_rows<T, P, WITH_LOGPDF, HAS_V>(mu[b], S[b], row1, res[b][0], V[b][0], row2, ry, Om, ab);
        }
      } else {
        T wa[MA][P], Va[MAS], wm[M][P];
        MT ra[MA];
        rows(C, b, jl[b], wm);
        RD_UNROLL for (int k = 0; k < MAS; ++k) Va[k] = T(0);
        RD_UNROLL for (int r = 0; r < M; ++r) {
          RD_UNROLL for (int j = 0; j < P; ++j) wa[r][j] = wm[r][j];
          ra[r] = res[b][r];
          RD_UNROLL for (int s = r; s < M; ++s) Va[sidx<MA>(r, s)] = V[b][sidx<M>(r, s)];
        }
        RD_UNROLL for (int r = 0; r < NOBS; ++r) {
          MT a = (MT)__ldg(o.obs_data + (i * NB + b) * NOBS + r);
          RD_UNROLL for (int j = 0; j < P; ++j) {
            wa[M + r][j] = __ldg(o.obs_weight + ((i * NB + b) * NOBS + r) * P + j);
            a = rd_fma(-(MT)wa[M + r][j], mu[b][j], a);
          }
          ra[M + r] = a;
          RD_UNROLL for (int s = r; s < NOBS; ++s)
            Va[sidx<MA>(M + r, M + s)] = __ldg(o.obs_var + ((i * NB + b) * NOBS + r) * NOBS + s);
        }
        update<T, P, MA, WITH_LOGPDF>(mu[b], S[b], wa, ra, Va, ab);
      }
    }
  }

  // pure observation update (fenrir backward pass, fenrir.py:160-179): rows D_i, offset 0, noise Omega_i
  template <int NOBS>
  RD_DEV void update_y(const ObsArgs<T>& o, int i, LogPdfAcc<T>& acc) {
    constexpr int OS = NOBS * (NOBS + 1) / 2;
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T wa[NOBS][P], Va[OS];
      MT ra[NOBS];
      RD_UNROLL for (int r = 0; r < NOBS; ++r) {
        MT a = (MT)__ldg(o.obs_data + (i * NB + b) * NOBS + r);
        RD_UNROLL for (int j = 0; j < P; ++j) {
          wa[r][j] = __ldg(o.obs_weight + ((i * NB + b) * NOBS + r) * P + j);
          a = rd_fma(-(MT)wa[r][j], mu[b][j], a);
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
#ifndef RODEO_DALTON_MINB
#define RODEO_DALTON_MINB 16
#endif
// Three launch geometries (CommonArgs::dalton_geometry).  The filters differ only on the n_obs observation steps, where
// the joint one takes the augmented (ODE row + observation row) update, about three times the plain one; with both kinds
// interleaved in a warp every observation step executes BOTH paths for all 32 lanes (8 % more FP64 instructions).
//   0  32-thread CTAs, lanes 2k / 2k+1 are the two filters of a theta and meet in a shuffle (the NVRTC launcher);
//   1  64-thread CTAs, warp 0 = joint filters of 32 thetas, warp 1 = their marginal filters, meeting in shared memory
//      (float32 output: the difference must be formed in double before it is rounded).  The lighter marginal warp idles
//      until its partner is done: 0.845 ms on BASELINE configs[1];
//   2  32-thread CTAs, the first half of the grid runs joint filters, the second half marginal ones (heavier CTAs first),
//      and each adds +/- its log-density to the zero-initialised output with one atomicAdd per theta.  0 + a - b and
//      0 - b + a are both exactly fl(a - b), so the result is bitwise the one of the other geometries (float64 output).
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(64, RODEO_DALTON_MINB / 2)
dalton_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> Cg,
              const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ loglik) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, M = F::M, JC = F::JC, MS = F::MS;
  __shared__ double marg[32];
  const int geo = a.dalton_geometry;
  const int lane = threadIdx.x & 31;
  const i64 half = gridDim.x >> 1;
  const i64 tid = (i64)(geo == 2 && (i64)blockIdx.x >= half ? (i64)blockIdx.x - half : (i64)blockIdx.x) * 32 + lane;
  const bool joint = geo == 2 ? (i64)blockIdx.x < half : (geo == 1 ? threadIdx.x < 32 : (tid & 1) == 0);
  i64 idx = geo == 0 ? (tid >> 1) : tid;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;                 // keep the whole warp in the loop for the final shuffle
  const PriorConsts<T, Model::NB, Model::P, Model::M, QK> pc(Cg, a, idx);
  const FilterConsts<T, Model::NB, Model::P, Model::M>& C = pc.C;
  typedef typename F::MT MT;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  F f;
  f.init(a.ode_init + idx * NB * P);
  f.load_scale(a, idx);
  LogPdfParts<T, NB> acc;      // per-block partial sums: bitwise the result of dalton_bl_kernel (rodeo_core.cuh)
  acc.init();

  // log p(Y_0 | X_0) when the first observation sits on t_min (dalton.py:207-215); joint filter only
  int i = 0;
  if (__ldg(o.obs_ind) == 0) {
    if (joint) {
      constexpr int OS = NOBS * (NOBS + 1) / 2;
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        T res[NOBS], Om[OS];
        RD_UNROLL for (int r = 0; r < NOBS; ++r) {
          MT m = MT(0);
          RD_UNROLL for (int j = 0; j < P; ++j) m = rd_fma((MT)__ldg(o.obs_weight + (b * NOBS + r) * P + j), f.mu[b][j], m);
          res[r] = (T)((MT)__ldg(o.obs_data + b * NOBS + r) - m);
          RD_UNROLL for (int s = r; s < NOBS; ++s) Om[sidx<NOBS>(r, s)] = __ldg(o.obs_var + (b * NOBS + r) * NOBS + s);
        }
        auto&& ab = acc_at(acc, b);
        logpdf_terms<T, NOBS>(Om, res, ab);
      }
    }
    i = 1;
  }
  // traced out-of-range gathers clamp: obs_ind[min(i, n_obs-1)]
  int next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));

  // One step n -> n+1; io >= 0: the joint filter's step that also conditions on observation io (dalton.py:136-149).
  // Each filter is linearised at its own prediction (dalton.py:116-132, 168-184).
  auto step = [&](int n, int io) {
    const MT t = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, n, a.n_steps) : MT(0);
    T jl[NB][M][JC], V[NB][MS], zc[NB][JC];
    MT res[NB][M];
    f.predict_all(C);
    f.template interr_normals<2>(a, idx, n, joint ? 0 : 1, zc);
    f.interrogate(C, q, t, zc, jl, res, V);
    if (io >= 0) f.template update_zy<NOBS, true>(C, jl, res, V, o, io, acc);
    else f.template update_z<true>(C, jl, res, V, acc);
    if ((n & 3) == 3) acc.renorm();       // fold the exponents of the running products every 4th step
  };
  // The reference's scan tests t + 1 == obs_ind[i] at every step (dalton.py:163).  Here the steps between two
  // observations run in an inner loop without that test (the loop body is then branch-free: no observation bookkeeping,
  // no reconvergence point on the dependency chain of a lone warp).  Same semantics, including the corner cases: an
  // observation index that is not ahead of the current step (duplicates, obs_ind[0] == 0 consumed above, the clamped
  // index after the last observation) never matches again.
  int n = 0;
  const int N = a.n_steps;
  for (;;) {
    const int stop = next_obs - 1;                         // the step with n + 1 == next_obs
    const bool has_obs = stop >= n && stop < N;
    const int end = has_obs ? stop : N;
    for (; n < end; ++n) step(n, -1);
    if (!has_obs) break;
    step(n, joint ? (i < o.n_obs ? i : o.n_obs - 1) : -1);
    ++n; ++i;
    next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));
  }
  const MT mine = acc.value();
  // logdens_joint - logdens_marg (dalton.py:235)
  if constexpr (sizeof(T) == 8) {
    if (geo == 2) {
      if (live) atomicAdd(loglik + idx, joint ? mine : -mine);
      return;
    }
  }
  MT other;
  if (geo == 1) {
    if (!joint) marg[lane] = mine;
    __syncthreads();
    other = marg[lane];
  } else {
    other = __shfl_xor_sync(0xffffffffu, mine, 1);
  }
  if (joint && live) loglik[idx] = (T)(mine - other);
}

// ------------------------------------------------------------------------------------------------------------------
// History for the backward sweeps (solve_mv / solve_sim / fenrir): checkpoints in HBM, segments in shared memory
// ------------------------------------------------------------------------------------------------------------------
// The smoothers need filt[n] for n = N-1 .. 0 (pred[n+1] is one predict away from filt[n] and is never stored).
// Writing every filt[n] to HBM and reading it back costs 2 x NSTATE x 8 B per theta*step (288 B for FitzHugh-Nagumo,
// against 192 B of compulsory solve_mv output), so instead
//   * the forward sweep stores only every K-th filtered state ("checkpoint", theta-innermost => coalesced), and
//   * the backward sweep works segment by segment: reload checkpoint jK, re-run the K-1 forward steps of the
//     segment into a per-warp shared-memory buffer, then smooth the segment backwards out of shared memory.
// History traffic drops to 2 x NSTATE x 8 / K bytes per theta*step for (K-1)/K of an extra forward step.
// The same buffer then stages the segment's outputs so that a warp writes them to HBM as contiguous runs of
// K time rows per theta (the reference layout is theta-outermost: a thread-per-theta store would scatter 8-byte
// writes 38 KB apart).
//
// Buffer layout, per warp:  buf[slot s][state k][lane] with a lane pitch of 33 doubles: conflict-free both for the
// per-thread accesses (fixed (s,k), lane varies) and for the copy-out (fixed lane, k varies).
constexpr int SEG_PITCH = 33;

// checkpoint interval: the largest K <= 8 whose buffer fits 18 KB per warp (12 one-warp CTAs per SM)
__host__ __device__ constexpr int seg_len(int nstate) {
  return (69 / nstate) < 1 ? 1 : ((69 / nstate) > 8 ? 8 : (69 / nstate));
}
__host__ __device__ constexpr int nstate_of(int nb, int p) { return nb * (p + p * (p + 1) / 2); }

// KC: checkpoint interval of the HBM history.  KC == K: one checkpoint per segment, the rest is recomputed (solve_mv,
// which is HBM-bound on its 192 B/theta*step of compulsory output).  KC == 1: every filtered state is stored and a
// segment is simply loaded back (solve_sim / fenrir: FP64-bound -- the draws or the backward filter dominate -- with
// little or no output, so re-running forward steps, and the chkrebtii random draws in them, would cost more than the
// history traffic).
template <typename T, class F>
struct SegBuf {
  static constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  static constexpr int K = seg_len(NSTATE);
  static constexpr int BYTES = K * NSTATE * SEG_PITCH * (int)sizeof(T);
  T* base;
  int lane;
  RD_DEV T& at(int s, int k) { return base[(s * NSTATE + k) * SEG_PITCH + lane]; }
  // means arrive in the mean type (double) and are staged / stored as T
  template <typename MT>
  RD_DEV void put(int s, const MT (&mu)[NB][P], const T (&S)[NB][NS]) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i) at(s, b * P + i) = (T)mu[b][i];
      RD_UNROLL for (int k = 0; k < NS; ++k) at(s, NB * P + b * NS + k) = S[b][k];
    }
  }
  template <typename MT>
  RD_DEV void get(int s, MT (&mu)[NB][P], T (&S)[NB][NS]) {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = (MT)at(s, b * P + i);
      RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = at(s, NB * P + b * NS + k);
    }
  }
  // Copy `rows` staged time rows starting at time n0 to a (B, N+1, ROW) output, ROW = NB*P (means / draws) or
  // NB*P*P (full covariances expanded from the packed slots).  Per theta the rows form ONE contiguous run of
  // rows*ROW elements, so the warp walks the 32 thetas and, for each, consecutive lanes store consecutive elements
  // of the run.  A lane's position r in the run -- hence its source slot/state -- is the same for every theta:
  // the (s, k) decode is done once per call, and the loop body is one LDS, one STG and two pointer bumps.
  // LOWER: the slots hold lower-triangular factors (square-root Kalman): expand to a full matrix with a zero
  // upper triangle instead of mirroring
  template <bool VAR, bool LOWER = false>
  RD_DEV void copy_out(T* __restrict__ out, i64 theta0, i64 B, int n_rows_total, int n0, int rows) {
    constexpr int ROW = VAR ? NB * P * P : NB * P;
    constexpr int NIT = (K * ROW + 31) / 32;
    const int run = rows * ROW;
    int src[NIT];
    RD_UNROLL for (int it = 0; it < NIT; ++it) {
      const int r = lane + 32 * it;
      const int s = r / ROW, e = r - s * ROW;
      int k = e;
      bool zero = false;
      if (VAR) {
        const int b = e / (P * P), ij = e - b * (P * P), i = ij / P, j = ij - i * P;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        if (LOWER) { k = NB * P + b * NS + i * (i + 1) / 2 + j; zero = j > i; }
        else k = NB * P + b * NS + lo * P - (lo * (lo - 1)) / 2 + (hi - lo);
      }
      src[it] = r < run ? (zero ? -2 : (s * NSTATE + k) * SEG_PITCH) : -1;
    }
    const i64 stride = (i64)n_rows_total * ROW;
    T* dst = out + (theta0 * (i64)n_rows_total + n0) * ROW + lane;
    const int nth = (B - theta0) < 32 ? (int)(B - theta0) : 32;
    RD_UNROLL4 for (int th = 0; th < nth; ++th) {
      RD_UNROLL for (int it = 0; it < NIT; ++it) {
        if (src[it] >= 0) dst[32 * it] = base[src[it] + th];
        else if (LOWER && src[it] == -2) dst[32 * it] = T(0);
      }
      dst += stride;
    }
  }
};

// asynchronous global -> shared copies of one element (LDGSTS: no register staging, completes in the background)
template <typename T>
RD_DEV void cp_async_elem(T* smem, const T* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}
RD_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
RD_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Every-state history (KC == 1): start copying filt[j*K .. j*K + cnt - 1] of theta idx into slots 0 .. cnt-1 of `buf`
// (row 0 = (ode_init, 0) is not in the history: the caller fills that slot).  Each lane copies, and later reads, only
// its own column of the buffer, so cp_async_wait_all() by the same lane is all the synchronisation needed.
template <typename T, class F>
RD_DEV void segment_fetch_async(const T* __restrict__ stash, i64 ldb, i64 idx, int j, int cnt, SegBuf<T, F>& buf) {
  constexpr int NSTATE = SegBuf<T, F>::NSTATE, K = SegBuf<T, F>::K;
  for (int s = 0; s < cnt; ++s) {
    const int n = j * K + s;
    if (n >= 1) {
      const T* g = stash + (i64)(n - 1) * NSTATE * ldb + idx;
      RD_UNROLL for (int k = 0; k < NSTATE; ++k) cp_async_elem(&buf.at(s, k), g + (i64)k * ldb);
    }
  }
  cp_async_commit();
}

// checkpoint j (= filt[j*K], j >= 1) of theta idx:  stash[((j-1) * NSTATE + k) * ldb + idx]
template <typename T, class F>
RD_DEV void ckpt_store(T* __restrict__ stash, i64 ldb, i64 idx, int j, const F& f) {
  constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  T* s = stash + (i64)(j - 1) * NSTATE * ldb + idx;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) s[(i64)(b * P + i) * ldb] = (T)f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) s[(i64)(NB * P + b * NS + k) * ldb] = f.S[b][k];
  }
}
template <typename T, class F>
RD_DEV void ckpt_load(const T* __restrict__ stash, i64 ldb, i64 idx, int j, F& f) {
  constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  const T* s = stash + (i64)(j - 1) * NSTATE * ldb + idx;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) f.mu[b][i] = s[(i64)(b * P + i) * ldb];
    RD_UNROLL for (int k = 0; k < NS; ++k) f.S[b][k] = s[(i64)(NB * P + b * NS + k) * ldb];
  }
}

// L2 prefetch of history entry j (no registers held while the lines travel from HBM)
template <typename T, class F>
RD_DEV void ckpt_prefetch(const T* __restrict__ stash, i64 ldb, i64 idx, int j) {
  constexpr int NSTATE = F::NB * (F::P + F::NS);
  const T* s = stash + (i64)(j - 1) * NSTATE * ldb + idx;
  RD_UNROLL for (int k = 0; k < NSTATE; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(s + (i64)k * ldb));
}

// one forward step n -> n+1 of a plain (no log-density) filter; `hook` adds the observation rows at observation steps
template <typename T, class Model, int INTERR, int QK>
RD_DEV void forward_step(const FilterConsts<T, Model::NB, Model::P, Model::M>& C, const CommonArgs<T>& a,
                         const typename Fwd<T, Model, INTERR, QK>::Par& q, i64 idx, int n,
                         Fwd<T, Model, INTERR, QK>& f, const ObsHook<T>* hook = nullptr,
                         const typename Fwd<T, Model, INTERR, QK>::MT* t_given = nullptr,
                         const typename Fwd<T, Model, INTERR, QK>::MT* frc_given = nullptr) {
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef typename F::MT MT;
  constexpr int NB = F::NB, M = F::M, JC = F::JC, MS = F::MS;
  LogPdfAcc<T> dummy;
  const MT t = Model::USES_TIME ? (t_given != nullptr ? *t_given : step_time<MT>(a.t_min, a.t_max, n, a.n_steps)) : MT(0);
  T jl[NB][M][JC], V[NB][MS], zc[NB][JC];
  MT res[NB][M];
  f.predict_all(C);
  f.template interr_normals<1>(a, idx, n, 0, zc);
  f.interrogate(C, q, t, zc, jl, res, V, frc_given);
  int io = -1;
  if (hook != nullptr) io = __ldg(hook->step_obs + n);
  if (io >= 0) f.template update_zy<1, false>(C, jl, res, V, hook->o, io, dummy);
  else f.template update_z<false>(C, jl, res, V, dummy);
}

// forward sweep 0 -> N storing filt[n] for every n that is a multiple of KC (history entry n / KC); on exit f holds
// filt[N]
template <typename T, class Model, int INTERR, int QK, int KC>
RD_DEV void forward_with_checkpoints(const FilterConsts<T, Model::NB, Model::P, Model::M>& C, const CommonArgs<T>& a,
                                     const typename Fwd<T, Model, INTERR, QK>::Par& q, i64 idx, bool live,
                                     Fwd<T, Model, INTERR, QK>& f, T* __restrict__ stash, i64 ldb,
                                     const ObsHook<T>* hook = nullptr) {
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef typename F::MT MT;
  int to_ckpt = KC, j = 0;
  // the time stamp of step n+1 (an IEEE division, solve.py:74) is formed while step n runs, off its dependency chain
  MT t_next = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, 0, a.n_steps) : MT(0);
  for (int n = 0; n < a.n_steps; ++n) {
    const MT t_now = t_next;
    if (Model::USES_TIME) t_next = step_time<MT>(a.t_min, a.t_max, n + 1, a.n_steps);
    forward_step<T, Model, INTERR, QK>(C, a, q, idx, n, f, hook, &t_now);
    if (--to_ckpt == 0) {
      to_ckpt = KC;
      ++j;
      if (live && n + 1 < a.n_steps) ckpt_store<T, F>(stash, ldb, idx, j, f);
    }
  }
}

// Rebuild filt[j*K .. j*K + cnt - 1] of segment j into the buffer slots 0 .. cnt-1.
//   KC == K: load checkpoint j (filt[0] = (ode_init, 0) for j == 0) and re-run the cnt-1 forward steps;
//   KC == 1: load every state of the segment from the history.
template <typename T, class Model, int INTERR, int QK, int KC>
RD_DEV void rebuild_segment(const FilterConsts<T, Model::NB, Model::P, Model::M>& C, const CommonArgs<T>& a,
                            const typename Fwd<T, Model, INTERR, QK>::Par& q, i64 idx, int j, int cnt,
                            Fwd<T, Model, INTERR, QK>& f, const T* __restrict__ stash, i64 ldb,
                            SegBuf<T, Fwd<T, Model, INTERR, QK>>& buf, const ObsHook<T>* hook = nullptr) {
  typedef Fwd<T, Model, INTERR, QK> F;
  constexpr int NB = F::NB, P = F::P, K = SegBuf<T, F>::K;
  if constexpr (KC == 1) {
    for (int s = 0; s < cnt; ++s) {
      const int n = j * K + s;
      if (n == 0) f.init(a.ode_init + idx * NB * P);
      else ckpt_load<T, F>(stash, ldb, idx, n, f);
      buf.put(s, f.mu, f.S);
    }
  } else {
    if (j == 0) f.init(a.ode_init + idx * NB * P);            // filt[0] = (ode_init, 0)
    else ckpt_load<T, F>(stash, ldb, idx, j, f);
    buf.put(0, f.mu, f.S);
    for (int s = 1; s < cnt; ++s) {
      forward_step<T, Model, INTERR, QK>(C, a, q, idx, j * K + s - 1, f, hook);
      buf.put(s, f.mu, f.S);
    }
  }
}

// prefetch what rebuild_segment(j) will load (called one segment ahead, before the copy-out of segment j+1)
template <typename T, class F, int KC>
RD_DEV void prefetch_segment(const T* __restrict__ stash, i64 ldb, i64 idx, int j, int K) {
  if (KC == 1) {
    for (int s = 0; s < K; ++s)
      if (j * K + s >= 1) ckpt_prefetch<T, F>(stash, ldb, idx, j * K + s);
  } else if (j >= 1) {
    ckpt_prefetch<T, F>(stash, ldb, idx, j);
  }
}

// full-matrix / vector stores of one time row (used for the single row N; everything else is staged)
template <typename T, int NB, int P, typename MT>
RD_DEV void store_mean_row(T* __restrict__ out, const MT (&mu)[NB][P]) {
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i) out[b * P + i] = (T)mu[b][i];
}
template <typename T, int NB, int P>
RD_DEV void store_var_row(T* __restrict__ out, const T (&S)[NB][P * (P + 1) / 2]) {
  RD_UNROLL for (int b = 0; b < NB; ++b)
    RD_UNROLL for (int i = 0; i < P; ++i)
      RD_UNROLL for (int j = 0; j < P; ++j) out[(b * P + i) * P + j] = S[b][sym<P>(i, j)];
}

extern __shared__ double rodeo_dyn_smem[];

// ------------------------------------------------------------------------------------------------------------------
// solve_mv: forward filter, then the mean/variance smoother  (reference src/rodeo/solve.py:208-302)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, class Model, int INTERR, int QK, bool OBS = false>
__global__ void __launch_bounds__(32)
solve_mv_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> Cg,
                const CommonArgs<T> a, T* __restrict__ stash, i64 ldb,
                T* __restrict__ mean_out, T* __restrict__ var_out, const ObsHook<T> oh = ObsHook<T>()) {
  const ObsHook<T>* hook = OBS ? &oh : nullptr;
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef SegBuf<T, F> Buf;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, K = Buf::K;
  const i64 theta0 = (i64)blockIdx.x * 32;
  i64 idx = theta0 + threadIdx.x;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;                 // the whole warp takes part in the cooperative copy-out
  const PriorConsts<T, Model::NB, Model::P, Model::M, QK> pc(Cg, a, idx);
  const FilterConsts<T, Model::NB, Model::P, Model::M>& C = pc.C;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  Buf buf{reinterpret_cast<T*>(rodeo_dyn_smem), (int)threadIdx.x};
  F f;
  f.init(a.ode_init + idx * NB * P);
  f.load_scale(a, idx);
  forward_with_checkpoints<T, Model, INTERR, QK, K>(C, a, q, idx, live, f, stash, ldb, hook);

  // smoothed[N] = filt[N]   (solve.py:279-282)
  typedef typename F::MT MT;
  MT ms[NB][P];
  T Ss[NB][NS];
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) ms[b][i] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) Ss[b][k] = f.S[b][k];
  }
  if (live && mean_out != nullptr) store_mean_row<T, NB, P>(mean_out + (idx * (i64)(N + 1) + N) * (NB * P), ms);
  if (live && var_out != nullptr) store_var_row<T, NB, P>(var_out + (idx * (i64)(N + 1) + N) * (NB * P * P), Ss);

  // rows N-1 .. 0, segment by segment
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;          // rows n0 .. n0+cnt-1  (all <= N-1)
    rebuild_segment<T, Model, INTERR, QK, K>(C, a, q, idx, j, cnt, f, stash, ldb, buf, hook);
    for (int s = cnt - 1; s >= 0; --s) {
      if (n0 + s == 0) break;                              // row 0 stays (ode_init, 0): never smoothed (solve.py:295-301)
      buf.get(s, f.mu, f.S);                               // filt[n]
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        MT mp[P], dm[P];
        T Sp[NS], G[P][P], Ct[P][P];
        predict<T, P, QK>(C.Q[b], C.R[b], f.rs[b], f.mu[b], f.S[b], mp, Sp);          // pred[n+1]
        smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
        // mu_s = mu_f + G (mu_s' - mu_p) ;  S_s = S_f + G (S_s' - S_p) G^T    (standard.py:213-216)
        T D[NS];
        RD_UNROLL for (int i = 0; i < P; ++i) dm[i] = ms[b][i] - mp[i];
        RD_UNROLL for (int k = 0; k < NS; ++k) D[k] = Ss[b][k] - Sp[k];
        RD_UNROLL for (int i = 0; i < P; ++i) {
          MT m = f.mu[b][i];
          RD_UNROLL for (int jj = 0; jj < P; ++jj) m = rd_fma((MT)G[i][jj], dm[jj], m);
          ms[b][i] = m;
        }
        RD_UNROLL for (int k = 0; k < NS; ++k) Ss[b][k] = f.S[b][k];
        add_GDGt<T, P>(G, D, Ss[b]);
      }
      buf.put(s, ms, Ss);                                  // stage the output row in place of filt[n]
    }
    if (j > 0) prefetch_segment<T, F, K>(stash, ldb, idx, j - 1, K);     // travels during the copy-out
    __syncwarp();
    if (mean_out != nullptr) buf.template copy_out<false>(mean_out, theta0, a.B, N + 1, n0, cnt);
    if (var_out != nullptr) buf.template copy_out<true>(var_out, theta0, a.B, N + 1, n0, cnt);
    __syncwarp();
  }
}


// ==================================================================================================================
// solve_mv with one lane per (theta, block)
// ==================================================================================================================
// solve_mv is bound by the write-back of its 192 B/theta*step of output, and the reference layout is theta-outermost:
// a warp can only write, per theta, the K time rows it has staged (K * 144 B of covariances).  A store-only
// micro-benchmark of that pattern (tools/micro/copyout_pattern.cu) reaches 3.2 TB/s at K = 3 but 4.3 TB/s at K = 6:
// the longer the per-theta run, the better HBM takes it.  K is capped by shared memory per warp, i.e. by the number
// of thetas a warp carries -- so for n_block >= 2 the blocks of a theta are spread over adjacent lanes: a warp then
// carries 32 / n_block thetas and stages n_block times more rows in the same buffer.  The block Kalman recursions are
// independent between blocks (reference: jax.vmap over n_block, src/rodeo/solve.py:62,81,263); only the ODE
// right-hand side couples them, through the JCOLS leading entries of each block's predicted mean, which the lanes of a
// theta exchange by shuffle once per forward step.  It also halves the work quantum per warp (finer load balance over
// the SM sub-partitions) and the registers per lane.
#ifndef RODEO_BL_SMEM
// staging bytes per warp.  FitzHugh-Nagumo (n_block 2, 18 states, 16 thetas per warp): 22,100 B -> K = 9 rows, 1.3 KB
// covariance runs, 10 resident warps per SM.  Measured on B200 for 65,536 thetas x 800 steps after the shared-memory
// accesses became conflict-free: K = 7 / 8 / 9 / 10 -> 3.09 / 3.02 / 2.84 / 3.03 ms (K = 5 / 6: 3.24): longer per-theta
// runs buy HBM efficiency until too few warps are left to keep the FP64 pipe busy.
#define RODEO_BL_SMEM 22100
#endif
// float32 (mixed arithmetic, half the output bytes) is bound by instruction issue rather than by HBM and prefers more
// resident warps: K = 7 / 9 / 16 -> 2.27 / 2.44 / 3.15 ms on the same batch
#ifndef RODEO_BL_SMEM_F32
#define RODEO_BL_SMEM_F32 8700
#endif
__host__ __device__ constexpr int seg_len_bl(int nstate, int nb, int elem_bytes = 8) {
  const int k = (elem_bytes == 8 ? RODEO_BL_SMEM : RODEO_BL_SMEM_F32) / (nstate * (32 / nb + 1) * elem_bytes);
  return k < 1 ? 1 : (k > 12 ? 12 : k);
}

template <typename T, class Model, int INTERR, int QK>
struct BlockLane {
  static constexpr int NB = Model::NB, P = Model::P, M = Model::M, JC = Model::JCOLS;
  static constexpr int NS = P * (P + 1) / 2, MS = M * (M + 1) / 2, NSTATE = NB * (P + NS);
  static constexpr bool UNITW = (QK == QK_UNIT_UPPER) && (M == 1);
  static constexpr int WK = Model::WCOL;
  static constexpr bool HAS_J = (INTERR == INTERR_KRAMER);
  static constexpr int TW = 32 / NB;                 // thetas per warp
  static constexpr int PITCH = TW + 1;
  static constexpr int K = seg_len_bl(NSTATE, NB, (int)sizeof(T));
  static constexpr int BYTES = K * NSTATE * PITCH * (int)sizeof(T);
  typedef typename MeanOf<T>::type MT;      // means, ODE evaluation and residuals: always double (rodeo_core.cuh)
  typedef FilterConsts<T, NB, P, M> Consts;
  typedef typename Model::template Par<MT> Par;

  MT mu[P];
  T S[NS], rs;
  T Q[P][P], R[NS], W[M][P];     // this lane's block of the shared constants, in registers
  int b, gb;                     // block index; this theta's slot tl in the warp (block c of it lives in lane c*TW + tl)

  RD_DEV void load_consts(const Consts& C) {
    RD_UNROLL for (int c = 0; c < NB; ++c)
      if (c == b) {
        RD_UNROLL for (int i = 0; i < P; ++i)
          RD_UNROLL for (int j = 0; j < P; ++j) Q[i][j] = C.Q[c][i][j];
        RD_UNROLL for (int k = 0; k < NS; ++k) R[k] = C.R[c][k];
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int j = 0; j < P; ++j) W[r][j] = C.W[c][r][j];
      }
  }
  RD_DEV void init(const T* x0) {
    RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = (MT)x0[b * P + i];
    RD_UNROLL for (int k = 0; k < NS; ++k) S[k] = T(0);
  }
  RD_DEV static int km(int bb, int i) { return bb * P + i; }
  RD_DEV static int kv(int bb, int k) { return NB * P + bb * NS + k; }

  // one forward step n -> n+1 of this lane's block (predict, interrogate, update); all 32 lanes must call it
  RD_DEV void step(const CommonArgs<T>& a, const Par& q, i64 idx, int n, const ObsHook<T>* hook = nullptr) {
    LogPdfAcc<T> dummy;
    int io = -1;
    if (hook != nullptr) io = __ldg(hook->step_obs + n);
    step_lp<false, 1>(a, q, idx, n, hook != nullptr ? &hook->o : nullptr, io, dummy, 0);
  }

  // The same with the forecast log-density of the update added to `acc` (WITH_LOGPDF) and, for io >= 0, the
  // observation rows of observation io (dalton zy_update, dalton.py:136-149).  NSTREAM / stream: layout of the injected
  // interrogation normals and the Philox stream (dalton: 2 streams, joint = 0, marginal = 1).
  template <bool WITH_LOGPDF, int NSTREAM, class ACC>
  RD_DEV void step_lp(const CommonArgs<T>& a, const Par& q, i64 idx, int n, const ObsArgs<T>* o, int io, ACC& acc,
                      int stream) {
    {
      MT mp[P];
      T Sp[NS];
      predict<T, P, QK>(Q, R, rs, mu, S, mp, Sp);
      RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = mp[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) S[k] = Sp[k];
    }
    const MT t = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, n, a.n_steps) : MT(0);
    MT xo[JC];
    if constexpr (INTERR == INTERR_CHKREBTII) {
      T zc[JC];
      if (a.z_interr != nullptr) {
        const T* z = a.z_interr + ((idx * a.n_steps + n) * NSTREAM + stream) * (NB * P) + b * P;
        RD_UNROLL for (int j = 0; j < JC; ++j) zc[j] = z[j];
      } else {
        // the same stream layout as the thread-per-theta kernels: normal k = b*JC + j of the step's vector; only
        // the Philox pair(s) that hold this block's normals are generated
        philox_normal_range<T, JC>(a.key0, a.key1, a.particle_offset + idx, n,
                                   stream == 0 ? TAG_INTERR_A : TAG_INTERR_B, b * JC, zc);
      }
      T A[P][P];
      psd_factor<T, P>(S, A);
      RD_UNROLL for (int j = 0; j < JC; ++j) {
        MT xa = mu[j];
        RD_UNROLL for (int k = 0; k <= j; ++k) xa = rd_fma((MT)A[j][k], (MT)zc[k], xa);
        xo[j] = xa;
      }
    } else {
      RD_UNROLL for (int j = 0; j < JC; ++j) xo[j] = mu[j];
    }
    // the right-hand side couples the blocks: gather every block's visible columns from the theta's lane group
    MT x[NB][JC];
    RD_UNROLL for (int c = 0; c < NB; ++c)
      RD_UNROLL for (int j = 0; j < JC; ++j) x[c][j] = __shfl_sync(0xffffffffu, xo[j], c * TW + gb);
    MT f[NB][M], jl[NB][M][JC];
    if constexpr (HAS_J) {
      eval_f_jac<Model, MT>(q, t, x, f, jl);
    } else {
      Model::template rhs<MT, MT>(q, t, x, f);
      RD_UNROLL for (int c = 0; c < NB; ++c)
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int j = 0; j < JC; ++j) jl[c][r][j] = MT(0);
    }
    MT fo[M];
    T jo[M][JC];
    RD_UNROLL for (int r = 0; r < M; ++r) {
      fo[r] = f[0][r];
      RD_UNROLL for (int j = 0; j < JC; ++j) jo[r][j] = (T)jl[0][r][j];
      RD_UNROLL for (int c = 1; c < NB; ++c) {
        fo[r] = (b == c) ? f[c][r] : fo[r];
        RD_UNROLL for (int j = 0; j < JC; ++j) jo[r][j] = (b == c) ? (T)jl[c][r][j] : jo[r][j];
      }
    }
    if (io >= 0) {
      // augmented update with the observation row of this block (dalton.py:136-149): scalar ODE row + one obs row with
      // uncorrelated noises = two scalar updates (rodeo_core.cuh update_two_rows; the same call as Fwd::update_zy)
      if constexpr (M == 1) {
        constexpr bool HAS_V = (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII);
        T D[P];
        const MT ry = (MT)__ldg(o->obs_data + (io * NB + b));
        RD_UNROLL for (int j = 0; j < P; ++j) D[j] = __ldg(o->obs_weight + (io * NB + b) * P + j);
        const T Om = __ldg(o->obs_var + (io * NB + b));
        const DenseRow<T, P> row2{D};
        if constexpr (UNITW) {
          const MT res = sub_exact(fo[0], mu[WK]);
          const T V0 = HAS_V ? S[sidx<P>(WK, WK)] : T(0);
          const UnitRow<T, P, JC, WK, HAS_J> row1{jo[0]};
          update_two_rows<T, P, WITH_LOGPDF, HAS_V>(mu, S, row1, res, V0, row2, ry, Om, acc);
        } else {
          T wm[P];
          MT res = fo[0];
          RD_UNROLL for (int j = 0; j < P; ++j) {
            wm[j] = (HAS_J && j < JC) ? W[0][j] - jo[0][j] : W[0][j];
            res = rd_fma(-(MT)W[0][j], mu[j], res);
          }
          T V0 = T(0);
          if constexpr (HAS_V) {
            RD_UNROLL for (int i = 0; i < P; ++i) {
              T u = S[sym<P>(i, 0)] * W[0][0];
              RD_UNROLL for (int j = 1; j < P; ++j) u = rd_fma(S[sym<P>(i, j)], W[0][j], u);
              V0 = i == 0 ? W[0][0] * u : rd_fma(W[0][i], u, V0);
            }
          }
          const DenseRow<T, P> row1{wm};
          update_two_rows<T, P, WITH_LOGPDF, HAS_V>(mu, S, row1, res, V0, row2, ry, Om, acc);
        }
      }
    } else if constexpr (UNITW) {
      const MT res = sub_exact(fo[0], mu[WK]);
      const T V = (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII) ? S[sidx<P>(WK, WK)] : T(0);
      update_unit_row<T, P, JC, WK, WITH_LOGPDF, HAS_J, (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII)>(
          mu, S, jo[0], res, V, acc);
    } else {
      T wm[M][P], V[MS];
      MT res[M];
      RD_UNROLL for (int r = 0; r < M; ++r) {
        MT acc0 = fo[r];
        RD_UNROLL for (int j = 0; j < P; ++j) {
          wm[r][j] = (HAS_J && j < JC) ? W[r][j] - jo[r][j] : W[r][j];
          acc0 = rd_fma(-(MT)W[r][j], mu[j], acc0);
        }
        res[r] = acc0;
      }
      if constexpr (INTERR == INTERR_RODEO || INTERR == INTERR_CHKREBTII) {
        T u[M][P];
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int i = 0; i < P; ++i) {
            T acc0 = S[sym<P>(i, 0)] * W[r][0];
            RD_UNROLL for (int j = 1; j < P; ++j) acc0 = rd_fma(S[sym<P>(i, j)], W[r][j], acc0);
            u[r][i] = acc0;
          }
        RD_UNROLL for (int r = 0; r < M; ++r)
          RD_UNROLL for (int s = r; s < M; ++s) {
            T acc0 = W[r][0] * u[s][0];
            RD_UNROLL for (int i = 1; i < P; ++i) acc0 = rd_fma(W[r][i], u[s][i], acc0);
            V[sidx<M>(r, s)] = acc0;
          }
      } else {
        RD_UNROLL for (int k = 0; k < MS; ++k) V[k] = T(0);
      }
      update<T, P, M, WITH_LOGPDF>(mu, S, wm, res, V, acc);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------
// dalton with one lane per (theta, filter, block)
// ------------------------------------------------------------------------------------------------------------------
// dalton_kernel gives a thread one filter of one theta with all of its blocks.  When the batch is too small to fill the
// GPU's FP64 pipes -- 8,192 thetas per GPU is what BASELINE configs[1] leaves each of 8 GPUs: 512 warps for 592 SM
// sub-partitions -- its run time is the serial dependency chain of one thread (measured 474 cycles per step against 250
// issue cycles), so the blocks of a filter are spread over lanes as in solve_mv_bl_kernel: a warp carries 32 / n_block
// thetas, every lane runs the recursion of ONE block, and the lanes of a theta exchange the ODE-visible entries of
// their predicted means by shuffle once per step.  Same grid split as geometry 2 of dalton_kernel (first half of the
// grid joint filters, second half marginal ones, combined by atomicAdd into the zeroed output).  The per-block partial
// sums are gathered by shuffle at the end and combined exactly as LogPdfParts::value() does, so the result is BITWISE
// the one of dalton_kernel: which kernel the host picks (by batch size) never shows in the numbers.
#ifndef RODEO_DALTON_BL_MINB
#define RODEO_DALTON_BL_MINB 12
#endif
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32, RODEO_DALTON_BL_MINB)
dalton_bl_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                 const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ loglik) {
  static_assert(NOBS == 1, "block-lane dalton: one observation row per block");
  typedef BlockLane<T, Model, INTERR, QK> L;
  typedef typename L::MT MT;
  constexpr int NB = L::NB, P = L::P, TW = L::TW;
  const int lane = threadIdx.x;
  int b = lane / TW, tl = lane - b * TW;            // block-major lanes, see solve_mv_bl_kernel
  const bool lane_ok = b < NB;
  if (!lane_ok) { tl = TW - 1; b = NB - 1; }
  const i64 half = gridDim.x >> 1;
  const bool joint = (i64)blockIdx.x < half;
  const i64 theta0 = (joint ? (i64)blockIdx.x : (i64)blockIdx.x - half) * TW;
  i64 idx = theta0 + tl;
  const bool live = lane_ok && idx < a.B;
  if (idx >= a.B) idx = a.B - 1;
  const typename L::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  L f;
  f.b = b; f.gb = tl;
  f.load_consts(C);
  f.rs = a.r_scale != nullptr ? a.r_scale[idx * NB + b] : T(1);
  f.init(a.ode_init + idx * NB * P);
  LogPdfPart<T> part;
  LogPdfCtl ctl;
  part.init(); ctl.init();
  LogPdfPartRef<T> acc{part, ctl};

  // log p(Y_0 | X_0) when the first observation sits on t_min (dalton.py:207-215); joint filter only
  int i = 0;
  if (__ldg(o.obs_ind) == 0) {
    if (joint) {
      T res[1], Om[1];
      MT m = MT(0);
      RD_UNROLL for (int j = 0; j < P; ++j) m = rd_fma((MT)__ldg(o.obs_weight + b * P + j), f.mu[j], m);
      res[0] = (T)((MT)__ldg(o.obs_data + b) - m);
      Om[0] = __ldg(o.obs_var + b);
      logpdf_terms<T, 1>(Om, res, acc);
    }
    i = 1;
  }
  int next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));
  // steps between two observations run in a branch-free inner loop: see dalton_kernel
  int n = 0;
  const int N = a.n_steps;
  for (;;) {
    const int stop = next_obs - 1;
    const bool has_obs = stop >= n && stop < N;
    const int end = has_obs ? stop : N;
    for (; n < end; ++n) {
      f.template step_lp<true, 2>(a, q, idx, n, &o, -1, acc, joint ? 0 : 1);
      if ((n & 3) == 3) part.renorm(ctl);
    }
    if (!has_obs) break;
    f.template step_lp<true, 2>(a, q, idx, n, &o, joint ? (i < o.n_obs ? i : o.n_obs - 1) : -1, acc, joint ? 0 : 1);
    if ((n & 3) == 3) part.renorm(ctl);
    ++n; ++i;
    next_obs = __ldg(o.obs_ind + (i < o.n_obs ? i : o.n_obs - 1));
  }
  // every block's partial sums, in block order, then the one formula both kernels share
  LogPdfPart<T> all[NB];
  LogPdfCtl tot;
  tot.init();
  RD_UNROLL for (int c = 0; c < NB; ++c) {
    const int src = c * TW + tl;
    all[c].quad = __shfl_sync(0xffffffffu, part.quad, src);
    all[c].prod = __shfl_sync(0xffffffffu, part.prod, src);
    tot.esum += __shfl_sync(0xffffffffu, ctl.esum, src);
    tot.cnt += __shfl_sync(0xffffffffu, ctl.cnt, src);
    tot.sgn |= __shfl_sync(0xffffffffu, ctl.sgn, src);
  }
  const MT mine = combine_logpdf<T, NB>(all, tot);
  if (live && b == 0 && lane < TW) atomicAdd(loglik + idx, (T)(joint ? mine : -mine));
}

// output stores: written once, never read back by the kernel
template <typename T>
RD_DEV void store_out(T* p, T v) {
#if defined(RODEO_STORE_CS)
  __stcs(p, v);
#elif defined(RODEO_STORE_WT)
  __stwt(p, v);
#else
  *p = v;
#endif
}

// copy `rows` staged time rows of `nth` thetas (buffer [slot][state][theta], pitch PITCH) to a (B, N+1, ROW) output;
// same scheme as SegBuf::copy_out
template <typename T, int NB, int P, int K, int PITCH, bool VAR>
RD_DEV void seg_copy_out(const T* __restrict__ base, int lane, T* __restrict__ out, i64 theta0, int nth,
                         int n_rows_total, int n0, int rows) {
  constexpr int NS = P * (P + 1) / 2, NSTATE = NB * (P + NS);
  constexpr int ROW = VAR ? NB * P * P : NB * P;
  constexpr int NIT = (K * ROW + 31) / 32;
  const int run = rows * ROW;
  int src[NIT];
  RD_UNROLL for (int it = 0; it < NIT; ++it) {
    const int r = lane + 32 * it;
    const int s = r / ROW, e = r - s * ROW;
    int k = e;
    if (VAR) {
      const int bb = e / (P * P), ij = e - bb * (P * P), i = ij / P, j = ij - i * P;
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      k = NB * P + bb * NS + lo * P - (lo * (lo - 1)) / 2 + (hi - lo);
    }
    src[it] = r < run ? (s * NSTATE + k) * PITCH : -1;
  }
  const i64 stride = (i64)n_rows_total * ROW;
  T* dst = out + (theta0 * (i64)n_rows_total + n0) * ROW + lane;
  RD_UNROLL4 for (int th = 0; th < nth; ++th) {
    RD_UNROLL for (int it = 0; it < NIT; ++it)
      if (src[it] >= 0) store_out(dst + 32 * it, base[src[it] + th]);
    dst += stride;
  }
}

// (no minimum-blocks bound: an explicit 1 lets ptxas take 162 registers, which is harmless for float64 -- shared memory
// caps the residency at 10 CTAs per SM -- but costs the float32 instantiation a third of its resident warps; caps of
// 18 / 20 CTAs per SM (96 registers, spills) are 6 % / 12 % slower)
template <typename T, class Model, int INTERR, int QK, bool OBS = false>
__global__ void __launch_bounds__(32)
solve_mv_bl_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                   const CommonArgs<T> a, T* __restrict__ stash, i64 ldb,
                   T* __restrict__ mean_out, T* __restrict__ var_out, const ObsHook<T> oh = ObsHook<T>()) {
  const ObsHook<T>* hook = OBS ? &oh : nullptr;
  typedef BlockLane<T, Model, INTERR, QK> L;
  constexpr int NB = L::NB, P = L::P, NS = L::NS, K = L::K, TW = L::TW, NSTATE = L::NSTATE, PITCH = L::PITCH;
  const int lane = threadIdx.x;
  // block-major lanes: lane = b * TW + tl.  A half-warp then holds consecutive thetas of ONE block, so its 64-bit
  // shared-memory accesses buf[(s * NSTATE + k(b, .)) * PITCH + tl] fall in 16 distinct bank pairs (theta-major lanes
  // put two blocks' rows, 3 * PITCH apart, into each half-warp: every LDS / STS took 4 wavefronts instead of 2)
  int b = lane / TW, tl = lane - b * TW;
  const bool lane_ok = b < NB;             // 32 % NB trailing lanes shadow the last group; their stores are masked
  if (!lane_ok) { tl = TW - 1; b = NB - 1; }
  const i64 theta0 = (i64)blockIdx.x * TW;
  i64 idx = theta0 + tl;
  const bool live = lane_ok && idx < a.B;
  if (idx >= a.B) idx = a.B - 1;
  const typename L::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  T* buf = reinterpret_cast<T*>(rodeo_dyn_smem);
  L f;
  f.b = b; f.gb = tl;
  f.load_consts(C);
  f.rs = a.r_scale != nullptr ? a.r_scale[idx * NB + b] : T(1);
  const T* x0 = a.ode_init + idx * NB * P;
  f.init(x0);
  typedef typename L::MT MT;
  auto put = [&](int s, const MT (&m)[P], const T (&Sv)[NS]) {
    RD_UNROLL for (int i = 0; i < P; ++i) buf[(s * NSTATE + L::km(b, i)) * PITCH + tl] = (T)m[i];
    RD_UNROLL for (int k = 0; k < NS; ++k) buf[(s * NSTATE + L::kv(b, k)) * PITCH + tl] = Sv[k];
  };
  auto get = [&](int s, MT (&m)[P], T (&Sv)[NS]) {
    RD_UNROLL for (int i = 0; i < P; ++i) m[i] = (MT)buf[(s * NSTATE + L::km(b, i)) * PITCH + tl];
    RD_UNROLL for (int k = 0; k < NS; ++k) Sv[k] = buf[(s * NSTATE + L::kv(b, k)) * PITCH + tl];
  };
  // history entry j (= filt[j*K]) of this lane's block
  auto ck = [&](int j, int k) -> i64 { return ((i64)(j - 1) * NSTATE + k) * ldb + idx; };

  // forward sweep with one checkpoint per segment
  {
    int to_ckpt = K, j = 0;
    for (int n = 0; n < N; ++n) {
      f.step(a, q, idx, n, hook);
      if (--to_ckpt == 0) {
        to_ckpt = K; ++j;
        if (live && n + 1 < N) {
          RD_UNROLL for (int i = 0; i < P; ++i) stash[ck(j, L::km(b, i))] = (T)f.mu[i];
          RD_UNROLL for (int k = 0; k < NS; ++k) stash[ck(j, L::kv(b, k))] = f.S[k];
        }
      }
    }
  }
  // smoothed[N] = filt[N]   (solve.py:279-282)
  MT ms[P];
  T Ss[NS];
  RD_UNROLL for (int i = 0; i < P; ++i) ms[i] = f.mu[i];
  RD_UNROLL for (int k = 0; k < NS; ++k) Ss[k] = f.S[k];
  if (live && mean_out != nullptr)
    RD_UNROLL for (int i = 0; i < P; ++i) mean_out[(idx * (i64)(N + 1) + N) * (NB * P) + b * P + i] = (T)ms[i];
  if (live && var_out != nullptr)
    RD_UNROLL for (int i = 0; i < P; ++i)
      RD_UNROLL for (int jj = 0; jj < P; ++jj)
        var_out[(idx * (i64)(N + 1) + N) * (NB * P * P) + (b * P + i) * P + jj] = Ss[sym<P>(i, jj)];

  const int nth = (a.B - theta0) < TW ? (int)(a.B - theta0) : TW;
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;          // rows n0 .. n0+cnt-1  (all <= N-1)
    if (j == 0) {
      f.init(x0);                                           // filt[0] = (ode_init, 0)
    } else {
      RD_UNROLL for (int i = 0; i < P; ++i) f.mu[i] = (MT)stash[ck(j, L::km(b, i))];
      RD_UNROLL for (int k = 0; k < NS; ++k) f.S[k] = stash[ck(j, L::kv(b, k))];
    }
    put(0, f.mu, f.S);
    for (int s = 1; s < cnt; ++s) {
      f.step(a, q, idx, n0 + s - 1, hook);
      put(s, f.mu, f.S);
    }
    for (int s = cnt - 1; s >= 0; --s) {
      if (n0 + s == 0) break;                              // row 0 stays (ode_init, 0): never smoothed
      get(s, f.mu, f.S);                                   // filt[n]
      MT mp[P], dm[P];
      T Sp[NS], G[P][P], Ct[P][P];
      predict<T, P, QK>(f.Q, f.R, f.rs, f.mu, f.S, mp, Sp);               // pred[n+1]
      smooth_gain<T, P, QK>(f.Q, f.S, Sp, G, Ct);
      T D[NS];
      RD_UNROLL for (int i = 0; i < P; ++i) dm[i] = ms[i] - mp[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) D[k] = Ss[k] - Sp[k];
      RD_UNROLL for (int i = 0; i < P; ++i) {
        MT m = f.mu[i];
        RD_UNROLL for (int jj = 0; jj < P; ++jj) m = rd_fma((MT)G[i][jj], dm[jj], m);
        ms[i] = m;
      }
      RD_UNROLL for (int k = 0; k < NS; ++k) Ss[k] = f.S[k];
      add_GDGt<T, P>(G, D, Ss);
      put(s, ms, Ss);                                      // stage the output row in place of filt[n]
    }
    if (j > 1) {                                           // next segment's checkpoint travels during the copy-out
      RD_UNROLL for (int i = 0; i < P; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(stash + ck(j - 1, L::km(b, i))));
      RD_UNROLL for (int k = 0; k < NS; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(stash + ck(j - 1, L::kv(b, k))));
    }
    __syncwarp();
    if (mean_out != nullptr) seg_copy_out<T, NB, P, K, PITCH, false>(buf, lane, mean_out, theta0, nth, N + 1, n0, cnt);
    if (var_out != nullptr) seg_copy_out<T, NB, P, K, PITCH, true>(buf, lane, var_out, theta0, nth, N + 1, n0, cnt);
    __syncwarp();
  }
}



// ==================================================================================================================
// solve_mv with kalman_type = "square-root"  (reference src/rodeo/solve.py:208-302 with kalman_funs = square_root)
// ==================================================================================================================
// Thread per theta, generic (no structure exploitation): the reference's square-root path is the numerically robust
// one, not the fast one.  Fwd<..>::S holds the packed LOWER factor L; FilterConsts::R holds the packed lower factor
// of the prior variance (the caller passes prior_pars = (Q, cholesky(R)), docs/examples/higher_order.md:108-112).
template <typename T, class Model, int INTERR>
RD_DEV void forward_step_sqrt(const FilterConsts<T, Model::NB, Model::P, Model::M>& C, const CommonArgs<T>& a,
                              const typename Fwd<T, Model, INTERR, QK_DENSE>::Par& q, i64 idx, int n,
                              Fwd<T, Model, INTERR, QK_DENSE>& f) {
  typedef typename Fwd<T, Model, INTERR, QK_DENSE>::MT MT;      // means / right-hand side: double also for T = float
  constexpr int NB = Model::NB, P = Model::P, M = Model::M, JC = Model::JCOLS, NS = P * (P + 1) / 2;
  static_assert(M == 1, "square-root path: scalar measurements per block");
  constexpr bool CHK = (INTERR == INTERR_CHKREBTII);
  const MT t = Model::USES_TIME ? step_time<MT>(a.t_min, a.t_max, n, a.n_steps) : MT(0);
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    MT mp[P];
    T Lp[NS];
    sqrt_predict<T, P>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Lp);
    RD_UNROLL for (int i = 0; i < P; ++i) f.mu[b][i] = mp[i];
    RD_UNROLL for (int k = 0; k < NS; ++k) f.S[b][k] = Lp[k];
  }
  MT x[NB][JC];
  T vrow[NB][CHK ? P : 1];
  if constexpr (CHK) {
    // interrogate.py:36-47 ("square-root" branch), literally: var_meas = W L is an (m, p) block and
    // x_state = mean + var_meas @ z, whose single entry broadcasts over every state entry of the block
    T z[NB * P];
    if (a.z_interr != nullptr) {
      const T* zp = a.z_interr + (idx * a.n_steps + n) * (NB * P);
      RD_UNROLL for (int k = 0; k < NB * P; ++k) z[k] = zp[k];
    } else {
      philox_normals<T, NB * P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_INTERR_A, z);
    }
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      MT shift = MT(0);
      RD_UNROLL for (int c = 0; c < P; ++c) {
        T v = T(0);
        RD_UNROLL for (int k = c; k < P; ++k) v = rd_fma(C.W[b][0][k], f.S[b][lidx(k, c)], v);
        vrow[b][c] = v;
        shift = rd_fma((MT)v, (MT)z[b * P + c], shift);
      }
      RD_UNROLL for (int j = 0; j < JC; ++j) x[b][j] = f.mu[b][j] + shift;
    }
  } else {
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      vrow[b][0] = T(0);
      RD_UNROLL for (int j = 0; j < JC; ++j) x[b][j] = f.mu[b][j];
    }
  }
  MT fv[NB][M], jl[NB][M][JC];
  if constexpr (INTERR == INTERR_KRAMER) {
    eval_f_jac<Model, MT>(q, t, x, fv, jl);
  } else {
    Model::template rhs<MT, MT>(q, t, x, fv);
    RD_UNROLL for (int b = 0; b < NB; ++b)
      RD_UNROLL for (int j = 0; j < JC; ++j) jl[b][0][j] = MT(0);
  }
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    T w[P];
    MT res = fv[b][0];
    RD_UNROLL for (int j = 0; j < P; ++j) {
      w[j] = (INTERR == INTERR_KRAMER && j < JC) ? C.W[b][0][j] - (T)jl[b][0][j] : C.W[b][0][j];
      res = rd_fma(-(MT)C.W[b][0][j], f.mu[b][j], res);
    }
    sqrt_update_row<T, P, CHK ? P : 0>(f.mu[b], f.S[b], w, res, vrow[b]);
  }
}

template <typename T, class Model, int INTERR>
__global__ void __launch_bounds__(32)
solve_mv_sqrt_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                     const CommonArgs<T> a, T* __restrict__ stash, i64 ldb,
                     T* __restrict__ mean_out, T* __restrict__ var_out) {
  typedef Fwd<T, Model, INTERR, QK_DENSE> F;
  typedef SegBuf<T, F> Buf;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, K = Buf::K;
  const i64 theta0 = (i64)blockIdx.x * 32;
  i64 idx = theta0 + threadIdx.x;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  Buf buf{reinterpret_cast<T*>(rodeo_dyn_smem), (int)threadIdx.x};
  F f;
  f.init(a.ode_init + idx * NB * P);
  {
    int to_ckpt = K, j = 0;
    for (int n = 0; n < N; ++n) {
      forward_step_sqrt<T, Model, INTERR>(C, a, q, idx, n, f);
      if (--to_ckpt == 0) {
        to_ckpt = K; ++j;
        if (live && n + 1 < N) ckpt_store<T, F>(stash, ldb, idx, j, f);
      }
    }
  }
  typename F::MT ms[NB][P];
  T Ls[NB][NS];
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) ms[b][i] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) Ls[b][k] = f.S[b][k];
  }
  if (live) {
    store_mean_row<T, NB, P>(mean_out + (idx * (i64)(N + 1) + N) * (NB * P), ms);
    T* vr = var_out + (idx * (i64)(N + 1) + N) * (NB * P * P);
    RD_UNROLL for (int b = 0; b < NB; ++b)
      RD_UNROLL for (int i = 0; i < P; ++i)
        RD_UNROLL for (int j = 0; j < P; ++j) vr[(b * P + i) * P + j] = lget<T, P>(Ls[b], i, j);
  }
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    if (j == 0) f.init(a.ode_init + idx * NB * P);
    else ckpt_load<T, F>(stash, ldb, idx, j, f);
    buf.put(0, f.mu, f.S);
    for (int s = 1; s < cnt; ++s) {
      forward_step_sqrt<T, Model, INTERR>(C, a, q, idx, n0 + s - 1, f);
      buf.put(s, f.mu, f.S);
    }
    for (int s = cnt - 1; s >= 0; --s) {
      if (n0 + s == 0) break;
      buf.get(s, f.mu, f.S);                               // filt[n] (mean, lower factor)
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        typename F::MT mp[P];
        T Lp[NS];
        sqrt_predict<T, P>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Lp);      // pred[n+1]
        sqrt_smooth_mv<T, P>(C.Q[b], C.R[b], f.mu[b], f.S[b], mp, Lp, ms[b], Ls[b]);
      }
      buf.put(s, ms, Ls);
    }
    __syncwarp();
    buf.template copy_out<false>(mean_out, theta0, a.B, N + 1, n0, cnt);
    buf.template copy_out<true, true>(var_out, theta0, a.B, N + 1, n0, cnt);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// solve_sim with one lane per (theta, block)
// ------------------------------------------------------------------------------------------------------------------
// Same lane mapping as solve_mv_bl_kernel.  The sampling smoother is FP64-bound (three reciprocals, a 3x3 factor and
// the Box-Muller normals per block-step) and its batches are smaller (C5: 32,768 particles per GPU = 1.7 one-theta
// warps per SM sub-partition), so spreading the blocks over lanes is mainly about parallelism.  Every filtered state
// is kept in HBM (per-lane loads are 9 doubles, prefetched one step ahead in registers); shared memory only stages
// the draws so that they leave as runs of K time rows per theta.
// (Checkpointing every second state and re-running the odd forward steps, which pays for the one-lane-per-theta kernel,
// does not here: 2.09 -> 2.26 ms on BASELINE configs[4] -- at 3.5 warps per scheduler the extra instructions cost more
// than the saved traffic.)

// 14 resident one-warp CTAs per SM (what the 13 KB staging buffer allows): BASELINE configs[4] gives one GPU 32,768
// particles = 2,048 warps = 13.8 per SM; at 12 CTAs per SM (148 registers) the launch was 1.15 waves.  With the bound
// ptxas settles on 125 registers without spilling: 2.64 -> 2.23 ms on that configuration (a plain 144-register cap,
// which also keeps everything resident, is no faster than before: the gain is the tighter schedule, not the residency)
template <typename T, class Model, int INTERR, int QK, bool OBS = false>
__global__ void __launch_bounds__(32, RODEO_SIM_BL_MINB)
solve_sim_bl_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                    const CommonArgs<T> a, const T* __restrict__ z_smooth, T* __restrict__ stash, i64 ldb,
                    T* __restrict__ x_out, const ObsHook<T> oh = ObsHook<T>(),
                    const SimLoglik<T> sl = SimLoglik<T>()) {
  const ObsHook<T>* hook = OBS ? &oh : nullptr;
  typedef BlockLane<T, Model, INTERR, QK> L;
  constexpr int NB = L::NB, P = L::P, NS = L::NS, TW = L::TW, NSTATE = L::NSTATE, PITCH = L::PITCH;
  constexpr int K = 16, ROW = NB * P;                    // staged time rows: K * ROW * PITCH elements per warp
  const int lane = threadIdx.x;
  int b = lane / TW, tl = lane - b * TW;   // block-major lanes, see solve_mv_bl_kernel
  const bool lane_ok = b < NB;
  if (!lane_ok) { tl = TW - 1; b = NB - 1; }
  const i64 theta0 = (i64)blockIdx.x * TW;
  i64 idx = theta0 + tl;
  const bool live = lane_ok && idx < a.B;
  if (idx >= a.B) idx = a.B - 1;
  const typename L::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  T* buf = reinterpret_cast<T*>(rodeo_dyn_smem);
  L f;
  f.b = b; f.gb = tl;
  f.load_consts(C);
  f.rs = a.r_scale != nullptr ? a.r_scale[idx * NB + b] : T(1);
  const T* x0 = a.ode_init + idx * NB * P;
  f.init(x0);
  // history entry n (= filt[n], 1 <= n <= N-1) of this lane's block
  auto hs = [&](int n, int k) -> i64 { return ((i64)(n - 1) * NSTATE + k) * ldb + idx; };

  for (int n = 0; n < N; ++n) {
    f.step(a, q, idx, n, hook);
    if (live && n + 1 < N) {
      RD_UNROLL for (int i = 0; i < P; ++i) stash[hs(n + 1, L::km(b, i))] = (T)f.mu[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) stash[hs(n + 1, L::kv(b, k))] = f.S[k];
    }
  }

  auto normals = [&](int n, T (&z)[P]) {
    if (z_smooth != nullptr) {
      const T* zp = z_smooth + (idx * (i64)(N + 1) + n) * (NB * P) + b * P;
      RD_UNROLL for (int k = 0; k < P; ++k) z[k] = zp[k];
    } else {
      philox_normals<T, P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_SMOOTH, z, b * ((P + 3) / 4));
    }
  };
  typedef typename L::MT MT;
  auto draw = [&](const MT (&m)[P], const T (&Cv)[NS], const T (&z)[P], MT (&xo)[P]) {
    T A[P][P];
    psd_factor<T, P>(Cv, A);
    RD_UNROLL for (int i = 0; i < P; ++i) {
      MT acc = m[i];
      RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)A[i][k], (MT)z[k], acc);
      xo[i] = acc;
    }
  };

  // terminal draw from N(mu_f[N], S_f[N])  (solve.py:182-186)
  MT x[P];
  SimLoglikAcc<T, MT> la;
  la.init(sl, N + 1);
  {
    T z[P];
    MT xn[P];
    normals(N, z);
    draw(f.mu, f.S, z, xn);
    RD_UNROLL for (int i = 0; i < P; ++i) x[i] = xn[i];
    if (live && x_out != nullptr)
      RD_UNROLL for (int i = 0; i < P; ++i) x_out[(idx * (i64)(N + 1) + N) * ROW + b * P + i] = (T)x[i];
    la.row(sl, N, NB, b, b + 1, [&](int) { return x[0]; });
  }

  const int nth = (a.B - theta0) < TW ? (int)(a.B - theta0) : TW;
  // software pipeline: filt[n-1] is loaded into registers while row n is being processed
  T nmu[P], nS[NS];
  auto load = [&](int n) {
    if (n >= 1) {
      RD_UNROLL for (int i = 0; i < P; ++i) nmu[i] = stash[hs(n, L::km(b, i))];
      RD_UNROLL for (int k = 0; k < NS; ++k) nS[k] = stash[hs(n, L::kv(b, k))];
    }
  };
  load(N - 1);
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    for (int s = cnt - 1; s >= 0; --s) {
      const int n = n0 + s;
      if (n == 0) {                                        // row 0 = ode_init: x0 is known, not sampled
        if (x_out != nullptr)
          RD_UNROLL for (int i = 0; i < P; ++i) buf[(s * ROW + b * P + i) * PITCH + tl] = x0[b * P + i];
        la.row(sl, 0, NB, b, b + 1, [&](int) { return (MT)x0[b * P]; });
        break;
      }
      RD_UNROLL for (int i = 0; i < P; ++i) f.mu[i] = (MT)nmu[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) f.S[k] = nS[k];
      load(n - 1);
      T z[P];
      MT xn[P];
      normals(n, z);
      MT mp[P], m[P];
      T Sp[NS], G[P][P], Ct[P][P], Cv[NS];
      predict<T, P, QK>(f.Q, f.R, f.rs, f.mu, f.S, mp, Sp);               // pred[n+1]
      smooth_gain<T, P, QK>(f.Q, f.S, Sp, G, Ct);
      // m = mu_f + G (x' - mu_p) ;  C = S_f - G (S_f Q^T)^T      (standard.py:251-254)
      RD_UNROLL for (int i = 0; i < P; ++i) {
        MT acc = f.mu[i];
        RD_UNROLL for (int jj = 0; jj < P; ++jj) acc = rd_fma((MT)G[i][jj], x[jj] - mp[jj], acc);
        m[i] = acc;
      }
      cond_var<T, P>(f.S, G, Ct, Cv);
      draw(m, Cv, z, xn);
      RD_UNROLL for (int i = 0; i < P; ++i) x[i] = xn[i];
      if (x_out != nullptr)
        RD_UNROLL for (int i = 0; i < P; ++i) buf[(s * ROW + b * P + i) * PITCH + tl] = (T)xn[i];
      la.row(sl, n, NB, b, b + 1, [&](int) { return x[0]; });
    }
    if (x_out != nullptr) {
      __syncwarp();
      // copy rows n0 .. n0+cnt-1: per theta one contiguous run of cnt*ROW elements
      constexpr int NIT = (K * ROW + 31) / 32;
      const int run = cnt * ROW;
      const i64 stride = (i64)(N + 1) * ROW;
      T* dst = x_out + (theta0 * (i64)(N + 1) + n0) * ROW + lane;
      RD_UNROLL4 for (int th = 0; th < nth; ++th) {
        RD_UNROLL for (int it = 0; it < NIT; ++it) {
          const int r = lane + 32 * it;
          if (r < run) dst[32 * it] = buf[r * PITCH + th];
        }
        dst += stride;
      }
      __syncwarp();
    }
  }
  if (sl.out != nullptr) {
    // sum the blocks of a theta in block order (lane c * TW + tl holds block c)
    MT tot = MT(0);
    RD_UNROLL for (int c = 0; c < NB; ++c) tot += __shfl_sync(0xffffffffu, la.ll, c * TW + tl);
    if (live && b == 0 && lane < TW) sl.out[idx] = (T)tot;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// solve_sim: forward filter, then the sampling smoother  (reference src/rodeo/solve.py:125-205)
// ------------------------------------------------------------------------------------------------------------------
// z_smooth: optional injected normals (B, N+1, NB, P); row N feeds the terminal draw, rows 1..N-1 the backward
// draws.  Without it the draws come from Philox keyed by (key, particle, step).
// 14 resident CTAs per SM for models of up to 3 blocks (a thread holding more would spill heavily): Lorenz63 goes from
// 166 registers to 128 with 116 B spilled and BASELINE configs[2] from 37.6 to 32.8 ms -- all 2,048 warps resident at once
// and, as for solve_sim_bl_kernel, a tighter schedule
#ifndef RODEO_SIM_T_RECOMPUTE
#define RODEO_SIM_T_RECOMPUTE 1
#endif
#ifndef RODEO_SIM_T_MINB
#define RODEO_SIM_T_MINB 14
#endif
template <typename T, class Model, int INTERR, int QK, bool OBS = false>
__global__ void __launch_bounds__(32, Model::NB <= 3 ? RODEO_SIM_T_MINB : 1)
solve_sim_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> Cg,
                 const CommonArgs<T> a, const T* __restrict__ z_smooth, T* __restrict__ stash, i64 ldb,
                 T* __restrict__ x_out, const ObsHook<T> oh = ObsHook<T>(), const SimLoglik<T> sl = SimLoglik<T>()) {
  const ObsHook<T>* hook = OBS ? &oh : nullptr;
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef SegBuf<T, F> Buf;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, K = Buf::K;
  const i64 theta0 = (i64)blockIdx.x * 32;
  i64 idx = theta0 + threadIdx.x;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const PriorConsts<T, Model::NB, Model::P, Model::M, QK> pc(Cg, a, idx);
  const FilterConsts<T, Model::NB, Model::P, Model::M>& C = pc.C;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  Buf buf{reinterpret_cast<T*>(rodeo_dyn_smem), (int)threadIdx.x};
  F f;
  f.init(a.ode_init + idx * NB * P);
  f.load_scale(a, idx);
  // History: one checkpoint per shared-memory segment, the rest re-run (KC == K), as solve_mv does.  The sampling
  // solvers are bound by HBM traffic too -- storing and re-loading every filtered state is 2 x 8 x NSTATE bytes per
  // theta*step against 8 x NB x P of output -- and re-running a forward step reproduces its chkrebtii draw exactly
  // because the random numbers are counter-based.
  constexpr int KC = RODEO_SIM_T_RECOMPUTE ? K : 1;
  forward_with_checkpoints<T, Model, INTERR, QK, KC>(C, a, q, idx, live, f, stash, ldb, hook);

  auto normals = [&](int n, T (&z)[NB * P]) {
    if (z_smooth != nullptr) {
      const T* zp = z_smooth + (idx * (i64)(N + 1) + n) * (NB * P);
      RD_UNROLL for (int k = 0; k < NB * P; ++k) z[k] = zp[k];
    } else {
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        T zb[P];
        philox_normals<T, P>(a.key0, a.key1, a.particle_offset + idx, n, TAG_SMOOTH, zb, b * ((P + 3) / 4));
        RD_UNROLL for (int k = 0; k < P; ++k) z[b * P + k] = zb[k];
      }
    }
  };
  typedef typename F::MT MT;
  auto draw = [&](int b, const MT (&m)[P], const T (&Cv)[NS], const T (&z)[NB * P], MT (&x)[NB][P]) {
    T A[P][P];
    psd_factor<T, P>(Cv, A);
    RD_UNROLL for (int i = 0; i < P; ++i) {
      MT acc = m[i];
      RD_UNROLL for (int k = 0; k <= i; ++k) acc = rd_fma((MT)A[i][k], (MT)z[b * P + k], acc);
      x[b][i] = acc;
    }
  };

  // terminal draw from N(mu_f[N], S_f[N])  (solve.py:182-186)
  MT x[NB][P];
  {
    T z[NB * P];
    normals(N, z);
    MT xn[NB][P];
    RD_UNROLL for (int b = 0; b < NB; ++b) draw(b, f.mu[b], f.S[b], z, xn);
    RD_UNROLL for (int b = 0; b < NB; ++b)
      RD_UNROLL for (int i = 0; i < P; ++i) x[b][i] = xn[b][i];
    if (live && x_out != nullptr) store_mean_row<T, NB, P>(x_out + (idx * (i64)(N + 1) + N) * (NB * P), x);
  }
  SimLoglikAcc<T, MT> la;
  la.init(sl, N + 1);
  auto x_of = [&](int bb) {                                // x_t[bb][0] with bb a run-time block index
    MT v = x[0][0];
    RD_UNROLL for (int c = 1; c < NB; ++c) v = bb == c ? x[c][0] : v;
    return v;
  };
  la.row(sl, N, NB, 0, NB, x_of);

  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    rebuild_segment<T, Model, INTERR, QK, KC>(C, a, q, idx, j, cnt, f, stash, ldb, buf, hook);
    for (int s = cnt - 1; s >= 0; --s) {
      if (n0 + s == 0) {                                   // row 0 = ode_init: x0 is known, not sampled (solve.py:202-204)
        la.row(sl, 0, NB, 0, NB, [&](int bb) { return (MT)a.ode_init[idx * NB * P + bb * P]; });
        break;
      }
      buf.get(s, f.mu, f.S);                               // filt[n]
      T z[NB * P];
      normals(n0 + s, z);
      MT xn[NB][P];
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        MT mp[P], m[P];
        T Sp[NS], G[P][P], Ct[P][P], Cv[NS];
        predict<T, P, QK>(C.Q[b], C.R[b], f.rs[b], f.mu[b], f.S[b], mp, Sp);          // pred[n+1]
        smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
        // m = mu_f + G (x' - mu_p) ;  C = S_f - G (S_f Q^T)^T      (standard.py:251-254)
        RD_UNROLL for (int i = 0; i < P; ++i) {
          MT acc = f.mu[b][i];
          RD_UNROLL for (int jj = 0; jj < P; ++jj) acc = rd_fma((MT)G[i][jj], x[b][jj] - mp[jj], acc);
          m[i] = acc;
        }
        cond_var<T, P>(f.S[b], G, Ct, Cv);
        draw(b, m, Cv, z, xn);
      }
      RD_UNROLL for (int b = 0; b < NB; ++b)
        RD_UNROLL for (int i = 0; i < P; ++i) { x[b][i] = xn[b][i]; buf.at(s, b * P + i) = (T)xn[b][i]; }
      la.row(sl, n0 + s, NB, 0, NB, x_of);
    }
    if (j > 0) prefetch_segment<T, F, KC>(stash, ldb, idx, j - 1, K);
    __syncwarp();
    if (x_out != nullptr) buf.template copy_out<false>(x_out, theta0, a.B, N + 1, n0, cnt);
    __syncwarp();
  }
  if (sl.out != nullptr && live) sl.out[idx] = (T)la.ll;
}

// ------------------------------------------------------------------------------------------------------------------
// fenrir: forward filter, then a Kalman filter on the backward Markov chain with the Gaussian observations
// (reference src/rodeo/inference/fenrir.py:86-328)
// ------------------------------------------------------------------------------------------------------------------
#ifndef RODEO_FENRIR_MINB
#define RODEO_FENRIR_MINB 1
#endif
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32, RODEO_FENRIR_MINB)
fenrir_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> Cg,
              const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ stash, i64 ldb,
              T* __restrict__ loglik) {
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef SegBuf<T, F> Buf;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, K = Buf::K;
  i64 idx = (i64)blockIdx.x * 32 + threadIdx.x;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const PriorConsts<T, Model::NB, Model::P, Model::M, QK> pc(Cg, a, idx);
  const FilterConsts<T, Model::NB, Model::P, Model::M>& C = pc.C;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  // two history buffers: while segment j is processed out of one, segment j-1 travels HBM -> shared memory into the
  // other (cp.async), so the serial chain of a theta never waits for its own history
  Buf bufs[2] = {Buf{reinterpret_cast<T*>(rodeo_dyn_smem), (int)threadIdx.x},
                 Buf{reinterpret_cast<T*>(rodeo_dyn_smem) + Buf::BYTES / (int)sizeof(T), (int)threadIdx.x}};
  F f;
  f.init(a.ode_init + idx * NB * P);
  f.load_scale(a, idx);
  forward_with_checkpoints<T, Model, INTERR, QK, 1>(C, a, q, idx, live, f, stash, ldb);
  __threadfence();                                        // the history stores above precede the asynchronous reads below
  {
    const int jt = (N - 1) / K;
    segment_fetch_async<T, F>(stash, ldb, idx, jt, (N - jt * K) < K ? (N - jt * K) : K, bufs[jt & 1]);
  }

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
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K;
    const int cnt = (N - n0) < K ? (N - n0) : K;
    Buf& buf = bufs[j & 1];
    cp_async_wait_all();                                  // segment j has landed
    if (j > 0) segment_fetch_async<T, F>(stash, ldb, idx, j - 1, K, bufs[(j - 1) & 1]);
    else { f.init(a.ode_init + idx * NB * P); buf.put(0, f.mu, f.S); }      // filt[0] = (ode_init, 0)
    for (int s = cnt - 1; s >= 0; --s) {
      const int t = n0 + s;
      buf.get(s, f.mu, f.S);                              // filt[t]  (t = 0: (ode_init, 0))
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        typename F::MT mp[P], nm[P];
        T Sp[NS], G[P][P], Ct[P][P], Cv[NS];
        predict<T, P, QK>(C.Q[b], C.R[b], f.rs[b], f.mu[b], f.S[b], mp, Sp);          // pred[t+1]
        smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
        cond_var<T, P>(f.S[b], G, Ct, Cv);
        // backward chain X_t = A X_{t+1} + bvec + N(0, Cv), A = G, bvec = mu_f - G mu_p   (standard.py:366-370)
        // predict the backward filter through it (fenrir.py:151-157)
        RD_UNROLL for (int r = 0; r < P; ++r) {
          typename F::MT acc2 = f.mu[b][r];
          RD_UNROLL for (int jj = 0; jj < P; ++jj)
            acc2 = rd_fma((typename F::MT)G[r][jj], bk.mu[b][jj] - mp[jj], acc2);
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
      acc.ld.renorm();
    }
    __syncwarp();
  }
  if (live) loglik[idx] = (T)acc.value();
}


// ------------------------------------------------------------------------------------------------------------------
// fenrir, warp-specialised backward sweep
// ------------------------------------------------------------------------------------------------------------------
// BASELINE configs[3] (p = 4, 16,384 thetas = 512 warps for 592 SM sub-partitions) runs fenrir_kernel as ONE warp per
// sub-partition: its time is the serial chain of a theta.  ncu (profiles/r02_fenrir_before_ncu.txt): a backward step
// issues ~300 FP64 instructions (2 cycles each) behind a 4-pivot LDL^T whose reciprocals depend on one another, and 23 %
// of the samples sit in the history fetch.  But only a small part of a backward step is sequential in t: the parameters
// of the backward Markov chain  X_t = A_t X_{t+1} + b_t + N(0, C_t)  (smooth_cond, standard.py:366-370) depend on
// filt[t] alone, which the forward sweep has already written.  So the CTA is three warps over the same 32 thetas:
//   warp 0      forward filter (history to HBM), then the CONSUMER of the backward sweep: predicts the backward filter
//               through (A_t, b_t, C_t), applies the observations, accumulates the log-density (fenrir.py:151-179);
//   warps 1..NP PRODUCERS: producer p takes the steps t = N-1-p, N-1-p-NP, ...: load filt[t] (two of its steps ahead),
//               predict, LDL^T gain, conditional variance -> (A_t, mu_f, mu_p, C_t) into a 2 NP-slot shared-memory ring.
// Hand-over by named barriers (bar.arrive / bar.sync with 64 participants: one producer warp + the consumer), the
// producer / consumer pattern of the PTX manual.  NP = 2: measured on BASELINE configs[3] (B200, 16,384 thetas, N = 2,000)
// fenrir_kernel 2.65 ms, NP = 2: 2.01 ms, NP = 3: 2.87 ms (164 registers x 128 threads leave 3 CTAs per SM: two waves)
// or 2.51 ms capped at 128 registers (spills).  What is left: the forward sweep is one warp per CTA and costs ~670
// cycles per step against 320 issue cycles (the double-precision sin of the right-hand side, the history stores'
// address arithmetic), and a produced step costs ~1,750 cycles (476 issue cycles + FP64-pipe contention between the
// 10 warps an SM now holds) against ~300 for the consumer's.  The arithmetic is the same __device__ functions in the same order as
// fenrir_kernel: the log-likelihoods are bitwise identical.
#ifndef RODEO_FENRIR_NP
#define RODEO_FENRIR_NP 2
#endif
// Tuning knobs, measured on BASELINE configs[3] with the forcing buffer in place (B200): NP = 2, SL = 2, PF = 2 (the
// defaults, 168 registers) 1.97 ms; NP = 2, PF = 1 (148 registers) 2.07 ms; NP = 3, SL = 1, PF = 1 (128 registers, 40 B
// spilled, 4 CTAs of 4 warps per SM) 2.18 ms; NP = 3, SL = 1, PF = 2 (128 registers, 344 B spilled) 2.35 ms.  A third
// producer does not help although the consumer waits for producers 38 % of the time: the four warps a sub-partition
// then holds all compete for its one FP64 pipe ("math pipe throttle" is the producers' top stall reason).
#ifndef RODEO_FENRIR_SL
#define RODEO_FENRIR_SL 2          // ring slots per producer
#endif
#ifndef RODEO_FENRIR_PF
#define RODEO_FENRIR_PF 2          // history states a producer holds in registers ahead of the one it works on (1 or 2)
#endif
#ifndef RODEO_FENRIR_FCH
#define RODEO_FENRIR_FCH 16        // steps per chunk of the forcing buffer (models with Forcing<Model>::HAS)
#endif
// 4 CTAs per SM (BASELINE configs[3] needs 3.46 to hold its 512 CTAs in one wave): 170 registers at 96 threads.  Without
// the bound ptxas takes 174 and the launch falls into two waves (measured 3.11 ms instead of 2.01)
// (models with more than 20 state entries per theta would spill under that bound and keep ptxas's own choice)
#ifndef RODEO_FENRIR_WS_MINB
#define RODEO_FENRIR_WS_MINB(NSTATE) ((RODEO_FENRIR_NP <= 3 && (NSTATE) <= 20) ? 4 : 1)
#endif
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32 * (1 + RODEO_FENRIR_NP),
                                  RODEO_FENRIR_WS_MINB(Model::NB * (Model::P + Model::P * (Model::P + 1) / 2)))
fenrir_ws_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                 const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ stash, i64 ldb,
                 T* __restrict__ loglik) {
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef typename F::MT MT;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, NSTATE = NB * (P + NS);
  static_assert(sizeof(T) == 8, "the ring stages the means as T: float64 only");
  constexpr int NCH = NB * (P * P + P + NS);             // per theta and step in the ring: A_t, mu_f, C_t per block ...
  constexpr int NP = RODEO_FENRIR_NP, SL = RODEO_FENRIR_SL, RING = SL * NP;   // ... followed by a second region with mu_p: RING * NB * P
  static_assert(2 * RING + 1 <= 16, "named barriers");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  i64 idx = (i64)blockIdx.x * 32 + lane;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const int N = a.n_steps;
  T* ring = reinterpret_cast<T*>(rodeo_dyn_smem);        // [slot][k][lane]
  auto slot_of = [&](int t) { return (N - 1 - t) % RING; };
  auto bar_sync = [](int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); };
  auto bar_arrive = [](int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); };
  // barrier ids: 1 + slot = "slot is full", 1 + RING + slot = "slot is empty"

  F f;
  F bk;                                                   // backward-filter state (consumer)
  f.load_scale(a, idx);
  // Models with a state-independent forcing term (Forcing<Model>, rodeo_models.cuh): during the forward sweep the first
  // producer warp, otherwise idle, evaluates forcing(t_n) -- a double-precision sin with its table loads for the
  // second-order example, 40 % of the forward warp's time when it computed it itself -- into a shared-memory buffer
  // behind the ring, NFB chunks of FCH steps.  Measured on BASELINE configs[3]: 2.19 -> 1.95 ms; two helper warps and
  // four chunks give the same time (the forward warp then waits for 16 % instead of 43 % of the sweep, but the three
  // warps of a sub-partition share its FP64 pipe and the filter step itself slows down by as much).
  constexpr bool FRC = Forcing<Model>::HAS;
  constexpr int FCH = RODEO_FENRIR_FCH, NFB = 2;          // chunk c lives in buffer c % NFB
  static_assert(NFB <= RING, "the chunk barriers reuse the ring's ids (the ring is idle during the forward sweep)");
  MT* frc_buf = reinterpret_cast<MT*>(ring + (i64)RING * (NCH + NB * P) * 32);      // [NFB][FCH][32]
  // chunk barriers: 1 + buffer = "full", 1 + RING + buffer = "consumed"; every arrive is matched by a sync, so the ids are
  // back in their initial state when the backward sweep starts using them for the ring
  if constexpr (FRC) {
    if (warp == 1) {
      const typename F::Par q1 = load_par<Model, T>(a.theta + idx * Model::NTHETA);
      for (int c = 0; c * FCH < N; ++c) {
        if (c >= NFB) bar_sync(1 + RING + c % NFB);                                   // chunk c - NFB has been consumed
        MT* dst = frc_buf + (c % NFB) * FCH * 32 + lane;
        for (int s = 0; s < FCH && c * FCH + s < N; ++s)
          dst[s * 32] = Forcing<Model>::template eval<MT>(q1, step_time<MT>(a.t_min, a.t_max, c * FCH + s, N));
        bar_arrive(1 + c % NFB);
      }
    }
  }
  if (warp == 0) {
    const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
    f.init(a.ode_init + idx * NB * P);
    if constexpr (FRC) {
      const int nch = (N + FCH - 1) / FCH;
      MT t_next = step_time<MT>(a.t_min, a.t_max, 0, N);
      for (int c = 0; c < nch; ++c) {
        bar_sync(1 + c % NFB);
        const MT* src = frc_buf + (c % NFB) * FCH * 32 + lane;
        const int n_hi = (c + 1) * FCH < N ? (c + 1) * FCH : N;
        for (int n = c * FCH; n < n_hi; ++n) {
          const MT t_now = t_next;
          t_next = step_time<MT>(a.t_min, a.t_max, n + 1, N);
          const MT frc = src[(n - c * FCH) * 32];
          forward_step<T, Model, INTERR, QK>(C, a, q, idx, n, f, nullptr, &t_now, &frc);
          if (live && n + 1 < N) ckpt_store<T, F>(stash, ldb, idx, n + 1, f);
        }
        if (c + NFB < nch) bar_arrive(1 + RING + c % NFB);
      }
    } else {
      forward_with_checkpoints<T, Model, INTERR, QK, 1>(C, a, q, idx, live, f, stash, ldb);
    }
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      RD_UNROLL for (int i = 0; i < P; ++i) bk.mu[b][i] = f.mu[b][i];
      RD_UNROLL for (int k = 0; k < NS; ++k) bk.S[b][k] = f.S[b][k];
    }
  }
  __syncthreads();                                        // the history of this CTA's thetas is complete and visible

  if (warp > 0) {
    // ---- producers
    const int pid = warp - 1;
    auto load_filt = [&](int t, MT (&mu)[NB][P], T (&S)[NB][NS]) {
      if (t >= 1) {
        const T* s = stash + (i64)(t - 1) * NSTATE * ldb + idx;
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = (MT)s[(i64)(b * P + i) * ldb];
          RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = s[(i64)(NB * P + b * NS + k) * ldb];
        }
      } else {                                            // filt[0] = (ode_init, 0)
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          RD_UNROLL for (int i = 0; i < P; ++i) mu[b][i] = (MT)a.ode_init[idx * NB * P + b * P + i];
          RD_UNROLL for (int k = 0; k < NS; ++k) S[b][k] = T(0);
        }
      }
    };
    // software pipeline over this producer's steps: PF loads in flight ahead of the step being computed
    constexpr int PF = RODEO_FENRIR_PF;
    static_assert(PF == 1 || PF == 2, "producer prefetch depth");
    MT nmu[PF][NB][P];
    T nS[PF][NB][NS];
    int t = N - 1 - pid;
    if (t >= 0) load_filt(t, nmu[0], nS[0]);
    if (PF == 2 && t - NP >= 0) load_filt(t - NP, nmu[PF - 1], nS[PF - 1]);
    int produced = 0;
    for (; t >= 0; t -= NP) {
      const int cur = PF == 2 ? (produced & 1) : 0;
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        RD_UNROLL for (int i = 0; i < P; ++i) f.mu[b][i] = cur ? nmu[PF - 1][b][i] : nmu[0][b][i];
        RD_UNROLL for (int k = 0; k < NS; ++k) f.S[b][k] = cur ? nS[PF - 1][b][k] : nS[0][b][k];
      }
      if (t - PF * NP >= 0) {                              // refill the buffer just consumed
        if (cur) load_filt(t - PF * NP, nmu[PF - 1], nS[PF - 1]);
        else load_filt(t - PF * NP, nmu[0], nS[0]);
      }
      if (PF == 1 && t - 2 * NP >= 0) {                    // the load after that one: requested into L2
        const T* sp = stash + (i64)(t - 2 * NP - 1) * NSTATE * ldb + idx;
        if (t - 2 * NP >= 1)
          RD_UNROLL for (int k = 0; k < NSTATE; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (i64)k * ldb));
      }
      const int slot = slot_of(t);
      if (produced >= SL) bar_sync(1 + RING + slot);       // the consumer has released this slot (each producer owns SL)
      ++produced;
      T* r = ring + (i64)slot * NCH * 32 + lane;
      RD_UNROLL for (int b = 0; b < NB; ++b) {
        MT mp[P];
        T Sp[NS], G[P][P], Ct[P][P], Cv[NS];
        predict<T, P, QK>(C.Q[b], C.R[b], f.rs[b], f.mu[b], f.S[b], mp, Sp);          // pred[t+1]
        smooth_gain<T, P, QK>(C.Q[b], f.S[b], Sp, G, Ct);
        cond_var<T, P>(f.S[b], G, Ct, Cv);
        T* rb = r + (i64)b * (P * P + P + NS) * 32;
        RD_UNROLL for (int i = 0; i < P; ++i)
          RD_UNROLL for (int j = 0; j < P; ++j) rb[(i * P + j) * 32] = G[i][j];
        // the consumer forms mu_f + G (m - mu_p) exactly as fenrir_kernel does: it gets mu_f and mu_p rather than b_t
        RD_UNROLL for (int i = 0; i < P; ++i) rb[(P * P + i) * 32] = (T)f.mu[b][i];
        RD_UNROLL for (int k = 0; k < NS; ++k) rb[(P * P + P + k) * 32] = Cv[k];
        T* rp = ring + (i64)RING * NCH * 32 + ((i64)slot * NB * P + b * P) * 32 + lane;
        RD_UNROLL for (int i = 0; i < P; ++i) rp[i * 32] = (T)mp[i];
      }
      bar_arrive(1 + slot);
    }
    return;
  }

  // ---- consumer
  LogPdfAcc<T> acc;
  acc.init();
  int i = o.n_obs - 1;
  auto obs_at = [&](int k) { return __ldg(o.obs_ind + (k < 0 ? k + o.n_obs : k)); };
  if (obs_at(i) >= N) {                                   // terminal point update (fenrir.py:196-220)
    bk.template update_y<NOBS>(o, i, acc);
    --i;
  }
  int next_obs = obs_at(i);
  for (int t = N - 1; t >= 0; --t) {
    const int slot = slot_of(t);
    bar_sync(1 + slot);
    const T* r = ring + (i64)slot * NCH * 32 + lane;
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      const T* rb = r + (i64)b * (P * P + P + NS) * 32;
      const T* rp = ring + (i64)RING * NCH * 32 + ((i64)slot * NB * P + b * P) * 32 + lane;
      T G[P][P], Cv[NS];
      MT mf[P], mp[P], nm[P];
      RD_UNROLL for (int ii = 0; ii < P; ++ii)
        RD_UNROLL for (int j = 0; j < P; ++j) G[ii][j] = rb[(ii * P + j) * 32];
      RD_UNROLL for (int ii = 0; ii < P; ++ii) { mf[ii] = (MT)rb[(P * P + ii) * 32]; mp[ii] = (MT)rp[ii * 32]; }
      RD_UNROLL for (int k = 0; k < NS; ++k) Cv[k] = rb[(P * P + P + k) * 32];
      // predict the backward filter through the chain (fenrir.py:151-157): same expressions as fenrir_kernel
      RD_UNROLL for (int rr = 0; rr < P; ++rr) {
        MT acc2 = mf[rr];
        RD_UNROLL for (int jj = 0; jj < P; ++jj) acc2 = rd_fma((MT)G[rr][jj], bk.mu[b][jj] - mp[jj], acc2);
        nm[rr] = acc2;
      }
      add_GDGt<T, P>(G, bk.S[b], Cv);
      RD_UNROLL for (int rr = 0; rr < P; ++rr) bk.mu[b][rr] = nm[rr];
      RD_UNROLL for (int k = 0; k < NS; ++k) bk.S[b][k] = Cv[k];
    }
    bar_arrive(1 + RING + slot);                          // the slot may be refilled
    if (next_obs == t) {
      bk.template update_y<NOBS>(o, i < 0 ? i + o.n_obs : i, acc);
      --i;
      next_obs = obs_at(i);
    }
    acc.ld.renorm();
  }
  if (live) loglik[idx] = (T)acc.value();
}

// ------------------------------------------------------------------------------------------------------------------
// fenrir.solve_mv: posterior mean / variance p(X_{0:N} | Z_{1:N}, Y_{0:M}) by the Fenrir construction
// (reference src/rodeo/inference/fenrir.py:86-259 `_backward` stacks, :333-401 `_smooth_mv`, :404-457 `solve_mv`)
// ------------------------------------------------------------------------------------------------------------------
// Three sweeps per theta, thread per theta:
//   1. forward ODE filter, filt[n] kept in history H1;
//   2. backward filter over the smoothing Markov chain X_t = A_t X_{t+1} + b_t + N(0, C_t) with the observations,
//      its filtered states bfilt[t] kept in history H2 (rows 0 and 1 of the output are bfilt[0], bfilt[1]);
//   3. RTS pass over that chain forward in time: row k+1 from row k, bfilt[k+1] and bpred[k], where
//      (A_k, b_k, C_k) and bpred[k] = predict(bfilt[k+1]; A_k, b_k, C_k) are recomputed from filt[k] instead of stored.
template <typename T, class Model, int INTERR, int QK, int NOBS>
__global__ void __launch_bounds__(32)
fenrir_solve_mv_kernel(const __grid_constant__ FilterConsts<T, Model::NB, Model::P, Model::M> C,
                       const CommonArgs<T> a, const ObsArgs<T> o, T* __restrict__ h1, T* __restrict__ h2, i64 ldb,
                       T* __restrict__ mean_out, T* __restrict__ var_out) {
  typedef Fwd<T, Model, INTERR, QK> F;
  typedef SegBuf<T, F> Buf;
  constexpr int NB = F::NB, P = F::P, NS = F::NS, K = Buf::K;
  const i64 theta0 = (i64)blockIdx.x * 32;
  i64 idx = theta0 + threadIdx.x;
  const bool live = idx < a.B;
  if (!live) idx = a.B - 1;
  const typename F::Par q = load_par<Model, T>(a.theta + idx * Model::NTHETA);
  const int N = a.n_steps;
  Buf buf{reinterpret_cast<T*>(rodeo_dyn_smem), (int)threadIdx.x};
  F f;
  f.init(a.ode_init + idx * NB * P);
  f.load_scale(a, idx);
  forward_with_checkpoints<T, Model, INTERR, QK, 1>(C, a, q, idx, live, f, h1, ldb);     // H1 entry n = filt[n], 1..N-1

  // the chain parameters of step t from filt[t]:  A = G, bvec = mu_f - G mu_p, Cv = S_f - G (S_f Q^T)^T
  auto chain = [&](int b, const F& fl, T (&G)[P][P], T (&mp)[P], T (&Cv)[NS]) {
    T Sp[NS], Ct[P][P];
    predict<T, P, QK>(C.Q[b], C.R[b], fl.rs[b], fl.mu[b], fl.S[b], mp, Sp);
    smooth_gain<T, P, QK>(C.Q[b], fl.S[b], Sp, G, Ct);
    cond_var<T, P>(fl.S[b], G, Ct, Cv);
  };
  auto load_filt = [&](int t, F& fl) {
    if (t >= 1) ckpt_load<T, F>(h1, ldb, idx, t, fl);
    else fl.init(a.ode_init + idx * NB * P);
  };

  // ---- sweep 2: backward filter, bfilt[t] -> H2 entry t (t = 1..N)
  F bk;
  bk.load_scale(a, idx);
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    RD_UNROLL for (int i = 0; i < P; ++i) bk.mu[b][i] = f.mu[b][i];
    RD_UNROLL for (int k = 0; k < NS; ++k) bk.S[b][k] = f.S[b][k];
  }
  LogPdfAcc<T> acc;
  acc.init();
  int i = o.n_obs - 1;
  auto obs_at = [&](int k) { return __ldg(o.obs_ind + (k < 0 ? k + o.n_obs : k)); };
  if (obs_at(i) >= N) { bk.template update_y<NOBS>(o, i, acc); --i; }
  if (live) ckpt_store<T, F>(h2, ldb, idx, N, bk);
  int next_obs = obs_at(i);
  for (int t = N - 1; t >= 0; --t) {
    load_filt(t, f);
    RD_UNROLL for (int b = 0; b < NB; ++b) {
      T G[P][P], mp[P], Cv[NS], nm[P];
      chain(b, f, G, mp, Cv);
      RD_UNROLL for (int r = 0; r < P; ++r) {
        T s = f.mu[b][r];
        RD_UNROLL for (int jj = 0; jj < P; ++jj) s = rd_fma(G[r][jj], bk.mu[b][jj] - mp[jj], s);
        nm[r] = s;
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
    acc.ld.renorm();
    if (t >= 1) { if (live) ckpt_store<T, F>(h2, ldb, idx, t, bk); }
    else if (live) {                                       // row 0 = bfilt[0]
      store_mean_row<T, NB, P>(mean_out + idx * (i64)(N + 1) * (NB * P), bk.mu);
      store_var_row<T, NB, P>(var_out + idx * (i64)(N + 1) * (NB * P * P), bk.S);
    }
  }

  // ---- sweep 3: forward-in-time RTS over the backward chain; rows 1..N staged K at a time
  T ms[NB][P], Ss[NB][NS];
  F bf;                                                    // bfilt[k+1]
  for (int n0 = 1; n0 <= N; n0 += K) {
    const int cnt = (N + 1 - n0) < K ? (N + 1 - n0) : K;
    for (int s = 0; s < cnt; ++s) {
      const int row = n0 + s;
      ckpt_load<T, F>(h2, ldb, idx, row, bf);
      if (row == 1) {                                      // row 1 = bfilt[1]  (fenrir.py:373-376, 394-399)
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          RD_UNROLL for (int i2 = 0; i2 < P; ++i2) ms[b][i2] = bf.mu[b][i2];
          RD_UNROLL for (int k = 0; k < NS; ++k) Ss[b][k] = bf.S[b][k];
        }
      } else {
        const int k = row - 1;                             // chain step X_k = A_k X_{k+1} + ...
        load_filt(k, f);
        RD_UNROLL for (int b = 0; b < NB; ++b) {
          T A[P][P], mp[P], Cv[NS], bpm[P];
          chain(b, f, A, mp, Cv);
          // bpred[k] = predict(bfilt[k+1]; A, bvec, Cv)
          RD_UNROLL for (int r = 0; r < P; ++r) {
            T sacc = f.mu[b][r];
            RD_UNROLL for (int jj = 0; jj < P; ++jj) sacc = rd_fma(A[r][jj], bf.mu[b][jj] - mp[jj], sacc);
            bpm[r] = sacc;
          }
          add_GDGt<T, P>(A, bf.S[b], Cv);                  // Cv <- bvar_pred[k]
          // smooth_mv(next = row k, filt = bfilt[k+1], pred = bpred[k], wgt_state = A)   (standard.py:210-216)
          //   G2 = bS_f[k+1] A^T bS_p[k]^{-1}
          T Ct[P][P], G2[P][P], L[P][P], rD[P];
          RD_UNROLL for (int r = 0; r < P; ++r)
            RD_UNROLL for (int c = 0; c < P; ++c) {
              T sacc = T(0);
              RD_UNROLL for (int kk = 0; kk < P; ++kk) sacc = rd_fma(bf.S[b][sym<P>(r, kk)], A[c][kk], sacc);
              Ct[r][c] = sacc;
            }
          ldlt<T, P>(Cv, L, rD);
          RD_UNROLL for (int r = 0; r < P; ++r) {
            T x[P];
            RD_UNROLL for (int c = 0; c < P; ++c) x[c] = Ct[r][c];
            ldlt_solve<T, P>(L, rD, x);
            RD_UNROLL for (int c = 0; c < P; ++c) G2[r][c] = x[c];
          }
          T dm[P], D[NS];
          RD_UNROLL for (int r = 0; r < P; ++r) dm[r] = ms[b][r] - bpm[r];
          RD_UNROLL for (int kk = 0; kk < NS; ++kk) D[kk] = Ss[b][kk] - Cv[kk];
          RD_UNROLL for (int r = 0; r < P; ++r) {
            T m = bf.mu[b][r];
            RD_UNROLL for (int c = 0; c < P; ++c) m = rd_fma(G2[r][c], dm[c], m);
            ms[b][r] = m;
          }
          RD_UNROLL for (int kk = 0; kk < NS; ++kk) Ss[b][kk] = bf.S[b][kk];
          add_GDGt<T, P>(G2, D, Ss[b]);
        }
      }
      buf.put(s, ms, Ss);
    }
    __syncwarp();
    buf.template copy_out<false>(mean_out, theta0, a.B, N + 1, n0, cnt);
    buf.template copy_out<true>(var_out, theta0, a.B, N + 1, n0, cnt);
    __syncwarp();
  }
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
