/*
 * ude_oracle.h -- CPU restatement of the reference's UDE hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path
 * (universal_differential_equations_b200/, include/, the C-ABI .so) may
 * include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs (as the checker / the
 * reported CPU baseline, never as the thing shipped).
 *
 * What is restated (reference = ChrisRackauckas/universal_differential_equations
 * @ d622d92; file:line relative to /root/reference):
 *   - UDE right-hand sides
 *       LV      LotkaVolterra/scenario_1.jl:69-76, scenario_2.jl:90-98,
 *               hudson_bay.jl:85-91
 *       SEIR    SEIR_exposure/seir_exposure.jl:117-130 (constants :33)
 *       FKPP    FisherKPP/Fisher-KPP-CNN.jl:111-126, LotkaVolterra/scenario_3.jl:103-114
 *   - embedded dense chains (Lux.Chain / FastChain / Flux.Chain) and their flat
 *     parameter layout: scenario_1.jl:59-66, seir_exposure.jl:114-115,
 *     hudson_bay.jl:77-82, Fisher-KPP-CNN.jl:92-109
 *   - solve(prob, Tsit5()/Vern7(); saveat, abstol, reltol) as called at
 *     scenario_1.jl:82-88, seir_exposure.jl:137-141, Fisher-KPP-CNN.jl:136
 *   - sensealg = InterpolatingAdjoint(autojacvec=ReverseDiffVJP())
 *     seir_exposure.jl:71,140; Fisher-KPP-CNN.jl:136
 *   - L2 losses scenario_1.jl:91-94, seir_exposure.jl:144-147
 *
 * The arithmetic of those calls lives in third-party Julia packages that are
 * NOT vendored in the reference (only pinned in each directory's Manifest.toml):
 *   OrdinaryDiffEq 6.19.2 / 5.32.0 / 5.43.0  (Tsit5, Vern7 perform_step!, PI controller)
 *   DiffEqSensitivity 6.79.0 / 6.9.0 / 6.33.0 (InterpolatingAdjoint)
 *   Lux 0.4.11, DiffEqFlux 1.7.1/1.24.0, Flux 0.10.3/0.11.1 (Dense layers)
 * Their published algorithms are restated here: Tsitouras 5(4) (Tsitouras
 * 2011) with its free 4th-order interpolant, Verner's "most efficient" 7(6)
 * pair, and the continuous adjoint with discrete loss jumps.  The tableau
 * constants are checked against the copy OrdinaryDiffEq itself serialized
 * into the reference's LotkaVolterra/results/Scenario_1_recovery_0.005.jld2
 * (tests/golden/scenario_1.npz: tsit5_consts, vern7_consts).
 *
 * Pinning status: Julia is absent from this image, so the reference cannot be
 * executed.  The oracle is pinned against the artefacts the reference commits
 * (KAT-1..8 in SURVEY.md section 8c; tests/test_oracle_golden.py): MLP forward,
 * forward solves, loss values, and -- through an ADAM replay of the stored
 * loss history -- the parameter gradient, all for the Lotka-Volterra family.
 * InterpolatingAdjoint numerics for SEIR / Fisher-KPP have no committed
 * artefact: for those models PARITY IS UNPINNED (oracle-vs-autograd only).
 */
#ifndef UDE_ORACLE_H
#define UDE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { UDE_MODEL_LV = 0, UDE_MODEL_SEIR = 1, UDE_MODEL_FKPP = 2, UDE_MODEL_NODE = 3, UDE_MODEL_SEIR_NODE = 4 };
enum { UDE_ACT_IDENTITY = 0, UDE_ACT_TANH = 1, UDE_ACT_RBF = 2 };
enum { UDE_TSIT5 = 0, UDE_VERN7 = 1 };

#define UDE_MAX_LAYERS 6
#define UDE_MAX_WIDTH 128
#define UDE_MAX_STATE 256

typedef struct {
    int32_t model;                       /* UDE_MODEL_* */
    int32_t d;                           /* state dimension (2 LV, 7 SEIR, Nx FKPP) */
    int32_t n_layers;                    /* dense layers in the embedded chain */
    int32_t widths[UDE_MAX_LAYERS + 1];  /* widths[0] = in, widths[n_layers] = out */
    int32_t acts[UDE_MAX_LAYERS];        /* activation of each layer */
    int32_t n_prefix;                    /* trainable physics scalars stored before the chain in theta */
    int32_t n_suffix;                    /* trainable scalars after the chain (FKPP: w1,w2,w3,b,D0 = 5) */
    int32_t n_consts;
    double consts[16];                   /* fixed physics constants */
} ude_model;

/* number of entries of the flat parameter vector theta = [prefix | chain | suffix] */
size_t ude_num_params(const ude_model *m);

/* ---- double precision ---- */
void ude_mlp_forward_f64(const ude_model *m, const double *theta_chain, const double *x, double *y);
void ude_rhs_f64(const ude_model *m, const double *theta, const double *u, double *du);
/* dlam = (df/du)^T lam ;  gtheta += w * (df/dtheta)^T lam */
void ude_rhs_vjp_f64(const ude_model *m, const double *theta, const double *u, const double *lam,
                     double *dlam, double *gtheta, double w);
/* fixed-step solve, n_steps steps of size dt from t0; state saved every `save_every` steps
 * (including t0): out[(n_steps/save_every+1)][d].  dense (optional, may be NULL):
 * [n_steps][stages][d] stage derivatives k_i for the interpolant. returns 0 or -1 on non-finite. */
int ude_solve_fixed_f64(const ude_model *m, const double *theta, const double *u0, double dt,
                        int n_steps, int solver, int save_every, double *out, double *dense);
/* adaptive Tsit5/Vern7 with OrdinaryDiffEq's default PI controller; saveat[] strictly increasing,
 * saveat[0] = t0.  Tsit5 saves by interpolation (does not alter the steps), Vern7 steps onto the
 * save points (tstops).  returns number of accepted steps, or -1. */
int ude_solve_adaptive_f64(const ude_model *m, const double *theta, const double *u0,
                           const double *saveat, int n_save, double abstol, double reltol,
                           int solver, double *out, int *n_rejected);
/* continuous interpolating adjoint on the fixed grid of ude_solve_fixed (Tsit5 only):
 * dLdout[n_save][d] are the loss jumps.  grad_theta[P] is ACCUMULATED (+=), grad_u0[d] is set. */
void ude_adjoint_fixed_f64(const ude_model *m, const double *theta, const double *out,
                           const double *dense, double dt, int n_steps, int save_every,
                           const double *dLdout, double *grad_theta, double *grad_u0);
/* exact gradient of the discrete fixed-step Tsit5 scheme by reverse accumulation through the stages (what
 * ForwardDiffSensitivity, scenario_1.jl:86, computes in forward mode); same arguments as ude_adjoint_fixed */
void ude_adjoint_discrete_fixed_f64(const ude_model *m, const double *theta, const double *out,
                                    const double *dense, double dt, int n_steps, int save_every,
                                    const double *dLdout, double *grad_theta, double *grad_u0);
/* ensemble loss + gradient, trajectory-fastest layouts: u0[d][N], y[n_save][d][N], wmask[d]
 * L = sum_n sum_i sum_k wmask[k] (u_k(t_i) - y)^2.  out (optional) [n_save][d][N].
 * grad_theta[P] (set), grad_u0 (optional) [d][N].  OpenMP over trajectories. */
double ude_ensemble_loss_grad_f64(const ude_model *m, const double *theta, const double *u0,
                                  const double *y, const double *wmask, size_t N, double dt,
                                  int n_steps, int save_every, double *out, double *grad_theta,
                                  double *grad_u0, int n_threads);

/* RKC2 (second-order Runge-Kutta-Chebyshev, the closed-form ROCK2-class stabilised explicit method): fixed step, `stages`
 * stages per step; ustep / fstep [n_steps+1][d] (u_n and f(u_n)) feed the cubic-Hermite dense output of the adjoint, which
 * integrates lambda backwards with RKC2 itself.  grad_theta is ACCUMULATED. */
int ude_rkc2_coefficients(int s, double *mt1, double *mu, double *nu, double *mt, double *ga, double *c, double *w, double *beta);
int ude_solve_rkc2_fixed_f64(const ude_model *m, const double *theta, const double *u0, double dt, int n_steps, int stages,
                             int save_every, double *out, double *ustep, double *fstep);
void ude_adjoint_rkc2_fixed_f64(const ude_model *m, const double *theta, const double *ustep, const double *fstep, double dt,
                                int n_steps, int stages, int save_every, const double *dLdout, double *grad_theta, double *grad_u0);
int ude_solve_rkc2_fixed_f32(const ude_model *m, const float *theta, const float *u0, float dt, int n_steps, int stages,
                             int save_every, float *out, float *ustep, float *fstep);
void ude_adjoint_rkc2_fixed_f32(const ude_model *m, const float *theta, const float *ustep, const float *fstep, float dt,
                                int n_steps, int stages, int save_every, const float *dLdout, float *grad_theta, float *grad_u0);

/* ---- single precision (same semantics; state, parameters and arithmetic in float) ---- */
void ude_mlp_forward_f32(const ude_model *m, const float *theta_chain, const float *x, float *y);
void ude_rhs_f32(const ude_model *m, const float *theta, const float *u, float *du);
void ude_rhs_vjp_f32(const ude_model *m, const float *theta, const float *u, const float *lam,
                     float *dlam, float *gtheta, float w);
int ude_solve_fixed_f32(const ude_model *m, const float *theta, const float *u0, float dt,
                        int n_steps, int solver, int save_every, float *out, float *dense);
int ude_solve_adaptive_f32(const ude_model *m, const float *theta, const float *u0,
                           const float *saveat, int n_save, float abstol, float reltol,
                           int solver, float *out, int *n_rejected);
void ude_adjoint_fixed_f32(const ude_model *m, const float *theta, const float *out,
                           const float *dense, float dt, int n_steps, int save_every,
                           const float *dLdout, float *grad_theta, float *grad_u0);
void ude_adjoint_discrete_fixed_f32(const ude_model *m, const float *theta, const float *out,
                                    const float *dense, float dt, int n_steps, int save_every,
                                    const float *dLdout, float *grad_theta, float *grad_u0);
double ude_ensemble_loss_grad_f32(const ude_model *m, const float *theta, const float *u0,
                                  const float *y, const float *wmask, size_t N, float dt,
                                  int n_steps, int save_every, float *out, float *grad_theta,
                                  float *grad_u0, int n_threads);

/* adaptive Tsit5 forward that records every accepted step (tgrid[nacc+1], ustep[(nacc+1)*d], dense[nacc*7*d]);
 * returns nacc or -1.  The two adjoints below consume that record:
 *   replay   -- one backward Tsit5 step per accepted forward step, split at the save times (what the GPU does);
 *   adaptive -- reference-style error-controlled backward solve of [lambda; mu] with tstops at the save times. */
int ude_solve_adaptive_dense_f64(const ude_model *m, const double *theta, const double *u0, const double *saveat, int n_save,
                                  double abstol, double reltol, int max_steps, double *out, double *tgrid, double *ustep, double *dense);
void ude_adjoint_replay_f64(const ude_model *m, const double *theta, const double *saveat, int n_save, const double *tgrid,
                             const double *ustep, const double *dense, int nacc, const double *dLdout, double *grad_theta, double *grad_u0);
int ude_adjoint_adaptive_f64(const ude_model *m, const double *theta, const double *saveat, int n_save, const double *tgrid,
                              const double *ustep, const double *dense, int nacc, const double *dLdout, double abstol, double reltol,
                              double *grad_theta, double *grad_u0);
int ude_solve_adaptive_dense_f32(const ude_model *m, const float *theta, const float *u0, const float *saveat, int n_save,
                                  float abstol, float reltol, int max_steps, float *out, float *tgrid, float *ustep, float *dense);
void ude_adjoint_replay_f32(const ude_model *m, const float *theta, const float *saveat, int n_save, const float *tgrid,
                             const float *ustep, const float *dense, int nacc, const float *dLdout, float *grad_theta, float *grad_u0);
int ude_adjoint_adaptive_f32(const ude_model *m, const float *theta, const float *saveat, int n_save, const float *tgrid,
                              const float *ustep, const float *dense, int nacc, const float *dLdout, float abstol, float reltol,
                              float *grad_theta, float *grad_u0);

/* tableau access for the known-answer test against OrdinaryDiffEq's serialized constants */
void ude_tsit5_constants(double *c56);  /* same order as the .jld2 block: c1..c6,a21..a76,bt1..7,r11..r74 */
void ude_vern7_constants(double *c58);  /* c2..c8, nonzero a_ij row by row, b1,b4..b9, bt1,bt4..bt10 */

#ifdef __cplusplus
}
#endif
#endif
