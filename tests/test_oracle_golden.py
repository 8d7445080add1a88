"""Pins the CPU oracle against everything the reference commits for this path.

The reference has no test suite for the UDE solve/adjoint; its known answers are the result files
its own scripts wrote (tests/golden/*.npz, extracted by tools/make_golden.py).  Julia is not available,
so these artefacts -- not a live run of the reference -- are what pins the oracle.
"""
import numpy as np
import pytest

from helpers import theta_scenario1_init

RBF3 = ("rbf", "rbf", "rbf", "identity")


def test_tableaus_match_ordinarydiffeq_serialized_constants(O, golden):
    """KAT-7: Tsit5 and Vern7 constants equal the copy OrdinaryDiffEq serialized into the reference's .jld2."""
    g = golden["scenario_1"]
    assert np.array_equal(O.tsit5_constants(), g["tsit5_consts"][:56])
    assert np.array_equal(O.vern7_constants(), g["vern7_consts"][:58])
    c = O.tsit5_constants()
    b = c[21:27]
    cc = np.array([0.0, *c[0:5]])
    assert abs(b.sum() - 1) < 1e-15 and abs(b @ cc - 0.5) < 1e-15 and abs(b @ cc**2 - 1 / 3) < 1e-15
    assert abs(c[27:34].sum()) < 1e-15  # sum(btilde) = 0


def test_kat1_mlp_forward_scenario1(O, golden):
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    Y = np.stack([O.mlp_forward(m, g["theta_trained"], g["Xhat"][:, i]) for i in range(61)], 1)
    assert np.abs(Y - g["Yhat"]).max() < 1e-13


def test_kat2_forward_solve_scenario1(O, golden):
    """Stored X-hat = predict(theta_trained, Xn[:,1], 0:0.05:3) (reference: Vern7, tol 1e-6)."""
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    ts = np.linspace(0, 3, 61)
    for solver, tol in ((O.TSIT5, 1e-12), (O.VERN7, 1e-10)):
        out, nacc, nrej = O.solve_adaptive(m, g["theta_trained"], g["X"][:, 0], ts, tol, tol, solver=solver)
        assert np.abs(out.T - g["Xhat"]).max() < 5e-7
    out, _, _ = O.solve_adaptive(m, g["theta_trained"], g["X"][:, 0], ts, 1e-6, 1e-6, solver=O.VERN7)
    assert np.abs(out.T - g["Xhat"]).max() < 5e-7
    out = O.solve_fixed(m, g["theta_trained"], g["X"][:, 0], 0.05 / 4, 240, solver=O.TSIT5, save_every=4)
    assert np.abs(out.T - g["Xhat"]).max() < 5e-7


def _loss_grad_s1(O, m, th, X, sub=16, discrete=False):
    dt, ns = 0.1 / sub, 30 * sub
    out, dense = O.solve_fixed(m, th, X[:, 0], dt, ns, save_every=sub, want_dense=True)
    r = out - X.T
    adj = O.adjoint_discrete if discrete else O.adjoint_fixed
    gth, gu = adj(m, th, out, dense, dt, ns, 2 * r, save_every=sub)
    return (r**2).sum(), gth


def test_kat3_loss_values_scenario1(O, golden):
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    l0, _ = _loss_grad_s1(O, m, theta_scenario1_init(g), g["X"])
    l1, _ = _loss_grad_s1(O, m, g["theta_trained"], g["X"])
    assert abs(l0 - g["losses"][0]) < 1e-7 * g["losses"][0]
    assert abs(l1 - g["losses"][-1]) < 1e-5 * g["losses"][-1]


@pytest.mark.parametrize("discrete", [False, True])
def test_kat4_gradient_pinned_by_adam_replay(O, golden, discrete):
    """The reference's stored loss history losses[k] = L(theta_k) under ADAM(0.1) (scenario_1.jl:114) is
    reproduced by replaying ADAM with the ORACLE'S interpolating-adjoint gradient: this pins d L / d theta
    (the reference computed it with ForwardDiffSensitivity; any correct gradient must agree)."""
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    th = theta_scenario1_init(g)
    mm, vv = np.zeros_like(th), np.zeros_like(th)
    b1, b2, eta, eps = 0.9, 0.999, 0.1, 1e-8
    for it in range(1, 7):
        l, gr = _loss_grad_s1(O, m, th, g["X"], discrete=discrete)   # discrete = ForwardDiffSensitivity's quantity (scenario_1.jl:86)
        assert abs(l - g["losses"][it - 1]) < 2e-6 * g["losses"][it - 1], (it, l, g["losses"][it - 1])  # reference solver tol 1e-6
        mm = b1 * mm + (1 - b1) * gr
        vv = b2 * vv + (1 - b2) * gr * gr
        th = th - eta * (mm / (1 - b1**it)) / (np.sqrt(vv / (1 - b2**it)) + eps)


def test_adjoint_matches_finite_differences(O, golden):
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    th = theta_scenario1_init(g)
    _, gr = _loss_grad_s1(O, m, th, g["X"], sub=8)
    for i in (0, 7, 20, 50, 86):
        e = np.zeros_like(th)
        e[i] = 1e-6
        fd = (_loss_grad_s1(O, m, th + e, g["X"], sub=8)[0] - _loss_grad_s1(O, m, th - e, g["X"], sub=8)[0]) / 2e-6
        assert abs(fd - gr[i]) < 1e-5 * max(1.0, abs(fd))


def test_kat5_scenario2_trainable_decay_rate(O, golden):
    """theta = [delta; U] (scenario_2.jl:87): MLP and forward solve over tspan (0, 6)."""
    g = golden["scenario_2"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3, n_prefix=1)
    th = g["theta_trained"]
    assert abs(th[0] - 1.78527) < 1e-4
    Y = np.stack([O.mlp_forward(m, th[1:], g["Xhat"][:, i]) for i in range(121)], 1)
    assert np.abs(Y - g["Yhat"]).max() < 1e-12
    out, _, _ = O.solve_adaptive(m, th, g["X"][:, 0], np.linspace(0, 6, 121), 1e-11, 1e-11)
    assert np.abs(out.T - g["Xhat"]).max() < 1e-4


def test_kat6_scenario3_pde_rhs(O, golden):
    """Fisher-KPP UPDE on the 26-point grid, Float32 (scenario_3.jl): pointwise net + stencil."""
    g = golden["scenario_3"]
    m = O.fkpp_model(26, (1, 5, 5, 5, 1), RBF3)
    th = g["theta_trained"].astype(np.float64)
    assert O.num_params(m) == 81
    R = np.array([[O.mlp_forward(m, th, np.array([g["Xhat"][i, k]], dtype=np.float64))[0] for k in range(11)] for i in range(26)])
    assert np.abs(R - g["Rhat"]).max() < 5e-6
    out, _, _ = O.solve_adaptive(m, th, g["X"][:, 0].astype(np.float64), np.linspace(0, 5, 11), 1e-9, 1e-9)
    assert np.abs(out.T - g["Xhat"]).max() < 3e-4
    loss = ((out.T - g["X"]) ** 2).sum() + abs(th[76:79].sum())
    assert abs(loss - g["losses"][-1]) < 2e-3 * g["losses"][-1] + 2e-5


def test_kat8_hudson_bay_fastchain_layout(O, golden):
    """FastChain(rbf, rbf, tanh, linear) with two trainable rates in front (hudson_bay.jl:77-91)."""
    g = golden["hudson_bay"]
    m = O.lv_model((2, 5, 5, 5, 2), ("rbf", "rbf", "tanh", "identity"), n_prefix=2)
    th = g["theta_trained"].astype(np.float64)
    Y = np.stack([O.mlp_forward(m, th[2:], g["Xhat"][:, i]) for i in range(41)], 1)
    assert np.abs(Y - g["Yhat"]).max() < 1e-6
    out, _, _ = O.solve_adaptive(m, th, g["X"][:, 0].astype(np.float64), g["tsample"], 1e-10, 1e-10)
    assert np.abs(out.T - g["Xhat"]).max() < 1e-4


@pytest.mark.parametrize("model", ["lv", "seir", "fkpp", "seir_node"])
def test_rhs_vjp_matches_finite_differences(O, model):
    rng = np.random.default_rng(3)
    if model == "lv":
        m = O.lv_model((2, 8, 8, 2), ("tanh", "rbf", "identity"), n_prefix=2)
        u = np.array([0.7, 2.3])
    elif model == "seir":
        m = O.seir_model((3, 16, 16, 1))
        u = np.array([1.2e7, 80.0, 90.0, 130.0, 1.4e7, 20.0, 250.0])
    elif model == "seir_node":   # dudt_node (seir_exposure.jl:52-64)
        m = O.seir_node_model((7, 12, 12, 7), ("tanh", "tanh", "identity"))
        u = np.array([1.2e7, 80.0, 90.0, 130.0, 1.4e7, 20.0, 250.0])
    else:
        m = O.fkpp_model(9, (1, 6, 6, 1), ("tanh", "tanh", "identity"))
        u = rng.uniform(0, 1, 9)
    P = O.num_params(m)
    th = 0.3 * rng.standard_normal(P)
    lam = rng.standard_normal(m.d)
    dlam, gth = O.rhs_vjp(m, th, u, lam)
    for k in range(m.d):
        e = np.zeros(m.d)
        h = 1e-6 * max(1.0, abs(u[k]))
        e[k] = h
        fd = lam @ (O.rhs(m, th, u + e) - O.rhs(m, th, u - e)) / (2 * h)
        assert abs(fd - dlam[k]) < 5e-6 * max(1.0, abs(fd)), (k, fd, dlam[k])
    for i in rng.choice(P, 12, replace=False):
        e = np.zeros(P)
        e[i] = 1e-4   # SEIR right-hand sides are O(1e6): a larger step keeps the round-off of the difference small
        fd = lam @ (O.rhs(m, th + e, u) - O.rhs(m, th - e, u)) / 2e-4
        assert abs(fd - gth[i]) < 5e-6 * max(1.0, abs(fd)), (i, fd, gth[i])


def test_interpolating_adjoint_vs_autograd_through_the_scheme(O):
    """Independent cross-oracle: torch-fp64 autograd through a pure-torch Tsit5 (discrete adjoint) agrees
    with the oracle's continuous interpolating adjoint to truncation error (~3e-8 at dt=0.1)."""
    torch = pytest.importorskip("torch")
    from helpers import glorot_theta, synthetic_ensemble
    N = 16
    theta = glorot_theta((2, 32, 32, 2), seed=1).astype(np.float64)
    u0, y = synthetic_ensemble(N)
    m = O.lv_model()
    l, g, gu = O.ensemble_loss_grad(m, theta, u0, y, np.ones(2), 0.1, 30)
    c = O.tsit5_constants()
    A = np.zeros((7, 7))
    p = 6
    for i in range(1, 7):
        for j in range(i):
            A[i, j] = c[p]
            p += 1
    th = torch.tensor(theta, requires_grad=True)
    U0 = torch.tensor(u0.astype(np.float64), requires_grad=True)

    def rhs(u):
        W1 = th[0:64].reshape(2, 32).T; b1 = th[64:96]
        W2 = th[96:1120].reshape(32, 32).T; b2 = th[1120:1152]
        W3 = th[1152:1216].reshape(32, 2).T; b3 = th[1216:1218]
        h = torch.tanh(W1 @ u + b1[:, None]); h = torch.tanh(W2 @ h + b2[:, None]); o = W3 @ h + b3[:, None]
        return torch.stack([1.3 * u[0] + o[0], -1.8 * u[1] + o[1]])
    u = U0
    loss = ((u - torch.tensor(y[0].astype(np.float64))) ** 2).sum()
    for s in range(30):
        ks = [rhs(u)]
        for i in range(1, 7):
            ks.append(rhs(u + 0.1 * sum(A[i, j] * ks[j] for j in range(i))))
            if i == 5:
                pass
        u = u + 0.1 * sum(A[6, j] * ks[j] for j in range(6))
        loss = loss + ((u - torch.tensor(y[s + 1].astype(np.float64))) ** 2).sum()
    loss.backward()
    assert abs(float(loss) - l) < 1e-10 * l
    assert np.linalg.norm(th.grad.numpy() - g) < 1e-6 * np.linalg.norm(g)
    assert np.abs(U0.grad.numpy() - gu).max() < 1e-6 * np.abs(gu).max()
    # the oracle's DISCRETE adjoint (reverse accumulation through the stages = ForwardDiffSensitivity's result,
    # scenario_1.jl:86) is the same quantity as autograd through the scheme: equal to round-off
    ld, gd, gud = O.ensemble_loss_grad_discrete(m, theta, u0.astype(np.float64), y.astype(np.float64), np.ones(2), 0.1, 30)
    assert abs(ld - float(loss)) < 1e-12 * ld
    assert np.linalg.norm(th.grad.numpy() - gd) < 1e-11 * np.linalg.norm(gd)
    assert np.abs(U0.grad.numpy() - gud).max() < 1e-11 * np.abs(gud).max()


def test_fp32_oracle_close_to_fp64(O):
    from helpers import glorot_theta, synthetic_ensemble
    theta = glorot_theta((2, 32, 32, 2), seed=1)
    u0, y = synthetic_ensemble(256)
    m = O.lv_model()
    l64, g64, gu64 = O.ensemble_loss_grad(m, theta.astype(np.float64), u0, y, np.ones(2), 0.1, 30)
    l32, g32, gu32 = O.ensemble_loss_grad(m, theta, u0, y, np.ones(2, np.float32), 0.1, 30)
    assert abs(l32 - l64) < 1e-5 * l64
    assert np.linalg.norm(g32 - g64) < 1e-4 * np.linalg.norm(g64)


def test_replay_adjoint_matches_error_controlled_backward_solve(O, golden):
    """Adaptive solves: the adjoint that replays the accepted forward steps (what the GPU kernels do) agrees with a
    reference-style error-controlled backward solve of the augmented state [lambda; mu] (tstops at the save times)
    and with a fine fixed-grid adjoint, at the level of the solver tolerance."""
    g = golden["scenario_1"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3)
    th = theta_scenario1_init(g)
    X, ts = g["X"], g["t"]
    sub = 32
    out, dense = O.solve_fixed(m, th, X[:, 0], 0.1 / sub, 30 * sub, save_every=sub, want_dense=True)
    g_true, gu_true = O.adjoint_fixed(m, th, out, dense, 0.1 / sub, 30 * sub, 2 * (out - X.T), save_every=sub)
    n = np.linalg.norm
    for tol, bound in ((1e-6, 5e-7), (1e-8, 2e-8)):
        outa, rec = O.solve_adaptive_dense(m, th, X[:, 0], ts, tol, tol)
        assert np.abs(outa - out).max() < 20 * tol
        dl = 2 * (outa - X.T)
        g_rep, gu_rep = O.adjoint_replay(m, th, ts, rec, dl)
        g_ada, gu_ada, nback = O.adjoint_adaptive(m, th, ts, rec, dl, tol, tol)
        assert nback > rec[3]                      # the backward solve takes its own (more) steps
        assert n(g_rep - g_ada) < bound * n(g_ada) and n(g_rep - g_true) < bound * n(g_true)
        assert n(gu_rep - gu_ada) < bound * n(gu_ada)


def test_rkc2_oracle_coefficients_convergence_and_adjoint(O):
    """RKC2 (second-order Runge-Kutta-Chebyshev; the closed-form ROCK2-class stabilised explicit method, north star /
    Climate/NeuralPDE/npde.jl:61): recurrence coefficients are consistent (c_s = 1, sum w = 1, sum w c = 1/2, stability
    interval ~ 0.65 s^2), the forward solve of the stiff Fisher-KPP grid converges with order 2 towards Tsit5, and the
    interpolating adjoint (lambda stepped backwards by RKC2, cubic-Hermite dense output) converges with order 2 to the
    Tsit5 interpolating-adjoint gradient."""
    from helpers import glorot_theta
    for s in (2, 5, 16, 64):
        c = O.rkc2_coefficients(s)
        assert abs(c["c"][s] - 1) < 1e-12 and abs(c["w"].sum() - 1) < 1e-12 and abs((c["w"] * c["c"][:s]).sum() - 0.5) < 1e-12
        assert 0.45 * s * s < c["beta"] < 0.67 * s * s + 1   # -> 0.65 s^2 for large s (1.96 at s = 2)
    Nx = 64
    widths = (1, 16, 16, 1)
    m = O.fkpp_model(Nx, widths, ("tanh", "tanh", "identity"))
    D0 = 0.01 * (Nx - 1) ** 2
    th = np.concatenate([glorot_theta(widths, seed=3), [1.0, -2.0, 1.0, 0.0, D0]]).astype(np.float64)
    x = np.linspace(0, 1, Nx)
    u0 = 0.5 * (np.tanh((x - 0.35) / 0.05) - np.tanh((x - 0.65) / 0.05))
    y = np.tile(u0[None], (6, 1)) * 0.9

    def tsit(dt):
        n = int(round(1.0 / dt))
        out, dense = O.solve_fixed(m, th, u0, dt, n, save_every=n // 5, want_dense=True)
        g, gu = O.adjoint_fixed(m, th, out, dense, dt, n, 2 * (out - y), save_every=n // 5)
        return out, g, gu
    ref, g_ref, gu_ref = tsit(2e-3)
    errs, gerrs = [], []
    for dt, s in ((0.05, 6), (0.025, 5), (0.0125, 4)):
        n = int(round(1.0 / dt))
        assert O.rkc2_coefficients(s)["beta"] > dt * (4 * D0 + 2)          # inside the stability interval
        out, us, fs = O.solve_rkc2(m, th, u0, dt, n, s, save_every=n // 5, want_record=True)
        g, gu = O.adjoint_rkc2(m, th, us, fs, dt, n, s, 2 * (out - y), save_every=n // 5)
        errs.append(np.abs(out - ref).max())
        gerrs.append(np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref))
    assert errs[0] < 2e-3 and errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5, errs
    assert gerrs[0] < 1e-3 and gerrs[1] < gerrs[0] / 3 and gerrs[2] < gerrs[1] / 3, gerrs
    # Tsit5 itself is unstable at these steps: dt * rho = 8 > 3.3
    try:
        blown = np.abs(O.solve_fixed(m, th, u0, 0.05, 20)).max() > 1e3
    except FloatingPointError:
        blown = True
    assert blown


def test_kat9_scenario2_multiple_shooting_loss_history_pins_trainable_rate_gradient(O, golden):
    """KAT-9.  scenario_2.jl:57-72,112-123,141: theta = [delta; U] (delta = the predator's linear decay rate, trainable), loss =
    1e-3 mean(theta[2:end]^2) + sum over 5 shooting segments of [ sum (x data - x_hat)^2 + |y(end) - y_hat(end)| ], ADAM(0.1).
    Replaying ADAM with the ORACLE's interpolating-adjoint gradient reproduces the reference's stored loss history (which the
    reference produced with Vern7 @ 1e-6 + ForwardDiffSensitivity): pins the gradient w.r.t. a physics parameter in front of
    the chain (n_prefix = 1), the cotangent of a non-L2 loss term, and the regulariser."""
    g = golden["scenario_2"]
    t, X, losses = g["t"], g["X"], g["losses"]
    m = O.lv_model((2, 5, 5, 5, 2), RBF3, n_prefix=1)
    ty = np.arange(t[0], t[-1] + 1e-9, 6 / 5)
    segs = [np.where((t >= ty[i] - 1e-9) & (t <= ty[i + 1] + 1e-9))[0] for i in range(len(ty) - 1)]
    assert [len(sg) for sg in segs] == [13] * 5

    def loss_grad(th, sub=16, want_grad=True):
        l = 1e-3 * np.sum(th[1:] ** 2) / len(th[1:])
        gth = np.zeros_like(th)
        gth[1:] = 2e-3 * th[1:] / len(th[1:])
        for idx in segs:
            xs, y0, y1, n, dt = X[0, idx], X[1, idx[0]], X[1, idx[-1]], len(idx) - 1, 0.1 / sub
            out, dense = O.solve_fixed(m, th, np.array([xs[0], y0]), dt, n * sub, save_every=sub, want_dense=True)
            l += np.sum((xs - out[:, 0]) ** 2) + abs(y1 - out[-1, 1])
            if want_grad:
                cot = np.zeros_like(out)
                cot[:, 0] = 2 * (out[:, 0] - xs)
                cot[-1, 1] = -np.sign(y1 - out[-1, 1])
                gth += O.adjoint_fixed(m, th, out, dense, dt, n * sub, cot, save_every=sub)[0]
        return l, gth
    th = g["theta_init"].copy()
    mm, vv = np.zeros_like(th), np.zeros_like(th)
    for it in range(1, 7):
        l, gr = loss_grad(th)
        assert abs(l - losses[it - 1]) < 5e-7 * losses[it - 1], (it, l, losses[it - 1])
        mm = 0.9 * mm + 0.1 * gr
        vv = 0.999 * vv + 0.001 * gr * gr
        th = th - 0.1 * (mm / (1 - 0.9 ** it)) / (np.sqrt(vv / (1 - 0.999 ** it)) + 1e-8)
    l_end, _ = loss_grad(g["theta_trained"], want_grad=False)
    assert abs(l_end - losses[-1]) < 2e-4 * losses[-1]


def test_kat10_hudson_bay_final_loss(O, golden):
    """KAT-10.  hudson_bay.jl:82-91,112-115: two trainable rates in front of a FastChain with rbf, rbf, tanh; the script's
    'equivalent L2 loss' sum(abs2, Xn - X_hat) / 21 + 1e-3 mean(theta[3:end]^2) at the stored trained parameters equals the
    last entry of the stored loss history (Float32 data; reference solve: Vern7 @ 1e-6)."""
    g = golden["hudson_bay"]
    t, X, th = g["t"].astype(np.float64), g["X"].astype(np.float64), g["theta_trained"].astype(np.float64)
    m = O.lv_model((2, 5, 5, 5, 2), ("rbf", "rbf", "tanh", "identity"), n_prefix=2)
    out, _, _ = O.solve_adaptive(m, th, X[:, 0], t, 1e-9, 1e-9)
    loss = np.sum((X - out.T) ** 2) / X.shape[1] + 1e-3 * np.sum(th[2:] ** 2) / len(th[2:])
    assert abs(loss - g["losses"][-1]) < 5e-5 * g["losses"][-1]


def test_kat11_scenario3_fisher_kpp_loss_history_pins_the_pde_gradient(O, golden):
    """KAT-11.  scenario_3.jl:96-145 (the Lux twin of FisherKPP/Fisher-KPP-CNN.jl): theta = (chain 1-5-5-5-1 rbf, 4 stencil
    parameters, D0), objective = sum(abs2, pred - Xn) + abs(sum of the three stencil taps), ADAM(0.1) x 10 from the stored
    initial parameters (taps = 0, D0 = 6.5).  Replaying ADAM with the ORACLE's interpolating-adjoint gradient reproduces all ten
    stored ADAM losses to Float32 storage precision: pins the Fisher-KPP gradient -- chain, stencil weights and D0 -- against
    the reference's own run (which used ForwardDiffSensitivity through Vern7)."""
    g = golden["scenario_3"]
    X, losses = g["X"].astype(np.float64), g["losses"]
    m = O.fkpp_model(26, (1, 5, 5, 5, 1), RBF3)

    def loss_grad(th, sub=40):
        dt, n = 0.5 / sub, 10 * sub
        out, dense = O.solve_fixed(m, th, X[:, 0], dt, n, save_every=sub, want_dense=True)
        r = out - X.T
        sw = th[76:79].sum()
        gth = O.adjoint_fixed(m, th, out, dense, dt, n, 2 * r, save_every=sub)[0].copy()
        gth[76:79] += np.sign(sw)
        return (r ** 2).sum() + abs(sw), gth
    th = g["theta_init"].astype(np.float64)
    assert np.all(th[76:80] == 0) and th[80] == 6.5
    mm, vv = np.zeros_like(th), np.zeros_like(th)
    for it in range(1, 11):
        l, gr = loss_grad(th)
        assert abs(l - losses[it - 1]) < 1e-5 * losses[it - 1], (it, l, losses[it - 1])
        mm = 0.9 * mm + 0.1 * gr
        vv = 0.999 * vv + 0.001 * gr * gr
        th = th - 0.1 * (mm / (1 - 0.9 ** it)) / (np.sqrt(vv / (1 - 0.999 ** it)) + 1e-8)
