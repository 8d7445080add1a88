"""Shared builders for the tests: reference-shaped models and the synthetic ensemble of BASELINE config 2."""
import numpy as np


def theta_scenario1_init(g):
    """ComponentVector{Float64}(p) of scenario_1.jl:113 from the stored Float32 Lux init (zero biases)."""
    return np.concatenate([g["W1_init"].ravel("F"), np.zeros(5), g["W2_init"].ravel("F"), np.zeros(5),
                           g["W3_init"].ravel("F"), np.zeros(5), g["W4_init"].ravel("F"), np.zeros(2)]).astype(np.float64)


def glorot_theta(widths, seed=1):
    rng = np.random.default_rng(seed)
    parts = []
    for nin, nout in zip(widths[:-1], widths[1:]):
        lim = np.sqrt(6.0 / (nin + nout))
        parts += [rng.uniform(-lim, lim, size=(nout, nin)).astype(np.float32).ravel("F"), np.zeros(nout, np.float32)]
    return np.concatenate(parts)


def lv_true_solution(u0, ts, p=(1.3, 0.9, 0.8, 1.8)):
    """Tight solve of the true Lotka-Volterra system lotka! (scenario_1.jl:30-34) for data generation."""
    from scipy.integrate import solve_ivp
    a, b, c, d = p
    out = np.empty((len(ts), 2, u0.shape[1]))
    for n in range(u0.shape[1]):
        s = solve_ivp(lambda t, u: [a * u[0] - b * u[1] * u[0], c * u[0] * u[1] - d * u[1]], (ts[0], ts[-1]), u0[:, n],
                      t_eval=ts, rtol=1e-10, atol=1e-10, method="DOP853")
        out[:, :, n] = s.y.T
    return out


def synthetic_ensemble(N, n_steps=30, dt=0.1, seed=0, cheap_targets=True):
    """u0[2, N] ~ U(0.2,1) x U(2,5) and targets y[n_save, 2, N].

    cheap_targets: true-LV targets from a vectorised RK4 at dt/8 (error ~1e-8) instead of per-trajectory DOP853."""
    rng = np.random.default_rng(seed)
    u0 = np.stack([rng.uniform(0.2, 1.0, N), rng.uniform(2.0, 5.0, N)])
    a, b, c, d = 1.3, 0.9, 0.8, 1.8

    def f(u):
        return np.stack([a * u[0] - b * u[1] * u[0], c * u[0] * u[1] - d * u[1]])
    sub = 8
    h = dt / sub
    y = np.empty((n_steps + 1, 2, N))
    u = u0.copy()
    y[0] = u
    for s in range(n_steps):
        for _ in range(sub):
            k1 = f(u); k2 = f(u + 0.5 * h * k1); k3 = f(u + 0.5 * h * k2); k4 = f(u + h * k3)
            u = u + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        y[s + 1] = u
    y += 5e-3 * y.mean(axis=(0, 2), keepdims=True) * rng.standard_normal(y.shape)  # scenario_1.jl:51-53
    return u0.astype(np.float32), y.astype(np.float32)
