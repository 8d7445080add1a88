"""ctypes binding of the CPU oracle (oracle/libude_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(universal_differential_equations_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libude_oracle.so")

MODEL_LV, MODEL_SEIR, MODEL_FKPP, MODEL_NODE, MODEL_SEIR_NODE = 0, 1, 2, 3, 4
ACT_IDENTITY, ACT_TANH, ACT_RBF = 0, 1, 2
TSIT5, VERN7 = 0, 1
_ACT = {"identity": 0, "tanh": 1, "rbf": 2}


class UdeModel(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("d", C.c_int32), ("n_layers", C.c_int32),
        ("widths", C.c_int32 * 7), ("acts", C.c_int32 * 6),
        ("n_prefix", C.c_int32), ("n_suffix", C.c_int32), ("n_consts", C.c_int32),
        ("consts", C.c_double * 16),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, a few seconds)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("ude_oracle.c", "ude_oracle_impl.h", "ude_oracle.h", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ude_num_params.restype = C.c_size_t
        for suf in ("_f32", "_f64"):
            getattr(_lib, "ude_ensemble_loss_grad" + suf).restype = C.c_double
            getattr(_lib, "ude_solve_fixed" + suf).restype = C.c_int
            getattr(_lib, "ude_solve_adaptive" + suf).restype = C.c_int
    return _lib


def make_model(model, d, widths, acts, n_prefix=0, n_suffix=0, consts=()):
    m = UdeModel()
    m.model, m.d, m.n_layers = model, d, len(widths) - 1
    for i, w in enumerate(widths):
        m.widths[i] = w
    for i, a in enumerate(acts):
        m.acts[i] = _ACT[a] if isinstance(a, str) else int(a)
    m.n_prefix, m.n_suffix, m.n_consts = n_prefix, n_suffix, len(consts)
    for i, c in enumerate(consts):
        m.consts[i] = float(c)
    return m


def lv_model(widths=(2, 32, 32, 2), acts=("tanh", "tanh", "identity"), n_prefix=0, consts=(1.3, 1.8)):
    """LV UDE (scenario_1.jl:69-73): consts = (p_[1], p_[4]) = (1.3, 1.8) (scenario_1.jl:39)."""
    return make_model(MODEL_LV, 2, widths, acts, n_prefix=n_prefix, consts=consts)


SEIR_CONSTS = (10.0, 0.5944, 0.4239, 1117.3, 0.02, 1 / 3, 1 / 5, 0.2, 1 / 11.2)  # seir_exposure.jl:33


def seir_model(widths=(3, 64, 64, 1), acts=("tanh", "tanh", "identity")):
    return make_model(MODEL_SEIR, 7, widths, acts, consts=SEIR_CONSTS)


def seir_node_model(widths=(7, 64, 64, 64, 7), acts=("tanh", "tanh", "tanh", "identity")):
    """The script's black-box baseline dudt_node (seir_exposure.jl:52-64)."""
    return make_model(MODEL_SEIR_NODE, 7, widths, acts, consts=SEIR_CONSTS)


def fkpp_model(nx, widths, acts):
    return make_model(MODEL_FKPP, nx, widths, acts, n_suffix=5)


def num_params(m):
    return int(lib().ude_num_params(C.byref(m)))


def _dt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", C.c_float
    if dtype == np.float64:
        return "_f64", C.c_double
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def mlp_forward(m, theta_chain, x):
    suf, _ = _dt(theta_chain.dtype)
    x = np.ascontiguousarray(x, dtype=theta_chain.dtype)
    y = np.empty(m.widths[m.n_layers], dtype=theta_chain.dtype)
    getattr(lib(), "ude_mlp_forward" + suf)(C.byref(m), _p(np.ascontiguousarray(theta_chain)), _p(x), _p(y))
    return y


def rhs(m, theta, u):
    suf, _ = _dt(theta.dtype)
    u = np.ascontiguousarray(u, dtype=theta.dtype)
    du = np.empty(m.d, dtype=theta.dtype)
    getattr(lib(), "ude_rhs" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u), _p(du))
    return du


def rhs_vjp(m, theta, u, lam, w=1.0):
    suf, ct = _dt(theta.dtype)
    u = np.ascontiguousarray(u, dtype=theta.dtype)
    lam = np.ascontiguousarray(lam, dtype=theta.dtype)
    dlam = np.empty(m.d, dtype=theta.dtype)
    g = np.zeros(num_params(m), dtype=theta.dtype)
    getattr(lib(), "ude_rhs_vjp" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u), _p(lam), _p(dlam), _p(g), ct(w))
    return dlam, g


def solve_fixed(m, theta, u0, dt, n_steps, solver=TSIT5, save_every=1, want_dense=False):
    """Returns out[n_save, d] (and dense[n_steps, stages, d])."""
    suf, ct = _dt(theta.dtype)
    u0 = np.ascontiguousarray(u0, dtype=theta.dtype)
    n_save = n_steps // save_every + 1
    out = np.empty((n_save, m.d), dtype=theta.dtype)
    stages = 7 if solver == TSIT5 else 10
    dense = np.empty((n_steps, stages, m.d), dtype=theta.dtype) if want_dense else None
    rc = getattr(lib(), "ude_solve_fixed" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u0), ct(dt),
                                                 n_steps, solver, save_every, _p(out), _p(dense))
    if rc != 0:
        raise FloatingPointError("oracle: non-finite state")
    return (out, dense) if want_dense else out


def solve_adaptive(m, theta, u0, saveat, abstol, reltol, solver=TSIT5):
    """Returns out[n_save, d], n_accepted, n_rejected."""
    suf, ct = _dt(theta.dtype)
    u0 = np.ascontiguousarray(u0, dtype=theta.dtype)
    saveat = np.ascontiguousarray(saveat, dtype=theta.dtype)
    out = np.empty((len(saveat), m.d), dtype=theta.dtype)
    nrej = C.c_int(0)
    nacc = getattr(lib(), "ude_solve_adaptive" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u0), _p(saveat),
                                                      len(saveat), ct(abstol), ct(reltol), solver, _p(out), C.byref(nrej))
    if nacc < 0:
        raise FloatingPointError("oracle: adaptive solve failed")
    return out, nacc, nrej.value


def adjoint_fixed(m, theta, out, dense, dt, n_steps, dLdout, save_every=1):
    """Returns grad_theta[P], grad_u0[d]."""
    suf, ct = _dt(theta.dtype)
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.empty(m.d, dtype=theta.dtype)
    getattr(lib(), "ude_adjoint_fixed" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(np.ascontiguousarray(out)),
                                              _p(np.ascontiguousarray(dense)), ct(dt), n_steps, save_every,
                                              _p(np.ascontiguousarray(dLdout, dtype=theta.dtype)), _p(g), _p(gu))
    return g, gu


def rkc2_coefficients(s):
    """RKC2 recurrence coefficients for s stages: dict(mt1, mu, nu, mt, ga, c, w, beta) (arrays indexed by stage)."""
    arr = [np.zeros(s + 1) for _ in range(6)]
    mt1, beta = C.c_double(0), C.c_double(0)
    rc = lib().ude_rkc2_coefficients(s, C.byref(mt1), *[_p(a) for a in arr[:4]], _p(arr[4]), _p(arr[5]), C.byref(beta))
    if rc != 0:
        raise ValueError("2 <= s <= 64")
    return dict(mt1=mt1.value, mu=arr[0], nu=arr[1], mt=arr[2], ga=arr[3], c=arr[4], w=arr[5][:s], beta=beta.value)


def solve_rkc2(m, theta, u0, dt, n_steps, stages, save_every=1, want_record=False):
    """Fixed-step RKC2 solve: out[n_save, d] (and the record (ustep, fstep)[n_steps+1, d] the adjoint reads)."""
    suf, ct = _dt(theta.dtype)
    u0 = np.ascontiguousarray(u0, dtype=theta.dtype)
    out = np.empty((n_steps // save_every + 1, m.d), dtype=theta.dtype)
    us = np.empty((n_steps + 1, m.d), dtype=theta.dtype)
    fs = np.empty((n_steps + 1, m.d), dtype=theta.dtype)
    rc = getattr(lib(), "ude_solve_rkc2_fixed" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u0), ct(dt), n_steps, stages,
                                                      save_every, _p(out), _p(us), _p(fs))
    if rc != 0:
        raise FloatingPointError("oracle: RKC2 solve failed (%d)" % rc)
    return (out, us, fs) if want_record else out


def adjoint_rkc2(m, theta, ustep, fstep, dt, n_steps, stages, dLdout, save_every=1):
    """Interpolating adjoint of an RKC2 solve (lambda stepped backwards with RKC2, u(t) by cubic Hermite): grad_theta, grad_u0."""
    suf, ct = _dt(theta.dtype)
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.empty(m.d, dtype=theta.dtype)
    getattr(lib(), "ude_adjoint_rkc2_fixed" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(np.ascontiguousarray(ustep)),
                                                   _p(np.ascontiguousarray(fstep)), ct(dt), n_steps, stages, save_every,
                                                   _p(np.ascontiguousarray(dLdout, dtype=theta.dtype)), _p(g), _p(gu))
    return g, gu


def adjoint_discrete(m, theta, out, dense, dt, n_steps, dLdout, save_every=1):
    """Exact gradient of the discrete fixed-step Tsit5 scheme (ForwardDiffSensitivity's result): grad_theta[P], grad_u0[d]."""
    suf, ct = _dt(theta.dtype)
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.empty(m.d, dtype=theta.dtype)
    getattr(lib(), "ude_adjoint_discrete_fixed" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(np.ascontiguousarray(out)),
                                                       _p(np.ascontiguousarray(dense)), ct(dt), n_steps, save_every,
                                                       _p(np.ascontiguousarray(dLdout, dtype=theta.dtype)), _p(g), _p(gu))
    return g, gu


def ensemble_loss_grad_discrete(m, theta, u0, y, wmask, dt, n_steps, save_every=1):
    """Like ensemble_loss_grad with the discrete adjoint (python loop over trajectories; small N only)."""
    N = u0.shape[1]
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.zeros((m.d, N), dtype=theta.dtype)
    loss = 0.0
    for k in range(N):
        out, dense = solve_fixed(m, theta, u0[:, k].astype(theta.dtype), dt, n_steps, save_every=save_every, want_dense=True)
        r = (out - y[:, :, k]) * np.asarray(wmask)[None, :]
        loss += float((np.asarray(wmask)[None, :] * (out - y[:, :, k]) ** 2).sum())
        gk, guk = adjoint_discrete(m, theta, out, dense, dt, n_steps, 2 * r, save_every=save_every)
        g += gk
        gu[:, k] = guk
    return loss, g, gu


def ensemble_loss_grad(m, theta, u0, y, wmask, dt, n_steps, save_every=1, n_threads=None, want_out=False, want_gu0=True):
    """u0[d, N], y[n_save, d, N] -> loss, grad_theta[P], grad_u0[d, N] (, out[n_save, d, N])."""
    suf, ct = _dt(theta.dtype)
    dtp = theta.dtype
    u0 = np.ascontiguousarray(u0, dtype=dtp)
    y = np.ascontiguousarray(y, dtype=dtp)
    wmask = np.ascontiguousarray(wmask, dtype=dtp)
    N = u0.shape[1]
    n_save = n_steps // save_every + 1
    assert y.shape == (n_save, m.d, N), (y.shape, (n_save, m.d, N))
    out = np.empty((n_save, m.d, N), dtype=dtp) if want_out else None
    g = np.empty(num_params(m), dtype=dtp)
    gu = np.empty((m.d, N), dtype=dtp) if want_gu0 else None
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    loss = getattr(lib(), "ude_ensemble_loss_grad" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u0), _p(y), _p(wmask),
                                                          C.c_size_t(N), ct(dt), n_steps, save_every, _p(out), _p(g), _p(gu),
                                                          int(n_threads))
    res = [loss, g, gu]
    if want_out:
        res.append(out)
    return tuple(res)


def tsit5_constants():
    c = np.empty(56)
    lib().ude_tsit5_constants(_p(c))
    return c


def vern7_constants():
    c = np.empty(58)
    lib().ude_vern7_constants(_p(c))
    return c


def solve_adaptive_dense(m, theta, u0, saveat, abstol, reltol, max_steps=4096):
    """Adaptive Tsit5 recording the accepted steps -> (out[n_save,d], rec) with rec = (tgrid, ustep, dense, nacc)."""
    suf, ct = _dt(theta.dtype)
    dtp = theta.dtype
    u0 = np.ascontiguousarray(u0, dtype=dtp)
    saveat = np.ascontiguousarray(saveat, dtype=dtp)
    out = np.empty((len(saveat), m.d), dtype=dtp)
    tgrid = np.empty(max_steps + 1, dtype=dtp)
    ustep = np.empty((max_steps + 1, m.d), dtype=dtp)
    dense = np.empty((max_steps, 7, m.d), dtype=dtp)
    f = getattr(lib(), "ude_solve_adaptive_dense" + suf)
    f.restype = C.c_int
    nacc = f(C.byref(m), _p(np.ascontiguousarray(theta)), _p(u0), _p(saveat), len(saveat), ct(abstol), ct(reltol), max_steps,
             _p(out), _p(tgrid), _p(ustep), _p(dense))
    if nacc < 0:
        raise FloatingPointError("oracle: adaptive dense solve failed")
    return out, (tgrid, ustep, dense, nacc)


def adjoint_replay(m, theta, saveat, rec, dLdout):
    suf, ct = _dt(theta.dtype)
    tgrid, ustep, dense, nacc = rec
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.empty(m.d, dtype=theta.dtype)
    saveat = np.ascontiguousarray(saveat, dtype=theta.dtype)
    getattr(lib(), "ude_adjoint_replay" + suf)(C.byref(m), _p(np.ascontiguousarray(theta)), _p(saveat), len(saveat), _p(tgrid), _p(ustep),
                                               _p(dense), nacc, _p(np.ascontiguousarray(dLdout, dtype=theta.dtype)), _p(g), _p(gu))
    return g, gu


def adjoint_adaptive(m, theta, saveat, rec, dLdout, abstol, reltol):
    """Reference-style error-controlled backward solve of [lambda; mu]; returns grad_theta, grad_u0, n_backward_steps."""
    suf, ct = _dt(theta.dtype)
    tgrid, ustep, dense, nacc = rec
    g = np.zeros(num_params(m), dtype=theta.dtype)
    gu = np.empty(m.d, dtype=theta.dtype)
    saveat = np.ascontiguousarray(saveat, dtype=theta.dtype)
    f = getattr(lib(), "ude_adjoint_adaptive" + suf)
    f.restype = C.c_int
    ns = f(C.byref(m), _p(np.ascontiguousarray(theta)), _p(saveat), len(saveat), _p(tgrid), _p(ustep), _p(dense), nacc,
           _p(np.ascontiguousarray(dLdout, dtype=theta.dtype)), ct(abstol), ct(reltol), _p(g), _p(gu))
    if ns < 0:
        raise FloatingPointError("oracle: adaptive adjoint failed")
    return g, gu, ns
