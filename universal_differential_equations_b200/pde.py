"""Host-side mirror of the terminal-PDE call surface of highdim_pde/lambaem.jl:18-34 (NeuralNetDiffEq's NNPDENS) over the
C ABI's b200ude_bsde_* entry points (include/b200ude.h).  No CPU fallback; nothing here imports the oracle.

    prob   = TerminalPDEProblem(HJBTerminal(0.5, 0.5), HJBNonlinearity(lam), ZeroDrift(), ConstantDiffusion(sqrt(2)), x0, tspan)
    pdealg = NNPDENS(u0_chain, sigmaT_grad_u_chain, opt=ADAM(0.03))
    ans    = solve(prob, pdealg, verbose=True, maxiters=500, trajectories=m, alg=LambaEM(), dt=1/20)

A C ABI cannot take closures, so g, f, mu, sigma are NAMED forms (the script's Hamilton-Jacobi-Bellman family); the chains are
Flux-style `Chain(Dense(...), ...)` shape descriptions whose widths must be the script's architecture.
"""
import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .sciml import ADAM, Chain, FastChain


@dataclass
class HJBTerminal:
    """g(X) = log(a + b |X|^2)   (lambaem.jl:14)"""
    a: float = 0.5
    b: float = 0.5


@dataclass
class HJBNonlinearity:
    """f(X, u, sigmaT_grad_u, p, t) = -lam * sum(sigmaT_grad_u.^2)   (lambaem.jl:15)"""
    lam: float = 1.0


class ZeroDrift:
    """mu(X, p, t) = zero(X)   (lambaem.jl:16)"""


@dataclass
class ConstantDiffusion:
    """sigma(X, p, t) = Diagonal(s * ones(d))   (lambaem.jl:17)"""
    s: float = math.sqrt(2.0)


@dataclass
class TerminalPDEProblem:
    g: HJBTerminal
    f: HJBNonlinearity
    mu: ZeroDrift
    sigma: ConstantDiffusion
    x0: Sequence[float]
    tspan: tuple


@dataclass
class NNPDENS:
    u0: FastChain
    sigmaT_grad_u: FastChain
    opt: Optional[ADAM] = None


class LambaEM:
    """Euler-Maruyama; the device path takes the fixed step `dt` given to solve (default (t1 - t0) / 20)."""


relu = "relu"   # activation tag for Dense(in, out, relu) descriptions


def _widths(chain):
    w, a = list(chain.widths), list(chain.acts)
    if any(x != relu for x in a[:-1]) or a[-1] != "identity":
        raise ValueError("NNPDENS chains: relu hidden layers and a linear output layer (lambaem.jl:24-31)")
    return w


def initial_params_pde(pdealg: NNPDENS, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Flux's default Dense init (glorot_uniform weights, zero biases), theta = [u0 net; sigmaT_grad_u net]."""
    rng = rng or np.random.default_rng(0)
    parts = []
    for chain in (pdealg.u0, pdealg.sigmaT_grad_u):
        w = _widths(chain)
        for nin, nout in zip(w[:-1], w[1:]):
            lim = math.sqrt(6.0 / (nin + nout))
            parts += [rng.uniform(-lim, lim, (nout, nin)).ravel("F"), np.zeros(nout)]
    return np.concatenate(parts)


class BSDESolver:
    """Owns one b200ude_bsde handle."""

    def __init__(self, prob: TerminalPDEProblem, pdealg: NNPDENS, n_steps: int, max_paths: int, device=0, dtype=torch.float32, arithmetic=None):
        """dtype: element type of the parameters / results the caller sees.  arithmetic: element type the device computes in; by default
        float64 whenever the fused fp64 sweeps apply (widths <= 110 / 111) -- they are faster than the fp32 library-GEMM path and the
        float32 caller simply gets correctly rounded double results -- else dtype.  arithmetic=torch.float32 forces the fp32 path."""
        d = len(prob.x0)
        wu, wz = _widths(pdealg.u0), _widths(pdealg.sigmaT_grad_u)
        hls = wu[1]
        if wu != [d, hls, hls, 1] or wz != [d + 1, hls, hls, hls, d]:
            raise ValueError(f"NNPDENS chains must be d->h->h->1 and d+1->h->h->h->d (got {wu}, {wz})")
        if not isinstance(prob.g, HJBTerminal) or not isinstance(prob.f, HJBNonlinearity) or not isinstance(prob.mu, ZeroDrift) \
                or not isinstance(prob.sigma, ConstantDiffusion):
            raise ValueError("TerminalPDEProblem: only the Hamilton-Jacobi-Bellman family of lambaem.jl has a device path")
        if float(prob.tspan[0]) != 0.0:
            raise ValueError("tspan must start at 0")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.io_dtype = dtype
        if arithmetic is None:
            arithmetic = torch.float64 if (d <= 110 and hls <= 111) else dtype
        dtype = arithmetic
        self.dtype = dtype
        self.np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self.d, self.hls, self.n_steps = d, hls, n_steps
        x0 = (C.c_double * d)(*[float(v) for v in prob.x0])
        desc = _lib.BsdeDesc(struct_size=C.sizeof(_lib.BsdeDesc), device=self.device.index or 0,
                             dtype=_lib.F64 if dtype == torch.float64 else _lib.F32, dim=d, hidden=hls, n_steps=n_steps,
                             T=float(prob.tspan[1]), lam=float(prob.f.lam), sigma=float(prob.sigma.s), g_a=float(prob.g.a), g_b=float(prob.g.b),
                             x0=x0, max_paths=max_paths)
        self._L = _lib.lib()
        h = C.c_void_p()
        _lib.check_bsde(None, self._L.b200ude_bsde_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.P = int(self._L.b200ude_bsde_num_params(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200ude_bsde_destroy(self._h)
            self._h = None

    __del__ = close

    def set_params(self, theta):
        if isinstance(theta, torch.Tensor):
            t = theta.detach().to(self.device, self.dtype).contiguous()
            self._fence()
            _lib.check_bsde(self._h, self._L.b200ude_bsde_set_params(self._h, t.data_ptr(), t.numel(), _lib.DEVICE))
        else:
            a = np.ascontiguousarray(theta, self.np_dtype)
            _lib.check_bsde(self._h, self._L.b200ude_bsde_set_params(self._h, a.ctypes.data, a.size, _lib.HOST))

    def get_params(self) -> np.ndarray:
        a = np.empty(self.P, self.np_dtype)
        _lib.check_bsde(self._h, self._L.b200ude_bsde_get_params(self._h, a.ctypes.data, a.size, _lib.HOST))
        return a.astype(np.float64 if self.io_dtype == torch.float64 else np.float32, copy=False)

    def _fence(self):
        """The b200ude_bsde_* calls run on the handle's own (non-blocking) stream and are synchronous to the host on return; work torch has
        queued on ITS stream for tensors handed to them (a gradient coming out of an all-reduce, recycled allocator blocks) must be
        finished first."""
        torch.cuda.current_stream(self.device).synchronize()

    def loss_gradient(self, n_paths, seed, path_offset=0, total_paths=0):
        """(loss, grad [P] device tensor, u0(x0)) at the handle's theta."""
        out = torch.empty(2, device=self.device, dtype=self.dtype)
        grad = torch.empty(self.P, device=self.device, dtype=self.dtype)
        self._fence()
        _lib.check_bsde(self._h, self._L.b200ude_bsde_loss_gradient(self._h, n_paths, seed, path_offset, total_paths, out.data_ptr(), grad.data_ptr(),
                                                                    out.data_ptr() + out.element_size()))
        o = out.cpu()
        return float(o[0]), grad.to(self.io_dtype), float(o[1])

    def last_train_ms(self) -> float:
        return float(self._L.b200ude_bsde_last_train_ms(self._h))

    def last_sweep_ms(self):
        """(forward sweep, cotangent sweep, weight-gradient GEMMs) device ms of the last loss_gradient call (fused fp64 handles)."""
        ms = (C.c_double * 3)()
        _lib.check_bsde(self._h, self._L.b200ude_bsde_last_sweep_ms(self._h, ms))
        return tuple(ms)

    def _adam(self, opt):
        return _lib.Adam(struct_size=C.sizeof(_lib.Adam), eta=opt.eta, beta1=opt.beta[0], beta2=opt.beta[1], eps=opt.eps, loss_scale=1.0, l2_reg=0.0)

    def adam_step(self, opt: ADAM, grad: torch.Tensor):
        """One ADAM update with a device gradient (e.g. after an all-reduce over path shards)."""
        g = grad.to(self.device, self.dtype).contiguous()
        self._fence()
        a = self._adam(opt)
        _lib.check_bsde(self._h, self._L.b200ude_bsde_adam_step(self._h, C.byref(a), g.data_ptr()))

    def train_adam(self, opt: ADAM, n_paths, iters, seed0=1):
        """`iters` iterations on the device; returns (loss history, u0(x0) history) as device tensors."""
        hist = torch.empty(2, iters, device=self.device, dtype=self.dtype)
        self._fence()
        a = self._adam(opt)
        _lib.check_bsde(self._h, self._L.b200ude_bsde_train_adam(self._h, C.byref(a), n_paths, iters, seed0, hist[0].data_ptr(), hist[1].data_ptr()))
        return hist[0].to(self.io_dtype), hist[1].to(self.io_dtype)


def solve(prob: TerminalPDEProblem, pdealg: NNPDENS, *, verbose=False, maxiters=500, trajectories=100, alg=None, dt=None, theta0=None, seed=1,
          device=0, dtype=torch.float32, **_tolerances):
    """`solve(prob, pdealg; verbose, maxiters, trajectories, alg = LambaEM(), ...)` (lambaem.jl:33-34): returns u0(x0) after training.
    Step-size tolerances are accepted and ignored (fixed-step Euler-Maruyama, `dt`)."""
    T = float(prob.tspan[1])
    n_steps = 20 if dt is None else max(1, int(round(T / dt)))
    s = BSDESolver(prob, pdealg, n_steps, trajectories, device=device, dtype=dtype)
    try:
        s.set_params(initial_params_pde(pdealg) if theta0 is None else theta0)
        opt = pdealg.opt or ADAM(0.001)
        losses, u0s = s.train_adam(opt, trajectories, maxiters, seed0=seed)
        if verbose:
            lh = losses.cpu().numpy()
            for i in range(0, maxiters, max(1, maxiters // 10)):
                print(f"Current loss is: {lh[i]:.6g}")
        # the answer is u0(x0) at the FINAL parameters
        _, _, ans = s.loss_gradient(trajectories, seed + maxiters)
        solve.last = {"losses": losses, "u0_history": u0s, "theta": s.get_params()}
        return ans
    finally:
        s.close()
