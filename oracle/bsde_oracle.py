"""CPU restatement (numpy, float64) of the deep-BSDE solve the reference runs in highdim_pde/lambaem.jl -- TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does).

Reference call (highdim_pde/lambaem.jl:8-34):
    prob   = TerminalPDEProblem(g, f, mu, sigma, x0, tspan)        g(X) = log(0.5 + 0.5 |X|^2), f = -lambda |sigma^T grad u|^2,
                                                                   mu = 0, sigma = sqrt(2) I  (Hamilton-Jacobi-Bellman, d = 100)
    pdealg = NNPDENS(u0, sigmaT_grad_u, opt = ADAM(0.03))          u0: d -> hls -> hls -> 1, sigma^T grad u: d+1 -> hls x3 -> d, relu
    solve(prob, pdealg; maxiters = 500, trajectories = m, alg = LambaEM(), ...)
The arithmetic lives in un-vendored packages (NeuralNetDiffEq 1.1.0, StochasticDiffEq 6.16.0; highdim_pde/Manifest.toml:440,641) [EXT]:
NNPDENS integrates the (d+1)-dimensional forward SDE  dX = mu dt + sigma dW,  du = -f(X, u, z, p, t) dt + z . dW  with
z = sigmaT_grad_u([X; t]) and u(0) = u0(x0), and minimises  mean over trajectories of (g(X_T) - u_T)^2  with ADAM; the answer
is u0(x0).  Restated here with a FIXED-step Euler-Maruyama discretisation (the script's LambaEM is EM with an adaptive step;
for this problem X is integrated exactly by EM at any step and only the u-quadrature depends on dt), for the family
    mu = 0,  sigma = s I,  f = -lambda |z|^2,  g(X) = log(a + b |X|^2).
PARITY UNPINNED at the level of individual iterations: NeuralNetDiffEq / StochasticDiffEq are not in /root/reference, the script commits
no result file, and its Brownian paths come from Julia's RNG -- there is no vector of the reference's own to compare a loss or a gradient
with.  What IS pinned is the reference's own (and only) test, lambaem.jl:36-48 -- |u0(x0) - u_analytic| / |u0(x0)| < 0.2 against a Monte-Carlo
of the closed-form HJB solution -- is what tests/ run against both this oracle and the GPU path.  Brownian increments come from a
counter-based generator (Philox4x32-10 + Box-Muller) that the device code implements identically, so loss and gradient of one
iteration can be compared path by path.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (vectorised)."""
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def normals(seed, step, n_paths, d, path0=0):
    """dW / sqrt(dt): standard normals [d, n_paths] of paths path0 .. path0 + n_paths - 1; component block q = 4 consecutive
    components from one Philox call with counter (path, q, step, 0) and key (seed low, seed high); Box-Muller on (x + 0.5) 2^-32."""
    nq = (d + 3) // 4
    path = (np.arange(n_paths, dtype=np.uint64) + np.uint64(path0))[None, :].repeat(nq, 0)
    q = np.arange(nq, dtype=np.uint64)[:, None].repeat(n_paths, 1)
    r = philox4x32(path, q, np.full_like(path, step), np.zeros_like(path), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [(x.astype(np.float64) + 0.5) * 2.0 ** -32 for x in r]
    out = np.empty((4 * nq, n_paths))
    for h, (a, b) in enumerate(((u[0], u[1]), (u[2], u[3]))):
        rad = np.sqrt(-2.0 * np.log(a))
        out[2 * h::4] = rad * np.cos(2.0 * np.pi * b)
        out[2 * h + 1::4] = rad * np.sin(2.0 * np.pi * b)
    return out[:d]


def num_params(d, hls):
    pu = d * hls + hls + hls * hls + hls + hls + 1
    pz = (d + 1) * hls + hls + 2 * (hls * hls + hls) + hls * d + d
    return pu, pz


def unpack(theta, widths):
    """[(W (out x in), b)] from the flat vector: per layer vec(W) column-major, then b (Flux.destructure order)."""
    layers, o = [], 0
    for nin, nout in zip(widths[:-1], widths[1:]):
        W = theta[o:o + nin * nout].reshape(nin, nout).T
        o += nin * nout
        layers.append((W, theta[o:o + nout]))
        o += nout
    return layers


def mlp_fwd(layers, x):
    """x [n_in, M] -> activations list (post-relu of hidden layers, linear last).  relu(a) = max(0, a) with derivative 1 AT the
    tie a = 0, as Flux 0.9 / Tracker differentiate max(zero(x), x) (d max(x, y)/dy = x > y ? 0 : 1) [EXT]: with the script's zero
    input x0 and zero initial biases every hidden pre-activation of the u0 net starts exactly at the tie, and this convention is
    what lets all its biases train from the first iteration.  The sign bit of a zero output records which side it came from
    (+0.0: a >= 0, derivative 1; -0.0: a < 0, derivative 0)."""
    acts = [x]
    for l, (W, b) in enumerate(layers):
        a = W @ acts[-1] + b[:, None]
        acts.append(np.where(a >= 0, a, -0.0) if l < len(layers) - 1 else a)
    return acts


def mlp_bwd(layers, acts, cot):
    """cot [n_out, M] -> (flat gradient in theta order, input cotangent)."""
    grads = [None] * len(layers)
    for l in range(len(layers) - 1, -1, -1):
        if l < len(layers) - 1:
            cot = cot * ~np.signbit(acts[l + 1])
        W, _ = layers[l]
        grads[l] = np.concatenate([(cot @ acts[l].T).T.ravel(), cot.sum(1)])
        cot = W.T @ cot
    return np.concatenate(grads), cot


def loss_and_grad(theta, d, hls, x0, T, n_steps, n_paths, seed, lam=1.0, s=np.sqrt(2.0), a=0.5, b=0.5, path0=0, n_total=None):
    """One NNPDENS iteration: loss = mean (g(X_T) - u_T)^2 and its gradient w.r.t. theta = [theta_u0; theta_z].
    path0 / n_total: this call evaluates the shard path0 .. path0 + n_paths - 1 of n_total paths (the mean's denominator), so that the
    shards of a multi-GPU job add up to the whole batch."""
    n_total = n_total or n_paths
    pu, pz = num_params(d, hls)
    Lu = unpack(theta[:pu], (d, hls, hls, 1))
    Lz = unpack(theta[pu:], (d + 1, hls, hls, hls, d))
    dt = T / n_steps
    X = np.repeat(np.asarray(x0, float)[:, None], n_paths, 1)
    au = mlp_fwd(Lu, np.asarray(x0, float)[:, None])
    u = np.full(n_paths, au[-1][0, 0])
    steps = []                                     # per step: activations and dW (Tracker keeps the whole tape; nothing is recomputed)
    for n in range(n_steps):
        acts = mlp_fwd(Lz, np.vstack([X, np.full((1, n_paths), n * dt)]))
        z = acts[-1]
        dW = np.sqrt(dt) * normals(seed, n, n_paths, d, path0)
        steps.append((acts, dW))
        u = u + lam * (z * z).sum(0) * dt + (z * dW).sum(0)
        X = X + s * dW
    nrm = a + b * (X * X).sum(0)
    r = np.log(nrm) - u
    loss = (r * r).sum() / n_total
    ubar = -2.0 * r / n_total
    gz = np.zeros(pz)
    for n in range(n_steps - 1, -1, -1):
        acts, dW = steps[n]
        zbar = ubar[None, :] * (2.0 * lam * acts[-1] * dt + dW)
        g, _ = mlp_bwd(Lz, acts, zbar)          # X does not depend on theta (mu = 0, constant sigma): no cotangent flows through X
        gz += g
    gu, _ = mlp_bwd(Lu, au, np.array([[ubar.sum()]]))
    return loss, np.concatenate([gu, gz]), au[-1][0, 0]


def analytic_hjb(x0, T, lam=1.0, n_mc=100000, seed=0):
    """u(x0, 0) = -(1/lambda) log E exp(-lambda g(x0 + sqrt(2 T) W))   (lambaem.jl:36-40)."""
    rng = np.random.default_rng(seed)
    x = np.asarray(x0, float)[None, :] + np.sqrt(2.0 * T) * rng.standard_normal((n_mc, len(x0)))
    gx = np.log(0.5 + 0.5 * (x * x).sum(1))
    return -np.log(np.mean(np.exp(-lam * gx))) / lam


def glorot(widths, rng):
    parts = []
    for nin, nout in zip(widths[:-1], widths[1:]):
        lim = np.sqrt(6.0 / (nin + nout))
        parts += [rng.uniform(-lim, lim, (nout, nin)).ravel("F"), np.zeros(nout)]
    return np.concatenate(parts)


def init_params(d, hls, seed=0):
    rng = np.random.default_rng(seed)
    return np.concatenate([glorot((d, hls, hls, 1), rng), glorot((d + 1, hls, hls, hls, d), rng)])


def train(theta, d, hls, x0, T, n_steps, n_paths, iters, eta=0.03, seed0=1, **kw):
    """ADAM(eta) (Flux defaults) on fresh Brownian paths per iteration (seed0 + it); returns theta, losses, u0(x0) history."""
    m = np.zeros_like(theta); v = np.zeros_like(theta)
    losses, u0s = [], []
    for it in range(1, iters + 1):
        l, g, u0 = loss_and_grad(theta, d, hls, x0, T, n_steps, n_paths, seed0 + it - 1, **kw)
        losses.append(l); u0s.append(u0)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        theta = theta - eta * (m / (1 - 0.9 ** it)) / (np.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    return theta, np.array(losses), np.array(u0s)
