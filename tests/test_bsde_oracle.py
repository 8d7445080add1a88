"""CPU checks of oracle/bsde_oracle.py (the deep-BSDE restatement of highdim_pde/lambaem.jl): generator known answers, the hand-written
reverse sweep against finite differences, and the reference's own acceptance test (lambaem.jl:36-48) at a CPU-sized dimension."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import bsde_oracle as bo  # noqa: E402


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = bo.philox4x32(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_normals_are_standard_and_independent_of_batch_size():
    z = bo.normals(seed=7, step=3, n_paths=4000, d=10)
    assert z.shape == (10, 4000)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.05 and abs(np.corrcoef(z[2], z[3])[0, 1]) < 0.05
    np.testing.assert_array_equal(z[:, :100], bo.normals(7, 3, 100, 10))          # counter-based: a path's draws do not depend on M
    assert not np.array_equal(z, bo.normals(7, 4, 4000, 10)) and not np.array_equal(z, bo.normals(8, 3, 4000, 10))


def test_gradient_matches_central_differences():
    d, hls, M, N = 6, 9, 16, 4
    rng = np.random.default_rng(0)
    pu, pz = bo.num_params(d, hls)
    theta = 0.4 * rng.standard_normal(pu + pz)          # generic biases: no relu ties
    x0 = rng.standard_normal(d)
    l0, g, _ = bo.loss_and_grad(theta, d, hls, x0, 1.0, N, M, seed=5)
    idx = rng.choice(theta.size, 40, replace=False)
    for i in idx:
        e = np.zeros_like(theta); e[i] = 1e-6
        fd = (bo.loss_and_grad(theta + e, d, hls, x0, 1.0, N, M, 5)[0] - bo.loss_and_grad(theta - e, d, hls, x0, 1.0, N, M, 5)[0]) / 2e-6
        assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_relu_tie_takes_derivative_one():
    # zero input and zero biases: every hidden pre-activation of the u0 net is exactly 0 (lambaem.jl:9 x0 = fill(0, d), Flux zero-bias init);
    # Flux 0.9 / Tracker give relu'(0) = 1 there, so the bias gradients of the u0 net are non-zero from the first iteration
    d, hls = 4, 5
    theta = bo.init_params(d, hls, seed=1)
    pu, _ = bo.num_params(d, hls)
    _, g, _ = bo.loss_and_grad(theta, d, hls, np.zeros(d), 1.0, 4, 8, seed=2)
    b1 = g[d * hls:d * hls + hls]
    assert np.all(b1 != 0.0)


def test_reference_acceptance_criterion_small_dimension():
    # lambaem.jl:36-48: error_l2 = |ans - analytic| / |ans| < 0.2, here at d = 10 so that the CPU suite stays short
    d, hls, M, N = 10, 20, 64, 20
    x0 = np.zeros(d)
    theta, losses, u0s = bo.train(bo.init_params(d, hls, 0), d, hls, x0, 1.0, N, M, iters=250, eta=0.03, seed0=1)
    ans = bo.loss_and_grad(theta, d, hls, x0, 1.0, N, M, seed=999)[2]
    ref = bo.analytic_hjb(x0, 1.0, n_mc=200000)
    assert abs(ans - ref) / abs(ans) < 0.2, (ans, ref)
    assert losses[-20:].mean() < 0.2 * losses[:5].mean()


def _golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "hjb_small.npz"))


def test_oracle_reproduces_its_frozen_vectors():
    g = _golden()
    d, hls, M, N = int(g["d"]), int(g["hls"]), int(g["M"]), int(g["N"])
    np.testing.assert_allclose(bo.normals(int(g["seed"]), 3, 8, d), g["normals_step3"], rtol=0, atol=1e-15)
    l, gr, u0 = bo.loss_and_grad(g["theta"], d, hls, g["x0"], 1.0, N, M, int(g["seed"]))
    assert abs(l - float(g["loss"])) <= 1e-13 * abs(l) and abs(u0 - float(g["u0"])) <= 1e-13
    np.testing.assert_allclose(gr, g["grad"], rtol=1e-11, atol=1e-14)
    l, gr, _ = bo.loss_and_grad(g["theta_tie"], d, hls, np.zeros(d), 1.0, N, M, int(g["seed_tie"]))
    assert abs(l - float(g["loss_tie"])) <= 1e-13 * abs(l)
    np.testing.assert_allclose(gr, g["grad_tie"], rtol=1e-11, atol=1e-14)


def _shard_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = _golden()
    d, hls, M, N = int(g["d"]), int(g["hls"]), int(g["M"]), int(g["N"])
    lo, hi = rank * M // world, (rank + 1) * M // world
    l, gr, _ = bo.loss_and_grad(g["theta"], d, hls, g["x0"], 1.0, N, hi - lo, int(g["seed"]), path0=lo, n_total=M)
    buf = torch.from_numpy(np.concatenate([gr, [l]]))
    dist.all_reduce(buf)                                  # the job's one collective: [grad; loss] summed over the path shards
    q.put((rank, buf.numpy().copy()))
    dist.destroy_process_group()


def test_path_shards_over_two_ranks_add_up_gloo():
    """Multi-GPU rule of the SDE path (path_offset / total_paths of b200ude_bsde_loss_gradient): disjoint Philox path counters, the mean
    over ALL paths; the all-reduced [grad; loss] equals the single-process evaluation."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    ps = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    g = _golden()
    want = np.concatenate([g["grad"], [float(g["loss"])]])
    for r in range(2):
        np.testing.assert_allclose(res[r], want, rtol=1e-11, atol=1e-14)


def test_python_mirror_validates_and_never_falls_back():
    """pde.py: wrong chain shapes / problem forms are ValueErrors before any native call; without a device the solve raises the library's
    ENODEVICE -- the mirror has no CPU path (and does not import the oracle)."""
    import math
    import pytest
    import torch
    import universal_differential_equations_b200 as ude
    from universal_differential_equations_b200 import pde, _lib
    d, hls = 6, 8
    prob = ude.TerminalPDEProblem(ude.HJBTerminal(), ude.HJBNonlinearity(1.0), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), np.zeros(d), (0.0, 1.0))
    u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
    sg = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
    with pytest.raises(ValueError):
        ude.BSDESolver(prob, ude.NNPDENS(u0, u0), 4, 8)                                   # second chain has the wrong shape
    with pytest.raises(ValueError):
        ude.BSDESolver(prob, ude.NNPDENS(ude.Chain(ude.Dense(d, hls, ude.tanh), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1)), sg), 4, 8)
    bad = ude.TerminalPDEProblem(ude.HJBTerminal(), ude.HJBNonlinearity(1.0), ude.ZeroDrift(), ude.ConstantDiffusion(1.0), np.zeros(d), (0.5, 1.0))
    with pytest.raises(ValueError):
        ude.BSDESolver(bad, ude.NNPDENS(u0, sg), 4, 8)                                   # tspan must start at 0
    assert pde.initial_params_pde(ude.NNPDENS(u0, sg)).size == sum(bo.num_params(d, hls))
    import ast
    mods = [n.module or "" for n in ast.walk(ast.parse(open(pde.__file__).read())) if isinstance(n, ast.ImportFrom)] + \
           [a.name for n in ast.walk(ast.parse(open(pde.__file__).read())) if isinstance(n, ast.Import) for a in n.names]
    assert not any("oracle" in m for m in mods), mods
    if not torch.cuda.is_available():
        with pytest.raises(_lib.B200UDEError) as e:
            ude.solve(prob, ude.NNPDENS(u0, sg, opt=ude.ADAM(0.03)), maxiters=2, trajectories=8, alg=ude.LambaEM())
        assert e.value.code == _lib.ENODEVICE


def test_oracle_reproduces_the_start_of_its_committed_training_curve():
    """tests/golden/hjb_script_curve.npz (the script's configuration, tools/make_golden_bsde_curve.py): first iterations re-run here."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hjb_script_curve.npz"))
    d, hls, M, N = int(g["d"]), int(g["hls"]), int(g["M"]), int(g["N"])
    _, losses, u0s = bo.train(bo.init_params(d, hls, int(g["init_seed"])), d, hls, np.zeros(d), 1.0, N, M, 6, eta=float(g["eta"]), seed0=int(g["seed0"]))
    np.testing.assert_allclose(losses, g["losses"][:6], rtol=1e-12)
    np.testing.assert_allclose(u0s, g["u0s"][:6], rtol=1e-12, atol=1e-14)
    # the curve itself does what the reference's test asks for: u0(x0) ends within 20 % of the analytic value
    assert abs(g["u0s"][-1] - bo.analytic_hjb(np.zeros(d), 1.0, n_mc=100000)) / abs(g["u0s"][-1]) < 0.2
