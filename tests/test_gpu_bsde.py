"""GPU parity tests of the terminal-PDE / SDE path (b200ude_bsde_*, highdim_pde/lambaem.jl) against oracle/bsde_oracle.py.

Both sides draw the same Brownian increments (Philox4x32-10 + Box-Muller in fp64, counter = (path, component block, step)), so one
iteration's loss and gradient are compared directly.  Tolerances:
  fp64 handle   loss, u0(x0) rel <= 1e-10; gradient ||g - g_o|| / ||g_o|| <= 1e-9   (summation order of the GEMMs only)
  fp32 handle   loss rel <= 2e-4; gradient rel-L2 <= 2e-3                            (fp32 GEMMs over n_steps + backward sweep)
  training      the reference's own criterion, lambaem.jl:42-48:  |ans - analytic| / |ans| < 0.2
"""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import bsde_oracle as bo  # noqa: E402


def _ude():
    import universal_differential_equations_b200 as ude
    return ude


def _problem(ude, d, hls, x0=None, lam=1.0, T=1.0):
    x0 = np.zeros(d) if x0 is None else x0
    prob = ude.TerminalPDEProblem(ude.HJBTerminal(0.5, 0.5), ude.HJBNonlinearity(lam), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), x0, (0.0, T))
    u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
    sg = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
    return prob, ude.NNPDENS(u0, sg, opt=ude.ADAM(0.03))


@pytest.mark.parametrize("d,hls,M,N", [(100, 110, 100, 20), (10, 20, 37, 5), (7, 9, 33, 3), (1, 4, 5, 2)])
def test_f64_loss_gradient_vs_oracle(d, hls, M, N):
    ude = _ude()
    rng = np.random.default_rng(d)
    x0 = 0.3 * rng.standard_normal(d)
    prob, alg = _problem(ude, d, hls, x0)
    theta = bo.init_params(d, hls, seed=3) + 0.05 * rng.standard_normal(sum(bo.num_params(d, hls)))
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    assert s.P == theta.size
    s.set_params(theta)
    loss, g, u0 = s.loss_gradient(M, seed=11)
    lo, go, u0o = bo.loss_and_grad(theta, d, hls, x0, 1.0, N, M, seed=11)
    g = g.cpu().numpy()
    assert abs(loss - lo) <= 1e-10 * abs(lo), (loss, lo)
    assert abs(u0 - u0o) <= 1e-10 * max(1.0, abs(u0o))
    assert np.linalg.norm(g - go) <= 1e-9 * np.linalg.norm(go)
    s.close()


def test_f64_frozen_vectors():
    """tests/golden/hjb_small.npz (tools/make_golden_bsde.py): the comparison does not run the oracle."""
    import os
    ude = _ude()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hjb_small.npz"))
    d, hls, M, N = int(g["d"]), int(g["hls"]), int(g["M"]), int(g["N"])
    for th, x0, seed, l_w, g_w, u_w in ((g["theta"], g["x0"], int(g["seed"]), float(g["loss"]), g["grad"], float(g["u0"])),
                                        (g["theta_tie"], np.zeros(d), int(g["seed_tie"]), float(g["loss_tie"]), g["grad_tie"], float(g["u0_tie"]))):
        prob, alg = _problem(ude, d, hls, x0)
        s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
        s.set_params(th)
        loss, gr, u0 = s.loss_gradient(M, seed=seed)
        assert abs(loss - l_w) <= 1e-10 * abs(l_w) and abs(u0 - u_w) <= 1e-10 * max(1.0, abs(u_w))
        assert np.linalg.norm(gr.cpu().numpy() - g_w) <= 1e-9 * np.linalg.norm(g_w)
        s.close()


def test_f64_relu_ties_follow_the_reference_convention():
    """x0 = 0 with Flux's zero-bias init (lambaem.jl:9,24-31): every hidden pre-activation of the u0 net is exactly 0; relu'(0) = 1."""
    ude = _ude()
    d, hls, M, N = 12, 16, 24, 4
    prob, alg = _problem(ude, d, hls)
    theta = bo.init_params(d, hls, seed=1)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    s.set_params(theta)
    loss, g, _ = s.loss_gradient(M, seed=2)
    lo, go, _ = bo.loss_and_grad(theta, d, hls, np.zeros(d), 1.0, N, M, seed=2)
    g = g.cpu().numpy()
    assert np.all(g[d * hls:d * hls + hls] != 0.0)
    assert abs(loss - lo) <= 1e-10 * abs(lo) and np.linalg.norm(g - go) <= 1e-9 * np.linalg.norm(go)
    s.close()


def test_f32_loss_gradient_vs_oracle():
    ude = _ude()
    d, hls, M, N = 100, 110, 100, 20
    prob, alg = _problem(ude, d, hls)
    theta = bo.init_params(d, hls, seed=0).astype(np.float32)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float32, arithmetic=torch.float32)   # the fp32 library-GEMM path of the ABI
    s.set_params(theta)
    loss, g, u0 = s.loss_gradient(M, seed=1)
    lo, go, u0o = bo.loss_and_grad(theta.astype(np.float64), d, hls, np.zeros(d), 1.0, N, M, seed=1)
    rel = np.linalg.norm(g.cpu().numpy() - go) / np.linalg.norm(go)
    print(f"bsde fp32: loss {loss:.6f} vs {lo:.6f}, grad rel-L2 {rel:.2e}")
    assert abs(loss - lo) <= 2e-4 * abs(lo)
    assert rel <= 2e-3
    s.close()


def test_f32_caller_gets_double_arithmetic_by_default():
    """dtype=float32 (the script's precision): parameters / results are float32, the device computes on the fused fp64 sweeps."""
    ude = _ude()
    d, hls, M, N = 100, 110, 64, 10
    prob, alg = _problem(ude, d, hls)
    theta = bo.init_params(d, hls, seed=0).astype(np.float32)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float32)
    assert s.dtype == torch.float64 and s.io_dtype == torch.float32
    s.set_params(theta)
    loss, g, _ = s.loss_gradient(M, seed=1)
    lo, go, _ = bo.loss_and_grad(theta.astype(np.float64), d, hls, np.zeros(d), 1.0, N, M, seed=1)
    assert g.dtype == torch.float32 and s.get_params().dtype == np.float32
    assert abs(loss - lo) <= 1e-9 * abs(lo)
    assert np.linalg.norm(g.cpu().numpy() - go) <= 2e-7 * np.linalg.norm(go)      # rounding of the result to float32 only
    s.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_wide_network_runs_on_the_library_gemm_path(dtype):
    """hidden > 111: no fused sweeps; cuBLAS GEMMs with the hand-written kernels between them, both precisions."""
    ude = _ude()
    d, hls, M, N = 6, 120, 40, 4
    rng = np.random.default_rng(3)
    x0 = 0.2 * rng.standard_normal(d)
    prob, alg = _problem(ude, d, hls, x0)
    theta = bo.init_params(d, hls, seed=5) + 0.02 * rng.standard_normal(sum(bo.num_params(d, hls)))
    s = ude.BSDESolver(prob, alg, N, M, dtype=dtype)
    assert s.dtype == dtype
    s.set_params(theta)
    loss, g, u0 = s.loss_gradient(M, seed=3)
    lo, go, u0o = bo.loss_and_grad(theta if dtype == torch.float64 else theta.astype(np.float32).astype(np.float64), d, hls, x0, 1.0, N, M, seed=3)
    tol_l, tol_g = (1e-10, 1e-9) if dtype == torch.float64 else (2e-4, 2e-3)
    assert abs(loss - lo) <= tol_l * abs(lo) and np.linalg.norm(g.cpu().numpy().astype(np.float64) - go) <= tol_g * np.linalg.norm(go)
    losses, _ = s.train_adam(ude.ADAM(0.03), M, 4, seed0=3)          # the on-device loop of this path; slot 0 = the loss just checked
    lh = losses.cpu().numpy().astype(np.float64)
    assert abs(lh[0] - lo) <= tol_l * abs(lo) and np.isfinite(lh).all()
    s.close()


def test_f64_adam_history_vs_oracle():
    ude = _ude()
    d, hls, M, N, iters = 20, 30, 50, 10, 12
    prob, alg = _problem(ude, d, hls)
    theta = bo.init_params(d, hls, seed=4)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    s.set_params(theta)
    losses, u0s = s.train_adam(ude.ADAM(0.03), M, iters, seed0=5)
    th_o, lo, uo = bo.train(theta, d, hls, np.zeros(d), 1.0, N, M, iters, eta=0.03, seed0=5)
    np.testing.assert_allclose(losses.cpu().numpy(), lo, rtol=1e-7)
    np.testing.assert_allclose(u0s.cpu().numpy(), uo, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(s.get_params(), th_o, rtol=1e-6, atol=1e-8)
    # a second call continues the ADAM state (step counter, moments) and the seed sequence the caller passes
    l2, _ = s.train_adam(ude.ADAM(0.03), M, 3, seed0=5 + iters)
    _, lo2, _ = bo.train(theta, d, hls, np.zeros(d), 1.0, N, M, iters + 3, eta=0.03, seed0=5)
    np.testing.assert_allclose(l2.cpu().numpy(), lo2[iters:], rtol=1e-6)
    s.close()


def test_f64_repeat_is_bitwise_identical():
    """Same parameters, seed and batch -> the same bits (fixed-order sums; the slab ring and the two-warp tiles change timing, not order)."""
    ude = _ude()
    d, hls, M, N = 100, 110, 333, 7
    prob, alg = _problem(ude, d, hls)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    s.set_params(bo.init_params(d, hls, seed=9))
    l1, g1, u1 = s.loss_gradient(M, seed=4)
    s.loss_gradient(M, seed=5)
    l2, g2, u2 = s.loss_gradient(M, seed=4)
    assert l1 == l2 and u1 == u2 and torch.equal(g1, g2)
    s.close()


def test_f64_training_curve_at_the_script_configuration_vs_oracle():
    """300 ADAM(0.03) iterations at lambaem.jl's own sizes (d = 100, hls = 110, m = 100, 20 steps), Brownian seeds 1, 2, ...: the loss and
    u0(x0) histories of the on-device loop against the oracle's committed curve (tests/golden/hjb_script_curve.npz,
    tools/make_golden_bsde_curve.py).  Measured agreement 1.4e-14 relative over the whole curve; the tolerance leaves five orders."""
    import os
    ude = _ude()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hjb_script_curve.npz"))
    d, hls, M, N, iters = int(g["d"]), int(g["hls"]), int(g["M"]), int(g["N"]), int(g["iters"])
    prob, alg = _problem(ude, d, hls)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    s.set_params(bo.init_params(d, hls, seed=int(g["init_seed"])))
    losses, u0s = s.train_adam(ude.ADAM(float(g["eta"])), M, iters, seed0=int(g["seed0"]))
    np.testing.assert_allclose(losses.cpu().numpy(), g["losses"], rtol=1e-9)
    np.testing.assert_allclose(u0s.cpu().numpy(), g["u0s"], rtol=1e-9, atol=1e-12)
    assert abs(np.linalg.norm(s.get_params()) - float(g["theta_final_norm"])) <= 1e-9 * float(g["theta_final_norm"])
    s.close()


def test_path_shards_add_up():
    """Multi-GPU sharding rule (path_offset / total_paths): two half batches give the whole batch's loss and gradient."""
    ude = _ude()
    d, hls, M, N = 10, 20, 64, 6
    prob, alg = _problem(ude, d, hls)
    theta = bo.init_params(d, hls, seed=2)
    s = ude.BSDESolver(prob, alg, N, M, dtype=torch.float64)
    s.set_params(theta)
    l, g, _ = s.loss_gradient(M, seed=9)
    la, ga, _ = s.loss_gradient(M // 2, seed=9, path_offset=0, total_paths=M)
    lb, gb, _ = s.loss_gradient(M // 2, seed=9, path_offset=M // 2, total_paths=M)
    assert abs(la + lb - l) <= 1e-12 * abs(l)
    assert torch.allclose(ga + gb, g, rtol=1e-10, atol=1e-14)
    s.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_script_call_hjb_acceptance(dtype):
    """highdim_pde/lambaem.jl line by line: d = 100, m = 100, hls = 110, ADAM(0.03), maxiters = 500; @test error_l2 < 0.2."""
    ude = _ude()
    d = 100
    x0 = np.zeros(d, np.float32)
    tspan = (0.0, 1.0)
    m = 100
    lam = 1.0
    prob = ude.TerminalPDEProblem(ude.HJBTerminal(0.5, 0.5), ude.HJBNonlinearity(lam), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), x0, tspan)
    hls = 10 + d
    opt = ude.ADAM(0.03)
    u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
    sigT = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
    pdealg = ude.NNPDENS(u0, sigT, opt=opt)
    ans = ude.solve(prob, pdealg, verbose=True, maxiters=500, trajectories=m, alg=ude.LambaEM(), pabstol=1e-2, reltol=1e-4, abstol=1e-4, dtype=dtype)
    analytical_ans = bo.analytic_hjb(x0, tspan[1], lam, n_mc=10 ** 5)
    error_l2 = math.sqrt((ans - analytical_ans) ** 2 / ans ** 2)
    print(f"Hamilton Jacobi Bellman Equation ({dtype}): numerical = {ans:.4f}, analytical = {analytical_ans:.4f}, error_l2 = {error_l2:.4f}")
    assert error_l2 < 0.2


def test_bsde_error_paths():
    ude = _ude()
    from universal_differential_equations_b200._lib import B200UDEError, EINVAL, ESTATE
    prob, alg = _problem(ude, 6, 8)
    s = ude.BSDESolver(prob, alg, 4, 16)
    with pytest.raises(B200UDEError) as e:
        s.loss_gradient(8, seed=1)
    assert e.value.code == ESTATE
    with pytest.raises(B200UDEError) as e:
        s.set_params(np.zeros(s.P + 1, np.float32))
    assert e.value.code == EINVAL
    s.set_params(bo.init_params(6, 8).astype(np.float32))
    with pytest.raises(B200UDEError) as e:
        s.loss_gradient(17, seed=1)
    assert e.value.code == EINVAL and "n_paths" in str(e.value)
    with pytest.raises(ValueError):
        ude.BSDESolver(prob, ude.NNPDENS(alg.u0, alg.u0), 4, 16)
    s.close()
