"""B200-native drop-in for the UDE training path of ChrisRackauckas/universal_differential_equations.

Only the hot path is here: ensemble forward solve of a universal differential equation
(known physics + embedded dense chain, Tsit5) and its interpolating-adjoint gradient,
as hand-written sm_100a kernels behind a C ABI (include/b200ude.h), plus a host-side
mirror of the reference's call surface (sciml.py).  See DESIGN.md.
"""
from . import _lib  # noqa: F401
from .sciml import (  # noqa: F401
    ADAM, BFGS, Chain, Dense, EnsembleProblem, FastChain, FastDense, FisherKPPUDE, ForwardDiffSensitivity,
    InterpolatingAdjoint, LotkaVolterraUDE, NeuralODE, ODEProblem, SEIRExposureUDE, SEIRNeuralODE, ReverseDiffVJP, RKC2, ROCK2, Tsit5, UDESolver, Vern7,
    concrete_solve, identity, initial_params, rbf, remake, sciml_train, sciml_train_l2, solve, tanh,
)
from .dist import PeerAllReduce, shard_range, allreduce_loss_grad  # noqa: F401
