"""Study for the next round: what a stabilised explicit solver of the ROCK2 class would buy on the
stiff 256-point Fisher-KPP grid.  ROCK2's coefficient tables are not in the reference, but second-order Runge-Kutta-Chebyshev
(RKC2, Sommeijer / Shampine / Verwer 1998) has closed-form coefficients and the same purpose.  CPU / numpy only.

Prints, for the BASELINE config-4 shape (Nx = 256, reaction chain 1-16-16-1, D0 = D / dx^2 = 650), the number of RHS evaluations
and the error at t = 1 of fixed-step Tsit5 at its stability limit and of RKC2 at several step sizes."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import glorot_theta

Nx = 256
D0 = 0.01 * (Nx - 1) ** 2
th = glorot_theta((1, 16, 16, 1), seed=3).astype(np.float64)
W1, b1 = th[0:16], th[16:32]
W2 = th[32:288].reshape(16, 16).T      # theta layout: W2[j, i] at 32 + i*16 + j
b2, W3, b3 = th[288:304], th[304:320], th[320]
NEVAL = 0


def rhs(u):
    global NEVAL
    NEVAL += 1
    h1 = np.tanh(np.outer(u, W1) + b1)
    h2 = np.tanh(h1 @ W2.T + b2)
    return h2 @ W3 + b3 + D0 * (np.roll(u, 1) - 2 * u + np.roll(u, -1))


A = np.zeros((7, 7))
A[1, :1] = [0.161]
A[2, :2] = [-0.008480655492356989, 0.335480655492357]
A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
A[6, :6] = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]


def tsit5(u, dt, n):
    k = [rhs(u)]
    for _ in range(n):
        ks = [k[0]]
        for i in range(1, 6):
            ks.append(rhs(u + dt * sum(A[i, j] * ks[j] for j in range(i))))
        u = u + dt * sum(A[6, j] * ks[j] for j in range(6))
        k = [rhs(u)]
    return u


def rkc_coeffs(s, eps=2.0 / 13.0):
    w0 = 1 + eps / s ** 2
    T = np.zeros(s + 1); dT = np.zeros(s + 1); d2T = np.zeros(s + 1)
    T[0], T[1], dT[1] = 1, w0, 1
    for j in range(2, s + 1):
        T[j] = 2 * w0 * T[j - 1] - T[j - 2]
        dT[j] = 2 * T[j - 1] + 2 * w0 * dT[j - 1] - dT[j - 2]
        d2T[j] = 4 * dT[j - 1] + 2 * w0 * d2T[j - 1] - d2T[j - 2]
    w1 = dT[s] / d2T[s]
    b = np.zeros(s + 1)
    for j in range(2, s + 1):
        b[j] = d2T[j] / dT[j] ** 2
    b[0] = b[1] = b[2]
    a = 1 - b * T
    beta = (w0 + 1) * d2T[s] / dT[s]
    return w0, w1, b, a, beta


def rkc2(u, dt, n, s):
    w0, w1, b, a, _ = rkc_coeffs(s)
    for _ in range(n):
        F0 = rhs(u)
        Y0, Y1 = u, u + b[1] * w1 * dt * F0
        for j in range(2, s + 1):
            mu, nu = 2 * b[j] * w0 / b[j - 1], -b[j] / b[j - 2]
            mut = 2 * b[j] * w1 / b[j - 1]
            gam = -a[j - 1] * mut
            Y2 = (1 - mu - nu) * u + mu * Y1 + nu * Y0 + mut * dt * rhs(Y1) + gam * dt * F0
            Y0, Y1 = Y1, Y2
        u = Y1
    return u


x = np.linspace(0, 1, Nx)
u0 = 0.5 * (np.tanh((x - 0.35) / 0.03) - np.tanh((x - 0.65) / 0.03))
T_end = 1.0
rho = 4 * D0 + 2.0     # spectral radius bound: stencil + |d reaction / du|
NEVAL = 0
ref = tsit5(u0, 2.5e-4, int(round(T_end / 2.5e-4)))
print(f"Nx={Nx} D0={D0:.1f} rho~{rho:.0f}; reference: Tsit5 dt=2.5e-4 ({NEVAL} RHS evaluations)")
dt = 1.25e-3   # 3.3 / rho ~ 1.27e-3: the largest stable Tsit5 step
NEVAL = 0
e = np.abs(tsit5(u0, dt, int(round(T_end / dt))) - ref).max()
print(f"Tsit5  dt={dt:.5f}: {NEVAL:6d} RHS evaluations, max error {e:.2e}")
for dt in (0.0125, 0.025, 0.05, 0.1):
    s = 2
    while rkc_coeffs(s)[4] < 1.05 * dt * rho:
        s += 1
    NEVAL = 0
    e = np.abs(rkc2(u0, dt, int(round(T_end / dt)), s) - ref).max()
    print(f"RKC2   dt={dt:.5f}: s={s:3d} stages, {NEVAL:6d} RHS evaluations, max error {e:.2e}")
