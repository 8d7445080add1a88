/*
 * c_abi_demo.c -- the drop-in boundary used from plain C99: no Python, no torch, only include/b200ude.h and
 * libb200ude.so (what a Julia `ccall`, a cgo or a JNI binding would do).
 *
 *   c_abi_demo <in.bin> <out.bin>
 * in.bin  : int32 N, int32 n_steps, float dt, float theta[1218], float u0[2][N], float data[n_steps+1][2][N]
 * out.bin : double loss, float grad_theta[1218], float grad_u0[2][N], float out[n_steps+1][2][N], int32 status[N],
 *           int32 repeat_identical (second call bit-equal to the first)
 * Workload: the LV UDE with the 2 -> 32 -> 32 -> 2 tanh chain (LotkaVolterra/scenario_1.jl:69-73 with BASELINE's widths),
 * Tsit5 fixed step, InterpolatingAdjoint, L2 loss -- through the host-buffer entry points.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200ude.h"

#define P 1218

static int die(const char *what, b200ude_handle *h, int32_t rc)
{
    fprintf(stderr, "%s failed (%d): %s\n", what, (int)rc, b200ude_last_error(h));
    return 1;
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 2; }
    int32_t N = 0, n_steps = 0;
    float dt = 0;
    if (fread(&N, 4, 1, fi) != 1 || fread(&n_steps, 4, 1, fi) != 1 || fread(&dt, 4, 1, fi) != 1) return 2;
    const size_t n_save = (size_t)n_steps + 1, nu = 2 * (size_t)N, nd = n_save * nu;
    float *theta = malloc(sizeof(float) * P), *u0 = malloc(sizeof(float) * nu), *data = malloc(sizeof(float) * nd);
    float *out = malloc(sizeof(float) * nd), *g = malloc(sizeof(float) * P), *gu = malloc(sizeof(float) * nu);
    float *g2 = malloc(sizeof(float) * P), *gu2 = malloc(sizeof(float) * nu);
    int32_t *status = malloc(sizeof(int32_t) * (size_t)N);
    if (fread(theta, 4, P, fi) != P || fread(u0, 4, nu, fi) != nu || fread(data, 4, nd, fi) != nd) return 2;
    fclose(fi);

    b200ude_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    d.device = 0;
    d.dtype = B200UDE_F32;
    d.model = B200UDE_MODEL_LV;
    d.state_dim = 2;
    d.n_layers = 3;
    d.widths[0] = 2; d.widths[1] = 32; d.widths[2] = 32; d.widths[3] = 2;
    d.acts[0] = B200UDE_ACT_TANH; d.acts[1] = B200UDE_ACT_TANH; d.acts[2] = B200UDE_ACT_IDENTITY;
    d.n_consts = 2;
    d.consts[0] = 1.3; d.consts[1] = 1.8;                 /* p_[1], p_[4] of scenario_1.jl:39 */
    d.solver = B200UDE_TSIT5;
    d.sensealg = B200UDE_INTERPOLATING_ADJOINT;
    d.t0 = 0.0; d.dt = dt; d.n_steps = n_steps; d.save_every = 1;
    d.max_trajectories = (uint64_t)N;

    b200ude_handle *h = NULL;
    int32_t rc = b200ude_create(&d, &h);
    if (rc) return die("b200ude_create", NULL, rc);
    if (b200ude_num_params(h) != P || b200ude_num_save(h) != n_save) return die("shape query", h, -1);

    rc = b200ude_solve_host(h, theta, u0, (size_t)N, out, status);
    if (rc) return die("b200ude_solve_host", h, rc);
    double loss = 0, loss2 = 0;
    rc = b200ude_loss_gradient_host(h, theta, u0, data, (size_t)N, &loss, g, gu);
    if (rc) return die("b200ude_loss_gradient_host", h, rc);
    rc = b200ude_loss_gradient_host(h, theta, u0, data, (size_t)N, &loss2, g2, gu2);
    if (rc) return die("b200ude_loss_gradient_host (repeat)", h, rc);
    const int32_t same = loss == loss2 && !memcmp(g, g2, sizeof(float) * P) && !memcmp(gu, gu2, sizeof(float) * nu);

    /* error behaviour: a size mismatch is a negative status with a message, not a crash */
    rc = b200ude_set_params(h, theta, P - 1, B200UDE_HOST, NULL);
    if (rc != B200UDE_EINVAL) return die("set_params(P-1) should be EINVAL", h, rc);

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 2; }
    fwrite(&loss, 8, 1, fo);
    fwrite(g, 4, P, fo);
    fwrite(gu, 4, nu, fo);
    fwrite(out, 4, nd, fo);
    fwrite(status, 4, (size_t)N, fo);
    fwrite(&same, 4, 1, fo);
    fclose(fo);
    printf("N=%d loss=%.9g device_bytes=%zu repeat_identical=%d\n", (int)N, loss, b200ude_device_bytes(h), (int)same);
    b200ude_destroy(h);
    return 0;
}
