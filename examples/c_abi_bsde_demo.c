/*
 * c_abi_bsde_demo.c -- the terminal-PDE entry points (b200ude_bsde_*, highdim_pde/lambaem.jl's NNPDENS solve) from plain C99.
 *
 *   c_abi_bsde_demo <in.bin> <out.bin>
 * in.bin  : int32 d, hls, n_steps, n_paths; uint64 seed; double x0[d]; double theta[P]
 * out.bin : double loss, u0, grad[P]; double loss_after[3] (three ADAM(0.03) iterations on the device)
 * One evaluation of loss = mean (g(X_T) - u_T)^2 and its gradient in fp64, host pointers in and out.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200ude.h"

static int die(const char *what, b200ude_bsde_handle *h, int32_t rc)
{
    fprintf(stderr, "%s failed (%d): %s\n", what, (int)rc, b200ude_bsde_last_error(h));
    return 1;
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 2; }
    int32_t hd[4];
    uint64_t seed = 0;
    if (fread(hd, 4, 4, fi) != 4 || fread(&seed, 8, 1, fi) != 1) return 2;
    const int32_t d = hd[0], hls = hd[1], n_steps = hd[2], n_paths = hd[3];
    double *x0 = malloc(sizeof(double) * (size_t)d);
    if (fread(x0, 8, (size_t)d, fi) != (size_t)d) return 2;

    b200ude_bsde_desc desc;
    memset(&desc, 0, sizeof desc);
    desc.struct_size = sizeof desc;
    desc.device = 0;
    desc.dtype = B200UDE_F64;
    desc.dim = d; desc.hidden = hls; desc.n_steps = n_steps;
    desc.T = 1.0; desc.lambda = 1.0; desc.sigma = 1.4142135623730951; desc.g_a = 0.5; desc.g_b = 0.5;   /* lambaem.jl:10-17 */
    desc.x0 = x0;
    desc.max_paths = (uint64_t)n_paths;
    b200ude_bsde_handle *h = NULL;
    int32_t rc = b200ude_bsde_create(&desc, &h);
    if (rc) return die("b200ude_bsde_create", NULL, rc);
    const size_t P = b200ude_bsde_num_params(h);
    double *theta = malloc(sizeof(double) * P), *grad = malloc(sizeof(double) * P);
    if (fread(theta, 8, P, fi) != P) return die("theta size", h, -1);
    fclose(fi);

    rc = b200ude_bsde_loss_gradient(h, (size_t)n_paths, seed, 0, 0, NULL, NULL, NULL);
    if (rc != B200UDE_ESTATE) return die("loss_gradient before set_params should be ESTATE", h, rc);
    rc = b200ude_bsde_set_params(h, theta, P, B200UDE_HOST);
    if (rc) return die("b200ude_bsde_set_params", h, rc);
    double loss = 0, u0 = 0;
    rc = b200ude_bsde_loss_gradient(h, (size_t)n_paths, seed, 0, 0, &loss, grad, &u0);
    if (rc) return die("b200ude_bsde_loss_gradient", h, rc);

    b200ude_adam opt;
    memset(&opt, 0, sizeof opt);
    opt.struct_size = sizeof opt;
    opt.eta = 0.03; opt.beta1 = 0.9; opt.beta2 = 0.999; opt.eps = 1e-8;
    double after[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {   /* host-driven: gradient -> ADAM step on the device -> next evaluation */
        rc = b200ude_bsde_train_adam(h, &opt, (size_t)n_paths, 1, seed + 1 + (uint64_t)i, NULL, NULL);
        if (rc) return die("b200ude_bsde_train_adam", h, rc);
        rc = b200ude_bsde_loss_gradient(h, (size_t)n_paths, seed, 0, 0, &after[i], NULL, NULL);
        if (rc) return die("b200ude_bsde_loss_gradient", h, rc);
    }

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 2; }
    fwrite(&loss, 8, 1, fo);
    fwrite(&u0, 8, 1, fo);
    fwrite(grad, 8, P, fo);
    fwrite(after, 8, 3, fo);
    fclose(fo);
    printf("d=%d paths=%d P=%zu loss=%.12g u0=%.12g\n", (int)d, (int)n_paths, P, loss, u0);
    b200ude_bsde_destroy(h);
    return 0;
}
