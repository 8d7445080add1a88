/*
 * ude_oracle.c -- CPU restatement of the reference's UDE hot path (C99 + OpenMP).
 * TEST INFRASTRUCTURE ONLY -- see ude_oracle.h for scope, citations and the
 * pinning status.  Build: make -C oracle   (-> oracle/libude_oracle.so)
 */
#include "ude_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define UDE_MAX_STAGES 10

typedef struct {
    int s, order, fsal;
    double c[UDE_MAX_STAGES];
    double A[UDE_MAX_STAGES][UDE_MAX_STAGES];
    double b[UDE_MAX_STAGES];
    double bt[UDE_MAX_STAGES];
} ude_tableau;

/* Tsitouras 5(4), "Runge-Kutta pairs of order 5(4) satisfying only the first column
 * simplifying assumption" (2011), as used by OrdinaryDiffEq's Tsit5 (call sites:
 * Fisher-KPP-CNN.jl:66,132,136; seir_exposure.jl:66,132; scenario_1.jl:191).
 * Order of this block = order of the constants OrdinaryDiffEq serialized into the
 * reference's Scenario_1_recovery_0.005.jld2 (checked by tests/test_oracle_golden.py). */
static const double tsit5_flat[56] = {
    /* c1..c6 */ 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0,
    /* a21 */ 0.161,
    /* a31 a32 */ -0.008480655492356989, 0.335480655492357,
    /* a41.. */ 2.8971530571054935, -6.359448489975075, 4.3622954328695815,
    /* a51.. */ 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525,
    /* a61.. */ 5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
    -0.028269050394068383,
    /* a71.. (= b) */ 0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
    -3.290069515436081, 2.324710524099774,
    /* btilde1..7 */ -0.001780011052225777, -0.0008164344596567469, 0.007880878010261995,
    -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152,
    /* r11 r12 r13 r14 */ 1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216,
    /* r22 r23 r24 */ 0.13169999999999998, -0.2234, 0.1017,
    /* r32.. */ 3.9302962368947516, -5.941033872131505, 2.490627285651253,
    /* r42.. */ -12.411077166933676, 30.33818863028232, -16.548102889244902,
    /* r52.. */ 37.50931341651104, -88.1789048947664, 47.37952196281928,
    /* r62.. */ -27.896526289197286, 65.09189467479366, -34.87065786149661,
    /* r72.. */ 1.5, -4.0, 2.5,
};
static const double *const ude_tsit5_r = tsit5_flat + 34;

/* Verner's "most efficient" 7(6) pair as used by OrdinaryDiffEq's Vern7 (call sites:
 * scenario_1.jl:41,84; seir_exposure.jl:37,69,138; hudson_bay.jl:99,116): c2..c8, then the
 * nonzero a_ij row by row (pattern below), then b1,b4..b9, then btilde1,btilde4..btilde10. */
static const double vern7_flat[58] = {
    0.005, 0.10888888888888888, 0.16333333333333333, 0.4555, 0.6095094489978381, 0.884, 0.925,
    0.005,
    -1.07679012345679, 1.185679012345679,
    0.04083333333333333, 0.1225,
    0.6389139236255726, -2.455672638223657, 2.272258714598084,
    -2.6615773750187572, 10.804513886456137, -8.3539146573962, 0.820487594956657,
    6.067741434696772, -24.711273635911088, 20.427517930788895, -1.9061579788166472,
    1.006172249242068,
    12.054670076253203, -49.75478495046899, 41.142888638604674, -4.461760149974004,
    2.042334822239175, -0.09834843665406107,
    10.138146522881808, -42.6411360317175, 35.76384003992257, -4.3480228403929075,
    2.0098622683770357, 0.3487490460338272, -0.27143900510483127,
    -45.030072034298676, 187.3272437654589, -154.02882369350186, 18.56465306347536,
    -7.141809679295079, 1.3088085781613787,
    0.04715561848627222, 0.25750564298434153, 0.26216653977412624, 0.15216092656738558,
    0.4939969170032485, -0.29430311714032503, 0.08131747232495111,
    0.002547011879931045, -0.00965839487279575, 0.04206470975639691, -0.0666822437469301,
    0.2650097464621281, -0.29430311714032503, 0.08131747232495111, -0.02029518466335628,
};

static ude_tableau tab_tsit5, tab_vern7;
static int tabs_ready = 0;

static void build_tableaus(void)
{
    if (tabs_ready) return;
    ude_tableau *t = &tab_tsit5;
    memset(t, 0, sizeof(*t));
    t->s = 7; t->order = 5; t->fsal = 1;
    for (int i = 0; i < 6; ++i) t->c[i + 1] = tsit5_flat[i];
    int p = 6;
    for (int i = 1; i < 7; ++i)
        for (int j = 0; j < i; ++j) t->A[i][j] = tsit5_flat[p++];
    for (int j = 0; j < 6; ++j) t->b[j] = t->A[6][j];
    for (int j = 0; j < 7; ++j) t->bt[j] = tsit5_flat[p++];

    t = &tab_vern7;
    memset(t, 0, sizeof(*t));
    t->s = 10; t->order = 7; t->fsal = 0;
    for (int i = 0; i < 7; ++i) t->c[i + 1] = vern7_flat[i];
    t->c[8] = 1.0; t->c[9] = 1.0;
    /* sparsity: row2:{1} row3:{1,2} row4:{1,3} row5:{1,3,4} row6:{1,3,4,5} row7:{1,3..6}
     * row8:{1,3..7} row9:{1,3..8} row10:{1,3..7}   (1-based stage numbers) */
    p = 7;
    t->A[1][0] = vern7_flat[p++];
    t->A[2][0] = vern7_flat[p++]; t->A[2][1] = vern7_flat[p++];
    for (int i = 3; i <= 8; ++i) {
        t->A[i][0] = vern7_flat[p++];
        for (int j = 2; j < i; ++j) t->A[i][j] = vern7_flat[p++];
    }
    t->A[9][0] = vern7_flat[p++];
    for (int j = 2; j <= 6; ++j) t->A[9][j] = vern7_flat[p++];
    t->b[0] = vern7_flat[p++];
    for (int j = 3; j <= 8; ++j) t->b[j] = vern7_flat[p++];
    t->bt[0] = vern7_flat[p++];
    for (int j = 3; j <= 9; ++j) t->bt[j] = vern7_flat[p++];
    tabs_ready = 1;
}

static const ude_tableau *ude_get_tableau(int solver)
{
    build_tableaus();
    return solver == UDE_VERN7 ? &tab_vern7 : &tab_tsit5;
}

void ude_tsit5_constants(double *c56) { memcpy(c56, tsit5_flat, sizeof(tsit5_flat)); }
void ude_vern7_constants(double *c58) { memcpy(c58, vern7_flat, sizeof(vern7_flat)); }

size_t ude_num_params(const ude_model *m)
{
    size_t o = (size_t)m->n_prefix + (size_t)m->n_suffix;
    for (int l = 0; l < m->n_layers; ++l)
        o += (size_t)m->widths[l] * m->widths[l + 1] + m->widths[l + 1];
    return o;
}

#define REAL double
#define SUF _f64
#define R_TANH tanh
#define R_EXP exp
#define R_FABS fabs
#define R_SQRT sqrt
#define R_POW pow
#define R_LOG10 log10
#include "ude_oracle_impl.h"
#undef REAL
#undef SUF
#undef R_TANH
#undef R_EXP
#undef R_FABS
#undef R_SQRT
#undef R_POW
#undef R_LOG10

#define REAL float
#define SUF _f32
#define R_TANH tanhf
#define R_EXP expf
#define R_FABS fabsf
#define R_SQRT sqrtf
#define R_POW powf
#define R_LOG10 log10f
#include "ude_oracle_impl.h"
