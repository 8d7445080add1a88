// vern7.cuh -- Verner's "most efficient" 7(6) pair (the update weights; 9 stages, the 10th only feeds the error
// estimate) as compile-time constants.  Replaces OrdinaryDiffEq's Vern7ConstantCache used by the reference at
// LotkaVolterra/scenario_1.jl:41,84, SEIR_exposure/seir_exposure.jl:37,69,138, hudson_bay.jl:99,116.  Values:
// J.H. Verner's published pair; checked against the copy OrdinaryDiffEq serialized into the reference's
// Scenario_1_recovery_0.005.jld2 through the oracle (tests/test_oracle_golden.py) and the GPU parity test.
#pragma once

namespace b200ude {

struct Vern7 {
    static constexpr int S = 9;   // stages that enter the solution update
    __host__ __device__ static constexpr double a(int i, int j)
    {
        switch (i * 16 + j) {
        case 1 * 16 + 0: return 0.005;
        case 2 * 16 + 0: return -1.07679012345679;
        case 2 * 16 + 1: return 1.185679012345679;
        case 3 * 16 + 0: return 0.04083333333333333;
        case 3 * 16 + 2: return 0.1225;
        case 4 * 16 + 0: return 0.6389139236255726;
        case 4 * 16 + 2: return -2.455672638223657;
        case 4 * 16 + 3: return 2.272258714598084;
        case 5 * 16 + 0: return -2.6615773750187572;
        case 5 * 16 + 2: return 10.804513886456137;
        case 5 * 16 + 3: return -8.3539146573962;
        case 5 * 16 + 4: return 0.820487594956657;
        case 6 * 16 + 0: return 6.067741434696772;
        case 6 * 16 + 2: return -24.711273635911088;
        case 6 * 16 + 3: return 20.427517930788895;
        case 6 * 16 + 4: return -1.9061579788166472;
        case 6 * 16 + 5: return 1.006172249242068;
        case 7 * 16 + 0: return 12.054670076253203;
        case 7 * 16 + 2: return -49.75478495046899;
        case 7 * 16 + 3: return 41.142888638604674;
        case 7 * 16 + 4: return -4.461760149974004;
        case 7 * 16 + 5: return 2.042334822239175;
        case 7 * 16 + 6: return -0.09834843665406107;
        case 8 * 16 + 0: return 10.138146522881808;
        case 8 * 16 + 2: return -42.6411360317175;
        case 8 * 16 + 3: return 35.76384003992257;
        case 8 * 16 + 4: return -4.3480228403929075;
        case 8 * 16 + 5: return 2.0098622683770357;
        case 8 * 16 + 6: return 0.3487490460338272;
        case 8 * 16 + 7: return -0.27143900510483127;
        // 10th stage: error estimate only
        case 9 * 16 + 0: return -45.030072034298676;
        case 9 * 16 + 2: return 187.3272437654589;
        case 9 * 16 + 3: return -154.02882369350186;
        case 9 * 16 + 4: return 18.56465306347536;
        case 9 * 16 + 5: return -7.141809679295079;
        case 9 * 16 + 6: return 1.3088085781613787;
        default: return 0.0;
        }
    }
    __host__ __device__ static constexpr double b(int j)
    {
        switch (j) {
        case 0: return 0.04715561848627222;
        case 3: return 0.25750564298434153;
        case 4: return 0.26216653977412624;
        case 5: return 0.15216092656738558;
        case 6: return 0.4939969170032485;
        case 7: return -0.29430311714032503;
        case 8: return 0.08131747232495111;
        default: return 0.0;
        }
    }
    // error weights: err = dt * sum_j bt(j) k_j over the 10 stages
    __host__ __device__ static constexpr double bt(int j)
    {
        switch (j) {
        case 0: return 0.002547011879931045;
        case 3: return -0.00965839487279575;
        case 4: return 0.04206470975639691;
        case 5: return -0.0666822437469301;
        case 6: return 0.2650097464621281;
        case 7: return -0.29430311714032503;
        case 8: return 0.08131747232495111;
        case 9: return -0.02029518466335628;
        default: return 0.0;
        }
    }
};

}  // namespace b200ude
