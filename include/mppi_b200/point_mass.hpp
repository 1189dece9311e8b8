// point_mass.hpp -- drop-in C++ shim: the reference's controller class on top of the
// B200 C ABI (include/mppi_b200.h).
//
// The reference's driver (src/main.cu:309-396) uses exactly these members of
// `class PointMassModel` (reference include/point_mass.hpp:23-44):
//     PointMassModel(n, steps, dt, state_dim, act_dim, verbose)        main.cu:311
//     memcpy_set_data(x, u, goal, w)                                    main.cu:316
//     get_u(u) / get_act(next_act) / get_inf(...) / set_x(x)            main.cu:327,330,361,371
//     delete model                                                      main.cu:396
// A translation unit that includes THIS header instead of the reference's point_mass.hpp
// compiles unchanged and links against libmppi_b200.so instead of point_mass.cu,
// point_mass_gpu.cu, cost.cu and mppi_utils.cu.
//
// Behaviour kept from the reference: all buffers are caller-owned host memory, every call
// is blocking, get_act leaves U already shifted, errors print and exit(1) like
// CUDA_CALL_CONST (reference include/mppi_utils.hpp:19-25).  Behaviour added: the
// quantities the reference hard-codes (lambda, sigma, Sigma^-1) and the config keys it
// parses but drops (`lambda`, `noise`, `init-act`, `max-a`) can be passed through
// `PointMassModel::Options`; the default Options reproduce the reference.
#ifndef MPPI_B200_POINT_MASS_HPP_
#define MPPI_B200_POINT_MASS_HPP_

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../mppi_b200.h"

#ifndef TOL
#define TOL 1e-6          /* reference include/point_mass.hpp:16 */
#endif

class PointMassModel {
public:
    struct Options {
        float lambda = 1.0f;
        const float *sigma = nullptr;      // [act_dim] or null -> 0.025
        const float *inv_sigma = nullptr;  // [act_dim] or null -> 1
        const float *init_act = nullptr;   // [act_dim]; non-null enables init-act re-init
        const float *max_act = nullptr;    // [act_dim]; non-null enables clamping
        unsigned long long seed = 0;
        unsigned flags = MPPI_FLAG_AUTO_CHAIN;   // MPPI_FLAG_* bits; default: the library picks
                                                 // the fastest kernel chain for the shard size
        int device = 0;
        const float *state_gain = nullptr; // [4] + act_gain [2]: MPPI_MODEL_LINEAR_AXIS (the gains
        const float *act_gain = nullptr;   // PointMassModelGpu::init takes) instead of the
                                           // double integrator formed from dt
        const int *devices = nullptr;      // non-null: shard K over these GPUs (one process)
        int num_devices = 0;
        int philox_rounds = 10;            // 7: Philox-4x32-7 (kernel chains only, see mppi_params)
    };

    PointMassModel(int nb_sim, int steps, float dt, int state_dim, int act_dim,
                   bool verbose = false)
        : PointMassModel(nb_sim, steps, dt, state_dim, act_dim, verbose, Options()) {}

    PointMassModel(int nb_sim, int steps, float dt, int state_dim, int act_dim, bool verbose,
                   const Options &o)
    {
        mppi_params p;
        check(mppi_params_default(&p), __LINE__);
        p.samples = nb_sim;
        p.horizon = steps;
        p.dt = dt;
        p.state_dim = state_dim;
        p.act_dim = act_dim;
        p.verbose = verbose ? 1 : 0;
        p.lambda = o.lambda;
        p.seed = o.seed;
        p.flags = o.flags;
        p.device = o.device;
        p.philox_rounds = o.philox_rounds;
        for (int a = 0; a < act_dim && a < MPPI_MAX_ACT; ++a) {
            if (o.sigma) p.sigma[a] = o.sigma[a];
            if (o.inv_sigma) p.inv_sigma[a] = o.inv_sigma[a];
            if (o.init_act) p.init_act[a] = o.init_act[a];
            if (o.max_act) p.max_act[a] = o.max_act[a];
        }
        if (o.state_gain && o.act_gain) {
            p.model = MPPI_MODEL_LINEAR_AXIS;
            for (int i = 0; i < 4; ++i) p.state_gain[i] = o.state_gain[i];
            p.act_gain[0] = o.act_gain[0];
            p.act_gain[1] = o.act_gain[1];
        }
        if (o.init_act) p.flags |= MPPI_FLAG_REINIT_INIT_ACT;
        if (o.max_act) p.flags |= MPPI_FLAG_CLAMP_ACTIONS;
        _n_sim = nb_sim; _steps = steps; _state_dim = state_dim; _act_dim = act_dim;
        if (o.devices && o.num_devices > 0)
            check(mppi_create_multi(&p, o.devices, o.num_devices, &_h), __LINE__);
        else
            check(mppi_create(&p, &_h), __LINE__);
    }

    ~PointMassModel() { if (_h) mppi_destroy(_h); }
    PointMassModel(const PointMassModel &) = delete;
    PointMassModel &operator=(const PointMassModel &) = delete;

    void get_act(float *next_act) { check(mppi_step(_h, next_act), __LINE__); }

    void memcpy_set_data(float *x, float *u, float *goal, float *w)
    {
        check(mppi_set_problem(_h, x, u, goal, w), __LINE__);
    }

    // not in the reference: the final state charged by a second Cost object (Cost::final_cost with
    // w_final [state_dim]); nullptr returns to the reference's single object
    void set_terminal_weights(const float *w_final)
    {
        check(mppi_set_terminal_weights(_h, w_final), __LINE__);
    }

    // not in the reference (README: "decouple x into q and q_dot"): the state as positions and
    // velocities; x = [q, q_dot] (src/point_mass_gpu.cu:97-106)
    void set_q(const float *q, const float *q_dot)
    {
        float x[2 * MPPI_MAX_ACT];
        for (int i = 0; i < _act_dim; ++i) { x[i] = q[i]; x[i + _act_dim] = q_dot[i]; }
        set_x(x);
    }

    // reference: copies x [K,(T+1),S] and e [K,T,A] (src/point_mass.cu:230-234)
    void memcpy_get_data(float *x_all, float *e)
    {
        check(mppi_get_info(_h, x_all, nullptr, e, nullptr, nullptr, nullptr, nullptr), __LINE__);
    }

    void get_inf(float *x, float *u, float *e, float *cost, float *beta, float *nabla,
                 float *weight)
    {
        check(mppi_get_info(_h, x, u, e, cost, beta, nabla, weight), __LINE__);
    }

    void set_x(float *x) { check(mppi_set_state(_h, x), __LINE__); }
    void get_u(float *u) { check(mppi_get_u(_h, u), __LINE__); }

    // declared but never defined in the reference (include/point_mass.hpp:34); defined here as
    // "the initial state the next step will start from" would need a getter in the ABI, so it
    // is left out on purpose: no caller exists (grep over src/).

    // ---- extensions (no reference counterpart)
    void set_noise(const float *e) { check(mppi_set_noise(_h, e), __LINE__); }
    void set_noise_mode(bool injected) { check(mppi_set_noise_mode(_h, injected ? 1 : 0), __LINE__); }
    void step_info(mppi_step_info *info) { check(mppi_get_step_info(_h, info), __LINE__); }
    // ControllerBase::next(x) of the reference's intended interface
    // (include/controller_base.hpp:9-17)
    void next(float *x, float *next_act) { set_x(x); get_act(next_act); }
    mppi_handle *handle() { return _h; }

private:
    void check(int rc, int line)
    {
        if (rc != MPPI_OK) {
            printf("API error failed %s:%d Returned: %d (%s)\n", __FILE__, line, rc,
                   mppi_last_error());
            exit(1);
        }
    }
    mppi_handle *_h = nullptr;
    int _n_sim = 0, _steps = 0, _state_dim = 0, _act_dim = 0;
};

#endif  // MPPI_B200_POINT_MASS_HPP_
