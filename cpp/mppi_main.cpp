// mppi_main.cpp -- the reference's driver loop (src/main.cu:220-399) on the B200 core.
//
//   mppi_main -c <config.yaml> [-t traj.csv] [-s step_prefix] [--plant ideal|mjcf] [--steps N] [--samples K]
//             [--horizon T] [--honour-config] [--seed S] [--devices 0,1,..] [--verify-config] [--quiet]
//             [--model ideal|mjcf] [--plant-us N] [--flags BITS [--exact-flags]] [--terminal-w w0,w1,..]
//             [--philox-rounds 7|10]
//
// --plant-us N makes the plant's turn take N microseconds of host time (the reference steps
// MuJoCo there); --flags adds MPPI_FLAG_* bits to the default MPPI_FLAG_AUTO_CHAIN;
// --terminal-w gives the final state of every rollout cost weights of its own (state_dim values).
//
// --model mjcf gives the controller the dynamics of the MJCF body instead of the reference's
// double integrator (MPPI_MODEL_LINEAR_AXIS: the damped, geared point mass of envs/*.xml
// discretised at the plant's control period) -- the model-mismatch experiment of the
// reference's src/model_missmatch.cpp with the mismatch removed.
//
// Same sequence as the reference: parse config, build the plant and the controller, get_x,
// zero action sequence, memcpy_set_data, then loop { get_u; time(get_act); simulate; get_x;
// record; set_x } until the plant says done, print "Average controller execution time", write
// the trajectory CSV in the reference's format (to_csv_traj, src/main.cu:32-57; header
// x,y,vx,vy,ux,uy,size_x,size_u for 2-D, generalised to x0..,v0..,u0.. otherwise).
// The controller is the reference's class name, provided by the shim header.
//
// --honour-config passes lambda / noise / init-act / max-a to the controller (the reference
// parses and drops them, src/main.cu:311); without it the reference-compatible preset runs.
// --verify-config restates the reference's parser self-test (src/main.cu:295-307,686-725) and
// needs no GPU.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../include/mppi_b200/point_mass.hpp"
#include "config.hpp"
#include "plant.hpp"

typedef std::chrono::steady_clock Clock;

static void to_csv_traj(const std::string &filename, const std::vector<std::vector<float>> &x,
                        const std::vector<std::vector<float>> &u, int A)
{
    std::ofstream out(filename);
    std::cout << "Saving traj to file...: " << std::endl;
    static const char *axes[] = {"x", "y", "z", "w"};
    for (int i = 0; i < A; ++i) out << axes[i] << ",";
    for (int i = 0; i < A; ++i) out << "v" << axes[i] << ",";
    for (int i = 0; i < A; ++i) out << "u" << axes[i] << ",";
    out << "size_x,size_u" << std::endl;
    for (size_t r = 0; r < u.size(); ++r) {
        for (int i = 0; i < 2 * A; ++i) out << x[r][i] << ",";
        for (int i = 0; i < A; ++i) out << u[r][i] << (i + 1 < A || r == 0 ? "," : "");
        if (r == 0) out << x.size() << "," << u.size();
        out << std::endl;
    }
    for (int i = 0; i < 2 * A; ++i) out << x[u.size()][i] << (i + 1 < 2 * A ? "," : "");
    out.close();
    std::cout << x.size() << " " << u.size() << std::endl;
}

// Per-step introspection dump, the reference's to_csv2 (src/main.cu:90-156) that its
// scripts/plot_csv.py reads: one line per (sample, time) with state, noise, and -- on the
// first sample's lines -- U after/before the update; cost and weight of sample n sit on
// line n.  Column names are the reference's for the 2-D case (sample,x,y,x_dot,y_dot,e_x,e_y,
// u[0],u[1],u_prev[0],u_prev[1],c,w) and are generalised with axis names otherwise.
static void to_csv2(const std::string &filename, const float *x, const float *u, const float *u_prev,
                    const float *e, const float *cost, const float *w, int sample, int size,
                    int s_dim, int a_dim)
{
    static const char *axes[] = {"x", "y", "z", "w"};
    std::cout << "Saving data to file...: " << std::flush;
    std::ofstream out(filename);
    out << "sample";
    for (int i = 0; i < a_dim; ++i) out << "," << axes[i];
    for (int i = 0; i < a_dim; ++i) out << "," << axes[i] << "_dot";
    for (int i = 0; i < a_dim; ++i) out << ",e_" << axes[i];
    for (int d = 0; d < a_dim; ++d) out << ",u[" << d << "]";
    for (int d = 0; d < a_dim; ++d) out << ",u_prev[" << d << "]";
    out << ",c,w" << std::endl;
    for (int i = 0; i < sample; i++) {
        for (int j = 0; j < size + 1; j++) {
            out << i;
            for (int d = 0; d < s_dim; ++d) out << "," << x[((size_t)i * (size + 1) + j) * s_dim + d];
            for (int d = 0; d < a_dim; ++d) {
                out << ",";
                if (j < size) out << e[((size_t)i * size + j) * a_dim + d];
                else out << " ";
            }
            for (int d = 0; d < a_dim; ++d) {
                out << ",";
                if (i < 1 && j < size) out << u[j * a_dim + d]; else out << " ";
            }
            for (int d = 0; d < a_dim; ++d) {
                out << ",";
                if (i < 1 && j < size) out << u_prev[j * a_dim + d]; else out << " ";
            }
            const long n = (long)i * (size + 1) + j;
            if (n < sample) out << "," << cost[n] << "," << w[n];
            out << std::endl;
        }
    }
    std::cout << "Done" << std::endl;
}

int main(int argc, char **argv)
{
    std::string config_file = "config/point_mass2d.yaml", traj_file, step_file, plant_name = "ideal";
    std::string model_name = "ideal";
    long max_steps = -1, samples_override = -1, horizon_override = -1;
    bool honour = false, verify = false, quiet = false;
    unsigned long long seed = 0;
    long plant_us = 0;             // emulated compute time of the plant's step (busy wait)
    unsigned extra_flags = 0;      // MPPI_FLAG_* bits on top of the default (auto chain)
    bool exact_flags = false;
    std::vector<float> terminal_w;
    int philox_rounds = 10;
    std::vector<int> devices;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) { std::cerr << "missing value for " << a << std::endl; exit(2); }
            return argv[++i];
        };
        if (a == "-c" || a == "--config") config_file = next();
        else if (a == "-t" || a == "--traj-save") traj_file = next();
        else if (a == "-s" || a == "--step-save") step_file = next();
        else if (a == "--plant") plant_name = next();
        else if (a == "--model") model_name = next();
        else if (a == "--steps") max_steps = std::stol(next());
        else if (a == "--samples") samples_override = std::stol(next());
        else if (a == "--horizon") horizon_override = std::stol(next());
        else if (a == "--seed") seed = std::stoull(next());
        else if (a == "--plant-us") plant_us = std::stol(next());
        else if (a == "--flags") extra_flags = (unsigned)std::stoul(next(), nullptr, 0);
        else if (a == "--exact-flags") exact_flags = true;   // --flags replaces the default
        else if (a == "--philox-rounds") philox_rounds = std::stoi(next());
        else if (a == "--terminal-w") {
            std::string list = next();
            size_t pos = 0;
            while (pos <= list.size()) {
                size_t c = list.find(',', pos);
                if (c == std::string::npos) c = list.size();
                if (c > pos) terminal_w.push_back(std::stof(list.substr(pos, c - pos)));
                pos = c + 1;
            }
        }
        else if (a == "--devices") {           // e.g. --devices 0,1,2,3 : K sharded over GPUs
            std::string list = next();
            size_t pos = 0;
            while (pos <= list.size()) {
                size_t c = list.find(',', pos);
                if (c == std::string::npos) c = list.size();
                if (c > pos) devices.push_back(std::stoi(list.substr(pos, c - pos)));
                pos = c + 1;
            }
        }
        else if (a == "--honour-config") honour = true;
        else if (a == "--verify-config") verify = true;
        else if (a == "--quiet") quiet = true;
        else { std::cerr << "unknown argument " << a << std::endl; return 2; }
    }

    mppi_cfg::Config cfg = mppi_cfg::parse(config_file);
    std::cout << "N " << cfg.samples << " steps: " << cfg.horizon << " State dim: " << cfg.state_dim
              << std::endl;
    if (verify) {
        if (!mppi_cfg::verify_test_config(cfg)) { std::cout << "Test FAILED" << std::endl; return 1; }
        std::cout << "Test passed" << std::endl;
        return 0;
    }
    int n = samples_override > 0 ? (int)samples_override : cfg.samples;
    int steps = horizon_override > 0 ? (int)horizon_override : cfg.horizon;
    const int state_dim = cfg.state_dim, act_dim = cfg.act_dim;

    std::vector<float> next_act(act_dim), init_state(state_dim), init_actions((size_t)steps * act_dim, 0.f);
    std::vector<float> u_prev((size_t)steps * act_dim);
    std::vector<std::vector<float>> x, u;

    PointMassEnv env(act_dim, plant_name == "mjcf" ? PointMassEnv::kMjcf : PointMassEnv::kIdeal,
                     cfg.dt, max_steps > 0 ? 1e30 : 10.0);

    PointMassModel::Options opt;
    opt.seed = seed;
    opt.philox_rounds = philox_rounds;
    opt.flags = exact_flags ? extra_flags : (opt.flags | extra_flags);
    if (!devices.empty()) { opt.devices = devices.data(); opt.num_devices = (int)devices.size(); }
    float state_gain[4], act_gain[2];
    if (model_name == "mjcf") {
        // m qdd = gear u - damping qd over one control period h of the mjcf plant (2 x 0.01 s)
        const double h = 0.02, mass = PointMassEnv::mjcf_mass(), gear = 10.0, damping = 0.1;
        const double k = damping / mass;
        state_gain[0] = 1.0f; state_gain[1] = (float)(h - 0.5 * h * h * k);
        state_gain[2] = 0.0f; state_gain[3] = (float)(1.0 - h * k);
        act_gain[0] = (float)(0.5 * h * h * gear / mass);
        act_gain[1] = (float)(h * gear / mass);
        opt.state_gain = state_gain;
        opt.act_gain = act_gain;
    } else if (model_name != "ideal") {
        std::cerr << "unknown --model " << model_name << std::endl;
        return 2;
    }
    if (honour) {
        opt.lambda = cfg.lambda;
        opt.sigma = cfg.noise.data();
        opt.init_act = cfg.init_act.data();
        opt.max_act = cfg.max_a.data();
    }
    PointMassModel *model = new PointMassModel(n, steps, cfg.dt, state_dim, act_dim, false, opt);

    env.get_x(init_state.data());
    model->memcpy_set_data(init_state.data(), init_actions.data(), cfg.goal.data(), cfg.cost_w.data());
    if (!terminal_w.empty()) {
        if ((int)terminal_w.size() != state_dim) {
            std::cerr << "--terminal-w needs " << state_dim << " values" << std::endl;
            return 2;
        }
        model->set_terminal_weights(terminal_w.data());
    }
    x.push_back(init_state);

    std::vector<double> lat_ms;
    bool done = false;
    long t = 0;
    while (!done) {
        model->get_u(u_prev.data());
        auto t1 = Clock::now();
        model->get_act(next_act.data());
        auto t2 = Clock::now();
        lat_ms.push_back(std::chrono::duration<double, std::milli>(t2 - t1).count());
        if (!quiet) {
            std::cout << "next_act: ";
            for (int i = 0; i < act_dim; i++) std::cout << next_act[i] << " ";
            std::cout << std::endl;
        }
        done = env.simulate(next_act.data());
        if (plant_us > 0) {
            // a real plant (the reference steps MuJoCo here, src/main.cu:335-337) takes time:
            // emulate it, so that latency is measured as a closed loop sees it
            const auto until = Clock::now() + std::chrono::microseconds(plant_us);
            while (Clock::now() < until) { }
        }
        env.get_x(init_state.data());
        u.push_back(next_act);
        x.push_back(init_state);
        if (!step_file.empty()) {        // save_step branch of the reference, src/main.cu:355-368
            std::vector<float> h_x((size_t)n * (steps + 1) * state_dim), h_u((size_t)steps * act_dim),
                h_e((size_t)n * steps * act_dim), cost(n), weight(n);
            float beta = 0, nabla = 0;
            model->get_inf(h_x.data(), h_u.data(), h_e.data(), cost.data(), &beta, &nabla, weight.data());
            to_csv2(step_file + std::to_string(t), h_x.data(), h_u.data(), u_prev.data(), h_e.data(),
                    cost.data(), weight.data(), n, steps, state_dim, act_dim);
        }
        model->set_x(init_state.data());
        t += 1;
        if (max_steps > 0 && t >= max_steps) done = true;
    }
    double total = 0;
    for (double v : lat_ms) total += v;
    std::vector<double> sorted = lat_ms;
    std::sort(sorted.begin(), sorted.end());
    std::cout << "Average controller execution time: " << total / t << std::endl;
    std::cout << "T: " << t << std::endl;
    std::cout << "Delta: " << total << std::endl;
    std::cout << "latency_ms p50: " << sorted[sorted.size() / 2]
              << " p99: " << sorted[std::min(sorted.size() - 1, (size_t)(0.99 * sorted.size()))]
              << " rollout-steps/s: " << (double)n * steps / (total / t * 1e-3) << std::endl;
    std::cout << "final state:";
    for (int i = 0; i < state_dim; ++i) std::cout << " " << init_state[i];
    std::cout << std::endl;
    if (!traj_file.empty()) to_csv_traj(traj_file, x, u, act_dim);
    delete model;
    return 0;
}
