// plant.hpp -- stand-in for the reference's MuJoCo plant wrapper.
//
// The reference drives a MuJoCo 2.0 point mass through `PointMassEnv`
// (include/mppi_env.hpp:21-34, src/PointMassEnv.cpp:115-198): simulate(u) applies the controls
// and advances the simulation by >= 1/60 s, get_x(x) returns [qpos, qvel], step(x,u) advances
// one timestep.  MuJoCo 2.0 (closed binary, expired key), GLFW and GLEW are not available, so
// this class keeps that surface and integrates the same rigid body analytically described by
// envs/point_mass{1,2,3}d.xml: a sphere r = 0.05 of density 1000 (default) on A slide joints,
// armature 0.01, damping 0.1, joint range +-1.4, motors of gear 10 with ctrl clamped to
// +-1, gravity 0, RK4 at 0.01 s.  Two models:
//   kMjcf  : that body, 2 substeps per simulate() (0.02 s >= 1/60 s, PointMassEnv.cpp:136-139)
//   kIdeal : the controller's own model, x' = G x + B u with dt (src/model_missmatch.cpp:26-38)
#ifndef MPPI_CPP_PLANT_HPP_
#define MPPI_CPP_PLANT_HPP_

#include <algorithm>
#include <cmath>
#include <vector>

class PointMassEnv {
public:
    enum Model { kMjcf, kIdeal };

    PointMassEnv(int act_dim, Model model, float ctrl_dt, double sim_end = 10.0)
        : A_(act_dim), model_(model), ctrl_dt_(ctrl_dt), sim_end_(sim_end), q_(act_dim, 0.0),
          v_(act_dim, 0.0)
    {
        mass_ = mjcf_mass();
    }

    // sphere r = 0.05 of density 1000 + armature 0.01 (envs/point_mass2d.xml:6-10,28-29)
    static double mjcf_mass()
    {
        const double kPi = 3.14159265358979323846;
        return 1000.0 * 4.0 / 3.0 * kPi * 0.05 * 0.05 * 0.05 + 0.01;
    }

    // returns true when the episode is over (reference: window closed or time > simend)
    bool simulate(const float *u)
    {
        if (time_ >= sim_end_) return true;
        if (model_ == kIdeal) {
            const double dt = ctrl_dt_;
            for (int i = 0; i < A_; ++i) {
                const double a = u[i];
                q_[i] = q_[i] + dt * v_[i] + dt * dt / 2.0 * a;
                v_[i] = v_[i] + dt * a;
            }
            time_ += dt;
        } else {
            const double start = time_;
            while (time_ - start < 1.0 / 60.0) substep(u);
        }
        return false;
    }

    void step(float *x, const float *u)
    {
        substep(u);
        get_x(x);
    }

    void get_x(float *x) const
    {
        for (int i = 0; i < A_; ++i) { x[i] = (float)q_[i]; x[i + A_] = (float)v_[i]; }
    }

    void set_x(const float *x)
    {
        for (int i = 0; i < A_; ++i) { q_[i] = x[i]; v_[i] = x[i + A_]; }
    }

    double time() const { return time_; }

private:
    // one RK4 step of  m qdd = gear*clamp(u) - damping*qd  at h = 0.01 s, joint range +-1.4
    void substep(const float *u)
    {
        const double h = 0.01, gear = 10.0, damping = 0.1, range = 1.4;
        for (int i = 0; i < A_; ++i) {
            const double f = gear * std::min(1.0, std::max(-1.0, (double)u[i]));
            auto acc = [&](double vel) { return (f - damping * vel) / mass_; };
            const double k1v = acc(v_[i]), k1q = v_[i];
            const double k2v = acc(v_[i] + 0.5 * h * k1v), k2q = v_[i] + 0.5 * h * k1v;
            const double k3v = acc(v_[i] + 0.5 * h * k2v), k3q = v_[i] + 0.5 * h * k2v;
            const double k4v = acc(v_[i] + h * k3v), k4q = v_[i] + h * k3v;
            q_[i] += h / 6.0 * (k1q + 2 * k2q + 2 * k3q + k4q);
            v_[i] += h / 6.0 * (k1v + 2 * k2v + 2 * k3v + k4v);
            if (q_[i] > range) { q_[i] = range; v_[i] = std::min(v_[i], 0.0); }
            if (q_[i] < -range) { q_[i] = -range; v_[i] = std::max(v_[i], 0.0); }
        }
        time_ += h;
    }

    int A_;
    Model model_;
    float ctrl_dt_;
    double sim_end_, time_ = 0.0, mass_;
    std::vector<double> q_, v_;
};

#endif
