// model.cuh -- the plug-in point of the rollout: dynamics and cost as device functors.
//
// A rollout kernel is instantiated per MODEL = Model<STRICT, Dynamics> (no virtual calls, no
// per-sample objects: the reference keeps a 168-byte PointMassModelGpu and a 48-byte Cost per
// sample, include/point_mass_gpu.hpp:46-90, include/cost.hpp:32-46).  A functor has a scalar
// form (one sample) and a packed FP32x2 form (a pair of samples, two lanes of one
// instruction) with the same operations in the same order, so both give the same bits.
//
//   Dynamics::axis(p, v, ue) -> (p', v')      one action dimension of
//       PointMassModelGpu::step (src/point_mass_gpu.cu:97-106):
//       p' = g0*p + g1*v + b0*(u+e),  v' = g2*p + g3*v + b1*(u+e)
//     DoubleIntegrator  the gains the reference's constructor hard-codes, {1,dt,0,1} /
//                       {dt^2/2, dt} (src/point_mass.cu:46-51): g0*p == p and g2*p+g3*v == v
//                       exactly, which saves three operations per axis and step;
//     LinearAxis        any gains (PointMassModelGpu::init takes them as arguments,
//                       src/point_mass_gpu.cu:25-39): damped / geared point masses such as
//                       the MJCF plant of envs/point_mass2d.xml, README "generalize the code".
//   QuadraticCost       Cost::step_cost / final_cost (src/cost.cu:42-64).
//
// Arithmetic.  STRICT rounds every product and sum separately in source order (== the
// reference's host build, oracle ORACLE_ARITH_STRICT); otherwise the fused operations are
// exactly those nvcc -fmad=true forms for the reference's device build (ORACLE_ARITH_FMA):
// fma(b0,ue, fma(g0,p, g1*v)), fma((x-g)*w, (x-g), res), fma(u*inv_s, e, res).
//
// To add a model: write a Dynamics (or Cost) functor with load/axis/axis2, add a Model tag
// to MPPI_DISPATCH_MODEL (kernels.cu / step.cu) and restate it in oracle/mppi_oracle.c.
#pragma once

#include "common.cuh"

namespace mppi {

// ---------------------------------------------------------------------------------
// Packed FP32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2).  One instruction performs the
// IEEE operation on two independent float lanes, so two samples advance per issue slot with
// exactly the bits the scalar code produces.  The FMA pipe rate per FLOP is unchanged
// (measured: 123 vs 119 FMA/clk/SM, tools/ubench/ffma2.cu); what is saved is issue slots,
// which is what the rollout -- and above all the fused sample+rollout kernel, whose Philox
// integer and MUFU work competes for the same issue ports -- is bound by.
// ---------------------------------------------------------------------------------
struct f2 { unsigned long long r; };
__device__ __forceinline__ f2 mk2(float lo, float hi)
{
    f2 o; asm("mov.b64 %0, {%1,%2};" : "=l"(o.r) : "f"(lo), "f"(hi)); return o;
}
__device__ __forceinline__ void un2(f2 a, float &lo, float &hi)
{
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.r));
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 o; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r)); return o;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 o; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r)); return o;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 o; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r)); return o;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 o; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o.r) : "l"(a.r), "l"(b.r), "l"(c.r)); return o;
}
// A product whose result feeds an ADD.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even though both carry an explicit .rn (it honours .rn for the scalar forms only; seen
// in the SASS, and in the parity tests as 1-2 ulp cost differences).  fma(a, b, +0) rounds
// exactly like the product and cannot be merged with the following add; it costs the same
// issue slot.  Only the sign of a zero product can differ, which no cost term can see.
__device__ __forceinline__ f2 mulp2(f2 a, f2 b)
{
    f2 z; z.r = 0ull;
    return fma2(a, b, z);
}

// ---------------------------------------------------------------------------------
// dynamics functors
// ---------------------------------------------------------------------------------
struct DoubleIntegrator {
    float dt, b0, b1;
    f2 dt2, b02, b12;

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        dt = p->g[1]; b0 = p->b[0]; b1 = p->b[1];
    }
    __device__ __forceinline__ void load2(const ProblemDev *__restrict__ p)
    {
        dt2 = mk2(p->g[1], p->g[1]); b02 = mk2(p->b[0], p->b[0]); b12 = mk2(p->b[1], p->b[1]);
    }
    template <bool STRICT>
    __device__ __forceinline__ void axis(float &p, float &v, float ue) const
    {
        if (STRICT) {
            p = __fadd_rn(__fadd_rn(p, __fmul_rn(dt, v)), __fmul_rn(b0, ue));
            v = __fadd_rn(v, __fmul_rn(b1, ue));
        } else {
            p = __fmaf_rn(b0, ue, __fadd_rn(p, __fmul_rn(dt, v)));
            v = __fmaf_rn(b1, ue, v);
        }
    }
    template <bool STRICT>
    __device__ __forceinline__ void axis2(f2 &p, f2 &v, f2 ue) const
    {
        if (STRICT) {
            p = add2(add2(p, mulp2(dt2, v)), mulp2(b02, ue));
            v = add2(v, mulp2(b12, ue));
        } else {
            p = fma2(b02, ue, add2(p, mulp2(dt2, v)));
            v = fma2(b12, ue, v);
        }
    }
};

struct LinearAxis {
    float g0, g1, g2, g3, b0, b1;
    f2 g02, g12, g22, g32, b02, b12;

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        g0 = p->g[0]; g1 = p->g[1]; g2 = p->g[2]; g3 = p->g[3]; b0 = p->b[0]; b1 = p->b[1];
    }
    __device__ __forceinline__ void load2(const ProblemDev *__restrict__ p)
    {
        g02 = mk2(p->g[0], p->g[0]); g12 = mk2(p->g[1], p->g[1]);
        g22 = mk2(p->g[2], p->g[2]); g32 = mk2(p->g[3], p->g[3]);
        b02 = mk2(p->b[0], p->b[0]); b12 = mk2(p->b[1], p->b[1]);
    }
    template <bool STRICT>
    __device__ __forceinline__ void axis(float &p, float &v, float ue) const
    {
        const float po = p, vo = v;          // both rows read the OLD state
        if (STRICT) {
            p = __fadd_rn(__fadd_rn(__fmul_rn(g0, po), __fmul_rn(g1, vo)), __fmul_rn(b0, ue));
            v = __fadd_rn(__fadd_rn(__fmul_rn(g2, po), __fmul_rn(g3, vo)), __fmul_rn(b1, ue));
        } else {
            p = __fmaf_rn(b0, ue, __fmaf_rn(g0, po, __fmul_rn(g1, vo)));
            v = __fmaf_rn(b1, ue, __fmaf_rn(g2, po, __fmul_rn(g3, vo)));
        }
    }
    template <bool STRICT>
    __device__ __forceinline__ void axis2(f2 &p, f2 &v, f2 ue) const
    {
        const f2 po = p, vo = v;
        if (STRICT) {
            p = add2(add2(mulp2(g02, po), mulp2(g12, vo)), mulp2(b02, ue));
            v = add2(add2(mulp2(g22, po), mulp2(g32, vo)), mulp2(b12, ue));
        } else {
            p = fma2(b02, ue, fma2(g02, po, mulp2(g12, vo)));
            v = fma2(b12, ue, fma2(g22, po, mulp2(g32, vo)));
        }
    }
};

// ---------------------------------------------------------------------------------
// cost functor: lambda * sum_a u_a*inv_s_a*e_a  +  sum_i (x_i-g_i)*w_i*(x_i-g_i)
// ---------------------------------------------------------------------------------
template <int A>
struct QuadraticCost {
    float lambda, goal[2 * A], w[2 * A];
    f2 lambda2, goal2[2 * A], w2[2 * A];

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        lambda = p->lambda;
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) { goal[i] = p->goal[i]; w[i] = p->w[i]; }
    }
    __device__ __forceinline__ void load2(const ProblemDev *__restrict__ p)
    {
        lambda2 = mk2(p->lambda, p->lambda);
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            goal2[i] = mk2(p->goal[i], p->goal[i]);
            w2[i] = mk2(p->w[i], p->w[i]);
        }
    }
    // state cost added onto res in index order (Cost::final_cost; second half of step_cost)
    template <bool STRICT>
    __device__ __forceinline__ float state(const float (&x)[2 * A], float res) const
    {
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const float d = __fsub_rn(x[i], goal[i]);
            if (STRICT) res = __fadd_rn(res, __fmul_rn(__fmul_rn(d, w[i]), d));
            else        res = __fmaf_rn(__fmul_rn(d, w[i]), d, res);
        }
        return res;
    }
    template <bool STRICT>
    __device__ __forceinline__ f2 state2(const f2 (&x)[2 * A], f2 res) const
    {
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const f2 d = sub2(x[i], goal2[i]);
            if (STRICT) res = add2(res, mulp2(mul2(d, w2[i]), d));
            else        res = fma2(mul2(d, w2[i]), d, res);
        }
        return res;
    }
    // terminal cost, Cost::final_cost (src/cost.cu:57-64), with the weights of the Cost object the
    // final state is charged by: wf -- the stage weights again in the reference (one object,
    // src/point_mass_gpu.cu:116), a second object's after mppi_set_terminal_weights.  Read from
    // memory here, once per rollout, so that they cost the T loop no registers.
    template <bool STRICT>
    __device__ __forceinline__ float terminal(const float (&x)[2 * A], const float *__restrict__ wf) const
    {
        float res = 0.0f;
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const float d = __fsub_rn(x[i], goal[i]);
            const float wi = wf[i];
            if (STRICT) res = __fadd_rn(res, __fmul_rn(__fmul_rn(d, wi), d));
            else        res = __fmaf_rn(__fmul_rn(d, wi), d, res);
        }
        return res;
    }
    template <bool STRICT>
    __device__ __forceinline__ f2 terminal2(const f2 (&x)[2 * A], const float *__restrict__ wf) const
    {
        f2 res = mk2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const f2 d = sub2(x[i], goal2[i]);
            const f2 wi = mk2(wf[i], wf[i]);
            if (STRICT) res = add2(res, mulp2(mul2(d, wi), d));
            else        res = fma2(mul2(d, wi), d, res);
        }
        return res;
    }
    // one term of the control cost, accumulated in action order
    template <bool STRICT>
    __device__ __forceinline__ float control(float res, float ui, float e) const
    {
        return STRICT ? __fadd_rn(res, __fmul_rn(ui, e)) : __fmaf_rn(ui, e, res);
    }
    template <bool STRICT>
    __device__ __forceinline__ f2 control2(f2 res, f2 ui, f2 e) const
    {
        return STRICT ? add2(res, mulp2(ui, e)) : fma2(ui, e, res);
    }
    // stage cost from the finished control sum and the NEW state (Cost::step_cost)
    template <bool STRICT>
    __device__ __forceinline__ float stage(float ctrl, const float (&x)[2 * A]) const
    {
        return state<STRICT>(x, __fmul_rn(ctrl, lambda));
    }
    template <bool STRICT>
    __device__ __forceinline__ f2 stage2(f2 ctrl, const f2 (&x)[2 * A]) const
    {
        // the product feeds an add when STRICT (see mulp2)
        return state2<STRICT>(x, STRICT ? mulp2(ctrl, lambda2) : mul2(ctrl, lambda2));
    }
};

// ---------------------------------------------------------------------------------
// model tag: what a rollout kernel is instantiated on
// ---------------------------------------------------------------------------------
template <bool STRICT_, class DYN_>
struct Model {
    static constexpr bool kStrict = STRICT_;
    using Dyn = DYN_;
};

// one sample: x <- f(x, u+e);  c += step_cost(x_new, u, e)
template <int A, class MODEL>
struct PointMass {
    static constexpr bool STRICT = MODEL::kStrict;
    typename MODEL::Dyn dyn;
    QuadraticCost<A> cost;

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        dyn.load(p);
        cost.load(p);
    }
    __device__ __forceinline__ float state_cost(const float (&x)[2 * A], float res) const
    {
        return cost.template state<STRICT>(x, res);
    }
    __device__ __forceinline__ float terminal_cost(const float (&x)[2 * A],
                                                   const ProblemDev *__restrict__ p) const
    {
        return cost.template terminal<STRICT>(x, p->wf);
    }
    __device__ __forceinline__ void step(float (&x)[2 * A], float &c, const float (&u)[A],
                                         const float (&ui)[A], const float (&e)[A]) const
    {
        float res = 0.0f;
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const float ue = __fadd_rn(u[i], e[i]);
            dyn.template axis<STRICT>(x[i], x[i + A], ue);
            res = cost.template control<STRICT>(res, ui[i], e[i]);
        }
        c = __fadd_rn(c, cost.template stage<STRICT>(res, x));
    }
};

// a PAIR of samples: same operations, same order, two lanes
template <int A, class MODEL>
struct PointMass2 {
    static constexpr bool STRICT = MODEL::kStrict;
    typename MODEL::Dyn dyn;
    QuadraticCost<A> cost;

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        dyn.load2(p);
        cost.load2(p);
    }
    __device__ __forceinline__ f2 state_cost(const f2 (&x)[2 * A], f2 res) const
    {
        return cost.template state2<STRICT>(x, res);
    }
    __device__ __forceinline__ f2 terminal_cost(const f2 (&x)[2 * A],
                                                const ProblemDev *__restrict__ p) const
    {
        return cost.template terminal2<STRICT>(x, p->wf);
    }
    __device__ __forceinline__ void step(f2 (&x)[2 * A], f2 &c, const f2 (&u)[A], const f2 (&ui)[A],
                                         const f2 (&e)[A]) const
    {
        f2 res = mk2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const f2 ue = add2(u[i], e[i]);
            dyn.template axis2<STRICT>(x[i], x[i + A], ue);
            res = cost.template control2<STRICT>(res, ui[i], e[i]);
        }
        c = add2(c, cost.template stage2<STRICT>(res, x));
    }
};

// vector load of SPT consecutive floats through the non-coherent, no-L1-allocate path
template <int SPT> struct EpsVec;
template <> struct EpsVec<1> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[1])
    {
        asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
    }
};
template <> struct EpsVec<2> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[2])
    {
        asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
};
template <> struct EpsVec<4> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[4])
    {
        asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
            : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    }
};

// U staging in shared memory: per time step one 16-byte aligned record
// {u_0, u_0*inv_s_0, u_1, u_1*inv_s_1, ...} so a thread fetches a whole step with one or two
// LDS.128 (broadcast) instead of 2A scalar loads.
// Packed variant for the FP32x2 path: {u_a, u_a, u_a*inv_s_a, u_a*inv_s_a} per action dim, so
// one LDS.128 yields the two lane-duplicated operands directly.
template <int A> struct UStage2 {
    static constexpr int kStride = 4 * A;                    // floats per step
    static __device__ __forceinline__ void fetch(const float *s, int t, f2 (&u)[A], f2 (&ui)[A])
    {
        const float4 *p = reinterpret_cast<const float4 *>(s + (size_t)t * kStride);
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float4 v = p[a];
            u[a] = mk2(v.x, v.y);
            ui[a] = mk2(v.z, v.w);
        }
    }
};

template <int A> struct UStage {
    static constexpr int kStride = (2 * A + 3) / 4 * 4;     // floats per step
    static __device__ __forceinline__ void fetch(const float *s, int t, float (&u)[A], float (&ui)[A])
    {
        const float4 *p = reinterpret_cast<const float4 *>(s + (size_t)t * kStride);
        const float4 v0 = p[0];
        u[0] = v0.x; ui[0] = v0.y;
        if (A >= 2) { u[1 % A] = v0.z; ui[1 % A] = v0.w; }
        if (A >= 3) {
            const float4 v1 = p[1];
            u[2 % A] = v1.x; ui[2 % A] = v1.y;
            if (A >= 4) { u[3 % A] = v1.z; ui[3 % A] = v1.w; }
        }
    }
};

}  // namespace mppi
