// model.cuh -- the point-mass dynamics and cost of the rollout (reference
// PointMassModelGpu::step / Cost::step_cost / final_cost, src/point_mass_gpu.cu:82-121,
// src/cost.cu:42-64) as device structs shared by every rollout kernel: scalar (PointMass),
// packed FP32x2 on a pair of samples (PointMass2), and the shared-memory staging of U.
#pragma once

#include "common.cuh"

namespace mppi {

template <int A, bool STRICT>
struct PointMass {
    float dt, b0, b1, lambda;
    float goal[2 * A], w[2 * A];

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        dt = p->g[1]; b0 = p->b[0]; b1 = p->b[1]; lambda = p->lambda;
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) { goal[i] = p->goal[i]; w[i] = p->w[i]; }
    }

    // state cost  sum_i (x_i-g_i)*w_i*(x_i-g_i)  added onto res in index order
    __device__ __forceinline__ float state_cost(const float (&x)[2 * A], float res) const
    {
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const float d = __fsub_rn(x[i], goal[i]);
            if (STRICT) res = __fadd_rn(res, __fmul_rn(__fmul_rn(d, w[i]), d));
            else        res = __fmaf_rn(__fmul_rn(d, w[i]), d, res);
        }
        return res;
    }

    // one step: x <- f(x, u+e);  c += step_cost(x_new, u, e)
    __device__ __forceinline__ void step(float (&x)[2 * A], float &c, const float (&u)[A],
                                         const float (&ui)[A], const float (&e)[A]) const
    {
        float res = 0.0f;
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const float ue = __fadd_rn(u[i], e[i]);
            const float p = x[i], v = x[i + A];
            if (STRICT) {
                x[i]     = __fadd_rn(__fadd_rn(p, __fmul_rn(dt, v)), __fmul_rn(b0, ue));
                x[i + A] = __fadd_rn(v, __fmul_rn(b1, ue));
                res = __fadd_rn(res, __fmul_rn(ui[i], e[i]));
            } else {
                x[i]     = __fmaf_rn(b0, ue, __fadd_rn(p, __fmul_rn(dt, v)));
                x[i + A] = __fmaf_rn(b1, ue, v);
                res = __fmaf_rn(ui[i], e[i], res);
            }
        }
        res = __fmul_rn(res, lambda);
        res = state_cost(x, res);
        c = __fadd_rn(c, res);
    }
};

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

// PointMass on a PAIR of samples: same operations, same order, two lanes.
template <int A, bool STRICT>
struct PointMass2 {
    f2 dt, b0, b1, lambda;
    f2 goal[2 * A], w[2 * A];

    __device__ __forceinline__ void load(const ProblemDev *__restrict__ p)
    {
        dt = mk2(p->g[1], p->g[1]); b0 = mk2(p->b[0], p->b[0]); b1 = mk2(p->b[1], p->b[1]);
        lambda = mk2(p->lambda, p->lambda);
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            goal[i] = mk2(p->goal[i], p->goal[i]);
            w[i] = mk2(p->w[i], p->w[i]);
        }
    }
    __device__ __forceinline__ f2 state_cost(const f2 (&x)[2 * A], f2 res) const
    {
#pragma unroll
        for (int i = 0; i < 2 * A; ++i) {
            const f2 d = sub2(x[i], goal[i]);
            if (STRICT) res = add2(res, mulp2(mul2(d, w[i]), d));
            else        res = fma2(mul2(d, w[i]), d, res);
        }
        return res;
    }
    __device__ __forceinline__ void step(f2 (&x)[2 * A], f2 &c, const f2 (&u)[A], const f2 (&ui)[A],
                                         const f2 (&e)[A]) const
    {
        f2 res = mk2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const f2 ue = add2(u[i], e[i]);
            const f2 p = x[i], v = x[i + A];
            if (STRICT) {
                x[i]     = add2(add2(p, mulp2(dt, v)), mulp2(b0, ue));
                x[i + A] = add2(v, mulp2(b1, ue));
                res = add2(res, mulp2(ui[i], e[i]));
            } else {
                x[i]     = fma2(b0, ue, add2(p, mulp2(dt, v)));
                x[i + A] = fma2(b1, ue, v);
                res = fma2(ui[i], e[i], res);
            }
        }
        res = STRICT ? mulp2(res, lambda) : mul2(res, lambda);   // feeds an add when STRICT
        res = state_cost(x, res);
        c = add2(c, res);
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
