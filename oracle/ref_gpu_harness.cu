/*
 * ref_gpu_harness.cu -- runs the REFERENCE's own GPU path (class PointMassModel of
 * src/point_mass.cu with point_mass_gpu.cu, cost.cu, mppi_utils.cu, recompiled for sm_100)
 * as a stand-alone process.  TEST INFRASTRUCTURE ONLY; built by `make -C oracle ref-gpu`
 * from the reference sources where they lie, output oracle/_ref/ref_gpu_run.
 *
 *   ref_gpu_run <in.bin> <out.bin>
 * in : int32 K,T,A,nsteps ; float dt, x0[S], U[T*A], goal[S], w[S]
 * out: float u_pre[T*A] (get_u before the last get_act), e[K*T*A], cost[K], beta, nabla,
 *      weight[K], u_post[T*A], next_act[A], ms[nsteps] (wall clock per get_act, as
 *      src/main.cu:329-332 times it)
 *
 * The loop is the reference driver's (src/main.cu:311-371): new PointMassModel,
 * memcpy_set_data, then per step get_u / get_act / set_x.  The device malloc heap is raised
 * first: the reference's init() allocates ~4*T*A+... bytes per sample with device-side
 * malloc and never sets the limit.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "point_mass.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror("open input"); return 2; }
    int hdr[4];
    float dt;
    if (fread(hdr, sizeof(int), 4, fi) != 4 || fread(&dt, sizeof(float), 1, fi) != 1) return 2;
    const int K = hdr[0], T = hdr[1], A = hdr[2], nsteps = hdr[3], S = 2 * A;
    std::vector<float> x0(S), U((size_t)T * A), goal(S), w(S);
    if (fread(x0.data(), 4, S, fi) != (size_t)S || fread(U.data(), 4, (size_t)T * A, fi) != (size_t)T * A ||
        fread(goal.data(), 4, S, fi) != (size_t)S || fread(w.data(), 4, S, fi) != (size_t)S) return 2;
    fclose(fi);

    const size_t heap = (size_t)K * ((size_t)T * A * 4 + 8 * S + 4 * A + 256) * 2 + (64u << 20);
    if (cudaDeviceSetLimit(cudaLimitMallocHeapSize, heap) != cudaSuccess) {
        fprintf(stderr, "cannot raise the device malloc heap to %zu bytes\n", heap);
        return 3;
    }
    PointMassModel *model = new PointMassModel(K, T, dt, S, A, false);
    model->memcpy_set_data(x0.data(), U.data(), goal.data(), w.data());

    std::vector<float> u_pre((size_t)T * A), next_act(A), ms(nsteps);
    for (int s = 0; s < nsteps; ++s) {
        model->get_u(u_pre.data());
        auto t1 = std::chrono::steady_clock::now();
        model->get_act(next_act.data());
        auto t2 = std::chrono::steady_clock::now();
        ms[s] = std::chrono::duration<float, std::milli>(t2 - t1).count();
        if (s + 1 < nsteps) model->set_x(x0.data());
    }
    std::vector<float> x((size_t)K * (T + 1) * S), u_post((size_t)T * A), e((size_t)K * T * A),
        cost(K), weight(K);
    float beta = 0, nabla = 0;
    model->get_inf(x.data(), u_post.data(), e.data(), cost.data(), &beta, &nabla, weight.data());
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 4; }

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) { perror("open output"); return 2; }
    fwrite(u_pre.data(), 4, u_pre.size(), fo);
    fwrite(e.data(), 4, e.size(), fo);
    fwrite(cost.data(), 4, cost.size(), fo);
    fwrite(&beta, 4, 1, fo);
    fwrite(&nabla, 4, 1, fo);
    fwrite(weight.data(), 4, weight.size(), fo);
    fwrite(u_post.data(), 4, u_post.size(), fo);
    fwrite(next_act.data(), 4, next_act.size(), fo);
    fwrite(ms.data(), 4, ms.size(), fo);
    fclose(fo);
    delete model;
    return 0;
}
