// Stand-alone check and timing of the Cholesky-QR phase of the packed Benettin kernel (pack::chol_factor / chol_solve):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I qgs_b200/csrc scripts/cholqr_harness.cu -o build/cholqr_harness
// Every block holds 7 matrices of 36 x m as the Benettin kernel does; REP factorisations per launch.
#include <cstdio>
#include <random>
#include <vector>

#include "tgls_pack.cuh"
using namespace qgsb;
using namespace qgsb::pack;

constexpr int NN = 36, GM = 7;

template <int N>
__global__ void __launch_bounds__(256, 1) k_chol(double *buf, int m, int stride, int rep, int *flag)
{
    extern __shared__ __align__(16) double smem_pack[];
    const int t = threadIdx.x, g = t / m, c = t - g * m;
    long long tl = 0, tf = 0, ts = 0;
    for (int it = 0; it < rep; ++it) {
        long long c0 = clock64();
        for (int i = threadIdx.x; i < GM * stride; i += blockDim.x) smem_pack[i] = buf[(size_t)blockIdx.x * GM * stride + i];
        __syncthreads();
        long long c1 = clock64();
        long long c2 = c1;
        const bool done = !cholqr_dispatch<N>(smem_pack, stride, 0, m, GM, 1L << 40);
        if (!done && t == 0) *flag = 1;
        long long c3 = clock64();
        tl += c1 - c0; tf += c2 - c1; ts += c3 - c2;
    }
#ifdef CHOL_PROF
    if (blockIdx.x == 0 && t == 0 && rep > 1) printf("factor: gram %lld  load columns %lld  cholesky %lld\n", chol_prof[0] / rep, chol_prof[1] / rep, chol_prof[2] / rep), chol_prof[0] = chol_prof[1] = chol_prof[2] = 0;
#endif
    if (blockIdx.x == 0 && t == 0 && rep > 1) printf("cycles per iteration: load %lld  factor %lld  solve %lld\n", tl / rep, tf / rep, ts / rep);
    for (int i = threadIdx.x; i < GM * stride; i += blockDim.x) buf[(size_t)blockIdx.x * GM * stride + i] = smem_pack[i];
}

int main(int argc, char **argv)
{
    const int m = argc > 1 ? atoi(argv[1]) : 36, blocks = 148, rep = 200;
    const Carve<NN> cv(0, m);
    const int stride = cv.total();
    std::vector<double> h((size_t)blocks * GM * stride, 0.), h0;
    std::mt19937_64 rng(1);
    std::normal_distribution<double> nd;
    for (int b = 0; b < blocks * GM; ++b) {
        double *fm = h.data() + (size_t)b * stride + cv.o_fm();
        for (int i = 0; i < NN; ++i)
            for (int c = 0; c < m; ++c) fm[i * m + c] = (i == c ? 1. : 0.) + 0.1 * nd(rng);
    }
    h0 = h;
    double *d;
    int *flag;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&flag, 4);
    cudaMemset(flag, 0, 4);
    const size_t smem = (size_t)GM * stride * 8;
    cudaFuncSetAttribute(k_chol<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    k_chol<NN><<<blocks, 256, smem>>>(d, m, stride, 1, flag);
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    int hf = 0;
    cudaMemcpy(&hf, flag, 4, cudaMemcpyDeviceToHost);
    printf("launch: %s, refused pivots: %d\n", cudaGetErrorString(cudaGetLastError()), hf);
    double eo = 0., er = 0.;
    const int ld = m == NN ? NN : m <= 16 ? 16 : m <= 24 ? 24 : 32;      // column capacity of the instantiation
    for (int b = 0; b < blocks * GM; ++b) {
        const double *Q = h.data() + (size_t)b * stride + cv.o_fm(), *R = h.data() + (size_t)b * stride + cv.o_facc();
        const double *A = h0.data() + (size_t)b * stride + cv.o_fm();
        for (int a = 0; a < m; ++a)
            for (int c = 0; c < m; ++c) {
                double s = 0.;
                for (int i = 0; i < NN; ++i) s += Q[i * m + a] * Q[i * m + c];
                eo = fmax(eo, fabs(s - (a == c)));
            }
        for (int i = 0; i < NN; ++i)
            for (int c = 0; c < m; ++c) {
                double s = 0.;
                for (int k = 0; k <= c; ++k) s += Q[i * m + k] * R[k * ld + c];
                er = fmax(er, fabs(s - A[i * m + c]));
            }
    }
    printf("m=%d  max|Q^T Q - I| = %.2e   max|Q R - A| = %.2e\n", m, eo, er);
    cudaMemcpy(d, h0.data(), h0.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_chol<NN><<<blocks, 256, smem>>>(d, m, stride, rep, flag);
    cudaEventRecord(e0);
    k_chol<NN><<<blocks, 256, smem>>>(d, m, stride, rep, flag);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%.3f us per block factorisation (7 members, incl. reloading 7 x %d doubles) = %.0f cycles at 1.965 GHz\n",
           ms * 1e3 / rep, stride, ms * 1e3 / rep * 1965.);
    return 0;
}
