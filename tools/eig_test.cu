// standalone GPU check of eig_kernel<6> on one known Hermitian matrix (debug tool)
#include "../misonet_b200/csrc/mvdr.cu"
#include "../misonet_b200/csrc/runtime.cu"
#include <vector>
#include <complex>
int main() {
    constexpr int M = 6, NV = 84;
    // A = X X^H / T with a fixed LCG
    const int T = 20;
    std::vector<std::complex<double>> X(M * T);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0 - 0.5; };
    for (auto &x : X) x = {rnd(), rnd()};
    std::vector<float> partial(NV, 0.f);
    int k = 0;
    for (int i = 0; i < M; ++i)
        for (int j = i; j < M; ++j) {
            std::complex<double> a = 0;
            for (int t = 0; t < T; ++t) a += X[i * T + t] * std::conj(X[j * T + t]);
            partial[4 * k] = (float)a.real(); partial[4 * k + 1] = (float)a.imag();
            partial[4 * k + 2] = (float)a.real(); partial[4 * k + 3] = (float)a.imag();
            printf("A %d %d %.9g %.9g\n", i, j, partial[4 * k] / (double)T, partial[4 * k + 1] / (double)T);
            ++k;
        }
    float *dp; double2 *ds;
    cudaMalloc(&dp, NV * 4); cudaMalloc(&ds, M * 16);
    cudaMemcpy(dp, partial.data(), NV * 4, cudaMemcpyHostToDevice);
    miso::eig_kernel<6><<<1, 64>>>(dp, ds, 1, 1, T, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    double2 h[M];
    cudaMemcpy(h, ds, M * 16, cudaMemcpyDeviceToHost);
    for (int m = 0; m < M; ++m) printf("d %d %.12g %.12g\n", m, h[m].x, h[m].y);
    return 0;
}
