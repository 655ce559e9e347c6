// csrc/fieldhead.cu compiled for the host (tests/emu/cuda_emu.h) -- TEST INFRASTRUCTURE ONLY.
// Entry points mirror the device ones, with the grid given explicitly so that tests can force many tiles per CTA.
#define BH_HOST_EMULATION 1
#include "../../bihome_b200/csrc/fieldhead.cu"

extern "C" {
void emu_moments(const float* x, double* partials, long long n_pix, int grid) {
    emu_launch(grid, bh::kMomThreads, [=] { bh::moments_kernel<16>(x, partials, n_pix); });
}
void emu_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* out, int B, int HW,
             int grid) {
    const long long n = static_cast<long long>(B) * HW;
    emu_launch(grid, 256, [=] { bh::fieldhead_fwd_kernel<16, 128>(x, W1, b1, W2, b2, out, n, HW); });
}
void emu_bwd(const float* x, const float* W1, const float* b1, const float* W2, const float* gOut, float* gx, float* partials,
             int B, int HW, int grid) {
    const long long n = static_cast<long long>(B) * HW;
    emu_launch(grid, bh::kFhThreads, [=] { bh::fieldhead_bwd_kernel<16, 128>(x, W1, b1, W2, gOut, gx, partials, n, HW); });
}
void emu_affine(const float* x, const float* a, const float* M, float* gx, long long n_pix, int accumulate, int grid) {
    emu_launch(grid, 256, [=] { bh::affine_acc_kernel<16>(x, a, M, gx, n_pix, accumulate); });
}
}

#ifdef BH_EMU_MAIN
// race hunt: run every kernel on a small problem under -fsanitize=thread; numbers are checked by the .so build
#include <cstdio>
#include <cstdlib>
int main() {
    const int B = 2, HW = 75, grid = 2;      // 150 pixels: 5 backward tiles (one partial, one straddling the samples)
    const long long n = static_cast<long long>(B) * HW;
    std::vector<float> x(n * 16), W1(128 * 16), b1(128), W2(256), b2(2), out(n * 2), g(n * 2), gx(n * 16), a(16), M(256);
    std::vector<float> parts(grid * (128 * 16 + 3 * 128 + 2));
    std::vector<double> mom(grid * (16 + 256));
    auto fill = [](std::vector<float>& v) { for (auto& e : v) e = static_cast<float>(rand()) / RAND_MAX - 0.4f; };
    fill(x); fill(W1); fill(b1); fill(W2); fill(b2); fill(g); fill(a); fill(M);
    emu_moments(x.data(), mom.data(), n, grid);
    emu_fwd(x.data(), W1.data(), b1.data(), W2.data(), b2.data(), out.data(), B, HW, grid);
    emu_bwd(x.data(), W1.data(), b1.data(), W2.data(), g.data(), gx.data(), parts.data(), B, HW, grid);
    emu_affine(x.data(), a.data(), M.data(), gx.data(), n, 1, grid);
    std::printf("ok %f %f %f\n", out[0], gx[0], mom[0]);
    return 0;
}
#endif
