// Host-side weight calibration for AADFF_MODE_ECON (no reference counterpart: the reference computes in fp32).
//
// In econ mode the late layers drop the Ah*Wl term, i.e. their weights are fp16 values.  Plain rounding of W
// leaves a max-abs image error of 6e-5 .. 1.2e-4 -- at the 1e-4 bar.  The weights are fixed and the network's
// input domain is a known box (x, y in [-1,1], z(depth), z(focus) in [0,1]; deeplens/psfnet.py:427-441), so
// the rounding can be chosen to minimise the *output* error  E |a^T (w - q)|^2  over the activations a that
// actually occur:  q is picked column by column and the rounding error of column k is pushed onto the columns
// not yet rounded through the inverse activation covariance (the OBQ/GPTQ recurrence).  Measured: 2.5-4x
// smaller image error than plain rounding (profiles/NOTES_r01.md), at zero run-time cost.
//
//   probes   : NC Halton points of the input box (bases 2,3,5,7), every 8th snapped to a face of the box
//   forward  : fp32 Linear+ReLU chain on the host (threads over probes)
//   per layer: H = A^T A / NC + damp * mean(diag) * I  (double),  U = chol(H^-1)^T (upper),
//              for k: q_k = fp16(w_k); e = (w_k - q_k) / U_kk; w_j -= e * U_kj (j > k)
#pragma once
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

namespace aadff {

constexpr int TC_ECON_FIRST_LAYER = 5;          // econ: layers L5.. (and the head) run on calibrated fp16 weights
constexpr int ECON_CALIB_PROBES = 2048;
constexpr double ECON_CALIB_DAMP = 1e-6;

inline double halton(unsigned i, unsigned base) {
    double f = 1.0, r = 0.0;
    while (i > 0) {
        f /= base;
        r += f * (i % base);
        i /= base;
    }
    return r;
}

template <typename F>
inline void parallel_for(int n, F fn) {
    const int nt = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 8, n}));
    if (nt == 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([=]() { fn((int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt)); });
    for (auto& t : th) t.join();
}

// lower Cholesky factor in place (row-major n x n, upper part ignored); false if not positive definite
inline bool chol_lower(std::vector<double>& a, int n) {
    for (int j = 0; j < n; ++j) {
        double d = a[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        a[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = a[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
            a[(size_t)i * n + j] = s / d;
        }
    }
    return true;
}

// out[n][k] (fp16-representable floats) for W [N][K] given the layer's input activations A [NC][K]
inline bool gptq_round_fp16(const float* W, int N, int K, const std::vector<float>& A, int NC, std::vector<float>& out) {
    std::vector<double> H((size_t)K * K, 0.0);
    parallel_for(K, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i)
            for (int p = 0; p < NC; ++p) {
                const double ai = A[(size_t)p * K + i];
                if (ai == 0.0) continue;
                const float* ap = &A[(size_t)p * K];
                double* hi = &H[(size_t)i * K];
                for (int j = 0; j <= i; ++j) hi[j] += ai * ap[j];
            }
    });
    double tr = 0.0;
    for (int i = 0; i < K; ++i) tr += H[(size_t)i * K + i] / NC;
    for (int i = 0; i < K; ++i) {
        for (int j = 0; j <= i; ++j) H[(size_t)i * K + j] /= NC;
        H[(size_t)i * K + i] += ECON_CALIB_DAMP * tr / K;
    }
    if (!chol_lower(H, K)) return false;                        // H = L L^T
    // Li = L^-1 (lower), Hinv = Li^T Li
    std::vector<double> Li((size_t)K * K, 0.0);
    for (int c = 0; c < K; ++c) {
        Li[(size_t)c * K + c] = 1.0 / H[(size_t)c * K + c];
        for (int i = c + 1; i < K; ++i) {
            double s = 0.0;
            for (int k = c; k < i; ++k) s -= H[(size_t)i * K + k] * Li[(size_t)k * K + c];
            Li[(size_t)i * K + c] = s / H[(size_t)i * K + i];
        }
    }
    std::vector<double> Hinv((size_t)K * K, 0.0);
    parallel_for(K, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = 0.0;
                for (int k = i; k < K; ++k) s += Li[(size_t)k * K + i] * Li[(size_t)k * K + j];
                Hinv[(size_t)i * K + j] = s;
            }
    });
    if (!chol_lower(Hinv, K)) return false;                     // Hinv = C C^T, U = C^T: U[k][j] = C[j][k]
    out.assign((size_t)N * K, 0.f);
    parallel_for(N, [&](int n0, int n1) {
        std::vector<double> w(K);
        for (int n = n0; n < n1; ++n) {
            for (int k = 0; k < K; ++k) w[k] = W[(size_t)n * K + k];
            for (int k = 0; k < K; ++k) {
                const float q = __half2float(__float2half_rn((float)w[k]));
                out[(size_t)n * K + k] = q;
                const double e = (w[k] - (double)q) / Hinv[(size_t)k * K + k];
                if (e != 0.0)
                    for (int j = k + 1; j < K; ++j) w[j] -= e * Hinv[(size_t)j * K + k];
            }
        }
    });
    return true;
}

// wq[l] = calibrated fp16-valued weights of layer l for l >= first_layer (empty where calibration failed)
inline void calibrate_econ(const float* const* W, const float* const* b, const int* dims, int n_layers, int first_layer,
                           std::vector<std::vector<float>>& wq) {
    const int NC = ECON_CALIB_PROBES;
    wq.assign(n_layers, {});
    std::vector<float> A((size_t)NC * 4);
    for (int p = 0; p < NC; ++p) {
        double c[4] = {2.0 * halton(p + 1, 2) - 1.0, 2.0 * halton(p + 1, 3) - 1.0, halton(p + 1, 5), halton(p + 1, 7)};
        if (p % 8 == 0) {                                        // a face of the box: the image border / depth clamps
            const int d = (p / 8) % 4;
            const bool up = ((p / 32) & 1) != 0;
            c[d] = (d < 2) ? (up ? 1.0 : -1.0) : (up ? 1.0 : 0.0);
        }
        for (int d = 0; d < 4; ++d) A[(size_t)p * 4 + d] = (float)c[d];
    }
    for (int l = 0; l < n_layers; ++l) {
        const int K = dims[l], N = dims[l + 1];
        if (l >= first_layer) {
            std::vector<float> q;
            if (gptq_round_fp16(W[l], N, K, A, NC, q)) wq[l] = std::move(q);
        }
        if (l == n_layers - 1) break;
        std::vector<float> Y((size_t)NC * N);
        parallel_for(NC, [&](int p0, int p1) {
            for (int p = p0; p < p1; ++p)
                for (int n = 0; n < N; ++n) {
                    const float* a = &A[(size_t)p * K];
                    const float* w = &W[l][(size_t)n * K];
                    float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // eight lanes: lets the compiler vectorise
                    int k = 0;
                    for (; k + 8 <= K; k += 8)
                        for (int u = 0; u < 8; ++u) part[u] += a[k + u] * w[k + u];
                    float s = 0.f;
                    for (; k < K; ++k) s += a[k] * w[k];
                    for (int u = 0; u < 8; ++u) s += part[u];
                    s += b[l][n];
                    Y[(size_t)p * N + n] = s > 0.f ? s : 0.f;
                }
        });
        A.swap(Y);
    }
}

}  // namespace aadff
