// Host-side unit test of exblas_b200/csrc/window.cuh (the register window of the superaccumulator).
// The header is plain IEEE double arithmetic + bit casts, so compiling it with g++ -ffp-contract=off
// runs exactly the arithmetic the CUDA kernels run.  One emulated thread walks a stream of products
// (or single summands) in groups of four with the kernels' policy: window path when every summand of
// the group is inside, otherwise the ordinary path + win_after_slow_group; drained every
// kWinFlushEvery summands and at the end.  The limbs must equal those of the ordinary path alone.
//
// Build + run: tests/test_window_host.py.   Prints "OK <cases>" or "FAIL ...".
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>
#include "../exblas_b200/csrc/window.cuh"

using namespace exb;

static unsigned hi_of(double x) { return (unsigned)(win_bits(x) >> 32); }

struct Gen {
    std::mt19937_64 rng;
    explicit Gen(uint64_t s) : rng(s) {}
    double value(int emin, int emax, bool signs) {
        std::uniform_real_distribution<double> m(1.0, 2.0);
        std::uniform_int_distribution<int> e(emin, emax);
        double v = std::ldexp(m(rng), e(rng));
        if (signs && (rng() & 1)) v = -v;
        return v;
    }
};

static bool same_limbs(long long* a, long long* b) {
    normalize(a);
    normalize(b);
    return std::memcmp(a, b, sizeof(long long) * kLimbs) == 0;
}

// products: returns true when the two accumulations agree
static bool run_products(const std::vector<double>& a, const std::vector<double>& x, double fail_prob, uint64_t seed,
                         long* fast_out) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    long long ref[kLimbs] = {0}, acc[kLimbs] = {0};
    Window w;
    win_reset(w);
    long fast = 0;
    auto emit = [&](double v) { accumulate_double(acc, v); };
    const size_t n = a.size() / 4 * 4;
    for (size_t i = 0; i < n; i += 4) {
        if (i % 1024 == 0) { normalize(ref); normalize(acc); }     // keep the plain limb arrays far from 2^63
        double p[4], e[4];
        unsigned hi[4];
        bool mine = true;
        for (int k = 0; k < 4; ++k) {
            p[k] = a[i + k] * x[i + k];
            e[k] = std::fma(a[i + k], x[i + k], -p[k]);
            accumulate_double(ref, p[k]);
            accumulate_double(ref, e[k]);
            hi[k] = hi_of(p[k]);
            mine = mine && win_holds(w, hi[k]);
        }
        const bool warp_ok = mine && !(u01(rng) < fail_prob);   // another lane of the warp may be outside
        if (warp_ok) {
            for (int k = 0; k < 4; ++k) win_add_product(w, p[k], e[k]);
            w.cnt += 4;
            fast += 4;
        } else {
            for (int k = 0; k < 4; ++k) {
                accumulate_double(acc, p[k]);
                accumulate_double(acc, e[k]);
            }
            win_after_slow_group<4>(w, mine, hi, false, emit);
        }
        if (w.cnt > (unsigned)kWinFlushEvery - 4u) {
            double out[4];
            win_drain(w, out);
            for (int k = 0; k < 4; ++k) emit(out[k]);
        }
    }
    double out[4];
    win_drain(w, out);
    for (int k = 0; k < 4; ++k) emit(out[k]);
    if (fast_out) *fast_out = fast;
    return same_limbs(ref, acc);
}

static bool run_singles(const std::vector<double>& a, double fail_prob, uint64_t seed, long* fast_out) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    long long ref[kLimbs] = {0}, acc[kLimbs] = {0};
    Window w;
    win_reset(w);
    long fast = 0;
    auto emit = [&](double v) { accumulate_double(acc, v); };
    const size_t n = a.size() / 4 * 4;
    for (size_t i = 0; i < n; i += 4) {
        if (i % 1024 == 0) { normalize(ref); normalize(acc); }
        unsigned hi[4];
        bool mine = true;
        for (int k = 0; k < 4; ++k) {
            accumulate_double(ref, a[i + k]);
            hi[k] = hi_of(a[i + k]);
            mine = mine && win_holds_sum(w, hi[k]);
        }
        const bool warp_ok = mine && !(u01(rng) < fail_prob);
        if (warp_ok) {
            for (int k = 0; k < 4; ++k) win_add_single(w, a[i + k]);
            w.cnt += 4;
            fast += 4;
        } else {
            for (int k = 0; k < 4; ++k) accumulate_double(acc, a[i + k]);
            win_after_slow_group<4>(w, mine, hi, true, emit);
        }
        if (w.cnt > (unsigned)kWinFlushEvery - 4u) {
            double out[4];
            win_drain_single(w, out);
            for (int k = 0; k < 4; ++k) emit(out[k]);
        }
    }
    double out[4];
    win_drain_single(w, out);
    for (int k = 0; k < 4; ++k) emit(out[k]);
    if (fast_out) *fast_out = fast;
    return same_limbs(ref, acc);
}

// W-digit product window (WindowP<W>): same emulation as run_products
template <int W>
static bool run_products_wide(const std::vector<double>& a, const std::vector<double>& x, double fail_prob, uint64_t seed,
                              long* fast_out) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    long long ref[kLimbs] = {0}, acc[kLimbs] = {0};
    WindowP<W> w;
    winp_reset(w);
    long fast = 0;
    auto emit = [&](double v) { accumulate_double(acc, v); };
    const size_t n = a.size() / 4 * 4;
    for (size_t i = 0; i < n; i += 4) {
        if (i % 1024 == 0) { normalize(ref); normalize(acc); }
        double p[4], e[4];
        unsigned hi[4];
        bool mine = true;
        for (int k = 0; k < 4; ++k) {
            p[k] = a[i + k] * x[i + k];
            e[k] = std::fma(a[i + k], x[i + k], -p[k]);
            accumulate_double(ref, p[k]);
            accumulate_double(ref, e[k]);
            hi[k] = hi_of(p[k]);
            mine = mine && winp_holds(w, hi[k]);
        }
        const bool warp_ok = mine && !(u01(rng) < fail_prob);
        if (warp_ok) {
            for (int k = 0; k < 4; ++k) winp_add_product(w, p[k], e[k]);
            w.cnt += 4;
            fast += 4;
        } else {
            for (int k = 0; k < 4; ++k) {
                accumulate_double(acc, p[k]);
                accumulate_double(acc, e[k]);
            }
            if (seed & 1) {
                winp_after_slow_group<W, 4>(w, mine, hi, emit);          // per-lane policy
            } else {                                                      // range-covering policy (one lane = the warp)
                int gmin = 4096, gmax = -4096;
                for (int k = 0; k < 4; ++k) {
                    const int E = (int)((hi[k] >> 20) & 0x7ffu);
                    if (E != 0 && E != 0x7ff) { gmin = std::min(gmin, E - 1023); gmax = std::max(gmax, E - 1023); }
                }
                winp_cover<W>(w, gmin, gmax, emit);
            }
        }
        if (w.cnt > (unsigned)kWinFlushEvery - 4u) {
            double out[W + 1];
            winp_drain(w, out);
            for (int k = 0; k <= W; ++k) emit(out[k]);
        }
    }
    double out[W + 1];
    winp_drain(w, out);
    for (int k = 0; k <= W; ++k) emit(out[k]);
    if (fast_out) *fast_out = fast;
    return same_limbs(ref, acc);
}

// W-digit single-summand window with the range-covering policy
template <int W>
static bool run_singles_wide(const std::vector<double>& a, double fail_prob, uint64_t seed, long* fast_out) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    long long ref[kLimbs] = {0}, acc[kLimbs] = {0};
    WindowP<W> w;
    winp_reset(w);
    long fast = 0;
    auto emit = [&](double v) { accumulate_double(acc, v); };
    const size_t n = a.size() / 4 * 4;
    for (size_t i = 0; i < n; i += 4) {
        if (i % 1024 == 0) { normalize(ref); normalize(acc); }
        unsigned hi[4];
        bool mine = true;
        for (int k = 0; k < 4; ++k) {
            accumulate_double(ref, a[i + k]);
            hi[k] = hi_of(a[i + k]);
            mine = mine && winp_holds(w, hi[k]);
        }
        const bool warp_ok = mine && !(u01(rng) < fail_prob);
        if (warp_ok) {
            for (int k = 0; k < 4; ++k) wins_add(w, a[i + k]);
            w.cnt += 4;
            fast += 4;
        } else {
            for (int k = 0; k < 4; ++k) accumulate_double(acc, a[i + k]);
            int gmin = 4096, gmax = -4096;
            for (int k = 0; k < 4; ++k) {
                const int E = (int)((hi[k] >> 20) & 0x7ffu);
                if (E != 0 && E != 0x7ff) { gmin = std::min(gmin, E - 1023); gmax = std::max(gmax, E - 1023); }
            }
            wins_cover<W>(w, gmin, gmax, emit);
        }
        if (w.cnt > (unsigned)kWinFlushEvery - 4u) {
            double out[W + 1];
            wins_drain(w, out);
            for (int k = 0; k <= W; ++k) emit(out[k]);
        }
    }
    double out[W + 1];
    wins_drain(w, out);
    for (int k = 0; k <= W; ++k) emit(out[k]);
    if (fast_out) *fast_out = fast;
    return same_limbs(ref, acc);
}

int main() {
    int cases = 0;
    long total_fast = 0, total_wide = 0, total_cover = 0, total_swide = 0;
    // (centre exponent, half width) of the factors; products then span about twice the width
    const int centres[] = {0, -300, 300, -430, 440, 37, -51};
    const int widths[] = {0, 1, 5, 12, 24, 25, 26, 40, 80, 300};
    for (int ci = 0; ci < 7; ++ci)
        for (int wi = 0; wi < 10; ++wi)
            for (int rep = 0; rep < 3; ++rep) {
                const int c = centres[ci], hw = widths[wi];
                Gen g(1000 * ci + 10 * wi + rep);
                const size_t n = rep == 0 ? 64 : (rep == 1 ? 1200 : 9000);
                std::vector<double> a(n), x(n);
                for (size_t i = 0; i < n; ++i) {
                    a[i] = g.value(c / 2 - hw / 2, c / 2 + hw / 2, true);
                    x[i] = g.value(c - c / 2 - hw / 2, c - c / 2 + hw / 2, rep != 1);
                }
                if (rep == 2) {   // zeros, an outlier far below and one far above, exact-product pairs
                    a[17] = 0.0;
                    x[40] = -0.0;
                    a[100] = std::ldexp(a[100], -90);
                    a[2000] = std::ldexp(a[2000], 70);
                    a[3000] = 3.0; x[3000] = std::ldexp(5.0, c);
                }
                long fast = 0;
                for (double fp : {0.0, 0.05}) {
                    long f3 = 0, f4 = 0, f5 = 0;
                    long g3 = 0, g4 = 0, g5 = 0;
                    if (!run_products_wide<3>(a, x, fp, 77, &f3) || !run_products_wide<4>(a, x, fp, 77, &f4) ||
                        !run_products_wide<5>(a, x, fp, 77, &f5) || !run_products_wide<3>(a, x, fp, 78, &g3) ||
                        !run_products_wide<4>(a, x, fp, 78, &g4) || !run_products_wide<5>(a, x, fp, 78, &g5)) {
                        printf("FAIL wide products centre=%d width=%d rep=%d fail_prob=%g\n", c, hw, rep, fp);
                        return 1;
                    }
                    if (fp == 0.0 && (f4 < f3 || f5 < f4)) {         // a wider window never admits fewer products
                        printf("FAIL wide products admit fewer: centre=%d width=%d rep=%d %ld %ld %ld\n", c, hw, rep, f3, f4, f5);
                        return 1;
                    }
                    total_wide += f5;
                    cases += 6;
                    total_cover += g5;
                    if (!run_products(a, x, fp, 77 + rep, &fast)) {
                        printf("FAIL products centre=%d width=%d rep=%d fail_prob=%g\n", c, hw, rep, fp);
                        return 1;
                    }
                    total_fast += fast;
                    ++cases;
                }
                // single summands over the same exponent ranges (twice the centre to move around the layout)
                std::vector<double> s(n);
                for (size_t i = 0; i < n; ++i) s[i] = g.value(2 * c - hw, 2 * c + hw, rep != 1);
                if (rep == 2) { s[5] = 0.0; s[900] = std::ldexp(s[900], -120); s[901] = std::ldexp(1.0, 2 * c + hw); }
                for (double fp : {0.0, 0.05}) {
                    long s2 = 0, s3 = 0, s4 = 0;
                    if (!run_singles_wide<2>(s, fp, 99 + rep, &s2) || !run_singles_wide<3>(s, fp, 99 + rep, &s3) ||
                        !run_singles_wide<4>(s, fp, 99 + rep, &s4)) {
                        printf("FAIL wide singles centre=%d width=%d rep=%d fail_prob=%g\n", 2 * c, hw, rep, fp);
                        return 1;
                    }
                    total_swide += s3;
                    cases += 3;
                    if (!run_singles(s, fp, 99 + rep, &fast)) {
                        printf("FAIL singles centre=%d width=%d rep=%d fail_prob=%g\n", 2 * c, hw, rep, fp);
                        return 1;
                    }
                    total_fast += fast;
                    ++cases;
                }
            }
    // worst-case digit growth: 1024 equal summands of the largest magnitude the window admits, both signs
    for (int sign = -1; sign <= 1; sign += 2) {
        std::vector<double> a(4096, sign * std::ldexp(1.9999999999999998, 24)), x(4096, std::ldexp(1.9999999999999998, 25));
        long fast = 0;
        if (!run_products(a, x, 0.0, 5, &fast) || fast < 4000) {
            printf("FAIL saturation sign=%d fast=%ld\n", sign, fast);
            return 1;
        }
        ++cases;
        std::vector<double> s(4096, sign * std::ldexp(1.9999999999999998, 25));
        if (!run_singles(s, 0.0, 5, &fast)) {
            printf("FAIL saturation singles sign=%d\n", sign);
            return 1;
        }
        ++cases;
    }
    // window edges: anchor with 8 summands at 2^10, then fill with summands at the top (b + 50) and the
    // bottom (b + 1) exponent the window admits, largest mantissas, one sign -- the largest digits possible
    for (int sign = -1; sign <= 1; sign += 2)
        for (int edge = 0; edge < 3; ++edge) {
            std::vector<double> a(8 + 4096), x(8 + 4096);
            for (size_t i = 0; i < a.size(); ++i) {
                const int ex = i < 8 ? 10 : (edge == 0 ? 35 : (edge == 1 ? -14 : ((i & 1) ? 35 : -14)));   // b = -15: window [-14, 35]
                a[i] = sign * std::ldexp(1.9999999999999998, ex - 1 - 3);
                x[i] = std::ldexp(1.9999999999999998, 3);
            }
            long fast = 0;
            if (!run_products(a, x, 0.0, 5, &fast) || fast < 4088) {
                printf("FAIL product edge=%d sign=%d fast=%ld\n", edge, sign, fast);
                return 1;
            }
            ++cases;
            std::vector<double> s(8 + 4096);
            for (size_t i = 0; i < s.size(); ++i) {
                const int ex = i < 8 ? 10 : (edge == 0 ? 35 : (edge == 1 ? -15 : ((i & 1) ? 35 : -15)));    // singles: [b, b + 50]
                s[i] = sign * std::ldexp(1.9999999999999998, ex);
            }
            if (!run_singles(s, 0.0, 5, &fast) || fast < 4088) {
                printf("FAIL single edge=%d sign=%d fast=%ld\n", edge, sign, fast);
                return 1;
            }
            ++cases;
        }
    // 5-digit window edges: anchor at 2^10 -> b = 37, admitted exponents [-66, 87]
    for (int sign = -1; sign <= 1; sign += 2)
        for (int edge = 0; edge < 3; ++edge) {
            std::vector<double> a(8 + 4096), x(8 + 4096);
            for (size_t i = 0; i < a.size(); ++i) {
                const int ex = i < 8 ? 10 : (edge == 0 ? 87 : (edge == 1 ? -66 : ((i & 1) ? 87 : -66)));
                a[i] = sign * std::ldexp(1.9999999999999998, ex - 1 - 3);
                x[i] = std::ldexp(1.9999999999999998, 3);
            }
            long fast = 0;
            if (!run_products_wide<5>(a, x, 0.0, 5, &fast) || fast < 4088) {
                printf("FAIL wide product edge=%d sign=%d fast=%ld\n", edge, sign, fast);
                return 1;
            }
            ++cases;
        }
    // 3-digit single-summand window edges: the first group (exponent 10) anchors lo = 10 - 51 = -41, hi = 61
    for (int sign = -1; sign <= 1; sign += 2)
        for (int edge = 0; edge < 3; ++edge) {
            std::vector<double> sv(8 + 4096);
            for (size_t i = 0; i < sv.size(); ++i) {
                const int ex = i < 8 ? 10 : (edge == 0 ? 61 : (edge == 1 ? -41 : ((i & 1) ? 61 : -41)));
                sv[i] = sign * std::ldexp(1.9999999999999998, ex);
            }
            long fast = 0;
            if (!run_singles_wide<3>(sv, 0.0, 5, &fast) || fast < 4088) {
                printf("FAIL wide single edge=%d sign=%d fast=%ld\n", edge, sign, fast);
                return 1;
            }
            ++cases;
        }
    printf("OK %d cases, %ld summands through the 3-digit window, %ld / %ld products through the 5-digit window (per-lane / range-covering anchoring), %ld summands through the 3-digit single window\n", cases, total_fast, total_wide, total_cover, total_swide);
    return 0;
}
