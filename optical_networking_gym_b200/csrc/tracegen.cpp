// tracegen.cpp -- host-side request-stream generator (no GPU code).
//
// The reference draws every request from one `random.Random` instance (CPython stdlib: MT19937)
// in the order of QRMSAEnv._next_service (reference envs/qrmsa.pyx:1079-1089) and _get_node_pair
// (qrmsa.pyx:1134-1148):
//     at  = float32(current_time + expovariate(1/mean_iat));  current_time = at
//     ht  = float32(expovariate(1/mean_holding))
//     src = choices(nodes, weights)            dst = choices(nodes, weights with src zeroed, renormalised)
//     bit_rate = choices(bit_rates, probs, k=1)        ("continuous": randint(lower, higher), qrmsa.pyx:246-254)
// This file restates the published algorithms those stdlib calls use so that a batch of envs can
// replay "a request trace the reference generated from the same seeds" without 10^6 draws/s
// CPython in the loop:
//   * MT19937 (Matsumoto & Nishimura 2002) with init_by_array seeding from the 32-bit little-endian
//     words of abs(seed), as CPython's Random.seed(int) does;
//   * random()      = (a*2^26 + b) / 2^53 with a = next32 >> 5, b = next32 >> 6;
//   * expovariate   = -log(1 - random()) / lambd;
//   * choices       = bisect_right(cum_weights, random() * cum_weights[-1], 0, n-1).
// Parity is pinned by tests/test_tracegen.py against traces recorded from the compiled reference.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/qrmsa_b200.h"

namespace {

struct MT19937 {
    uint32_t mt[624];
    int idx;

    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t *key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        int k = 624 > len ? 624 : len;
        for (; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (k = 623; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            i++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    void seed(uint64_t s) {
        uint32_t key[2] = {(uint32_t)(s & 0xffffffffu), (uint32_t)(s >> 32)};
        init_by_array(key, key[1] ? 2 : 1);
    }
    void refill() {
        const uint32_t U = 0x80000000u, L = 0x7fffffffu, A = 0x9908b0dfu;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            uint32_t y = (mt[kk] & U) | (mt[kk + 1] & L);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        for (; kk < 623; kk++) {
            uint32_t y = (mt[kk] & U) | (mt[kk + 1] & L);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        }
        uint32_t y = (mt[623] & U) | (mt[0] & L);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        idx = 0;
    }
    inline uint32_t next32() {
        if (idx >= 624) refill();
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline double random() {
        uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
};

inline int bisect_right(const double *a, double x, int lo, int hi) {
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (x < a[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

}  // namespace

struct qrmsa_tracegen {
    int n_envs, n_nodes, n_rates;
    double mean_holding;
    std::vector<MT19937> rng;
    std::vector<double> now;       // env clock (a float32 value held in a double, qrmsa.pyx:179)
    std::vector<double> lam_iat;   // 1 / mean_service_inter_arrival_time
    std::vector<double> src_cum, dst_cum, rate_cum;
    int randint_n = 0, randint_bits = 0;   // continuous bit rates: randint(lower, higher) -> _randbelow(higher - lower + 1)
};

extern "C" int qrmsa_tracegen_create(int n_envs, uint64_t base_seed, int n_nodes, int n_rates, const double *h_load,
                                     double mean_holding_time, const double *h_src_cum, const double *h_dst_cum,
                                     const double *h_rate_cum, qrmsa_tracegen **out) {
    if (!out || n_envs <= 0 || n_nodes < 2 || n_nodes > 255 || n_rates < 1 || n_rates > 255 || !h_load || !h_src_cum ||
        !h_dst_cum || !h_rate_cum || !(mean_holding_time > 0))
        return QRMSA_ERR_ARG;
    qrmsa_tracegen *g = new qrmsa_tracegen();
    g->n_envs = n_envs; g->n_nodes = n_nodes; g->n_rates = n_rates;
    // set_load takes the holding time as a C float (qrmsa.pyx:1124)
    g->mean_holding = (double)(float)mean_holding_time;
    g->rng.resize(n_envs);
    g->now.assign(n_envs, 0.0);
    g->lam_iat.resize(n_envs);
    for (int i = 0; i < n_envs; i++) {
        if (!(h_load[i] > 0)) { delete g; return QRMSA_ERR_ARG; }
        g->rng[i].seed(base_seed + (uint64_t)i);
        double mean_iat = 1.0 / (h_load[i] / g->mean_holding);  // qrmsa.pyx:1130
        g->lam_iat[i] = 1.0 / mean_iat;                         // qrmsa.pyx:1079
    }
    g->src_cum.assign(h_src_cum, h_src_cum + n_nodes);
    g->dst_cum.assign(h_dst_cum, h_dst_cum + (size_t)n_nodes * n_nodes);
    g->rate_cum.assign(h_rate_cum, h_rate_cum + n_rates);
    *out = g;
    return QRMSA_OK;
}

extern "C" int qrmsa_tracegen_set_randint_rates(qrmsa_tracegen *g, int lower, int higher) {
    if (!g || lower < 0 || higher < lower || higher - lower + 1 > 255) return QRMSA_ERR_ARG;
    g->randint_n = higher - lower + 1;
    g->randint_bits = 0;
    for (int n = g->randint_n; n; n >>= 1) g->randint_bits++;   // n.bit_length()
    return QRMSA_OK;
}

extern "C" int qrmsa_tracegen_next(qrmsa_tracegen *g, int n_requests, uint8_t *h_src, uint8_t *h_dst, uint8_t *h_rate,
                                   float *h_arrival, float *h_holding, int n_threads) {
    if (!g || n_requests < 0 || !h_src || !h_dst || !h_rate || !h_arrival || !h_holding) return QRMSA_ERR_ARG;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > g->n_envs) n_threads = g->n_envs;
    const int n_envs = g->n_envs, N = g->n_nodes, R = g->n_rates;
    const double lam_hold = 1.0 / g->mean_holding;  // qrmsa.pyx:1083
    // env-blocked so that the request-major output arrays are written in runs of consecutive envs
    auto work = [&](int e0, int e1) {
        const int BLK = 16;
        for (int b0 = e0; b0 < e1; b0 += BLK) {
            const int b1 = b0 + BLK < e1 ? b0 + BLK : e1;
            for (int r = 0; r < n_requests; r++) {
                const size_t row = (size_t)r * n_envs;
                for (int e = b0; e < b1; e++) {
                    MT19937 &rng = g->rng[e];
                    const double lam = g->lam_iat[e];
                    float at = (float)(g->now[e] + (-std::log(1.0 - rng.random()) / lam));
                    g->now[e] = (double)at;
                    float ht = (float)(-std::log(1.0 - rng.random()) / lam_hold);
                    int s = bisect_right(g->src_cum.data(), rng.random() * (g->src_cum[N - 1] + 0.0), 0, N - 1);
                    const double *dc = g->dst_cum.data() + (size_t)s * N;
                    int d = bisect_right(dc, rng.random() * (dc[N - 1] + 0.0), 0, N - 1);
                    int br;
                    if (g->randint_n) {
                        // Random.randint(a, b) = a + _randbelow_with_getrandbits(b - a + 1): k = n.bit_length();
                        // r = getrandbits(k) (one 32-bit output, top k bits) until r < n
                        uint32_t r32;
                        do { r32 = rng.next32() >> (32 - g->randint_bits); } while (r32 >= (uint32_t)g->randint_n);
                        br = (int)r32;
                    } else {
                        br = bisect_right(g->rate_cum.data(), rng.random() * (g->rate_cum[R - 1] + 0.0), 0, R - 1);
                    }
                    const size_t o = row + e;
                    h_arrival[o] = at; h_holding[o] = ht;
                    h_src[o] = (uint8_t)s; h_dst[o] = (uint8_t)d; h_rate[o] = (uint8_t)br;
                }
            }
        }
    };
    if (n_threads == 1) {
        work(0, n_envs);
    } else {
        std::vector<std::thread> th;
        int per = ((n_envs + n_threads - 1) / n_threads + 15) / 16 * 16;
        for (int t = 0; t < n_threads; t++) {
            int e0 = t * per, e1 = e0 + per > n_envs ? n_envs : e0 + per;
            if (e0 < e1) th.emplace_back(work, e0, e1);
        }
        for (auto &t : th) t.join();
    }
    return QRMSA_OK;
}

extern "C" void qrmsa_tracegen_destroy(qrmsa_tracegen *g) { delete g; }
