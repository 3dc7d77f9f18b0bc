// qrmsa_sampler.cuh -- masked categorical sampling over the action mask k_observation writes (BASELINE config 5).
//
// The consumer of the reference's mask is sb3_contrib's MaskablePPO, which asks the wrapper for `action_masks()`
// (wrappers/qrmsa_gym.py:74-75, examples/ONDM_2025/train_multi_masked_ppo.py:410-444) and samples
// a ~ Categorical(softmax(logits) restricted to mask == 1).  On the device that sample is one pass over the logits
// and the mask by the Gumbel-max identity: a = argmax_{mask} (logit_a + g_a), g_a = -log(-log u_a), u_a uniform in
// (0, 1) drawn from a Philox4x32-10 stream keyed by (seed, step) and counted by (env, action / 4) -- so the draw of
// (env, action) does not depend on the launch shape.  One warp per env row; a masked action costs one byte of a
// 16-byte load.  HBM-bound: the mask (1 B per action) is read once, a logit only where its mask byte is set, 8 B are
// written per env.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qrmsa {

__device__ __forceinline__ uint4 sampler_philox(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float sampler_logit(const float *p, size_t i) { return p[i]; }
__device__ __forceinline__ float sampler_logit(const __nv_bfloat16 *p, size_t i) { return __bfloat162float(p[i]); }

constexpr int SAMPLER_THREADS = 256;

// uniform in (0, 1): 24 random bits, centred ((r >> 8) + 0.5) / 2^24 -- never 0 or 1, so both logs are finite.
// __logf (MUFU.LG2 + one multiply; absolute error < 2^-21 away from 1) is plenty for sampling noise; the numpy
// restatement of the tests uses the exact log and accepts a swapped near tie.
__device__ __forceinline__ float sampler_gumbel(uint32_t r) {
    const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return -__logf(-__logf(u));
}

template <class T>
__global__ void __launch_bounds__(SAMPLER_THREADS)
    k_sample_masked(const T *__restrict__ logits, const uint8_t *__restrict__ mask, const int n_envs, const int n_actions,
                    const long long logit_stride, const long long mask_stride, const unsigned long long seed,
                    const unsigned long long step, long long *__restrict__ out) {
    const uint2 key = make_uint2((uint32_t)seed ^ (uint32_t)(step * 0x9E3779B97F4A7C15ull >> 32), (uint32_t)(seed >> 32) ^ (uint32_t)step);
    const int lane = threadIdx.x & 31;
    const int wpc = SAMPLER_THREADS / 32;
    // one WARP per env row: no block barrier, eight rows in flight per CTA
    for (int env = blockIdx.x * wpc + (threadIdx.x >> 5); env < n_envs; env += gridDim.x * wpc) {
        const T *lg = logits + (size_t)env * logit_stride;
        const uint8_t *mk = mask + (size_t)env * mask_stride;
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        auto consider = [&](int a, uint32_t w) {
            const float k = sampler_logit(lg, (size_t)a) + sampler_gumbel(w);
            if (k > best || (k == best && a < best_i)) { best = k; best_i = a; }
        };
        // one valid action on its own (row head / tail): its Gumbel noise is word (a & 3) of the Philox block of group a >> 2
        auto take = [&](int a) {
            const uint4 r = sampler_philox(make_uint4((uint32_t)(a >> 2), 0u, (uint32_t)env, 0u), key);
            consider(a, (a & 3) == 0 ? r.x : (a & 3) == 1 ? r.y : (a & 3) == 2 ? r.z : r.w);
        };
        // The mask row is scanned 16 bytes per lane and load (two loads in flight); rows start at any byte (n_actions is
        // odd), so up to 15 leading and 15 trailing bytes are read one by one.  Valid actions come in runs of consecutive
        // start slots: the 16-byte chunks that hold any are spread over the warp eight at a time, a lane taking one aligned
        // group of four actions -- one Philox block per group, drawn once.  Logits are only touched where the mask is set.
        const int head = min((int)((16u - (unsigned)(reinterpret_cast<uintptr_t>(mk) & 15u)) & 15u), n_actions);
        const int n16 = (n_actions - head) >> 4;
        const int tail0 = head + (n16 << 4);
        if (lane < head && mk[lane]) take(lane);
        const uint4 *mv = reinterpret_cast<const uint4 *>(mk + head);
        uint4 nxt = lane < n16 ? mv[lane] : make_uint4(0u, 0u, 0u, 0u);
        for (int c0 = 0; c0 < n16; c0 += 32) {
            const uint4 v = nxt;
            if (c0 + 32 + lane < n16) nxt = mv[c0 + 32 + lane];
            uint32_t m16 = 0u;
            if (c0 + lane < n16) {
                const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((ww[q] >> (8 * b)) & 0xffu) m16 |= 1u << (4 * q + b);
            }
            unsigned nz = __ballot_sync(0xffffffffu, m16 != 0u);
            while (nz) {
                // the (lane >> 2)-th chunk of the next eight that hold a valid action
                unsigned rest = nz;
                int src = -1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int l = rest ? __ffs(rest) - 1 : -1;
                    if (j == (lane >> 2)) src = l;
                    rest &= rest - 1u;
                }
                nz = rest;
                const uint32_t mm = __shfl_sync(0xffffffffu, m16, src < 0 ? 0 : src);
                // actions are counted from the row start, chunks from `head`: a chunk's four-action pieces straddle the
                // Philox groups unless head is a multiple of 4, so the noise word is picked per action from its own group
                const int a0 = head + ((c0 + src) << 4) + ((lane & 3) << 2);
                const uint32_t m4 = src < 0 ? 0u : (mm >> ((lane & 3) << 2)) & 0xfu;
                if (m4) {
                    const int g0 = a0 >> 2;
                    const uint4 r0 = sampler_philox(make_uint4((uint32_t)g0, 0u, (uint32_t)env, 0u), key);
                    uint4 r1 = r0;
                    if ((a0 & 3) && (m4 >> (4 - (a0 & 3))))      // some valid action of this piece lies in the next group
                        r1 = sampler_philox(make_uint4((uint32_t)(g0 + 1), 0u, (uint32_t)env, 0u), key);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (!((m4 >> b) & 1u)) continue;
                        const int a = a0 + b;
                        const uint4 &r = (a >> 2) == g0 ? r0 : r1;
                        consider(a, (a & 3) == 0 ? r.x : (a & 3) == 1 ? r.y : (a & 3) == 2 ? r.z : r.w);
                    }
                }
            }
        }
        if (lane < n_actions - tail0 && mk[tail0 + lane]) take(tail0 + lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ok > best || (ok == best && oi < best_i)) { best = ok; best_i = oi; }
        }
        // a row without a valid action cannot come from k_observation (the reject action is always valid,
        // qrmsa.pyx:766); it yields the last action, which is the reject action
        if (lane == 0) out[env] = best_i == 0x7fffffff ? (long long)(n_actions - 1) : (long long)best_i;
    }
}

}  // namespace qrmsa
