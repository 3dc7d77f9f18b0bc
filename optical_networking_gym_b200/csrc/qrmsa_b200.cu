// qrmsa_b200.cu -- C ABI (include/qrmsa_b200.h) over the sm_100a kernels of qrmsa_kernels.cuh.
// Host code: context / memory management, GN-table construction, launches.  No torch types.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "qrmsa_kernels.cuh"
#include "qrmsa_sampler.cuh"

using namespace qrmsa;

struct qrmsa_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    int threads = 512;
    int grid = 0;
    size_t ring_smem = 0;            // dynamic shared memory of the step kernel with the stream chunks only (0 = does not fit)
    size_t bm_smem = 0;              // dynamic shared memory of the step kernel with the bitmap rows staged (0 = does not fit)
    int staging = 2;                 // qrmsa_set_staging: 2 = rows + streams + paths when they fit, 1 = streams only, 0 = none
    int ring_hops_off = 0, bms_hops_off = 0;      // KParams.smem_pt_hops of the two staged variants
    int warp_off_bms = 0, warp_stride_bms = 0;   // KParams.smem_warp_* of the two staged variants
    int warp_off_ring = 0, warp_stride_ring = 0;
    size_t cta_smem = 0;   // k_step_highest_snr (and k_observation): one CTA per env
    int cta_grid = 0, cta_epc = 0, cta_env_smem = 0;
    // on-device request generator
    float *gen_clock = nullptr;
    double *gen_tables = nullptr;   // load[n_envs] | src_cum[N] | dst_cum[N*N] | rate_cum[R]
    unsigned long long gen_pos = 0;
    int n_groups = 1;
    int max_need = 1;
    KParams kp{};
    std::vector<void *> allocs;
    // staging for host-buffer entry points
    void *stage = nullptr;
    size_t stage_bytes = 0;
    int64_t *h_counters = nullptr;  // pinned
    std::vector<uint8_t> need, cls;
    std::vector<int> cls_n;
    const double *d_path_len_norm = nullptr;
    double inv_max_rate = 0.0;
    int obs_grid = 0, obs_epc = 0, obs_env_smem = 0;   // k_observation: envs per CTA, shared memory per env
    size_t obs_smem = 0;
    int obs2_grid = 0, obs2_epc = 0, obs2_env_smem = 0;   // k_observation_links (spectra up to 320 slots)
    int hs2_grid = 0, hs2_epc = 0;                        // k_step_highest_snr_links (same per-env shared-memory area)
    size_t hs2_smem = 0;
    size_t obs2_smem = 0;
    std::string err;
};

namespace {

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                             \
            return QRMSA_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)

template <class T>
int dev_alloc(qrmsa_ctx *ctx, T **out, size_t count) {
    void *p = nullptr;
    CK(cudaMalloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 16));
    ctx->allocs.push_back(p);
    *out = (T *)p;
    return QRMSA_OK;
}

template <class T>
int dev_upload(qrmsa_ctx *ctx, const T **out, const T *host, size_t count) {
    T *p = nullptr;
    int rc = dev_alloc(ctx, &p, count);
    if (rc) return rc;
    CK(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return QRMSA_OK;
}

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

int ensure_stage(qrmsa_ctx *ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return QRMSA_OK;
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    CK(cudaMalloc(&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return QRMSA_OK;
}

const double PHI_MOD[6] = {1.0, 1.0, 2.0 / 3.0, 17.0 / 25.0, 69.0 / 100.0, 13.0 / 21.0};  // osnr.pyx:38-41

}  // namespace

extern "C" const char *qrmsa_version(void) { return "qrmsa_b200 0.1 (sm_100a)"; }

extern "C" const char *qrmsa_strerror(int s) {
    switch (s) {
        case QRMSA_OK: return "ok";
        case QRMSA_ERR_ARG: return "invalid argument";
        case QRMSA_ERR_CUDA: return "CUDA runtime error";
        case QRMSA_ERR_UNSUPPORTED: return "unsupported configuration";
        case QRMSA_ERR_STATE: return "invalid call order";
        case QRMSA_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case QRMSA_ERR_ENV: return "an environment is in an error state";
        default: return "unknown status";
    }
}

extern "C" const char *qrmsa_last_error(const qrmsa_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" void qrmsa_destroy(qrmsa_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    delete ctx;
}

static int create_impl(qrmsa_ctx *ctx, const qrmsa_static_tables *t, int n_envs, int max_requests) {
    KParams &kp = ctx->kp;
    const int N = t->n_nodes, E = t->n_links, K = t->k_paths, M = t->n_mods, R = t->n_rates, S = t->n_slots;
    kp.n_envs = n_envs; kp.N = N; kp.E = E; kp.K = K; kp.M = M; kp.Mc = t->mods_to_consider; kp.R = R; kp.S = S;
    kp.W = (S + 31) / 32;
    kp.RW = row_words(S);
    kp.Hmax = t->max_hops;
    kp.D = 2 * S;
    kp.CAP = (int)round_up((size_t)(S + 1) / 2, 32);
    kp.T = max_requests;
    kp.n_req = 0;
    kp.group_size = n_envs;
    kp.f0 = t->frequency_start;
    kp.sb = t->slot_bandwidth_hz;

    // ---- validation of what the kernels implement
    if (S < 2 || S > 960) { ctx->err = "n_slots must be in 2..960 (one bitmap word per lane + the virtual slot)"; return QRMSA_ERR_UNSUPPORTED; }
    if (t->max_hops > 32 || t->max_hops < 1) { ctx->err = "paths longer than 32 hops"; return QRMSA_ERR_UNSUPPORTED; }
    if (M > 8 || R > 255 || N > 255 || E > 255 || K > 255) { ctx->err = "table dimension too large"; return QRMSA_ERR_UNSUPPORTED; }
    if (kp.Mc < 1 || kp.Mc > M) { ctx->err = "modulations_to_consider must be in 1..n_mods"; return QRMSA_ERR_UNSUPPORTED; }
    if ((long long)K * M * S >= (1 << 24)) { ctx->err = "action space exceeds 24 bits"; return QRMSA_ERR_UNSUPPORTED; }
    if (max_requests < 2 || max_requests > 16384) { ctx->err = "max_requests must be in 2..16384"; return QRMSA_ERR_UNSUPPORTED; }
    for (int l = 1; l < E; l++)
        if (t->link_alpha[l] != t->link_alpha[0]) { ctx->err = "per-link attenuation must be uniform (one G table)"; return QRMSA_ERR_UNSUPPORTED; }
    for (int m = 0; m < M; m++)
        if (t->mod_se[m] < 1 || t->mod_se[m] > 6) { ctx->err = "spectral efficiency outside 1..6 (phi table, osnr.pyx:38-41)"; return QRMSA_ERR_UNSUPPORTED; }

    // ---- slot classes
    ctx->need.assign(t->slots_needed, t->slots_needed + (size_t)R * M);
    std::vector<int> cls_n;
    for (uint8_t v : ctx->need) {
        if (v < 1) { ctx->err = "slots_needed must be >= 1"; return QRMSA_ERR_UNSUPPORTED; }
        bool seen = false;
        for (int c : cls_n) seen |= (c == v);
        if (!seen) cls_n.push_back(v);
        if (v > ctx->max_need) ctx->max_need = v;
    }
    for (size_t i = 0; i < cls_n.size(); i++)
        for (size_t j = i + 1; j < cls_n.size(); j++)
            if (cls_n[j] < cls_n[i]) std::swap(cls_n[i], cls_n[j]);
    const int NC = (int)cls_n.size();
    if (NC > 31) { ctx->err = "more than 31 distinct slot counts"; return QRMSA_ERR_UNSUPPORTED; }
    kp.NC = NC;
    kp.sentinel = (uint32_t)NC << 23;   // class NC: G row of zeros, PHIN 0
    ctx->cls_n = cls_n;
    ctx->cls.resize((size_t)R * M);
    for (size_t i = 0; i < ctx->need.size(); i++)
        for (int c = 0; c < NC; c++)
            if (cls_n[c] == ctx->need[i]) ctx->cls[i] = (uint8_t)c;
    kp.need_monotone = 1;
    for (int r = 0; r < R; r++)
        for (int m = 0; m + 1 < M; m++)
            if (ctx->need[(size_t)r * M + m] < ctx->need[(size_t)r * M + m + 1]) kp.need_monotone = 0;
    if (ctx->max_need + 1 > 97) { ctx->err = "a service (slots + guard) may span at most 4 bitmap words (number_slots <= 96)"; return QRMSA_ERR_UNSUPPORTED; }
    if ((S >> 5) + ((ctx->max_need + 1) >> 5) + 2 > 32) { ctx->err = "bitmap + shift distance exceed one warp"; return QRMSA_ERR_UNSUPPORTED; }

    // ---- GN tables (core/osnr.pyx:21-142), FP64, host libm
    const double beta_2 = -21.3e-27, gamma = 1.3e-3, h_plank = 6.626e-34, pi = M_PI;
    const double alpha = t->link_alpha[0], sb = t->slot_bandwidth_hz, P = t->launch_power_w;
    const double l_eff_a = 1.0 / (2.0 * alpha);
    const int D = kp.D;
    std::vector<double> G((size_t)(NC + 1) * D, 0.0), INV(D), PHIN(256, 0.0), W1(E), W2(E), SELF(NC), CN(NC), ASEC(NC), ACCT(M), ACCLO(M), ACCHI(M);
    for (int c = 0; c < NC; c++) {
        const double bw_r = sb * cls_n[c];
        for (int d = 0; d < D; d++) {
            const double df = (sb / 2.0) * d;
            G[(size_t)c * D + d] = asinh(pi * pi * fabs(beta_2) * l_eff_a * bw_r * (df + (bw_r / 2.0))) -
                                   asinh(pi * pi * fabs(beta_2) * l_eff_a * bw_r * (df - (bw_r / 2.0)));
        }
        for (int m = 0; m < M; m++) PHIN[(c << 3) | m] = PHI_MOD[t->mod_se[m] - 1] * bw_r;
        const double bw = bw_r;
        SELF[c] = asinh(pi * pi * fabs(beta_2) * (bw * bw) / (4.0 * alpha));
        CN[c] = pow(P / bw, 3.0) * (8.0 / (27.0 * pi * fabs(beta_2))) * (gamma * gamma) * bw / P;
        ASEC[c] = bw * h_plank / P;
    }
    INV[0] = 0.0;
    for (int d = 1; d < D; d++) INV[d] = 1.0 / ((sb / 2.0) * d);
    std::vector<double> leff(E), ex(E);
    for (int l = 0; l < E; l++) {
        const double len = t->link_span_len_m[l];
        leff[l] = (1.0 - exp(-2.0 * alpha * len)) / (2.0 * alpha);
        ex[l] = (exp(2.0 * alpha * len) - 1.0) * t->link_nf[l];
        W1[l] = t->link_n_spans[l] * leff[l];
        W2[l] = -(t->link_n_spans[l] * leff[l] * (5.0 / 3.0) * (leff[l] / len));  // stored negated
    }
    for (int m = 0; m < M; m++) {
        kp.mod_thr_nomargin[m] = t->mod_min_osnr[m];
        kp.acct0[m] = pow(10.0, -t->mod_min_osnr[m] / 10.0);                  // gsnr < minimum_osnr  <=>  acc > acct0
        kp.acct0_lo[m] = pow(10.0, -(t->mod_min_osnr[m] + 1e-3) / 10.0);
        kp.acct0_hi[m] = pow(10.0, -(t->mod_min_osnr[m] - 1e-3) / 10.0);
    }
    kp.feat = 0; kp.n_defrag = 0; kp.step_disrupted = nullptr;
    for (int m = 0; m < M; m++) {
        // accept iff gsnr_dB >= thr (heuristics.py:957-958)  <=>  acc = 1/GSNR <= 10^(-thr/10);
        // near-threshold flag: |gsnr_dB - thr| < 1e-3 dB
        const double thr = t->mod_min_osnr[m] + t->margin_db;
        ACCT[m] = pow(10.0, -thr / 10.0);
        ACCLO[m] = pow(10.0, -(thr + 1e-3) / 10.0);
        ACCHI[m] = pow(10.0, -(thr - 1e-3) / 10.0);
    }
    const size_t n_paths = (size_t)N * N * K;
    std::vector<double2> pgn(n_paths);
    for (size_t pi_ = 0; pi_ < n_paths; pi_++) {
        double pa = 0.0, pb = 0.0;
        const int hops = t->path_hops[pi_];
        if (hops > t->max_hops) { ctx->err = "path_hops exceeds max_hops"; return QRMSA_ERR_ARG; }
        for (int h = 0; h < hops; h++) {
            const int l = t->path_links[pi_ * t->max_hops + h];
            if (l >= E) { ctx->err = "link index out of range"; return QRMSA_ERR_ARG; }
            pa += t->link_n_spans[l] * ex[l];
            pb += W1[l];
        }
        pgn[pi_] = make_double2(pa, pb);
    }
    std::vector<int32_t> rate_milli(R);
    for (int r = 0; r < R; r++) rate_milli[r] = (int32_t)llround(t->bit_rates[r] * 1000.0);

    // ---- "prunable" paths: every neighbour term W1_l*G[c][d] + W2_l*PHIN[c,m]*INV[d] (W2 negated) is >= 0 for
    // all channel classes, modulations and admissible distances d >= n_c + 3 on all links of the path, so the
    // empty-network value bounds acc from below and hopeless modulations can be refused without a GN sum.
    std::vector<char> link_pos(E, 1);
    for (int l = 0; l < E; l++)
        for (int c = 0; c < NC && link_pos[l]; c++)
            for (int m = 0; m < M && link_pos[l]; m++)
                for (int d = cls_n[c] + 3; d < D; d++)
                    if (W1[l] * G[(size_t)c * D + d] + W2[l] * (PHIN[(c << 3) | m] * INV[d]) < 0.0) { link_pos[l] = 0; break; }
    std::vector<uint8_t> hops_dev(n_paths);
    for (size_t pi_ = 0; pi_ < n_paths; pi_++) {
        const int hops = t->path_hops[pi_];
        bool ok = hops > 0;
        for (int h = 0; h < hops; h++) ok = ok && link_pos[t->path_links[pi_ * t->max_hops + h]];
        hops_dev[pi_] = (uint8_t)(hops | (ok ? 0x80 : 0));
    }

    // compact path table for the step kernel's shared memory: u16 offset[n_paths] | u8 hops+flag[n_paths] | u8 links[...]
    std::vector<unsigned char> ptab;
    bool ptab_ok = true;
    {
        size_t total = 0;
        for (size_t pi_ = 0; pi_ < n_paths; pi_++) total += t->path_hops[pi_];
        kp.pt_hops_off = (int)round_up(2 * n_paths, 4);
        kp.pt_links_off = kp.pt_hops_off + (int)round_up(n_paths, 4);
        ptab.assign(round_up((size_t)kp.pt_links_off + total + 32, 16), 0);   // + 32: a full warp may read past the last path
        if (total > 0xffff) ptab_ok = false;
        size_t off = 0;
        for (size_t pi_ = 0; pi_ < n_paths && ptab_ok; pi_++) {
            const uint16_t o = (uint16_t)off;
            memcpy(&ptab[2 * pi_], &o, 2);
            ptab[kp.pt_hops_off + pi_] = hops_dev[pi_];
            for (int h = 0; h < t->path_hops[pi_]; h++) ptab[kp.pt_links_off + off++] = t->path_links[pi_ * t->max_hops + h];
        }
        kp.ptab_bytes = (int)ptab.size();
    }

    // ---- shared-memory image, fixed layout (qrmsa_kernels.cuh namespace lay)
    if (R * M > lay::MAX_RM || E > lay::MAX_E || NC > lay::MAX_NC || M > lay::MAX_M || R > lay::MAX_R) {
        ctx->err = "table dimension exceeds the shared-memory layout capacity";
        return QRMSA_ERR_UNSUPPORTED;
    }
    std::vector<unsigned char> blob((size_t)round_up(lay::G(D) + (size_t)(NC + 1) * D * 8, 16), 0);
    auto put = [&](int off, const void *src, size_t n) { memcpy(blob.data() + off, src, n); };
    put(lay::PHIN, PHIN.data(), PHIN.size() * 8);
    put(lay::W1, W1.data(), W1.size() * 8);
    put(lay::W2, W2.data(), W2.size() * 8);
    put(lay::SELF, SELF.data(), SELF.size() * 8);
    put(lay::CN, CN.data(), CN.size() * 8);
    put(lay::ASEC, ASEC.data(), ASEC.size() * 8);
    put(lay::ACCT, ACCT.data(), ACCT.size() * 8);
    put(lay::ACCLO, ACCLO.data(), ACCLO.size() * 8);
    put(lay::ACCHI, ACCHI.data(), ACCHI.size() * 8);
    put(lay::NEED, ctx->need.data(), ctx->need.size());
    put(lay::CLS, ctx->cls.data(), ctx->cls.size());
    put(lay::RATE, rate_milli.data(), rate_milli.size() * 4);
    put(lay::INV, INV.data(), INV.size() * 8);
    put(lay::G(D), G.data(), G.size() * 8);
    kp.blob_bytes = (int)blob.size();

    // ---- device properties and launch shape
    CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
    CK(cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    int smem_sm = 0;
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    if (kp.blob_bytes + 64 > ctx->smem_optin) {
        ctx->err = "GN tables (" + std::to_string(kp.blob_bytes) + " B) exceed shared memory per block";
        return QRMSA_ERR_UNSUPPORTED;
    }
    // one 1024-thread CTA per SM: the 32 warps share one copy of the GN tables, which leaves the rest of the
    // 228 KB for L1 (env state is re-read from L1/L2 across the steps of a launch)
    ctx->threads = MAX_THREADS;
    const int wpc = ctx->threads / 32;
    const int want = (n_envs + wpc - 1) / wpc;
    ctx->grid = want < ctx->sm_count ? want : ctx->sm_count;
    // the step kernels keep an 8-byte mbarrier in static shared memory next to the dynamic block
    cudaFuncAttributes fa{};
    CK(cudaFuncGetAttributes(&fa, k_step_policy<0, 0, 0, POLICY_FIRST_FIT>));
    const size_t smem_budget = (size_t)ctx->smem_optin - fa.sharedSizeBytes;
    auto opt_in = [&](const void *fn, size_t bytes) { return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess; };
    bool ok = true;
    ok &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_FIRST_FIT>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<640, 6, 5, POLICY_FIRST_FIT>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_FIRST_FIT>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_LOAD_BALANCING>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_LOAD_BALANCING>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_LB_FIRST_FIT>, kp.blob_bytes);
    ok &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_LB_FIRST_FIT>, kp.blob_bytes);
    if (!ok) { (void)cudaGetLastError(); ctx->err = "GN tables exceed shared memory per block"; return QRMSA_ERR_UNSUPPORTED; }
    // Bitmap rows staged in shared memory (one area per warp after the tables): every row read, commit and release
    // becomes an LDS/STS; a compact copy of the path table (hop counts + link ids) follows them.  Used whenever
    // tables + rows + paths fit what a CTA may have beside the kernel's static shared memory: nobel-eu/320 =
    // 96 + 84 + 31 KB, which leaves L1 only 28 KB for the lists, the position table and the trace -- measured 7.50e8
    // (all through L1, 156 KB) -> 7.85e8 (rows in shared memory, 60 KB of L1) -> 8.03e8 (rows + paths).
    ctx->bm_smem = 0;
    kp.smem_pt_off = kp.smem_pt_hops = kp.smem_pt_links = 0;
    {
        ctx->warp_off_bms = kp.blob_bytes;
        ctx->warp_stride_bms = WARP_STREAM_BYTES + E * kp.RW * 4;
        const int pt_off = ctx->warp_off_bms + wpc * ctx->warp_stride_bms;
        const size_t need = (size_t)pt_off + ptab.size();
        if (ptab_ok && need <= smem_budget) {
            bool fit = true;
            fit &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_FIRST_FIT, 1>, need);
            fit &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 1>, need);
            fit &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_LOAD_BALANCING, 1>, need);
            fit &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_LOAD_BALANCING, 1>, need);
            fit &= opt_in((const void *)k_step_policy<320, 6, 5, POLICY_LB_FIRST_FIT, 1>, need);
            fit &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_LB_FIRST_FIT, 1>, need);
            if (fit) {
                ctx->bm_smem = need;
                kp.smem_pt_off = pt_off;
                kp.smem_pt_hops = ctx->bms_hops_off = pt_off + kp.pt_hops_off;
                kp.smem_pt_links = pt_off + kp.pt_links_off;
            } else {
                (void)cudaGetLastError();   // the attribute call refused: fall back to the smaller variants below
            }
        }
    }
    // the stream chunks alone, 512 bytes per warp after the tables (configurations whose rows do not fit)
    ctx->ring_smem = 0;
    {
        size_t need = (size_t)kp.blob_bytes + (size_t)wpc * WARP_STREAM_BYTES;
        ctx->warp_off_ring = kp.blob_bytes;
        ctx->warp_stride_ring = WARP_STREAM_BYTES;
        // + the hop-count bytes of the path table (PathTab<2>)
        ctx->ring_hops_off = (int)need;
        need += round_up(n_paths, 16);
        if (need <= smem_budget) {
            bool fit = opt_in((const void *)k_step_policy<640, 6, 5, POLICY_FIRST_FIT, 2>, need);
            fit &= opt_in((const void *)k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 2>, need);
            if (fit) ctx->ring_smem = need;
            else (void)cudaGetLastError();
        }
    }
    CK(cudaFuncSetAttribute(k_step_action, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.blob_bytes));
    CK(cudaFuncSetAttribute(k_probe_gsnr, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.blob_bytes));
    CK(cudaFuncSetAttribute(k_build_schedule, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));

    // ---- CTA-per-env kernels (observation, highest-SNR policy): shared memory = tables + X[c2] + staged records
    int rc;
    ctx->cta_env_smem = (int)round_up(sizeof(ObsSmem) + (size_t)D * 8 + (size_t)S * 8 + (size_t)kp.Hmax * kp.CAP * 4, 16);
    {
        int epc = OBS_MAX_EPC;   // envs per CTA, each with its own staging area after the shared tables
        while (epc > 1 && kp.blob_bytes + epc * ctx->cta_env_smem > ctx->smem_optin) epc >>= 1;
        ctx->cta_smem = (size_t)kp.blob_bytes + (size_t)epc * ctx->cta_env_smem;
        if ((int)ctx->cta_smem <= ctx->smem_optin) {
            ctx->cta_epc = epc;
            CK(cudaFuncSetAttribute(k_step_highest_snr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->cta_smem));
            const int per_sm = std::max(1, std::min((int)((size_t)smem_sm / (ctx->cta_smem + 1024)), 2048 / (epc * OBS_ENV_THREADS)));
            ctx->cta_grid = std::min((n_envs + epc - 1) / epc, ctx->sm_count * per_sm);
        }
    }
    // link-major highest-SNR kernel: spectra up to 320 slots and at most 8 paths per pair (one bit per path in the link
    // sets); its static shared memory (the per-warp picks) comes off the budget
    if (D <= OBS2_MAX_D && K <= 8) {
        cudaFuncAttributes fh{};
        CK(cudaFuncGetAttributes(&fh, k_step_highest_snr_links));
        const size_t budget_h = (size_t)ctx->smem_optin - fh.sharedSizeBytes;
        ctx->obs2_env_smem = obs2_env_smem(K, D, kp.W, E);
        int e2 = OBS_MAX_EPC;
        while (e2 > 1 && (size_t)kp.blob_bytes + (size_t)e2 * ctx->obs2_env_smem > budget_h) e2 >>= 1;
        const size_t need2 = (size_t)kp.blob_bytes + (size_t)e2 * ctx->obs2_env_smem;
        if (need2 <= budget_h &&
            cudaFuncSetAttribute(k_step_highest_snr_links, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need2) == cudaSuccess) {
            ctx->hs2_epc = e2;
            ctx->hs2_smem = need2;
            const int per_sm = std::max(1, std::min((int)((size_t)smem_sm / (need2 + 1024)), 2048 / (e2 * OBS_ENV_THREADS)));
            ctx->hs2_grid = std::min((n_envs + e2 - 1) / e2, ctx->sm_count * per_sm);
        } else {
            (void)cudaGetLastError();
        }
    }
    // ---- observation kernel: route-length normalisation (qrmsa.pyx:676-690) and launch shape
    if (t->path_length_km && t->link_length_km) {
        double lo = t->link_length_km[0], hi = t->link_length_km[0];
        for (int l = 1; l < E; l++) { lo = std::min(lo, t->link_length_km[l]); hi = std::max(hi, t->link_length_km[l]); }
        std::vector<double> pln(n_paths);
        for (size_t i = 0; i < n_paths; i++) pln[i] = hi == lo ? 0.0 : (t->path_length_km[i] - lo) / (hi - lo);
        if ((rc = dev_upload(ctx, &ctx->d_path_len_norm, pln.data(), n_paths))) return rc;
        double mx = 0.0;
        for (int r = 0; r < R; r++) mx = std::max(mx, (double)rate_milli[r]);
        ctx->inv_max_rate = mx > 0 ? 1.0 / mx : 0.0;
        // up to OBS_MAX_EPC envs per CTA (each with its own X / record staging area after the shared tables)
        ctx->obs_env_smem = (int)round_up(sizeof(ObsSmem) + (size_t)D * 8 + (size_t)S * 8 + (size_t)kp.Hmax * kp.CAP * 4, 16);
        int epc = OBS_MAX_EPC;
        while (epc > 1 && kp.blob_bytes + epc * ctx->obs_env_smem > ctx->smem_optin) epc >>= 1;
        ctx->obs_smem = (size_t)kp.blob_bytes + (size_t)epc * ctx->obs_env_smem;
        if ((int)ctx->obs_smem <= ctx->smem_optin) {
            ctx->obs_epc = epc;
            CK(cudaFuncSetAttribute(k_observation, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->obs_smem));
            const int per_sm = std::max(1, std::min((int)((size_t)smem_sm / (ctx->obs_smem + 1024)), 2048 / (epc * OBS_ENV_THREADS)));
            ctx->obs_grid = std::min((n_envs + epc - 1) / epc, ctx->sm_count * per_sm);
        }
        // link-major variant: spectra up to 320 slots and at most 8 paths per pair (one bit per path in the link sets)
        if (D <= OBS2_MAX_D && K <= 8) {
            ctx->obs2_env_smem = obs2_env_smem(K, D, kp.W, E);
            int e2 = OBS_MAX_EPC;
            while (e2 > 1 && (size_t)kp.blob_bytes + (size_t)e2 * ctx->obs2_env_smem > smem_budget) e2 >>= 1;
            const size_t need2 = (size_t)kp.blob_bytes + (size_t)e2 * ctx->obs2_env_smem;
            if (need2 <= smem_budget &&
                cudaFuncSetAttribute(k_observation_links, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need2) == cudaSuccess) {
                ctx->obs2_epc = e2;
                ctx->obs2_smem = need2;
                const int per_sm = std::max(1, std::min((int)((size_t)smem_sm / (need2 + 1024)), 2048 / (e2 * OBS_ENV_THREADS)));
                ctx->obs2_grid = std::min((n_envs + e2 - 1) / e2, ctx->sm_count * per_sm);
            } else {
                (void)cudaGetLastError();
            }
        }
    }

    // ---- uploads and state
    if ((rc = dev_upload(ctx, &kp.path_hops, (const uint8_t *)hops_dev.data(), n_paths))) return rc;
    if ((rc = dev_upload(ctx, &kp.ptab, (const unsigned char *)ptab.data(), ptab.size()))) return rc;
    if ((rc = dev_upload(ctx, &kp.path_links, t->path_links, n_paths * t->max_hops))) return rc;
    if ((rc = dev_upload(ctx, &kp.path_gn, pgn.data(), n_paths))) return rc;
    {
        if ((rc = dev_alloc(ctx, &kp.work, 4))) return rc;
        if ((rc = dev_alloc(ctx, &kp.counted, (size_t)n_envs))) return rc;
    }
    if ((rc = dev_upload(ctx, &kp.blob, (const unsigned char *)blob.data(), blob.size()))) return rc;
    kp.bm_stride = (size_t)E * kp.RW;
    if ((rc = dev_alloc(ctx, &kp.bm, (size_t)n_envs * kp.bm_stride))) return rc;
    if ((rc = dev_alloc(ctx, &kp.lists, (size_t)n_envs * E * kp.CAP))) return rc;
    // list positions fit a byte up to 511 slots; above that there is no table and a release finds its records by search
    kp.pos_bytes = kp.CAP <= 256 ? 1 : 0;   // (search on nobel-eu/320 too: 8.94e8 vs 9.39e8 with the table)
    kp.pos_stride = (size_t)E * kp.CAP * kp.pos_bytes;
    // (a 256-byte stand-in without a table, so that the kernels' per-env pointer is still a global address)
    if ((rc = dev_alloc(ctx, &kp.pos, kp.pos_bytes ? (size_t)n_envs * kp.pos_stride : (size_t)256))) return rc;
    if ((rc = dev_alloc(ctx, &kp.trace, (size_t)n_envs * kp.T))) return rc;
    if ((rc = dev_alloc(ctx, &kp.perm, (size_t)n_envs * kp.T))) return rc;
    if ((rc = dev_alloc(ctx, &kp.estate, (size_t)n_envs))) return rc;
    if ((rc = dev_alloc(ctx, &kp.maxmod, (size_t)n_envs))) return rc;
    if ((rc = dev_alloc(ctx, &kp.counters, (size_t)QRMSA_N_COUNTERS * 1))) return rc;
    CK(cudaMemset(kp.counters, 0, sizeof(unsigned long long) * QRMSA_N_COUNTERS));
    CK(cudaMallocHost((void **)&ctx->h_counters, sizeof(int64_t) * QRMSA_N_COUNTERS));
    kp.gsnr_log = nullptr;
    k_reset<<<ctx->sm_count * 4, 256>>>(kp);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return QRMSA_OK;
}

extern "C" int qrmsa_create(const qrmsa_static_tables *t, int n_envs, int max_requests, int device, qrmsa_ctx **out) {
    if (!t || !out || n_envs <= 0) return QRMSA_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return QRMSA_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return QRMSA_ERR_NO_DEVICE;
    qrmsa_ctx *ctx = new qrmsa_ctx();
    ctx->device = device;
    int rc = create_impl(ctx, t, n_envs, max_requests);
    *out = ctx;  // returned even on failure so that qrmsa_last_error can be read; caller destroys it
    return rc;
}

extern "C" int qrmsa_set_groups(qrmsa_ctx *ctx, int n_groups) {
    if (!ctx || n_groups < 1 || ctx->kp.n_envs % n_groups) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    unsigned long long *c = nullptr;
    int rc = dev_alloc(ctx, &c, (size_t)QRMSA_N_COUNTERS * n_groups);
    if (rc) return rc;
    CK(cudaMemset(c, 0, sizeof(unsigned long long) * QRMSA_N_COUNTERS * n_groups));
    int64_t *h = nullptr;   // the new pinned buffer first: a failed allocation leaves the context as it was
    CK(cudaMallocHost((void **)&h, sizeof(int64_t) * QRMSA_N_COUNTERS * n_groups));
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    ctx->h_counters = h;
    ctx->kp.counters = c;
    ctx->n_groups = n_groups;
    ctx->kp.group_size = ctx->kp.n_envs / n_groups;
    return QRMSA_OK;
}

extern "C" int qrmsa_set_features(qrmsa_ctx *ctx, int measure_disruptions, int defragmentation, int n_defrag_services) {
    if (!ctx || n_defrag_services < 0) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    KParams &kp = ctx->kp;
    kp.feat = (measure_disruptions ? 1 : 0) | (defragmentation ? 2 : 0);
    kp.n_defrag = n_defrag_services;
    if ((kp.feat & 1) && !kp.step_disrupted) {
        int rc = dev_alloc(ctx, &kp.step_disrupted, (size_t)kp.n_envs);
        if (rc) return rc;
        CK(cudaMemset(kp.step_disrupted, 0, sizeof(int) * (size_t)kp.n_envs));
    }
    if (kp.feat) {
        CK(cudaFuncSetAttribute(k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.blob_bytes));
        CK(cudaFuncSetAttribute(k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.blob_bytes));
        CK(cudaFuncSetAttribute(k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.blob_bytes));
    }
    return QRMSA_OK;
}

extern "C" int qrmsa_get_max_modulation_idx_host(qrmsa_ctx *ctx, uint8_t *h_out, void *stream) {
    if (!ctx || !h_out) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(h_out, ctx->kp.maxmod, (size_t)ctx->kp.n_envs, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return QRMSA_OK;
}

extern "C" int qrmsa_get_step_disrupted_host(qrmsa_ctx *ctx, int32_t *h_out, void *stream) {
    if (!ctx || !h_out) return QRMSA_ERR_ARG;
    if (!ctx->kp.step_disrupted) { ctx->err = "measure_disruptions is not enabled"; return QRMSA_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(h_out, ctx->kp.step_disrupted, sizeof(int32_t) * (size_t)ctx->kp.n_envs, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return QRMSA_OK;
}

extern "C" int qrmsa_enable_gsnr_log(qrmsa_ctx *ctx, int enable) {
    if (!ctx) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (enable && !ctx->kp.gsnr_log) {
        int rc = dev_alloc(ctx, &ctx->kp.gsnr_log, (size_t)ctx->kp.n_envs * ctx->kp.T * 3);
        if (rc) return rc;
        CK(cudaMemset(ctx->kp.gsnr_log, 0, sizeof(double) * (size_t)ctx->kp.n_envs * ctx->kp.T * 3));
    }
    return QRMSA_OK;
}

extern "C" int qrmsa_reset(qrmsa_ctx *ctx, void *stream) {
    if (!ctx) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ctx->kp.n_req = 0;
    k_reset<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->kp);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(ctx->kp.counters, 0, sizeof(unsigned long long) * QRMSA_N_COUNTERS * ctx->n_groups, st));
    return QRMSA_OK;
}

extern "C" int qrmsa_cancel_pending_releases(qrmsa_ctx *ctx, void *stream) {
    if (!ctx) return QRMSA_ERR_ARG;
    if (ctx->kp.feat) { ctx->err = "not combinable with measure_disruptions / defragmentation"; return QRMSA_ERR_UNSUPPORTED; }
    if (ctx->kp.n_req < 1) { ctx->err = "no trace loaded"; return QRMSA_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    k_cancel_releases<<<(ctx->kp.n_envs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ctx->kp);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

static int build_schedule(qrmsa_ctx *ctx, int n_requests, cudaStream_t st);

extern "C" int qrmsa_load_trace(qrmsa_ctx *ctx, const uint8_t *d_src, const uint8_t *d_dst, const uint8_t *d_rate,
                                const float *d_arrival, const float *d_holding, int n_requests, void *stream) {
    if (!ctx || !d_src || !d_dst || !d_rate || !d_arrival || !d_holding) return QRMSA_ERR_ARG;
    if (n_requests < 1 || n_requests > ctx->kp.T) { ctx->err = "n_requests outside 1..max_requests"; return QRMSA_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    KParams &kp = ctx->kp;
    kp.n_req = n_requests;
    dim3 grid((kp.n_envs + 31) / 32, (n_requests + 31) / 32);
    k_ingest_trace<<<grid, 256, 0, st>>>(kp, d_src, d_dst, d_rate, d_arrival, d_holding, n_requests);
    CK(cudaGetLastError());
    return build_schedule(ctx, n_requests, st);
}

extern "C" int qrmsa_load_trace_host(qrmsa_ctx *ctx, const uint8_t *h_src, const uint8_t *h_dst, const uint8_t *h_rate,
                                     const float *h_arrival, const float *h_holding, int n_requests, void *stream) {
    if (!ctx || !h_src || !h_dst || !h_rate || !h_arrival || !h_holding) return QRMSA_ERR_ARG;
    if (n_requests < 1 || n_requests > ctx->kp.T) { ctx->err = "n_requests outside 1..max_requests"; return QRMSA_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)n_requests * ctx->kp.n_envs;
    const size_t nb = round_up(n, 256);
    int rc = ensure_stage(ctx, nb * 3 + nb * 8);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)ctx->stage;
    float *d_arr = (float *)base, *d_hold = (float *)(base + nb * 4);
    uint8_t *d_src = base + nb * 8, *d_dst = d_src + nb, *d_rate = d_dst + nb;
    CK(cudaMemcpyAsync(d_arr, h_arrival, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_hold, h_holding, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_src, h_src, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_dst, h_dst, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rate, h_rate, n, cudaMemcpyHostToDevice, st));
    return qrmsa_load_trace(ctx, d_src, d_dst, d_rate, d_arr, d_hold, n_requests, stream);
}

extern "C" int qrmsa_load_trace_host_strided(qrmsa_ctx *ctx, const uint8_t *h_src, const uint8_t *h_dst,
                                             const uint8_t *h_rate, const float *h_arrival, const float *h_holding,
                                             int n_requests, int64_t row_stride, void *stream) {
    if (!ctx || !h_src || !h_dst || !h_rate || !h_arrival || !h_holding || row_stride < ctx->kp.n_envs) return QRMSA_ERR_ARG;
    if (n_requests < 1 || n_requests > ctx->kp.T) { ctx->err = "n_requests outside 1..max_requests"; return QRMSA_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ne = (size_t)ctx->kp.n_envs;
    const size_t n = (size_t)n_requests * ne;
    const size_t nb = round_up(n, 256);
    // the staging buffer is sized once for the largest use of an episode so that it is never re-allocated
    // while copies of an earlier call are still in flight
    int rc = ensure_stage(ctx, std::max(nb * 11, (size_t)ctx->kp.T * ne * 11 + 4096));
    if (rc) return rc;
    unsigned char *base = (unsigned char *)ctx->stage;
    float *d_arr = (float *)base, *d_hold = (float *)(base + nb * 4);
    uint8_t *d_src = base + nb * 8, *d_dst = d_src + nb, *d_rate = d_dst + nb;
    const size_t rs = (size_t)row_stride;
    CK(cudaMemcpy2DAsync(d_arr, ne * 4, h_arrival, rs * 4, ne * 4, n_requests, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_hold, ne * 4, h_holding, rs * 4, ne * 4, n_requests, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_src, ne, h_src, rs, ne, n_requests, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_dst, ne, h_dst, rs, ne, n_requests, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(d_rate, ne, h_rate, rs, ne, n_requests, cudaMemcpyHostToDevice, st));
    return qrmsa_load_trace(ctx, d_src, d_dst, d_rate, d_arr, d_hold, n_requests, stream);
}

static int build_schedule(qrmsa_ctx *ctx, int n_requests, cudaStream_t st) {
    const KParams &kp = ctx->kp;
    int n_pad = 2;
    while (n_pad < n_requests) n_pad <<= 1;
    // barrier-heavy (one per compare-exchange stage): several small CTAs per SM overlap each other's waits
    const int cap = 256;
    int threads = n_pad / 2 < cap ? (n_pad / 2 < 32 ? 32 : n_pad / 2) : cap;
    int blocks = kp.n_envs < ctx->sm_count * 16 ? kp.n_envs : ctx->sm_count * 16;
    k_build_schedule<<<blocks, threads, (size_t)n_pad * 8, st>>>(kp, n_requests, n_pad);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

extern "C" int qrmsa_generate_trace(qrmsa_ctx *ctx, uint64_t seed, int restart, int64_t env_offset, const double *h_load,
                                    double mean_holding_time, const double *h_src_cum, const double *h_dst_cum,
                                    const double *h_rate_cum, int n_requests, void *stream) {
    if (!ctx || !h_load || !h_src_cum || !h_dst_cum || !h_rate_cum || !(mean_holding_time > 0) || env_offset < 0) return QRMSA_ERR_ARG;
    if (n_requests < 1 || n_requests > ctx->kp.T) { ctx->err = "n_requests outside 1..max_requests"; return QRMSA_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    KParams &kp = ctx->kp;
    const size_t ne = (size_t)kp.n_envs, N = (size_t)kp.N, R = (size_t)kp.R;
    for (size_t i = 0; i < ne; i++)
        if (!(h_load[i] > 0)) { ctx->err = "load must be positive"; return QRMSA_ERR_ARG; }
    if (!ctx->gen_clock) {
        int rc = dev_alloc(ctx, &ctx->gen_clock, ne);
        if (rc) return rc;
        if ((rc = dev_alloc(ctx, &ctx->gen_tables, ne + N + N * N + R))) return rc;
        restart = 1;
    }
    if (restart) {
        CK(cudaMemsetAsync(ctx->gen_clock, 0, ne * sizeof(float), st));
        ctx->gen_pos = 0;
    }
    double *d_load = ctx->gen_tables, *d_src = d_load + ne, *d_dst = d_src + N, *d_rate = d_dst + N * N;
    CK(cudaMemcpyAsync(d_load, h_load, ne * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_src, h_src_cum, N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_dst, h_dst_cum, N * N * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rate, h_rate_cum, R * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // the host tables may be pageable and short-lived
    kp.n_req = n_requests;
    // set_load takes the holding time as a C float (qrmsa.pyx:1124)
    k_generate_trace<<<(kp.n_envs + 7) / 8, 256, 0, st>>>(kp, seed, ctx->gen_pos, env_offset, d_load,
                                                              (double)(float)mean_holding_time, d_src, d_dst, d_rate,
                                                              ctx->gen_clock, n_requests);
    CK(cudaGetLastError());
    ctx->gen_pos += (unsigned long long)n_requests;
    return build_schedule(ctx, n_requests, st);
}

extern "C" int qrmsa_get_trace_host(qrmsa_ctx *ctx, int first, int count, uint8_t *h_src, uint8_t *h_dst, uint8_t *h_rate,
                                    float *h_arrival, float *h_holding, void *stream) {
    if (!ctx || !h_src || !h_dst || !h_rate || !h_arrival || !h_holding || first < 0 || count < 0 ||
        first + count > ctx->kp.n_req)
        return QRMSA_ERR_ARG;
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)count * ctx->kp.n_envs, nb = round_up(n, 256);
    int rc = ensure_stage(ctx, nb * 11);
    if (rc) return rc;
    unsigned char *base = (unsigned char *)ctx->stage;
    float *d_arr = (float *)base, *d_hold = (float *)(base + nb * 4);
    uint8_t *d_src = base + nb * 8, *d_dst = d_src + nb, *d_rate = d_dst + nb;
    dim3 grid((ctx->kp.n_envs + 31) / 32, (count + 31) / 32);
    k_gather_trace<<<grid, 256, 0, st>>>(ctx->kp, first, count, d_src, d_dst, d_rate, d_arr, d_hold);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_arrival, d_arr, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_holding, d_hold, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_src, d_src, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_dst, d_dst, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_rate, d_rate, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return QRMSA_OK;
}

extern "C" int qrmsa_step_heuristic(qrmsa_ctx *ctx, int policy, int n_steps, void *stream) {
    if (!ctx || n_steps < 0) return QRMSA_ERR_ARG;
    if (policy != QRMSA_POLICY_FIRST_FIT && policy != QRMSA_POLICY_LOAD_BALANCING && policy != QRMSA_POLICY_HIGHEST_SNR &&
        policy != QRMSA_POLICY_LB_FIRST_FIT) { ctx->err = "unknown policy"; return QRMSA_ERR_ARG; }
    if (ctx->kp.n_req < 2) { ctx->err = "no trace loaded"; return QRMSA_ERR_STATE; }
    if (n_steps == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int g = ctx->grid, th = ctx->threads, sm = ctx->kp.blob_bytes;
    // what the kernel keeps in shared memory beside the tables: rows + stream chunks + paths when they fit, else the
    // stream chunks, else nothing (qrmsa_set_staging lowers the level for the equivalence tests)
    const size_t bsm = ctx->staging >= 2 ? ctx->bm_smem : 0;
    const size_t rsm = (!bsm && ctx->staging >= 1) ? ctx->ring_smem : 0;
    KParams kp = ctx->kp;
    kp.smem_warp_off = bsm ? ctx->warp_off_bms : ctx->warp_off_ring;
    kp.smem_warp_stride = bsm ? ctx->warp_stride_bms : ctx->warp_stride_ring;
    kp.smem_pt_hops = bsm ? ctx->bms_hops_off : (rsm ? ctx->ring_hops_off : 0);
    // compile-time specialisations for the BASELINE configurations; anything else takes the generic kernel
    const bool c320 = kp.S == 320 && kp.M == 6 && kp.K == 5, c640 = kp.S == 640 && kp.M == 6 && kp.K == 5;
    if (kp.Mc < kp.M) {   // heuristics.py:36-54 composes the action from all modulations: with fewer digits it does not decode
        ctx->err = "the fused heuristics address all modulations; with modulations_to_consider < n_mods use qrmsa_step_action";
        return QRMSA_ERR_UNSUPPORTED;
    }
    if (kp.feat && policy != QRMSA_POLICY_FIRST_FIT) {
        ctx->err = "measure_disruptions / defragmentation are built for the first-fit policy and for qrmsa_step_action";
        return QRMSA_ERR_UNSUPPORTED;
    }
    if (policy == QRMSA_POLICY_FIRST_FIT && kp.feat) {
        CK(cudaMemsetAsync(kp.work, 0, 4, st));
        if (kp.feat == 1) k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 1><<<g, th, sm, st>>>(kp, n_steps);
        else if (kp.feat == 2) k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 2><<<g, th, sm, st>>>(kp, n_steps);
        else k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 0, 3><<<g, th, sm, st>>>(kp, n_steps);
    } else if (policy == QRMSA_POLICY_FIRST_FIT) {
        CK(cudaMemsetAsync(kp.work, 0, 4, st));
        if (c320 && bsm) k_step_policy<320, 6, 5, POLICY_FIRST_FIT, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else if (c320) k_step_policy<320, 6, 5, POLICY_FIRST_FIT><<<g, th, sm, st>>>(kp, n_steps);
        else if (c640 && rsm) k_step_policy<640, 6, 5, POLICY_FIRST_FIT, 2><<<g, th, rsm, st>>>(kp, n_steps);
        else if (c640) k_step_policy<640, 6, 5, POLICY_FIRST_FIT><<<g, th, sm, st>>>(kp, n_steps);
        else if (bsm) k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else if (rsm) k_step_policy<0, 0, 0, POLICY_FIRST_FIT, 2><<<g, th, rsm, st>>>(kp, n_steps);
        else k_step_policy<0, 0, 0, POLICY_FIRST_FIT><<<g, th, sm, st>>>(kp, n_steps);
    } else if (policy == QRMSA_POLICY_LB_FIRST_FIT) {
        CK(cudaMemsetAsync(kp.work, 0, 4, st));
        if (c320 && bsm) k_step_policy<320, 6, 5, POLICY_LB_FIRST_FIT, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else if (c320) k_step_policy<320, 6, 5, POLICY_LB_FIRST_FIT><<<g, th, sm, st>>>(kp, n_steps);
        else if (bsm) k_step_policy<0, 0, 0, POLICY_LB_FIRST_FIT, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else k_step_policy<0, 0, 0, POLICY_LB_FIRST_FIT><<<g, th, sm, st>>>(kp, n_steps);
    } else if (policy == QRMSA_POLICY_HIGHEST_SNR) {
        if (ctx->hs2_grid)
            k_step_highest_snr_links<<<ctx->hs2_grid, ctx->hs2_epc * OBS_ENV_THREADS, ctx->hs2_smem, st>>>(kp, n_steps, ctx->hs2_epc, ctx->obs2_env_smem);
        else if (!ctx->cta_grid) { ctx->err = "highest-SNR policy needs more shared memory than the device offers"; return QRMSA_ERR_UNSUPPORTED; }
        else k_step_highest_snr<<<ctx->cta_grid, ctx->cta_epc * OBS_ENV_THREADS, ctx->cta_smem, st>>>(kp, n_steps, ctx->cta_epc, ctx->cta_env_smem);
    } else {
        CK(cudaMemsetAsync(kp.work, 0, 4, st));
        if (c320 && bsm) k_step_policy<320, 6, 5, POLICY_LOAD_BALANCING, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else if (c320) k_step_policy<320, 6, 5, POLICY_LOAD_BALANCING><<<g, th, sm, st>>>(kp, n_steps);
        else if (bsm) k_step_policy<0, 0, 0, POLICY_LOAD_BALANCING, 1><<<g, th, bsm, st>>>(kp, n_steps);
        else k_step_policy<0, 0, 0, POLICY_LOAD_BALANCING><<<g, th, sm, st>>>(kp, n_steps);
    }
    CK(cudaGetLastError());
    k_count_decisions<<<ctx->sm_count * 8, 256, 0, st>>>(kp);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

extern "C" int qrmsa_set_staging(qrmsa_ctx *ctx, int level) {
    if (!ctx || level < 0 || level > 2) return QRMSA_ERR_ARG;
    ctx->staging = level;
    return QRMSA_OK;
}

extern "C" int qrmsa_step_first_fit(qrmsa_ctx *ctx, int n_steps, void *stream) {
    return qrmsa_step_heuristic(ctx, QRMSA_POLICY_FIRST_FIT, n_steps, stream);
}

extern "C" int qrmsa_step_action(qrmsa_ctx *ctx, const int64_t *d_action, float *d_reward, uint8_t *d_status,
                                 double *d_gsnr, uint8_t *d_terminated, void *stream) {
    if (!ctx || !d_action) return QRMSA_ERR_ARG;
    if (ctx->kp.n_req < 2) { ctx->err = "no trace loaded"; return QRMSA_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    k_step_action<<<ctx->grid, ctx->threads, ctx->kp.blob_bytes, (cudaStream_t)stream>>>(
        ctx->kp, (const long long *)d_action, d_reward, d_status, d_gsnr, d_terminated, ctx->kp.n_req);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

extern "C" int qrmsa_observation_dims(const qrmsa_ctx *ctx, int *obs_dim, int *n_actions) {
    if (!ctx) return QRMSA_ERR_ARG;
    if (obs_dim) *obs_dim = 1 + 2 + ctx->kp.K + 12 * ctx->kp.K * ctx->kp.Mc;   // qrmsa.pyx:323-328
    if (n_actions) *n_actions = ctx->kp.K * ctx->kp.Mc * ctx->kp.S + 1;        // qrmsa.pyx:319-321
    return QRMSA_OK;
}

extern "C" int qrmsa_observation(qrmsa_ctx *ctx, float *d_obs, uint8_t *d_mask, void *stream) {
    if (!ctx || !d_obs || !d_mask) return QRMSA_ERR_ARG;
    if (ctx->kp.n_req < 1) { ctx->err = "no trace loaded"; return QRMSA_ERR_STATE; }
    if (!ctx->d_path_len_norm) { ctx->err = "path_length_km / link_length_km were not given to qrmsa_create"; return QRMSA_ERR_STATE; }
    if (!ctx->obs_grid && !ctx->obs2_grid) { ctx->err = "observation kernel needs more shared memory than the device offers"; return QRMSA_ERR_UNSUPPORTED; }
    if (ctx->kp.Mc < ctx->kp.M && !ctx->obs2_grid) {
        ctx->err = "modulations_to_consider < n_mods is built into the observation kernel for spectra up to 320 slots only";
        return QRMSA_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(ctx->device));
    int obs_dim = 0, n_actions = 0;
    qrmsa_observation_dims(ctx, &obs_dim, &n_actions);
    if (ctx->obs2_grid)
        k_observation_links<<<ctx->obs2_grid, ctx->obs2_epc * OBS_ENV_THREADS, ctx->obs2_smem, (cudaStream_t)stream>>>(
            ctx->kp, ctx->d_path_len_norm, ctx->inv_max_rate, d_obs, d_mask, obs_dim, n_actions, ctx->obs2_epc, ctx->obs2_env_smem);
    else
        k_observation<<<ctx->obs_grid, ctx->obs_epc * OBS_ENV_THREADS, ctx->obs_smem, (cudaStream_t)stream>>>(
            ctx->kp, ctx->d_path_len_norm, ctx->inv_max_rate, d_obs, d_mask, obs_dim, n_actions, ctx->obs_epc, ctx->obs_env_smem);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

extern "C" int qrmsa_sample_masked_actions(const void *d_logits, int logits_dtype, const uint8_t *d_mask, int n_envs,
                                           int n_actions, int64_t logit_row_stride, int64_t mask_row_stride, uint64_t seed,
                                           uint64_t step, int64_t *d_action, int device, void *stream) {
    if (!d_logits || !d_mask || !d_action || n_envs < 1 || n_actions < 1 || logit_row_stride < n_actions ||
        mask_row_stride < n_actions || (logits_dtype != QRMSA_LOGITS_F32 && logits_dtype != QRMSA_LOGITS_BF16))
        return QRMSA_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return QRMSA_ERR_NO_DEVICE;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return QRMSA_ERR_CUDA;
    const int rows_per_cta = SAMPLER_THREADS / 32;          // a warp per row
    const int want = (n_envs + rows_per_cta - 1) / rows_per_cta;
    const int grid = want < sms * 8 ? want : sms * 8;       // up to 8 CTAs of 256 threads per SM, rows strided over the warps
    cudaStream_t st = (cudaStream_t)stream;
    if (logits_dtype == QRMSA_LOGITS_BF16)
        k_sample_masked<__nv_bfloat16><<<grid, SAMPLER_THREADS, 0, st>>>((const __nv_bfloat16 *)d_logits, d_mask, n_envs, n_actions,
                                                                        logit_row_stride, mask_row_stride, seed, step, (long long *)d_action);
    else
        k_sample_masked<float><<<grid, SAMPLER_THREADS, 0, st>>>((const float *)d_logits, d_mask, n_envs, n_actions, logit_row_stride,
                                                                 mask_row_stride, seed, step, (long long *)d_action);
    return cudaGetLastError() == cudaSuccess ? QRMSA_OK : QRMSA_ERR_CUDA;
}

extern "C" int qrmsa_get_actions(qrmsa_ctx *ctx, int first, int count, int32_t *d_out, void *stream) {
    if (!ctx || !d_out || first < 0 || count < 0 || first + count > ctx->kp.n_req) return QRMSA_ERR_ARG;
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    dim3 grid((ctx->kp.n_envs + 31) / 32, (count + 31) / 32);
    k_gather_actions<<<grid, 256, 0, (cudaStream_t)stream>>>(ctx->kp, first, count, d_out);
    CK(cudaGetLastError());
    return QRMSA_OK;
}

extern "C" int qrmsa_get_actions_host(qrmsa_ctx *ctx, int first, int count, int32_t *h_out, void *stream) {
    if (!ctx || !h_out || first < 0 || count < 0 || first + count > ctx->kp.n_req) return QRMSA_ERR_ARG;
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)count * ctx->kp.n_envs * 4;
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    rc = qrmsa_get_actions(ctx, first, count, (int32_t *)ctx->stage, stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_out, ctx->stage, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return QRMSA_OK;
}

extern "C" int qrmsa_get_actions_host_strided(qrmsa_ctx *ctx, int first, int count, int32_t *h_out, int64_t row_stride,
                                              void *stream) {
    if (!ctx || !h_out || first < 0 || count < 0 || first + count > ctx->kp.n_req || row_stride < ctx->kp.n_envs)
        return QRMSA_ERR_ARG;
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t ne = (size_t)ctx->kp.n_envs;
    int rc = ensure_stage(ctx, std::max((size_t)count * ne * 4, (size_t)ctx->kp.T * ne * 11 + 4096));
    if (rc) return rc;
    rc = qrmsa_get_actions(ctx, first, count, (int32_t *)ctx->stage, stream);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(h_out, (size_t)row_stride * 4, ctx->stage, ne * 4, ne * 4, count, cudaMemcpyDeviceToHost,
                         (cudaStream_t)stream));
    return QRMSA_OK;
}

extern "C" int qrmsa_get_env_log_host(qrmsa_ctx *ctx, int env, int first, int count, uint32_t *h_records4) {
    if (!ctx || !h_records4 || env < 0 || env >= ctx->kp.n_envs || first < 0 || count < 0 || first + count > ctx->kp.n_req)
        return QRMSA_ERR_ARG;
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_records4, ctx->kp.trace + (size_t)env * ctx->kp.T + first, (size_t)count * sizeof(uint4), cudaMemcpyDeviceToHost));
    return QRMSA_OK;
}

static int get_qot_component(qrmsa_ctx *ctx, int first, int count, int comp, double *h_out, void *stream) {
    if (!ctx || !h_out || first < 0 || count < 0 || first + count > ctx->kp.n_req) return QRMSA_ERR_ARG;
    if (!ctx->kp.gsnr_log) { ctx->err = "GSNR log not enabled"; return QRMSA_ERR_STATE; }
    if (count == 0) return QRMSA_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)count * ctx->kp.n_envs * 8;
    int rc = ensure_stage(ctx, bytes);
    if (rc) return rc;
    k_gather_gsnr<<<ctx->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(ctx->kp, first, count, comp, (double *)ctx->stage);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_out, ctx->stage, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return QRMSA_OK;
}

extern "C" int qrmsa_get_gsnr_host(qrmsa_ctx *ctx, int first, int count, double *h_out, void *stream) {
    return get_qot_component(ctx, first, count, 0, h_out, stream);
}

extern "C" int qrmsa_get_ase_nli_host(qrmsa_ctx *ctx, int first, int count, double *h_ase, double *h_nli, void *stream) {
    int rc = get_qot_component(ctx, first, count, 1, h_ase, stream);
    return rc ? rc : get_qot_component(ctx, first, count, 2, h_nli, stream);
}

extern "C" int qrmsa_counters(qrmsa_ctx *ctx, int64_t *h_out, void *stream) {
    if (!ctx || !h_out) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(int64_t) * QRMSA_N_COUNTERS * ctx->n_groups;
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->kp.counters, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(h_out, ctx->h_counters, bytes);
    return QRMSA_OK;
}

extern "C" int qrmsa_counters_device(qrmsa_ctx *ctx, int64_t **d_out) {
    if (!ctx || !d_out) return QRMSA_ERR_ARG;
    *d_out = (int64_t *)ctx->kp.counters;
    return QRMSA_OK;
}

extern "C" int qrmsa_env_state_host(qrmsa_ctx *ctx, int32_t *h_out, void *stream) {
    if (!ctx || !h_out) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->kp.n_envs;
    std::vector<int4> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), ctx->kp.estate, sizeof(int4) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    for (int i = 0; i < n; i++) {
        h_out[4 * i + 0] = tmp[i].x;
        h_out[4 * i + 1] = tmp[i].z;
        h_out[4 * i + 2] = tmp[i].x - tmp[i].z;
        h_out[4 * i + 3] = tmp[i].w;
    }
    return QRMSA_OK;
}

extern "C" int qrmsa_export_bitmaps(qrmsa_ctx *ctx, int first, int count, uint32_t *h_out) {
    if (!ctx || !h_out || first < 0 || count < 0 || first + count > ctx->kp.n_envs) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    const KParams &kp = ctx->kp;
    // rows are RW words apart on the device (bitmap words + padding + the count word); the export is dense
    CK(cudaMemcpy2D(h_out, (size_t)kp.W * 4, kp.bm + (size_t)first * kp.bm_stride, (size_t)kp.RW * 4, (size_t)kp.W * 4,
                    (size_t)count * kp.E, cudaMemcpyDeviceToHost));
    return QRMSA_OK;
}

extern "C" int qrmsa_export_slots(qrmsa_ctx *ctx, int env, int32_t *h_slots) {
    if (!ctx || !h_slots || env < 0 || env >= ctx->kp.n_envs) return QRMSA_ERR_ARG;
    const KParams &kp = ctx->kp;
    std::vector<uint32_t> words((size_t)kp.E * kp.W);
    int rc = qrmsa_export_bitmaps(ctx, env, 1, words.data());
    if (rc) return rc;
    for (int l = 0; l < kp.E; l++)
        for (int s = 0; s < kp.S; s++) h_slots[(size_t)l * kp.S + s] = (words[(size_t)l * kp.W + (s >> 5)] >> (s & 31)) & 1u;
    return QRMSA_OK;
}

extern "C" int qrmsa_export_link_list(qrmsa_ctx *ctx, int env, int link, int32_t *h_out3, int cap, int *n) {
    if (!ctx || !h_out3 || !n || env < 0 || env >= ctx->kp.n_envs || link < 0 || link >= ctx->kp.E) return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    const KParams &kp = ctx->kp;
    uint32_t c = 0;
    CK(cudaMemcpy(&c, kp.bm + (size_t)env * kp.bm_stride + (size_t)link * kp.RW + kp.RW - 1, 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> recs(kp.CAP);
    CK(cudaMemcpy(recs.data(), kp.lists + ((size_t)env * kp.E + link) * kp.CAP, (size_t)kp.CAP * 4, cudaMemcpyDeviceToHost));
    *n = (int)c;
    for (int i = 0; i < (int)c && i < cap; i++) {
        const uint32_t r = recs[i];
        const int nn = (r >> 12) & 0xff;
        h_out3[3 * i + 0] = ((int)(r & 0xfff) - nn) / 2;
        h_out3[3 * i + 1] = nn;
        h_out3[3 * i + 2] = (r >> 20) & 7;
    }
    return QRMSA_OK;
}

extern "C" int qrmsa_probe_qot(qrmsa_ctx *ctx, int env, int src, int dst, int p, int initial_slot, int number_slots,
                               double *h_gsnr_ase_nli_db) {
    if (!ctx || !h_gsnr_ase_nli_db || env < 0 || env >= ctx->kp.n_envs) return QRMSA_ERR_ARG;
    const KParams &kp = ctx->kp;
    if (src < 0 || src >= kp.N || dst < 0 || dst >= kp.N || p < 0 || p >= kp.K || initial_slot < 0 ||
        number_slots < 1 || initial_slot + number_slots > kp.S)
        return QRMSA_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_stage(ctx, 256);
    if (rc) return rc;
    k_probe_gsnr<<<1, 32, kp.blob_bytes>>>(kp, env, src, dst, p, initial_slot, number_slots, (double *)ctx->stage);
    CK(cudaGetLastError());
    CK(cudaMemcpy(h_gsnr_ase_nli_db, ctx->stage, 24, cudaMemcpyDeviceToHost));
    return QRMSA_OK;
}

extern "C" int qrmsa_probe_gsnr(qrmsa_ctx *ctx, int env, int src, int dst, int p, int initial_slot, int number_slots,
                                double *h_gsnr_db) {
    double v[3];
    if (!h_gsnr_db) return QRMSA_ERR_ARG;
    int rc = qrmsa_probe_qot(ctx, env, src, dst, p, initial_slot, number_slots, v);
    if (!rc) *h_gsnr_db = v[0];
    return rc;
}
