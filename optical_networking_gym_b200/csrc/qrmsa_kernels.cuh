// qrmsa_kernels.cuh -- sm_100a device code of the batched QRMSA environment step.
//
// Execution model: ONE WARP PER ENVIRONMENT.  Lane j holds word j of a 320/640-slot availability
// bitmap, so the AND along a path, the contiguous-block search and the commit/release masks are
// single warp-wide instructions; the GN-model sum deals the path's channel records to the lanes in
// groups of four (one 16-byte load each).  A warp keeps its env for all n_steps of a launch, so the
// env's state is pulled from HBM once per launch and then lives on chip: in the SM_ = 1 ("BMS") variant of the step
// kernel the env's link rows (bitmaps + channel counts), the current and next 128-byte chunk of its request and
// schedule streams (cp.async) and a compact copy of the path table sit in shared memory beside the GN tables
// (227 KB on nobel-eu/320); the channel lists, the position table and the trace go through L1/L2.  SM_ = 2 keeps
// only the stream chunks there (640 slots: the tables alone take 183 KB), SM_ = 0 nothing.  No tensor cores: no
// stage is a dense contraction.
//
// HBM layout per env (all per-env blocks are contiguous, 64-byte aligned):
//   bm      uint32 [E][RW]    link rows: words 0..W-1 = packed slots, 1 = free (reference:
//                             graph["available_slots"], int32[E][S], envs/qrmsa.pyx:302-305); word RW-1 = number of
//                             channels on the link; RW = 16 (one 64-byte line) up to 479 slots, else 32
//   lists   uint32 [E][CAP]   channel records  c2 | n<<12 | mod<<20 | cls<<23   (reference:
//                             topology[u][v]["running_services"], qrmsa.pyx:1305-1306)
//                             c2 = 2*initial_slot + n  (centre frequency in half-slots); entries at or past the
//                             count hold the filler record (class NC: table rows of zeros)
//   pos     uint8/16 [CAP][E] index, in the link's list, of the channel that starts in slot pair s>>1 (pair-major:
//                             the entries of one service on the links of its path share a line)
//   trace   uint4  [T]        {arrival f32, holding f32, src|dst<<8|rate<<16, action word}
//                             request table == service table == decision log
//   perm    uint64 [T]        float32(arrival+holding) << 32 | request id, ascending: the release schedule
//                             (stands in for the heapq of qrmsa.pyx:1327-1330, :1113-1122)
//   estate  int4              {current request, release pointer, accepted, error}
//   counted uint32            requests already covered by k_count_decisions
//
// GN model (core/osnr.pyx:21-142), factorised: spans of a link are identical and all
// frequencies sit on the half-slot grid, so with d = |c2_r - c2| (half-slots) the neighbour term is
//     asinh(Q n_r (d + n_r)) - asinh(Q n_r (d - n_r))  =: G[class(n_r)][d],   Q = pi^2 |beta2| sb^2 / (4 alpha)
//     phi_r * bw_r / |df|                               =  PHIN[class, mod] * INV[d]
// and   1/GSNR = ASEC[n] * fc * PA[path] + CN[n] * (SELF[n] * PB[path] + sum_links (W1_l sum_r G - W2_l sum_r PHIN INV))
// The accept test gsnr_dB >= threshold is made on the linear value (acc <= 10^(-thr/10)); log10 is only
// evaluated for the optional GSNR log.
// Tables G/INV/PHIN/W1/W2/SELF/CN/ASEC are built on the host in FP64 with the same libm the reference
// uses and staged in shared memory with one TMA bulk copy per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qrmsa_b200.h"

namespace qrmsa {

constexpr unsigned FULL = 0xffffffffu;
#ifndef QRMSA_STEP_THREADS
#define QRMSA_STEP_THREADS 1024
#endif
constexpr int MAX_THREADS = QRMSA_STEP_THREADS;  // threads per CTA of the step kernels (one CTA per SM)

enum EnvError : int {
    ENV_OK = 0,
    ENV_ERR_LIST_OVERFLOW = 1,
    ENV_ERR_RELEASE_NOT_FOUND = 2,
    ENV_ERR_LOW_GSNR = 3,  // step_action: the reference raises ValueError (qrmsa.pyx:925-929)
};

struct KParams {
    // dimensions
    int n_envs, N, E, K, M, Mc, R, S, W, RW, Hmax, NC, D, CAP, T, n_req, group_size;
    uint32_t sentinel;          // list filler record: class NC (all-zero G row, PHIN 0) -> contributes exactly 0
    double mod_thr_nomargin[8];  // Modulation.minimum_osnr (no margin): the observation's threshold (qrmsa.pyx:743-758)
    double acct0[8], acct0_lo[8], acct0_hi[8];   // 10^(-minimum_osnr/10) and its +-1e-3 dB band: disruption / defragmentation test
    int feat, n_defrag;          // bit 0 measure_disruptions, bit 1 defragmentation; n_defrag_services (qrmsa.pyx:206-237)
    int *step_disrupted;         // [n_envs] services found disrupted by the last decided request (nullable)
    uint8_t *maxmod;             // [n_envs] max_modulation_idx: n_mods-1 after reset; with modulations_to_consider < n_mods the
                                 // observation moves it with every request (qrmsa.pyx:437, :543-581) and step() decodes with it
    int need_monotone;  // slots_needed never decreases as the modulation index falls (true for SE-sorted tables)
    // static tables in global memory
    const uint8_t *path_hops;   // [N*N*K]
    const uint8_t *path_links;  // [N*N*K*Hmax]
    const double2 *path_gn;     // [N*N*K] {PA, PB}
    const unsigned char *blob;  // shared-memory image, 16-byte multiple
    const unsigned char *ptab;  // compact path table for shared memory: u16 offset[n_paths] | u8 hops[n_paths] | u8 links[sum hops]
    int ptab_bytes, pt_hops_off, pt_links_off;   // (16-byte multiple; byte offsets of the second and third part)
    // BMS step kernel, byte offsets in dynamic shared memory: warp w owns [smem_warp_off + w * smem_warp_stride, ...):
    // 512 bytes of request / schedule chunks, then its env's link rows; the path table starts at smem_pt_off
    int smem_warp_off, smem_warp_stride, smem_pt_off, smem_pt_hops, smem_pt_links;
    int *work;                  // env ticket counter of the step kernel (zeroed before each launch)
    uint8_t *pos;               // [n_envs][E][CAP] (u8 if CAP <= 256, else u16): list position of the channel that starts in
                                // slot pair s>>1 of the link -- lets a release drop its record without searching
    int pos_bytes;              // 1 or 2
    size_t pos_stride;          // bytes per env
    uint32_t *counted;          // [n_envs] requests already covered by k_count_decisions (or counted in-kernel)
    int blob_bytes;
    double f0, sb;
    // per-env state
    uint32_t *bm;
    uint32_t *lists;
    uint4 *trace;
    unsigned long long *perm;      // [n_envs][T] release schedule: float32 release key << 32 | request id, ascending
    int4 *estate;
    unsigned long long *counters;  // [n_groups][QRMSA_N_COUNTERS]
    double *gsnr_log;              // nullable, [n_envs][T][3] = GSNR, ASE-only, NLI-only in dB (osnr.pyx:138-140)
    size_t bm_stride;              // uint32 words per env (= E * RW)
};

// Shared-memory image of the static tables (byte offsets).  The layout is FIXED (capacities, not sizes) so
// that every table access compiles to an LDS with an immediate offset from the dynamic-shared-memory base;
// the only size-dependent offset is G's, which follows INV[D].
namespace lay {
constexpr int PHIN = 0;      // double[256]  phi(mod) * bw(class), index (class << 3) | mod
constexpr int W1 = 2048;     // double[256]  n_spans * l_eff                       per link
constexpr int W2 = 4096;     // double[256]  -n_spans * l_eff * (5/3) * l_eff / L  per link (negated)
constexpr int SELF = 6144;   // double[32]   self-channel asinh term               per slot class
constexpr int CN = 6400;     // double[32]   NLI prefactor / P                     per slot class
constexpr int ASEC = 6656;   // double[32]   bw * h / P                            per slot class
constexpr int ACCT = 6912;   // double[8]    10^(-thr/10)                          per modulation
constexpr int ACCLO = 6976;  // double[8]    10^(-(thr+1e-3)/10)
constexpr int ACCHI = 7040;  // double[8]    10^(-(thr-1e-3)/10)
constexpr int NEED = 7104;   // u8[512]      slots needed  [rate*M + m]
constexpr int CLS = 7616;    // u8[512]      slot class    [rate*M + m]
constexpr int RATE = 8128;   // i32[256]     bit rate in Mb/s
constexpr int INV = 9152;    // double[D]    1 / (d * slot_bw / 2)
constexpr int MAX_RM = 512, MAX_E = 256, MAX_NC = 32, MAX_M = 8, MAX_R = 256;
__host__ __device__ constexpr int G(int D) { return INV + 8 * D; }  // double[NC][D]
}  // namespace lay

extern __shared__ __align__(128) unsigned char qsmem[];

// Table reads.  On sm_100a the address of a __shared__ symbol is formed from SR_CgaCtaId (shared::cluster window);
// left to the compiler that S2R + LEA sequence is re-materialised at most uses inside the step loop.  `sb` holds
// the window address of the dynamic shared block once per thread (volatile asm: not re-materialisable) and every
// read is an LDS with register + immediate addressing.
struct Tab {
    uint32_t sb;
    __device__ __forceinline__ void init() {
        asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sb) : "l"(qsmem));
    }
    template <int OFF>
    __device__ __forceinline__ double f64(int idx) const {
        double v;
        asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(sb + (uint32_t)idx * 8u), "n"(OFF));
        return v;
    }
    template <int OFF>
    __device__ __forceinline__ int u8(int idx) const {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(sb + (uint32_t)idx), "n"(OFF));
        return (int)v;
    }
    __device__ __forceinline__ double PHIN(int i) const { return f64<lay::PHIN>(i); }
    __device__ __forceinline__ double W1(int i) const { return f64<lay::W1>(i); }
    __device__ __forceinline__ double W2(int i) const { return f64<lay::W2>(i); }
    __device__ __forceinline__ double SELF(int i) const { return f64<lay::SELF>(i); }
    __device__ __forceinline__ double CN(int i) const { return f64<lay::CN>(i); }
    __device__ __forceinline__ double ASEC(int i) const { return f64<lay::ASEC>(i); }
    __device__ __forceinline__ double ACCT(int i) const { return f64<lay::ACCT>(i); }
    __device__ __forceinline__ double ACCLO(int i) const { return f64<lay::ACCLO>(i); }
    __device__ __forceinline__ double ACCHI(int i) const { return f64<lay::ACCHI>(i); }
    __device__ __forceinline__ double INV(int i) const { return f64<lay::INV>(i); }
    // G follows INV[D]: entry i of G is entry D + i of the INV-based array
    __device__ __forceinline__ double G(int D, int i) const { return f64<lay::INV>(D + i); }
    __device__ __forceinline__ int need(int i) const { return u8<lay::NEED>(i); }
    __device__ __forceinline__ int cls(int i) const { return u8<lay::CLS>(i); }
    __device__ __forceinline__ int rate(int i) const {
        int v;
        asm("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(sb + (uint32_t)i * 4u), "n"(lay::RATE));
        return v;
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage the table blob into shared memory with the TMA (1-D bulk copy, completion on an mbarrier).
__device__ __forceinline__ void stage_tables(const KParams &p, uint64_t *mbar) {
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(p.blob_bytes) : "memory");
        const int CH = 32768;
        for (int off = 0; off < p.blob_bytes; off += CH) {
            int sz = p.blob_bytes - off < CH ? p.blob_bytes - off : CH;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(qsmem + off)),
                "l"(p.blob + off), "r"(sz), "r"(bar)
                : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, q;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
    }
}

// Words per link row: bitmap words 0..W-1 (1 = free), word RW-1 = number of channels on the link.  16 words
// (one 64-byte line, 4 lanes x 16 B) up to 479 slots, 32 words (8 lanes x 16 B) up to 991: the virtual slot at
// index S must fall before the count word.
__host__ __device__ constexpr int row_words(int S) { return S <= 479 ? 16 : 32; }
// Link rows are reached through row_ld / row_st: a plain pointer (global memory), or RowsS -- the warp's copy in shared
// memory (BMS), addressed from ONE pinned 32-bit register with the area's offset as an immediate.
constexpr int WARP_STREAM_BYTES = 512;   // request / schedule chunks ahead of the rows in a warp's shared-memory area
struct RowsS {
    uint32_t wb;   // shared-window address of the warp's area
};
__device__ __forceinline__ uint32_t row_ld(const uint32_t *bm, unsigned i) { return bm[i]; }
__device__ __forceinline__ void row_st(uint32_t *bm, unsigned i, uint32_t v) { bm[i] = v; }
__device__ __forceinline__ uint32_t row_ld(const RowsS &r, unsigned i) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(r.wb + 4u * i), "n"(WARP_STREAM_BYTES));
    return v;
}
__device__ __forceinline__ void row_st(const RowsS &r, unsigned i, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(r.wb + 4u * i), "n"(WARP_STREAM_BYTES), "r"(v) : "memory");
}
// Set / clear bits of a row word.  Rows in global memory: a reduction (SASS REDG.E.OR / .AND) -- nothing comes back, so
// the warp does not wait an L2 round trip for the old word (germany50/640: 6.26e8 -> 6.49e8 env-steps/s); the L1 line is
// invalidated by the reduction, later row loads of the warp see the new word.  Rows in shared memory: load + store.
template <bool SET>
__device__ __forceinline__ void row_rmw(uint32_t *bm, unsigned i, uint32_t mask) {
    if (SET) atomicOr(bm + i, mask); else atomicAnd(bm + i, ~mask);
}
template <bool SET>
__device__ __forceinline__ void row_rmw(const RowsS &r, unsigned i, uint32_t mask) {
    const uint32_t w = row_ld(r, i);
    row_st(r, i, SET ? (w | mask) : (w & ~mask));
}
__device__ __forceinline__ unsigned cnt_index(int l, int RW) { return (unsigned)(l * RW + RW - 1); }
__device__ __forceinline__ uint32_t *cnt_word(uint32_t *bm, int l, int RW) { return bm + (unsigned)(l * RW + RW - 1); }
__device__ __forceinline__ const uint32_t *cnt_word(const uint32_t *bm, int l, int RW) { return bm + (unsigned)(l * RW + RW - 1); }

// pos[link][s >> 1] = index of the channel record in the link's list.  A service and its guard slot cover at least
// one aligned slot pair exclusively, so the pair index identifies the service on that link.
// Laid out [pair][link]: the entries of one service on the links of its path share a 128-byte line (E <= 128), so a
// commit or a release touches one line of the table instead of one per hop.
__device__ __forceinline__ unsigned pos_index(const KParams &p, int l, int pair) {
    return (unsigned)(pair * p.E + l);
}
template <class DM>
__device__ __forceinline__ void pos_store(const DM &dm, const KParams &p, uint8_t *pos, int l, int pair, int v) {
    if (dm.has_pos()) pos[pos_index(p, l, pair)] = (uint8_t)v;   // (no table above 511 slots: release_by_search)
}
__device__ __forceinline__ int rec_pair(uint32_t rec) { return (int)(((rec & 0xfffu) - ((rec >> 12) & 0xffu)) >> 2); }

// Compile-time problem dimensions (0 = take them from KParams at run time).  Fixing S/M/K removes the
// integer divisions of the action decode and most address arithmetic from the per-step instruction stream.
template <int S_, int M_, int K_>
struct Dim {
    const KParams &p;
    __device__ __forceinline__ explicit Dim(const KParams &kp) : p(kp) {}
    __device__ __forceinline__ int S() const { return S_ ? S_ : p.S; }
    __device__ __forceinline__ int W() const { return S_ ? (S_ + 31) / 32 : p.W; }
    __device__ __forceinline__ int RW() const { return S_ ? row_words(S_) : p.RW; }
    __device__ __forceinline__ int D() const { return S_ ? 2 * S_ : p.D; }
    __device__ __forceinline__ int CAP() const { return S_ ? ((S_ + 1) / 2 + 31) / 32 * 32 : p.CAP; }
    __device__ __forceinline__ int M() const { return M_ ? M_ : p.M; }
    __device__ __forceinline__ int K() const { return K_ ? K_ : p.K; }
    // is there a position table (qrmsa_create: list indices fit a byte)?  Known at compile time in the specialised kernels,
    // which then carry only one of the two release paths
    __device__ __forceinline__ bool has_pos() const { return S_ ? (((S_ + 1) / 2 + 31) / 32 * 32 <= 256) : p.pos_bytes != 0; }
};

// word `lane` of (x >> b), x a multi-word bitmap spread over the lanes.  Lanes past the data hold 0 and
// qrmsa_create guarantees (S>>5) + ((max_need+1)>>5) + 2 <= 32, so every shuffle source that matters is in
// range and an out-of-range source (own value of a lane past the data) is 0.
__device__ __forceinline__ uint32_t shr_multi(uint32_t x, int b) {
    const int q = b >> 5;
    const uint32_t lo = __shfl_down_sync(FULL, x, q);
    const uint32_t hi = __shfl_down_sync(FULL, x, q + 1);
    return __funnelshift_r(lo, hi, b);
}

// bits [s, e) of the multi-word bitmap that fall in word j
__device__ __forceinline__ uint32_t range_mask(int s, int e, int j) {
    const int lo = max(s - (j << 5), 0), hi = min(e - (j << 5), 32);
    if (hi <= lo) return 0u;
    const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return upto_hi & ~((1u << lo) - 1u);
}

__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Path lookups (hop count + flag, link ids).  PathTab<0>: the global tables, read through L1.  PathTab<1>: the
// compact copy the step kernel keeps in shared memory (offset u16 | hops u8 | link ids u8, variable length): the two
// dependent table loads of every path visit and of every release become LDS, and ~19 KB of hot lines leave L1.
// PathTab<2>: only the hop-count bytes in shared memory (configurations whose rows and link ids do not fit, e.g.
// germany50/640: 12 KB of the 28 KB the 228 KB carve-out leaves unused beside tables + stream chunks) -- the first of
// the two dependent loads is an LDS, the link ids stay one load through L1/L2 (6.56e8 -> 6.64e8 env-steps/s).
template <int MODE>
struct PathTab {
    uint32_t base;   // shared-window address of the dynamic shared block (MODE != 0); the table's parts sit at the
                     // KParams offsets smem_pt_off / smem_pt_hops / smem_pt_links
    __device__ __forceinline__ int hops_flags(const KParams &p, int path) const {
        if (MODE != 0) {
            uint32_t v;
            asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)p.smem_pt_hops + (uint32_t)path));
            return (int)v;
        }
        return __ldg(p.path_hops + path);
    }
    // link id of hop `lane` (0 for lanes past the path)
    __device__ __forceinline__ int link(const KParams &p, int path, int lane, int hops) const {
        if (MODE == 1) {
            uint32_t off, v = 0;
            asm("ld.shared.u16 %0, [%1];" : "=r"(off) : "r"(base + (uint32_t)p.smem_pt_off + 2u * (uint32_t)path));
            if (lane < hops) asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)p.smem_pt_links + off + (uint32_t)lane));
            return (int)v;
        }
        return lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
    }
};

// qrmsa.pyx:1482-1512: AND of the path's link rows; plus one virtual free slot at index S, which turns the
// guard-band rule of qrmsa.pyx:529-540 ("n slots if the run touches the spectrum end, else n+1") into
// "n+1 consecutive free slots".  32/W link rows are fetched per pass (3 at W=10), then folded with shuffles.
template <class BM> struct RowUnroll { static constexpr int U = 4; };   // global rows: four row loads in flight per pass
template <> struct RowUnroll<RowsS> { static constexpr int U = 1; };     // shared memory: latency is short, keep it lean
template <class DM, class BM>
__device__ __forceinline__ uint32_t path_available(const DM &dm, const BM bm, int hops, int mylink, int lane) {
    const int W = dm.W(), S = dm.S(), RW = dm.RW();
    const int G = 32 / W;            // link rows per pass
    const int grp = lane / W, j = lane - grp * W;
    constexpr int U = RowUnroll<BM>::U;
    uint32_t av = 0xffffffffu;
#pragma unroll 1
    for (int i0 = 0; i0 < hops; i0 += U * G) {
        // the U loads of a pass are independent: issued back to back, folded afterwards (at 640 slots a pass is one
        // row, and a serial AND would make every hop its own L2 round trip)
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * G + grp;
            const int l = __shfl_sync(FULL, mylink, i & 31);
            v[u] = (grp < G && i < hops) ? row_ld(bm, (unsigned)(l * RW + j)) : 0xffffffffu;  // mutable state: plain (coherent) load
        }
#pragma unroll
        for (int u = 0; u < U; ++u) av &= v[u];
    }
    for (int g = 1; g < G; ++g) av &= __shfl_down_sync(FULL, av, g * W);
    if (lane >= W) av = 0u;
    if (lane == (S >> 5)) av |= 1u << (S & 31);
    return av;
}

// core/osnr.pyx:21-142 in the factorised form above:  acc = 1/GSNR = ase + cn * (selfpb + x)
//   ase + cn*selfpb is the value in an EMPTY network; x >= 0 is the neighbour sum (every term is >= 0 on
//   "prunable" paths, verified on the host), so a modulation whose empty-network value already misses its
//   threshold by more than 1e-3 dB can be refused without walking the channel lists.
struct GnBase {
    double ase, cn, selfpb;
    __device__ __forceinline__ double empty() const { return ase + cn * selfpb; }
    __device__ __forceinline__ double with(double x) const { return ase + cn * (selfpb + x); }
};

__device__ __forceinline__ GnBase gn_base(const KParams &p, const Tab &t, const double2 pg, int s, int n, int ncls) {
    const double fc = p.f0 + (p.sb * (double)s) + (p.sb * ((double)n / 2.0));  // heuristics.py:948-951
    GnBase b;
    b.ase = t.ASEC(ncls) * fc * pg.x;
    b.cn = t.CN(ncls);
    b.selfpb = t.SELF(ncls) * pg.y;
    return b;
}
__device__ __forceinline__ GnBase gn_base(const KParams &p, const Tab &t, int path, int s, int n, int ncls) {
    return gn_base(p, t, __ldg(p.path_gn + path), s, n, ncls);
}

// one channel record against a candidate centred at c2 half-slots (core/osnr.pyx:64-94, table form).
// SKIP: the candidate is a service that is already in the lists (measure_disruptions, defragment); its own record
// (centre and width in `self_key`) is left out, as calculate_osnr skips the service's own id (osnr.pyx:64-66)
template <bool SKIP = false>
__device__ __forceinline__ void gn_term(const Tab &t, const int D, const uint32_t rec, const int c2, double &s1, double &s2,
                                        const uint32_t self_key = 0u) {
    if (SKIP && (rec & 0xfffffu) == self_key) return;
    const int d = abs((int)(rec & 0xfffu) - c2);
    // INV[d] and G[class][d] from one scaled index: a_inv = base + 8 d, a_g = a_inv + (class + 1) * 8 D
    const uint32_t a_inv = t.sb + 8u * (uint32_t)d;
    const uint32_t a_g = a_inv + ((rec >> 23) + 1u) * (8u * (uint32_t)D);
    double g, inv;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(g) : "r"(a_g), "n"(lay::INV));
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(inv) : "r"(a_inv), "n"(lay::INV));
    s1 += g;
    s2 = fma(t.PHIN(rec >> 20), inv, s2);
}

// sum over the path's links and every channel on them (same value on every lane).
// The path's lists are cut into groups of 4 records (one 16-byte load) and the groups of ALL links are dealt out
// to the lanes in one flat sequence: lane t of pass b takes group b + t, which belongs to the hop whose group range
// [start, end) contains it.  A typical path (4 hops x 28 channels = 30 groups) is one pass with every lane busy.
// Lists are padded with the zero-contribution filler record, so the last group of a link needs no bounds test.
template <class DM, bool SKIP = false>
__device__ __forceinline__ double gn_neighbours(const DM &dm, const Tab &t, const uint32_t *lists, int hops, int mylink, int mycnt,
                                                int c2, int lane, uint32_t &terms, const uint32_t self_key = 0u) {
    const int D = dm.D(), CAP = dm.CAP();
    const int cnt = lane < hops ? mycnt : 0;
    terms += (uint32_t)__reduce_add_sync(FULL, cnt);
    // groups of this lane's hop (an empty link still gets one group of fillers, so that every hop owns a distinct
    // start), then an inclusive prefix sum over the hops
    const int mine = lane < hops ? max((cnt + 3) >> 2, 1) : 0;
    int end = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, end, o);
        if (lane >= o) end += v;
    }
    const int total = __shfl_sync(FULL, end, 31);
    const int start = end - mine;
    double x = 0.0;
#pragma unroll 1
    for (int b = 0; b < total; b += 32) {
        // hop of group g = b + lane: number of hops that start at or before g, minus one.  The starts inside this
        // pass are marked in a 32-bit mask (OR over the hop lanes), the ones before it are counted by a ballot.
        const int rel = start - b;
        const uint32_t mark = __reduce_or_sync(FULL, (lane < hops && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
        const int before = __popc(__ballot_sync(FULL, lane < hops && rel < 0));
        const int hop = before + __popc(mark & (0xffffffffu >> (31 - lane))) - 1;
        const int l = __shfl_sync(FULL, mylink, hop & 31);
        const int off = b + lane - __shfl_sync(FULL, start, hop & 31);
        if (b + lane < total) {
            const uint4 v = *reinterpret_cast<const uint4 *>(lists + (unsigned)(l * CAP) + 4 * off);
            double s1 = 0.0, s2 = 0.0;
            gn_term<SKIP>(t, D, v.x, c2, s1, s2, self_key);
            gn_term<SKIP>(t, D, v.y, c2, s1, s2, self_key);
            gn_term<SKIP>(t, D, v.z, c2, s1, s2, self_key);
            gn_term<SKIP>(t, D, v.w, c2, s1, s2, self_key);
            x = fma(t.W1(l), s1, x);
            x = fma(t.W2(l), s2, x);  // W2 is stored negated
        }
    }
    return warp_sum(x);
}

// Set (release) or clear (commit) bits [s, e) on every link of the path: lane -> (hop = lane>>2, word = lane&3),
// a service spans at most 4 bitmap words (number_slots + guard <= 97).
template <bool SET, class DM, class BM>
__device__ __forceinline__ void update_bitmaps(const DM &dm, const BM bm, int hops, int mylink, int s, int e,
                                               int lane) {
    const int W = dm.W(), RW = dm.RW();
    const int w0 = s >> 5;
    const int k = lane & 3;
    const int j = w0 + k;
    const uint32_t mask = j < W ? range_mask(s, e, j) : 0u;
#pragma unroll 1
    for (int i0 = 0; i0 < hops; i0 += 8) {
        const int i = i0 + (lane >> 2);
        const int l = __shfl_sync(FULL, mylink, i & 31);
        if (i < hops && mask) row_rmw<SET>(bm, (unsigned)(l * RW + j), mask);
    }
}

// qrmsa.pyx:1288-1325 (the release key of :1327-1330 is implicit in the precomputed schedule)
template <class DM, class BM>
__device__ __forceinline__ int commit(const DM &dm, const KParams &p, const BM bm, uint32_t *lists, uint8_t *pos, int hops,
                                      int mylink, int mycnt, int s, int n, uint32_t rec, int lane) {
    int e = s + n;
    if (e < dm.S()) e += 1;
    update_bitmaps<false>(dm, bm, hops, mylink, s, e, lane);
    int err = 0;
    if (lane < hops) {
        if (mycnt >= dm.CAP()) {
            err = 1;
        } else {
            lists[(unsigned)(mylink * dm.CAP() + mycnt)] = rec;
            pos_store(dm, p, pos, mylink, s >> 1, mycnt);
            row_st(bm, cnt_index(mylink, dm.RW()), (uint32_t)(mycnt + 1));
        }
    }
    return __any_sync(FULL, err);
}

// LPH[hops] = (32 / hops) | ceil(65536 / (32 / hops)) << 8: lanes per hop when the 32 lanes are split evenly between the
// hops of a path, and the multiplier that turns lane / lph into a multiply-shift (exact for lane < 32).
__constant__ uint32_t LPH[33] = {
    0x80020u, 0x80020u, 0x100010u, 0x199a0au, 0x200008u, 0x2aab06u, 0x333405u, 0x400004u, 0x400004u, 0x555603u, 0x555603u,
    0x800002u, 0x800002u, 0x800002u, 0x800002u, 0x800002u, 0x800002u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u,
    0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u, 0x1000001u,
    0x1000001u, 0x1000001u};

// Release without a position table (spectra above 511 slots, where a list index no longer fits a byte and a two-byte
// table would be 56 KB per env on germany50/640 -- more than rows and hot list lines together, never resident in L2 and a
// DRAM round trip in the middle of every release).  The record is FOUND instead: the lanes are split evenly between the
// hops, the lanes of a hop read its list four records per 16-byte load -- lines the GN sums keep warm -- and compare
// centre | width; the lane that finds the record swap-removes it (the list's last record is fetched beside the search).
// Returns 1 unless every hop found its record.
template <class DM, class BM>
__device__ __forceinline__ int release_by_search(const DM &dm, const KParams &p, const BM bm, uint32_t *lists, int hops,
                                                 int mylink, int s, int n, int lane) {
    const int CAP = dm.CAP(), RW = dm.RW();
    const uint32_t target = (uint32_t)(2 * s + n) | ((uint32_t)n << 12);   // centre and width identify the record
    const int cown = lane < hops ? (int)row_ld(bm, cnt_index(mylink, RW)) : 0;
    const uint32_t lut = LPH[hops];
    const int lph = (int)(lut & 0xffu);
    const int hop = (int)(((uint32_t)lane * (lut >> 8)) >> 16);   // lane / lph
    const int sub = lane - hop * lph;
    const int l = __shfl_sync(FULL, mylink, hop & 31);
    const int ch = __shfl_sync(FULL, cown, hop & 31);
    const int c = hop < hops ? ch : 0;                             // lanes left over when 32 % hops != 0 have no hop
    const int groups = (c + 3) >> 2;
    const int gmax = __reduce_max_sync(FULL, groups);
    uint32_t *lst = lists + (unsigned)(l * CAP);
    const uint32_t last = c > 0 ? lst[c - 1] : 0u;
    update_bitmaps<true>(dm, bm, hops, mylink, s, min(s + n + 1, dm.S()), lane);
    int fpos = -1;
    auto look = [&](const uint4 v, int g) {
        if ((v.x & 0xfffffu) == target) fpos = 4 * g;
        if ((v.y & 0xfffffu) == target) fpos = 4 * g + 1;
        if ((v.z & 0xfffffu) == target) fpos = 4 * g + 2;
        if ((v.w & 0xfffffu) == target) fpos = 4 * g + 3;
    };
    // two groups per lane and pass, both loads issued before either is looked at (a link of up to 8 * lph records is one
    // round trip: 64 records on a 4-hop path)
#pragma unroll 1
    for (int b = 0; b < gmax; b += 2 * lph) {
        const int g0 = b + sub, g1 = g0 + lph;
        uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
        if (g0 < groups) v0 = *reinterpret_cast<const uint4 *>(lst + 4 * g0);
        if (g1 < groups) v1 = *reinterpret_cast<const uint4 *>(lst + 4 * g1);
        look(v0, g0);
        look(v1, g1);
    }
    const bool found = fpos >= 0 && fpos < c;   // (fillers past the count have width 0: they never match)
    if (found) {
        lst[fpos] = last;
        lst[c - 1] = p.sentinel;   // entries past the count are always the zero-contribution filler
        row_st(bm, cnt_index(l, RW), (uint32_t)(c - 1));
    }
    const int n_found = __popc(__ballot_sync(FULL, found));
    __syncwarp();
    return n_found != hops;
}

// qrmsa.pyx:1332-1350: free [s, s+n+1) (clamped at S) on every link of the path, drop the channel record.
// Lane i handles hop i: the record's place in the link's list comes from the position table, so there is no search.
template <class DM, class BM, class PT = PathTab<0>>
__device__ __forceinline__ int release_service(const DM &dm, const KParams &p, const Tab &t, const BM bm,
                                               uint32_t *lists, uint8_t *pos, const uint4 rq, int lane,
                                               const PT pt = PT()) {
    const int S = dm.S(), M = dm.M(), CAP = dm.CAP();
    const uint32_t a = rq.w & QRMSA_ACTION_MASK;
    const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
    // action = (pi * M + rel) * S + s: two divisions (by constants in the specialised kernels), remainders by multiply-subtract
    const uint32_t q = a / (uint32_t)S;
    const int s = (int)(a - q * (uint32_t)S);
    const int pi = (int)(q / (uint32_t)M);
    const int rel = (int)q - pi * M;
    const int m = (M - 1) - rel;
    const int n = t.need(rate * M + m);
    const int path = (src * p.N + dst) * dm.K() + pi;
    const int hops = pt.hops_flags(p, path) & 0x7f;
    const int mylink = pt.link(p, path, lane, hops);
    if (!dm.has_pos()) return release_by_search(dm, p, bm, lists, hops, mylink, s, n, lane);
#ifdef QRMSA_VALIDATE_RELEASE
    const uint32_t target = (uint32_t)(2 * s + n) | ((uint32_t)n << 12);   // centre and width identify the record
#endif
    // The position entry and the list's last record are fetched together (the count is at hand), the rows are updated
    // while the two loads are in flight, and the swap-remove needs nothing else: the chain of a release is
    // request record -> {position, last record} -> stores.  QRMSA_VALIDATE_RELEASE (debug builds) also reads the record
    // the position names and compares its centre and width before removing it -- one more dependent load.
    int err = 0, c = 0, fpos = 0;
    uint32_t last = 0u;
    uint32_t *lst = lists + (unsigned)(mylink * CAP);
    const unsigned cw = cnt_index(mylink, dm.RW());
    if (lane < hops) {
        c = (int)row_ld(bm, cw);
        const unsigned pidx = pos_index(p, mylink, s >> 1);
        fpos = (int)pos[pidx];
        last = lst[max(c - 1, 0)];
    }
    update_bitmaps<true>(dm, bm, hops, mylink, s, min(s + n + 1, S), lane);
    if (lane < hops) {
        bool ok = fpos < c;
#ifdef QRMSA_VALIDATE_RELEASE
        ok = ok && (lst[fpos] & 0xfffffu) == target;
#endif
        if (!ok) {
            err = 1;
        } else {
            lst[fpos] = last;
            lst[c - 1] = p.sentinel;   // entries past the count are always the zero-contribution filler
            pos_store(dm, p, pos, mylink, rec_pair(last), fpos);
            row_st(bm, cw, (uint32_t)(c - 1));
        }
    }
    err = __any_sync(FULL, err);
    __syncwarp();
    return err;
}

// The request records and the schedule entries of an env are read in order.  Streams<false>: straight from global
// memory.  Streams<true> (BMS step kernel): the current and the next 128-byte chunk of both streams sit in a per-warp
// area of shared memory -- 16 records, then 32 schedule entries -- filled one chunk ahead by cp.async, so that the
// per-request reads are LDS and the fill's latency is eight requests (sixteen releases) away from its first use.
template <bool RING>
struct Streams {
    uint32_t base;   // shared-window address of this warp's 512-byte area (RING only)
    static constexpr int BYTES = WARP_STREAM_BYTES;
    __device__ __forceinline__ void fill_tr(const uint4 *tr, int chunk, int lane, int T) const {
        if (RING) {
            if (lane < 8) {
                const int i = min(chunk * 8 + lane, T - 1);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(((chunk & 1) << 7) + (lane << 4))),
                             "l"(tr + i) : "memory");
            }
        }
    }
    __device__ __forceinline__ void fill_pm(const unsigned long long *perm, int chunk, int lane, int T) const {
        if (RING) {
            if (lane < 16) {
                const int i = min(chunk * 16 + lane, T - 1);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(base + 256u + (uint32_t)(((chunk & 1) << 7) + (lane << 3))),
                             "l"(perm + i) : "memory");
            }
        }
    }
    __device__ __forceinline__ void wait() const {
        if (RING) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
        }
    }
    __device__ __forceinline__ uint4 rec(const uint4 *tr, int i) const {
        if (RING) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(base + (uint32_t)((i & 15) << 4)) : "memory");
            return v;
        }
        return tr[i];
    }
    __device__ __forceinline__ uint32_t arrival_bits(const uint4 *tr, int i) const {
        if (RING) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)((i & 15) << 4)) : "memory");
            return v;
        }
        return tr[i].x;
    }
    __device__ __forceinline__ unsigned long long entry(const unsigned long long *perm, int i) const {
        if (RING) {
            unsigned long long v;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(base + 256u + (uint32_t)((i & 31) << 3)) : "memory");
            return v;
        }
        return perm[i];
    }
};

struct Head {
    int id;     // request id at the head of the release schedule, -1 = exhausted
    float rel;  // its release time, float32(arrival + holding)  (qrmsa.pyx:1329)
};

template <class ST = Streams<false>>
__device__ __forceinline__ Head load_head(const KParams &p, const uint4 *tr, const unsigned long long *perm, int ptr,
                                          const ST sm = ST()) {
    // one 8-byte load: the schedule entry carries the release key, so the head's request record is only read
    // when the service is actually released
    Head h;
    h.id = -1;
    h.rel = 0.f;
    if (ptr < p.n_req) {
        const unsigned long long k = sm.entry(perm, ptr);
        h.id = (int)(unsigned)k;
        h.rel = __uint_as_float((unsigned)(k >> 32));
        prefetch_l1(tr + h.id);   // the record is read when the service is released, usually a few requests later
    }
    return h;
}

// --------------------------------------------------------------------------------------------------------
// measure_disruptions and defragmentation (qrmsa.pyx:937-952, :1545-1639).  Both walk the RUNNING services, which the
// device does not keep as objects: a running service is a request record whose action word says "accepted" and whose
// release has not happened yet.  Rows, lists and the path table are read from global memory (general-dimension
// kernels only): the switches multiply the work per request by 10-100, as they do in the reference.
// --------------------------------------------------------------------------------------------------------
struct SvcView {   // a provisioned service decoded from its request record
    int path, hops, link, s, n, m, ncls;
    uint32_t key;   // its channel record's centre | width << 12
};
template <class DM>
__device__ __forceinline__ SvcView decode_service(const DM &dm, const KParams &p, const Tab &t, const uint4 rq, int lane) {
    const int S = dm.S(), M = dm.M();
    const uint32_t a = rq.w & QRMSA_ACTION_MASK;
    const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
    SvcView v;
    v.s = a % S;
    v.m = (M - 1) - (int)((a / S) % M);
    v.n = t.need(rate * M + v.m);
    v.ncls = t.cls(rate * M + v.m);
    v.path = (src * p.N + dst) * dm.K() + (int)(a / (S * M));
    v.hops = __ldg(p.path_hops + v.path) & 0x7f;
    v.link = lane < v.hops ? __ldg(p.path_links + v.path * p.Hmax + lane) : 0;
    v.key = (uint32_t)(2 * v.s + v.n) | ((uint32_t)v.n << 12);
    return v;
}

// 1/GSNR of service v placed at start slot s (its own slot, or a defragmentation candidate), its own record skipped
template <class DM>
__device__ __forceinline__ double acc_of_service(const DM &dm, const KParams &p, const Tab &t, const uint32_t *bm,
                                                 const uint32_t *lists, const SvcView &v, int s, int lane) {
    const int cnt = lane < v.hops ? (int)bm[cnt_index(v.link, dm.RW())] : 0;
    uint32_t terms = 0;
    const double x = gn_neighbours<DM, true>(dm, t, lists, v.hops, v.link, cnt, 2 * s + v.n, lane, terms, v.key);
    return gn_base(p, t, v.path, s, v.n, v.ncls).with(x);
}

// qrmsa.pyx:937-952, after request `cur` has been provisioned (its action word is written).  Running services = the
// accepted requests among the schedule entries still ahead of the release pointer; the ones that share a link with the
// new service and are not yet marked are re-evaluated.  Returns how many were found disrupted now.
template <class DM>
__device__ __forceinline__ int measure_disruptions(const DM &dm, const KParams &p, const Tab &t, uint4 *tr,
                                                   const unsigned long long *perm, const uint32_t *bm, const uint32_t *lists,
                                                   int cur, int rel_ptr, int lane, uint32_t &flags) {
    const SvcView me = decode_service(dm, p, t, tr[cur], lane);
    int local = 0;
    for (int i0 = rel_ptr; i0 < p.n_req; i0 += 32) {
        const int i = i0 + lane;
        int id = -1;
        uint32_t w = 0u;
        if (i < p.n_req) {
            id = (int)(unsigned)perm[i];
            if (id <= cur) w = tr[id].w;
        }
        const bool cand = id >= 0 && id <= cur && (w & QRMSA_FLAG_ACCEPTED) &&
                          !(w & (QRMSA_FLAG_DISRUPTED | QRMSA_FLAG_RELEASE_CANCELLED));
        unsigned todo = __ballot_sync(FULL, cand);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int xid = __shfl_sync(FULL, id, sl);
            const uint4 rq = tr[xid];
            const SvcView v = decode_service(dm, p, t, rq, lane);
            bool share = false;
            for (int h = 0; h < me.hops; ++h) {
                const int l = __shfl_sync(FULL, me.link, h);
                share |= (lane < v.hops && v.link == l);
            }
            if (!__any_sync(FULL, share)) continue;
            const double acc = acc_of_service(dm, p, t, bm, lists, v, v.s, lane);
            if (acc > p.acct0_lo[v.m] && acc < p.acct0_hi[v.m]) flags |= QRMSA_FLAG_NEAR_THRESHOLD;
            if (acc > p.acct0[v.m]) {   // osnr < minimum_osnr
                local += 1;
                if (lane == 0) tr[xid].w = rq.w | QRMSA_FLAG_DISRUPTED;
            }
        }
    }
    __syncwarp();
    return local;
}

// qrmsa.pyx:1545-1639, called after the service with schedule key (rel_bits, rel_id) has been released.  The services
// still in the reference's heap are the accepted requests whose key sorts after it; they are visited in provisioning
// (= request) order.  Returns the number of services moved; flags receives the near-threshold mark.
template <class DM>
__device__ __forceinline__ int defragment(const DM &dm, const KParams &p, const Tab &t, uint4 *tr, uint32_t *bm, uint32_t *lists,
                                          uint8_t *pos, int cur, uint32_t rel_bits, int rel_id, int lane, uint32_t &flags) {
    const int S = dm.S();
    const int limit = p.n_defrag ? p.n_defrag : 1000000;
    int moved = 0, err = 0;
    for (int i0 = 0; i0 < cur && moved < limit; i0 += 32) {
        const int id = i0 + lane;
        uint4 rq = make_uint4(0u, 0u, 0u, 0u);
        bool act = false;
        if (id < cur) {
            rq = tr[id];
            if ((rq.w & (QRMSA_FLAG_ACCEPTED | QRMSA_FLAG_RELEASE_CANCELLED)) == QRMSA_FLAG_ACCEPTED) {
                const uint32_t key = __float_as_uint(__fadd_rn(__uint_as_float(rq.x), __uint_as_float(rq.y)));
                act = key > rel_bits || (key == rel_bits && id > rel_id);
            }
        }
        unsigned todo = __ballot_sync(FULL, act);
        while (todo && moved < limit) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1u;
            uint4 rx;
            rx.x = __shfl_sync(FULL, rq.x, sl); rx.y = __shfl_sync(FULL, rq.y, sl);
            rx.z = __shfl_sync(FULL, rq.z, sl); rx.w = __shfl_sync(FULL, rq.w, sl);
            const int xid = i0 + sl;
            const SvcView v = decode_service(dm, p, t, rx, lane);
            // valid starts for v.n slots on its path (the service itself still occupies its slots), below its own start
            uint32_t r = path_available(dm, bm, v.hops, v.link, lane);
            for (int a = 1, L = v.n + 1; a < L;) { const int b = min(a, L - a); r &= shr_multi(r, b); a += b; }
            r &= range_mask(0, v.s, lane);
            for (;;) {
                const unsigned any = __ballot_sync(FULL, r != 0u);
                if (!any) break;
                const int fl = __ffs(any) - 1;
                const uint32_t w = __shfl_sync(FULL, r, fl);
                const int c = (fl << 5) + __ffs(w) - 1;
                if (lane == fl) r &= r - 1u;
                const double acc = acc_of_service(dm, p, t, bm, lists, v, c, lane);
                if (acc > p.acct0_lo[v.m] && acc < p.acct0_hi[v.m]) flags |= QRMSA_FLAG_NEAR_THRESHOLD;
                if (acc > p.acct0[v.m]) continue;   // osnr < minimum_osnr: next candidate (qrmsa.pyx:1600-1604)
                // move: free the old slots and drop the record, then provision at c
                err |= release_service(dm, p, t, bm, lists, pos, rx, lane, PathTab<0>());
                const int cnt = lane < v.hops ? (int)bm[cnt_index(v.link, dm.RW())] : 0;
                const uint32_t rec = (uint32_t)(2 * c + v.n) | ((uint32_t)v.n << 12) | ((uint32_t)v.m << 20) | ((uint32_t)v.ncls << 23);
                err |= commit(dm, p, bm, lists, pos, v.hops, v.link, cnt, c, v.n, rec, lane);
                if (lane == 0) tr[xid].w = rx.w - (uint32_t)v.s + (uint32_t)c;   // the slot is the action's last digit
                __syncwarp();
                moved += 1;
                break;
            }
        }
    }
    return err ? -1 : moved;
}

// qrmsa.pyx:1067-1122 after a request has been decided: take the next request (clock := its arrival) and
// release every accepted service whose key is <= now.  Entries of not-yet-decided requests block the
// schedule exactly as they are absent from the reference heap.
struct DefragStats { uint32_t cycles, moved, flags; };
template <class DM, class BM, class PT = PathTab<0>, class ST = Streams<false>, int FEAT = 0>
__device__ __forceinline__ int advance_and_release(const DM &dm, const KParams &p, const Tab &t, uint4 *tr,
                                                   const unsigned long long *perm, const BM bm, uint32_t *lists, uint8_t *pos,
                                                   int &cur, int &rel_ptr, Head &head, int lane, uint32_t &n_rel,
                                                   const PT pt = PT(), const ST sm = ST(), DefragStats *ds = nullptr) {
    cur += 1;
    if ((cur & 7) == 0) {   // (no-ops without the ring) the chunk entered was requested eight requests ago
        sm.wait();
        sm.fill_tr(tr, (cur >> 3) + 1, lane, p.T);
    }
    const float now = __uint_as_float(sm.arrival_bits(tr, cur));
    int err = 0;
    while (head.id >= 0 && head.id < cur && head.rel <= now) {
        // (endpoints and action word in one 8-byte load: taken apart, the compiler fetches .z only after the test on .w)
        const uint2 zw = *reinterpret_cast<const uint2 *>(&tr[head.id].z);
        const uint4 rq = make_uint4(0u, 0u, zw.x, zw.y);
        if ((rq.w & (QRMSA_FLAG_ACCEPTED | QRMSA_FLAG_RELEASE_CANCELLED)) == QRMSA_FLAG_ACCEPTED) {
            err |= release_service(dm, p, t, bm, lists, pos, rq, lane, pt);
            n_rel += 1;
            if constexpr ((FEAT & 2) != 0) {
                if (p.n_defrag == 0 || (cur + 1) % p.n_defrag == 0) {   // qrmsa.pyx:1117-1119
                    const int mv = defragment(dm, p, t, tr, bm, lists, pos, cur, __float_as_uint(head.rel), head.id, lane, ds->flags);
                    ds->cycles += 1;
                    if (mv < 0) err |= 1; else ds->moved += (uint32_t)mv;
                }
            }
        }
        rel_ptr += 1;
        if ((rel_ptr & 15) == 0) {
            sm.wait();
            sm.fill_pm(perm, (rel_ptr >> 4) + 1, lane, p.T);
        }
        head = load_head(p, tr, perm, rel_ptr, sm);
    }
    return err;
}

#define QCNT(slot, v) cnt_reg += (lane == (slot)) ? (uint32_t)(v) : 0u

// One QoT-checked candidate: accept iff gsnr >= threshold (heuristics.py:957-958), decided on the linear
// value acc = 1/GSNR against ACCT[m] = 10^(-thr/10); |gsnr - thr| < 1e-3 dB <=> ACCLO[m] < acc < ACCHI[m].
__device__ __forceinline__ bool qot_ok(const Tab &t, int m, double acc, uint32_t &flags) {
    if (acc > t.ACCLO(m) && acc < t.ACCHI(m)) flags |= QRMSA_FLAG_NEAR_THRESHOLD;
    return acc <= t.ACCT(m);
}

// --------------------------------------------------------------------------------------------------------
// Fused heuristic + step, n_steps requests per env per launch (qrmsa.pyx:838-1065 for the step).
//   POLICY_FIRST_FIT       heuristic_shortest_available_path_first_fit_best_modulation (heuristics.py:923-966;
//                          shortest_available_path_lowest_spectrum_best_modulation :431-490 decides identically)
//   POLICY_LOAD_BALANCING  load_balancing_best_modulation (heuristics.py:547-627): among the k paths, the one
//                          with the lowest (occupied slots of the path availability) / hops that admits a
//                          modulation; best modulation + first-fit slot on it
//   POLICY_LB_FIRST_FIT    heuristic_load_balancing_first_fit (heuristics.py:202-270): the k paths ordered by the
//                          occupied fraction of their availability (ties: path index), then first-fit as above on
//                          the first path that admits a modulation
// --------------------------------------------------------------------------------------------------------
enum { POLICY_FIRST_FIT = 0, POLICY_LOAD_BALANCING = 1, POLICY_LB_FIRST_FIT = 3 };   // ids of include/qrmsa_b200.h

template <bool BMS>
struct RowHandle {
    typedef uint32_t *type;
    static __device__ __forceinline__ type make(uint32_t *g, uint32_t) { return g; }
};
template <>
struct RowHandle<true> {
    typedef RowsS type;
    static __device__ __forceinline__ type make(uint32_t *, uint32_t wb) { RowsS r; r.wb = wb; return r; }
};

// SM_: what the kernel keeps in shared memory beside the tables.  0 = nothing; 1 ("BMS") = per warp the env's link rows and
// the request / schedule stream chunks, plus the compact path table; 2 = the stream chunks only (configurations whose
// rows do not fit, e.g. 640 slots: the tables alone take 183 KB).
// FEAT: bit 0 measure_disruptions, bit 1 defragmentation (general-dimension, unstaged kernel only)
template <int S_, int M_, int K_, int POLICY, int SM_ = 0, int FEAT = 0>
__global__ void __launch_bounds__(MAX_THREADS, 1) k_step_policy(const KParams p, const int n_steps) {
    constexpr bool BMS = SM_ == 1, RING = SM_ != 0;
    static_assert(FEAT == 0 || (SM_ == 0 && POLICY == POLICY_FIRST_FIT), "the feature hooks read rows and lists from global memory");
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<S_, M_, K_> dm(p);
    const int S = dm.S(), M = dm.M(), K = dm.K();

    const int lane = threadIdx.x & 31;
    const int reject = K * M * S;
    // BMS: shared memory = tables | per-warp areas (stream chunks, then the env's link rows) | compact path table.
    // Everything is addressed from two pinned 32-bit registers (t.sb, wb) plus KParams offsets, which reach the
    // instructions as constant-bank operands -- no per-access re-derivation of warp index times stride.
    PathTab<SM_> pt;
    pt.base = t.sb;
    uint32_t wb = 0;
    if (BMS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.ptab);
        uint4 *dst = reinterpret_cast<uint4 *>(qsmem + p.smem_pt_off);
        for (int i = threadIdx.x; i < (p.ptab_bytes >> 4); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    if (SM_ == 2) {   // the hop-count bytes of the path table (PathTab<2>)
        const int n_paths = p.N * p.N * p.K;
        for (int i = threadIdx.x; i < n_paths; i += blockDim.x) qsmem[p.smem_pt_hops + i] = p.path_hops[i];
        __syncthreads();
    }
    if (RING) {
        wb = t.sb + (uint32_t)p.smem_warp_off + (uint32_t)((threadIdx.x >> 5) * p.smem_warp_stride);
        asm volatile("" : "+r"(wb));
    }
    Streams<RING> sm;   // request / schedule chunks of this warp's env: the first 512 bytes of the warp's area
    sm.base = wb;

    // envs are handed out by a ticket counter (zeroed before the launch): a warp that finishes early takes the next
    // env instead of idling behind a fixed share
    for (;;) {
        int env = 0;
        if (lane == 0) env = atomicAdd(p.work, 1);
        // broadcast by a REDUX: its result lives in a uniform register, so the compiler knows the env index (and
        // every per-env base address derived from it) is warp-uniform and keeps that arithmetic on the uniform datapath
        env = (int)__reduce_add_sync(FULL, (unsigned)env);
        if (env >= p.n_envs) break;
        int4 st = p.estate[env];
        if (st.w != ENV_OK) continue;
        int cur = st.x, rel_ptr = st.y, err = 0;
        uint4 *tr = p.trace + (size_t)env * p.T;
        const unsigned long long *perm = p.perm + (size_t)env * p.T;
        uint32_t *bm_global = p.bm + (size_t)env * p.bm_stride;
        typename RowHandle<BMS>::type bm = RowHandle<BMS>::make(bm_global, wb);
        if (BMS) {
            // BMS: the env's link rows (bitmaps + channel counts) live in shared memory for the whole launch -- every
            // row read, commit and release is an LDS/STS instead of a trip to L1/L2; 16-byte copies in and out
            const uint4 *src = reinterpret_cast<const uint4 *>(bm_global);
            for (int i = lane; i < (int)(p.bm_stride >> 2); i += 32) {
                const uint4 v = src[i];
                asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" ::"r"(wb + 16u * (uint32_t)i), "n"(WARP_STREAM_BYTES),
                             "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            __syncwarp();
        }
        uint32_t *lists = p.lists + (size_t)env * p.E * dm.CAP();
        uint8_t *pos = p.pos + (size_t)env * p.pos_stride;
        // the per-env base addresses are made opaque so that they stay in registers: left alone, the compiler
        // re-derives env * stride (two IMADs, an IMAD.WIDE, LEA + LEA.HI.X and the constant loads) at most accesses
        asm volatile("" : "+l"(tr));
        asm volatile("" : "+l"(perm));
        if (!BMS) asm volatile("" : "+l"(bm_global));
        asm volatile("" : "+l"(lists));
        asm volatile("" : "+l"(pos));
        __builtin_assume(__isGlobal(tr));
        __builtin_assume(__isGlobal(perm));
        if (!BMS) __builtin_assume(__isGlobal(bm_global));
        if (!BMS) bm = RowHandle<BMS>::make(bm_global, wb);
        __builtin_assume(__isGlobal(lists));
        __builtin_assume(__isGlobal(pos));
        sm.fill_tr(tr, cur >> 3, lane, p.T);
        sm.fill_tr(tr, (cur >> 3) + 1, lane, p.T);
        sm.fill_pm(perm, rel_ptr >> 4, lane, p.T);
        sm.fill_pm(perm, (rel_ptr >> 4) + 1, lane, p.T);
        sm.wait();
        Head head = load_head(p, tr, perm, rel_ptr, sm);
        uint32_t cnt_reg = 0;

        const int cur_end = min(cur + n_steps, p.n_req - 1);   // the last request of a trace is never decided
#pragma unroll 1
        while (cur < cur_end && !err) {
            const uint4 rq = sm.rec(tr, cur);
            const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
            const int pbase = (src * p.N + dst) * K;
            // lane m holds (slots needed, slot class) of modulation m for this request's bit rate
            const int mynd = lane < M ? (t.need(rate * M + lane) | (t.cls(rate * M + lane) << 8)) : 0;
            uint32_t flags = QRMSA_FLAG_DECIDED;
            int action = reject;
            double acc_ok = 1.0, ase_ok = 0.5;
            int blk_res = 0, blk_osnr = 0;
            bool found = false;
            // load balancing: best candidate so far (committed after all paths have been looked at)
            double lowest_load = 1e300;
            int best_pi = -1, best_m = 0, best_s = 0;

            // load-balancing first fit: lane j < K learns the rank of path j by (occupied slots, index)
            int my_rank = lane;
            if (POLICY == POLICY_LB_FIRST_FIT) {
                int my_occ = 0x7fffffff;   // a missing path sorts last and is skipped below
#pragma unroll 1
                for (int pj = 0; pj < K; ++pj) {
                    const int path = pbase + pj;
                    const int hops = pt.hops_flags(p, path) & 0x7f;
                    if (hops == 0) continue;
                    const int mylink = pt.link(p, path, lane, hops);
                    const uint32_t av = path_available(dm, bm, hops, mylink, lane);
                    int free_slots = __popc(lane == (S >> 5) ? av & ~(1u << (S & 31)) : av);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) free_slots += __shfl_xor_sync(FULL, free_slots, o);
                    if (lane == pj) my_occ = S - free_slots;   // np.sum(available == 0) (heuristics.py:222-224)
                    QCNT(QRMSA_CNT_LINKS_READ, hops);
                }
                my_rank = 0;
#pragma unroll 1
                for (int pj = 0; pj < K; ++pj) {
                    const int o = __shfl_sync(FULL, my_occ, pj);
                    my_rank += (o < my_occ || (o == my_occ && pj < lane)) ? 1 : 0;
                }
            }
#pragma unroll 1
            for (int kk = 0; kk < K && !(POLICY != POLICY_LOAD_BALANCING && found); ++kk) {
                // path visited k-th: its own index, or the path of rank k
                const int pi = POLICY == POLICY_LB_FIRST_FIT ? __ffs(__ballot_sync(FULL, lane < K && my_rank == kk)) - 1 : kk;
                const int path = pbase + pi;
                const int hp = pt.hops_flags(p, path);  // bit 7: every neighbour term of this path is >= 0
                const int hops = hp & 0x7f;
                if (hops == 0) continue;
                const bool prunable = (hp & 0x80) != 0;
                const int mylink = pt.link(p, path, lane, hops);
                const int mycnt = lane < hops ? (int)row_ld(bm, cnt_index(mylink, dm.RW())) : 0;
                // needed by the GN sum below; with the rows in shared memory L1 is down to 28 KB and a prefetched line
                // does not survive until its use (measured: 9.15e8 with, 9.29e8 without)
                if (!BMS && lane < hops) prefetch_l1(lists + (unsigned)(mylink * dm.CAP()));

                const uint32_t av = path_available(dm, bm, hops, mylink, lane);
                QCNT(QRMSA_CNT_LINKS_READ, hops);
                QCNT(QRMSA_CNT_PATHS_TRIED, 1);
                double path_load = 0.0;
                if (POLICY == POLICY_LOAD_BALANCING) {
                    // np.sum(available == 0) / len(path.links)  (heuristics.py:569-573); the virtual slot is not a slot
                    int free_slots = __popc(lane == (S >> 5) ? av & ~(1u << (S & 31)) : av);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) free_slots += __shfl_xor_sync(FULL, free_slots, o);
                    path_load = (double)(S - free_slots) / (double)hops;
                    if (path_load >= lowest_load) continue;
                }
                uint32_t r = av;
                int a = 1;
                bool counted = false;
                // Modulations that need the same number of slots share the candidate (same first-fit start, same
                // centre frequency, same bandwidth), hence the same GSNR: the search and the GN sum are done once per
                // RUN of such modulations, and the run's threshold tests are made side by side -- lane m compares against
                // the thresholds of modulation m, ballots say which modulations are hopeless in an empty network, which
                // pass, and which sit within 1e-3 dB.  The reference's loop order (best modulation first, stop at the
                // first that passes: heuristics.py:938-958) is read off the masks.
                uint32_t todo = (1u << M) - 1u;   // modulations not looked at yet (bit m)
                const int my_need = mynd & 0xff;
#pragma unroll 1
                while (todo) {
                    const int m_hi = 31 - __clz((int)todo);
                    const int nd = __shfl_sync(FULL, mynd, m_hi);
                    const int n = nd & 0xff, ncls = nd >> 8;
                    // the run: m_hi and the modulations right below it that need n slots too
                    const uint32_t same = __ballot_sync(FULL, lane < M && my_need == n);
                    const uint32_t gap = ~same & ((1u << m_hi) - 1u);
                    const uint32_t run = gap ? (same & todo & ~((2u << (31 - __clz((int)gap))) - 1u)) : (same & todo);
                    todo &= ~run;
                    const int L = n + 1;
                    if (L < a) { r = av; a = 1; }
                    while (a < L) {
                        const int b = min(a, L - a);
                        r &= shr_multi(r, b);
                        a += b;
                    }
                    const unsigned any = __ballot_sync(FULL, r != 0u);
                    if (!any) {
                        // no block of n+1 slots: the reference sets blocked_due_to_resources and tries the next
                        // modulation (heuristics.py:938-940); when the remaining ones all need >= n slots none of
                        // them can fit either, so the loop ends here with the same flags
                        blk_res = 1;
                        if (p.need_monotone) break;
                        continue;
                    }
                    const int fl = __ffs(any) - 1;
                    const uint32_t w = __shfl_sync(FULL, r, fl);
                    const int s = (fl << 5) + __ffs(w) - 1;
                    const GnBase gb = gn_base(p, t, path, s, n, ncls);
                    const bool mine = (run >> lane) & 1u;
                    // hopeless even in an empty network (only on paths where every neighbour term is >= 0)
                    const uint32_t dead = prunable ? __ballot_sync(FULL, mine && gb.empty() >= t.ACCHI(lane & 7)) : 0u;
                    const uint32_t live = run & ~dead;
                    if (!live) {
                        QCNT(QRMSA_CNT_GN_PRUNED, __popc(dead));
                        blk_osnr = 1;
                        if (POLICY == POLICY_FIRST_FIT) blk_res = 0;
                        continue;
                    }
                    uint32_t terms = 0;
                    const double x = gn_neighbours(dm, t, lists, hops, mylink, mycnt, 2 * s + n, lane, terms);
                    QCNT(QRMSA_CNT_GN_TERMS, terms);
                    if (!counted) { QCNT(QRMSA_CNT_RECORDS_READ, terms); counted = true; }
                    const double acc = gb.with(x);
                    // accept iff gsnr >= threshold (heuristics.py:957-958), decided on the linear value acc = 1/GSNR against
                    // ACCT[m] = 10^(-thr/10); |gsnr - thr| < 1e-3 dB <=> ACCLO[m] < acc < ACCHI[m]
                    const bool alive = (live >> lane) & 1u;
                    const uint32_t okm = __ballot_sync(FULL, alive && acc <= t.ACCT(lane & 7));
                    const uint32_t nearm = __ballot_sync(FULL, alive && acc > t.ACCLO(lane & 7) && acc < t.ACCHI(lane & 7));
                    // checked: every live modulation down to the first that passes; pruned: the dead ones above it
                    const uint32_t upto = okm ? ~((1u << (31 - __clz((int)okm))) - 1u) : 0xffffffffu;
                    QCNT(QRMSA_CNT_GN_EVALS, __popc(live & upto));
                    QCNT(QRMSA_CNT_GN_PRUNED, __popc(dead & upto));
                    if (nearm & live & upto) flags |= QRMSA_FLAG_NEAR_THRESHOLD;
                    if (okm) {
                        const int m = 31 - __clz((int)okm);
                        found = true;
                        acc_ok = acc;
                        ase_ok = gb.ase;
                        if (POLICY != POLICY_LOAD_BALANCING) {
                            action = pi * M * S + ((M - 1) - m) * S + s;
                            const uint32_t rec = (uint32_t)(2 * s + n) | ((uint32_t)n << 12) | ((uint32_t)m << 20) |
                                                 ((uint32_t)ncls << 23);
                            if (commit(dm, p, bm, lists, pos, hops, mylink, mycnt, s, n, rec, lane)) err = ENV_ERR_LIST_OVERFLOW;
                        } else {
                            lowest_load = path_load;
                            best_pi = pi; best_m = m; best_s = s;
                        }
                        break;
                    }
                    blk_osnr = 1;
                    if (POLICY == POLICY_FIRST_FIT) blk_res = 0;
                }
            }
            if (POLICY == POLICY_LOAD_BALANCING && best_pi >= 0) {
                const int path = pbase + best_pi;
                const int hops = pt.hops_flags(p, path) & 0x7f;
                const int mylink = pt.link(p, path, lane, hops);
                const int mycnt = lane < hops ? (int)row_ld(bm, cnt_index(mylink, dm.RW())) : 0;
                const int nd = __shfl_sync(FULL, mynd, best_m);
                const int n = nd & 0xff, ncls = nd >> 8;
                action = best_pi * M * S + ((M - 1) - best_m) * S + best_s;
                const uint32_t rec = (uint32_t)(2 * best_s + n) | ((uint32_t)n << 12) | ((uint32_t)best_m << 20) |
                                     ((uint32_t)ncls << 23);
                if (commit(dm, p, bm, lists, pos, hops, mylink, mycnt, best_s, n, rec, lane)) err = ENV_ERR_LIST_OVERFLOW;
            }
            // decided / accepted / rejected / bit rates / hops / modulation histogram / flags are counted from the
            // decision log by k_count_decisions after the launch (it also maintains the env's accepted total)
            if (found) {
                flags |= QRMSA_FLAG_ACCEPTED;
            } else {
                if (POLICY == POLICY_LOAD_BALANCING && blk_osnr) blk_res = 0;   // heuristics.py:624-626
                if (POLICY == POLICY_LB_FIRST_FIT) { blk_res = 1; blk_osnr = 0; }   // heuristics.py:270
                if (blk_res) flags |= QRMSA_FLAG_BLOCKED_RESOURCES;
                if (blk_osnr) flags |= QRMSA_FLAG_BLOCKED_OSNR;
            }
            if (lane == 0) {
                tr[cur].w = (uint32_t)action | flags;
                if (p.gsnr_log) {   // 10*log10(1/acc) for the total, ASE-only and NLI-only accumulators (osnr.pyx:133-140)
                    double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                    gl[0] = found ? -10.0 * log10(acc_ok) : 0.0;
                    gl[1] = found ? -10.0 * log10(ase_ok) : 0.0;
                    gl[2] = found ? -10.0 * log10(acc_ok - ase_ok) : 0.0;
                }
            }
            __syncwarp();
            if (FEAT & 1) {   // qrmsa.pyx:937-952
                int loc = 0;
                if (found) {
                    uint32_t f2 = 0u;
                    loc = measure_disruptions(dm, p, t, tr, perm, bm_global, lists, cur, rel_ptr, lane, f2);
                    if (f2 && lane == 0) tr[cur].w |= f2;
                    QCNT(QRMSA_CNT_DISRUPTED, loc);
                }
                if (lane == 0 && p.step_disrupted) p.step_disrupted[env] = loc;
                __syncwarp();
            }
            uint32_t n_rel = 0;
            DefragStats ds = {0u, 0u, 0u};
            if (advance_and_release<Dim<S_, M_, K_>, typename RowHandle<BMS>::type, PathTab<SM_>, Streams<RING>, FEAT>(
                    dm, p, t, tr, perm, bm, lists, pos, cur, rel_ptr, head, lane, n_rel, pt, sm, &ds))
                err = ENV_ERR_RELEASE_NOT_FOUND;
            QCNT(QRMSA_CNT_RELEASES, n_rel);
            if (FEAT & 2) {
                QCNT(QRMSA_CNT_DEFRAG_CYCLES, ds.cycles);
                QCNT(QRMSA_CNT_REALLOCATIONS, ds.moved);
                if (ds.flags && lane == 0) tr[cur - 1].w |= ds.flags;   // a near-threshold check while moving services
            }
        }
        if (BMS) {
            __syncwarp();
            uint4 *dst = reinterpret_cast<uint4 *>(bm_global);
            for (int i = lane; i < (int)(p.bm_stride >> 2); i += 32) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(wb + 16u * (uint32_t)i), "n"(WARP_STREAM_BYTES));
                dst[i] = v;
            }
            __syncwarp();   // the next env's rows overwrite the area
        }
        if (lane == 0) {   // estate.z (the accepted total) belongs to k_count_decisions
            int *es = reinterpret_cast<int *>(p.estate + env);
            *reinterpret_cast<int2 *>(es) = make_int2(cur, rel_ptr);
            es[3] = err;
        }
        if (cnt_reg) atomicAdd(p.counters + (size_t)(env / p.group_size) * QRMSA_N_COUNTERS + lane,
                               (unsigned long long)cnt_reg);
    }
}

// --------------------------------------------------------------------------------------------------------
// env.step(action) with an external action per env (qrmsa.pyx:838-1065), one request per launch.
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MAX_THREADS, 1)
    k_step_action(const KParams p, const long long *__restrict__ ext_action, float *__restrict__ o_reward,
                  uint8_t *__restrict__ o_status, double *__restrict__ o_gsnr, uint8_t *__restrict__ o_term,
                  const int episode_length) {
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<0, 0, 0> dm(p);

    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int gstride = gridDim.x * wpc;
    const int reject = p.K * p.Mc * p.S;          // the caller's reject action (k * modulations_to_consider * S)
    const int reject_log = p.K * p.M * p.S;       // the decision log addresses all n_mods modulations

    for (int env = gw; env < p.n_envs; env += gstride) {
        int4 st = p.estate[env];
        int cur = st.x, rel_ptr = st.y, accepted = st.z, err = st.w;
        int status = QRMSA_STEP_IDLE;
        float reward = 0.f;
        double g = 0.0, g_ase = 0.0, g_nli = 0.0;
        int term = 0;
        uint32_t cnt_reg = 0;
        if (err == ENV_OK && cur + 1 < p.n_req) {
            uint4 *tr = p.trace + (size_t)env * p.T;
            const unsigned long long *perm = p.perm + (size_t)env * p.T;
            uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
            uint32_t *lists = p.lists + (size_t)env * p.E * p.CAP;
            uint8_t *pos = p.pos + (size_t)env * p.pos_stride;
            const uint4 rq = tr[cur];
            const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
            const long long a64 = ext_action[env];
            uint32_t flags = QRMSA_FLAG_DECIDED;
            bool consume = true;
            int a_log = reject_log;
            if (a64 == reject || a64 < 0 || a64 > reject) {
                status = QRMSA_STEP_REJECT_ACTION;
                reward = -6.0f;  // qrmsa.pyx:992-995
            } else {
                const int a = (int)a64;
                const int s = a % p.S;
                const int rel = (a / p.S) % p.Mc;
                const int pi = (a / (p.S * p.Mc)) % p.K;
                const int mmax = (int)p.maxmod[env];
                const int m = (mmax > 1) ? mmax - rel : (p.Mc - 1) - rel;  // allowed_mods[rel], qrmsa.pyx:821-829
                a_log = pi * p.M * p.S + ((p.M - 1) - m) * p.S + s;        // the same allocation in the log's encoding
                const int n = t.need(rate * p.M + m);
                const int path = (src * p.N + dst) * p.K + pi;
                const int hops = __ldg(p.path_hops + path) & 0x7f;
                const int mylink = lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
                const int mycnt = lane < hops ? (int)*cnt_word(bm, mylink, dm.RW()) : 0;
                // is_path_free (qrmsa.pyx:1248-1264): [s, s+n (+1 guard if it ends before S)) free on every link
                bool free_ok = hops > 0 && s + n <= p.S;
                if (free_ok) {
                    const uint32_t av = path_available(dm, bm, hops, mylink, lane);
                    const int e = (s + n < p.S) ? s + n + 1 : s + n;
                    const uint32_t mask = lane < p.W ? range_mask(s, e, lane) : 0u;
                    free_ok = !__any_sync(FULL, (av & mask) != mask);
                }
                if (!free_ok) {
                    status = QRMSA_STEP_NOT_FREE;
                    consume = false;
                    // reward(): not accepted -> -3 * (1 + failed_ratio)  (qrmsa.pyx:1266-1271)
                    const double proc = (double)(cur + 1);
                    reward = (float)(-3.0 * (1.0 + (proc - (double)accepted) / proc));
                } else {
                    const int ncls = t.cls(rate * p.M + m);
                    uint32_t terms = 0;
                    const GnBase gb = gn_base(p, t, path, s, n, ncls);
                    const double acc = gb.with(gn_neighbours(dm, t, lists, hops, mylink, mycnt, 2 * s + n, lane, terms));
                    g = -10.0 * log10(acc);
                    g_ase = -10.0 * log10(gb.ase);
                    g_nli = -10.0 * log10(acc - gb.ase);
                    if (qot_ok(t, m, acc, flags)) {
                        const uint32_t rec = (uint32_t)(2 * s + n) | ((uint32_t)n << 12) | ((uint32_t)m << 20) |
                                             ((uint32_t)ncls << 23);
                        if (commit(dm, p, bm, lists, pos, hops, mylink, mycnt, s, n, rec, lane)) err = ENV_ERR_LIST_OVERFLOW;
                        flags |= QRMSA_FLAG_ACCEPTED;
                        accepted += 1;
                        status = QRMSA_STEP_ACCEPTED;
                        reward = 0.f;  // reward() falls off its end for accepted services (qrmsa.pyx:1266-1285)
                        QCNT(QRMSA_CNT_ACCEPTED, 1);
                        QCNT(QRMSA_CNT_RATE_PROVISIONED, t.rate(rate));
                        QCNT(QRMSA_CNT_HOPS_ACCEPTED, hops);
                        QCNT(QRMSA_CNT_MOD_HIST + m, 1);
                    } else {
                        status = QRMSA_STEP_LOW_GSNR;  // the reference raises here; the env is left untouched
                        consume = false;
                    }
                }
            }
            if (consume) {
                if (status == QRMSA_STEP_REJECT_ACTION) QCNT(QRMSA_CNT_REJECTED, 1);
                QCNT(QRMSA_CNT_DECIDED, 1);
                QCNT(QRMSA_CNT_RATE_REQUESTED, t.rate(rate));
                if (flags & QRMSA_FLAG_NEAR_THRESHOLD) QCNT(QRMSA_CNT_NEAR_THRESHOLD, 1);
                if (lane == 0) {
                    tr[cur].w = (uint32_t)(status == QRMSA_STEP_ACCEPTED ? a_log : reject_log) | flags;
                    if (p.gsnr_log) {
                        double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                        const bool acc_ = status == QRMSA_STEP_ACCEPTED;
                        gl[0] = acc_ ? g : 0.0; gl[1] = acc_ ? g_ase : 0.0; gl[2] = acc_ ? g_nli : 0.0;
                    }
                }
                __syncwarp();
                if (p.feat & 1) {   // qrmsa.pyx:937-952
                    int loc = 0;
                    if (status == QRMSA_STEP_ACCEPTED) {
                        uint32_t f2 = 0u;
                        loc = measure_disruptions(dm, p, t, tr, perm, bm, lists, cur, rel_ptr, lane, f2);
                        if (f2 && lane == 0) tr[cur].w |= f2;
                        if (f2 && !(flags & QRMSA_FLAG_NEAR_THRESHOLD)) QCNT(QRMSA_CNT_NEAR_THRESHOLD, 1);
                        QCNT(QRMSA_CNT_DISRUPTED, loc);
                    }
                    if (lane == 0 && p.step_disrupted) p.step_disrupted[env] = loc;
                    __syncwarp();
                }
                Head head = load_head(p, tr, perm, rel_ptr);
                uint32_t n_rel = 0;
                DefragStats ds = {0u, 0u, 0u};
                int rerr;
                if (p.feat & 2)
                    rerr = advance_and_release<Dim<0, 0, 0>, uint32_t *, PathTab<0>, Streams<false>, 2>(
                        dm, p, t, tr, perm, bm, lists, pos, cur, rel_ptr, head, lane, n_rel, PathTab<0>(), Streams<false>(), &ds);
                else
                    rerr = advance_and_release(dm, p, t, tr, perm, bm, lists, pos, cur, rel_ptr, head, lane, n_rel);
                if (rerr) err = ENV_ERR_RELEASE_NOT_FOUND;
                QCNT(QRMSA_CNT_RELEASES, n_rel);
                QCNT(QRMSA_CNT_DEFRAG_CYCLES, ds.cycles);
                QCNT(QRMSA_CNT_REALLOCATIONS, ds.moved);
                if (ds.flags && lane == 0) tr[cur - 1].w |= ds.flags;
                term = (cur + 1 == episode_length);  // episode_services_processed == episode_length (qrmsa.pyx:1056)
            }
            if (err) QCNT(QRMSA_CNT_ERRORS, 1);
            if (lane == 0) { p.estate[env] = make_int4(cur, rel_ptr, accepted, err); p.counted[env] = (uint32_t)cur; }
            if (cnt_reg) atomicAdd(p.counters + (size_t)(env / p.group_size) * QRMSA_N_COUNTERS + lane,
                                   (unsigned long long)cnt_reg);
        }
        if (lane == 0) {
            if (o_reward) o_reward[env] = reward;
            if (o_status) o_status[env] = (uint8_t)status;
            if (o_gsnr) o_gsnr[env] = g;
            if (o_term) o_term[env] = (uint8_t)term;
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// Observation vector + GSNR-validated action mask (QRMSAEnv.observation with gen_observation=True,
// qrmsa.pyx:583-781; calculate_osnr_observation, osnr.pyx:259-368).  One CTA per env, read-only on env state.
//
// The reference evaluates one GN sum per valid (path, modulation, start) -- and does it twice (features, then
// mask).  The neighbour sum depends only on the candidate's centre c2 = 2*start + n (half-slots), so per path the
// CTA first builds X[c2] for every c2 (thread per c2: consecutive threads read consecutive G/INV entries, so the
// shared-memory lookups are conflict-free), then every (modulation, start) is one lookup + log10.
// --------------------------------------------------------------------------------------------------------
// Valid-start bitmaps of every modulation of the open path, by one warp (lane j = word j of the availability): the
// doubling state carries over as the slot count grows.
__device__ __forceinline__ void valid_starts_all(const Tab &t, uint32_t av, int rate, int M, uint32_t (*validM)[32], int lane) {
    uint32_t r = av;
    int a = 1;
    for (int mi = 0; mi < M; ++mi) {
        const int L = t.need(rate * M + (M - 1) - mi) + 1;
        if (L < a) { r = av; a = 1; }
        while (a < L) { const int b = min(a, L - a); r &= shr_multi(r, b); a += b; }
        validM[mi][lane] = r;
        const unsigned any = __ballot_sync(FULL, r != 0u);
        if (lane == 31) validM[mi][31] = any ? 1u : 0u;   // word 31 is never spectrum (S <= 960): "some start is valid"
    }
}

// Is the neighbour sum at centre c2 (half-slots) needed, i.e. is c2 = 2*s + n for a valid start s of some modulation?
__device__ __forceinline__ bool centre_needed(const Tab &t, const uint32_t (*validM)[32], int rate, int M, int S, int c2) {
    bool need = false;
    for (int mi = 0; mi < M; ++mi) {
        const int d = c2 - t.need(rate * M + (M - 1) - mi);
        const int s = d >> 1;
        if (d >= 0 && !(d & 1) && s < S && ((validM[mi][s >> 5] >> (s & 31)) & 1u)) need = true;
    }
    return need;
}

constexpr int OBS_THREADS = 320;       // k_step_highest_snr: 10 warps per env, two CTAs per SM
constexpr int OBS_ENV_THREADS = 160;   // k_observation: 5 warps per env, up to OBS_MAX_EPC envs per CTA
constexpr int OBS_MAX_EPC = 4;
constexpr int OBS_MAX_IT = 6;          // 960 slots / OBS_ENV_THREADS
constexpr int OBS_NRED = 6;

struct ObsSmem {           // lives after the table blob in dynamic shared memory
    double red[OBS_THREADS / 32][OBS_NRED];
    double bcast[8];
    uint32_t av[32];       // path availability words (+ the virtual slot)
    uint32_t valid[32];    // valid-start bitmap of the current modulation
    int link[32], cnt[32];
    double w1[32], w2[32];
    uint32_t validM[8][32];                    // valid-start bitmaps of all modulations of the open path
    double part[8][OBS_THREADS / 32][8];       // per-warp partial statistics: [modulation | 7 = free blocks][warp][k]
};

// sums (or max for index 4, 5) of OBS_NRED per-thread values over the CTA; result broadcast to all threads
__device__ __forceinline__ void block_reduce6(ObsSmem *sm, double v[OBS_NRED], const unsigned max_mask) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < OBS_NRED; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(FULL, x, o);
            x = ((max_mask >> k) & 1u) ? fmax(x, y) : x + y;
        }
        if (lane == 0) sm->red[warp][k] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < OBS_NRED; ++k) {
            double x = lane < (int)(blockDim.x >> 5) ? sm->red[lane][k] : (((max_mask >> k) & 1u) ? -1e300 : 0.0);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double y = __shfl_xor_sync(FULL, x, o);
                x = ((max_mask >> k) & 1u) ? fmax(x, y) : x + y;
            }
            if (lane == 0) sm->bcast[k] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OBS_NRED; ++k) v[k] = sm->bcast[k];
    __syncthreads();
}

__global__ void __launch_bounds__(OBS_ENV_THREADS * OBS_MAX_EPC, 1)
    k_observation(const KParams p, const double *__restrict__ path_len_norm, const double inv_max_rate,
                  float *__restrict__ obs_out, uint8_t *__restrict__ mask_out, const int obs_dim, const int n_actions,
                  const int epc, const int env_smem) {
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<0, 0, 0> dm(p);
    const int S = p.S, W = p.W, M = p.M, K = p.K, D = p.D, CAP = p.CAP;
    // epc envs per CTA, OBS_ENV_THREADS threads each, synchronised by their own named barrier: while the warps of one
    // env wait for its slowest warp, the warps of the other envs keep the SM busy (the tables are shared)
    const int slot = threadIdx.x / OBS_ENV_THREADS;
    auto env_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(OBS_ENV_THREADS) : "memory"); };
    unsigned char *extra = qsmem + p.blob_bytes + (size_t)slot * env_smem;
    ObsSmem *sm = reinterpret_cast<ObsSmem *>(extra);
    double *X = reinterpret_cast<double *>(extra + sizeof(ObsSmem));          // [D]
    double *NRM = X + D;                                                        // [S] normalised GSNR per start
    uint32_t *rec = reinterpret_cast<uint32_t *>(NRM + p.S);                    // [Hmax][CAP]
    const int tid = threadIdx.x - slot * OBS_ENV_THREADS, lane = tid & 31, warp = tid >> 5;

    for (int env = blockIdx.x * epc + slot; env < p.n_envs; env += gridDim.x * epc) {
        const int4 st = p.estate[env];
        float *obs = obs_out + (size_t)env * obs_dim;
        uint8_t *mask = mask_out + (size_t)env * n_actions;
        const int cur = st.x < p.n_req ? st.x : p.n_req - 1;
        const uint4 rq = p.trace[(size_t)env * p.T + cur];
        const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
        const uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
        const uint32_t *lists = p.lists + (size_t)env * p.E * CAP;
        const int pbase = (src * p.N + dst) * K;
        if (tid == 0) {
            obs[0] = (float)((double)t.rate(rate) * inv_max_rate);                       // qrmsa.pyx:654-665
            obs[1] = (float)(p.N > 1 ? (double)src / (double)(p.N - 1) : 0.0);
            obs[2] = (float)(p.N > 1 ? (double)dst / (double)(p.N - 1) : 0.0);
            mask[n_actions - 1] = 1;                                                        // qrmsa.pyx:766
        }
        for (int pi = 0; pi < K; ++pi) {
            const int path = pbase + pi;
            const int hops = __ldg(p.path_hops + path) & 0x7f;
            if (tid == 0) obs[3 + pi] = hops ? (float)path_len_norm[path] : 0.f;
            if (hops == 0) {   // fewer than k paths for this pair: features stay -1, no valid action (qrmsa.pyx:697)
                for (int i = tid; i < M * 12; i += OBS_ENV_THREADS) obs[3 + K + pi * M * 12 + i] = -1.f;
                for (int i = tid; i < M * S; i += OBS_ENV_THREADS) mask[(size_t)pi * M * S + i] = 0;
                continue;
            }
            env_sync();
            if (tid < 32) {
                const int l = tid < hops ? __ldg(p.path_links + path * p.Hmax + tid) : 0;
                sm->link[tid] = l;
                sm->cnt[tid] = tid < hops ? (int)bm[(unsigned)(l * p.RW + p.RW - 1)] : 0;
                sm->w1[tid] = t.W1(l);
                sm->w2[tid] = t.W2(l);
                const uint32_t a = path_available(dm, bm, hops, l, lane);
                sm->av[tid] = a;
                valid_starts_all(t, a, rate, M, sm->validM, lane);
            }
            env_sync();
            // stage the channel records of the path's links
            for (int i = 0; i < hops; ++i) {
                const int c = sm->cnt[i];
                const uint32_t *lst = lists + (unsigned)(sm->link[i] * CAP);
                for (int q = tid; q < c; q += OBS_ENV_THREADS) rec[i * CAP + q] = lst[q];
            }
            env_sync();
            // X[c2]: neighbour sum for a candidate centred at c2 half-slots -- only where some modulation has a valid
            // start with that centre (a loaded network leaves most centres unused); a warp whose 32 centres are all
            // unused skips the sum, the others keep consecutive centres on consecutive lanes (conflict-free lookups)
            for (int c0 = 0; c0 < D; c0 += OBS_ENV_THREADS) {
                const int c2 = c0 + tid;
                const bool need = c2 < D && centre_needed(t, sm->validM, rate, M, S, c2);
                if (!__any_sync(FULL, need)) continue;
                if (need) {
                    double x = 0.0;
                    for (int i = 0; i < hops; ++i) {
                        const int c = sm->cnt[i];
                        const uint32_t *r = rec + i * CAP;
                        double s1 = 0.0, s2 = 0.0;
                        for (int q = 0; q < c; ++q) gn_term(t, D, r[q], c2, s1, s2);
                        x = fma(sm->w1[i], s1, x);
                        x = fma(sm->w2[i], s2, x);
                    }
                    X[c2] = x;
                }
            }
            // Statistics are reduced once per path: every warp leaves its partial sums for the free blocks and for each
            // modulation in shared memory, then warp mi finishes modulation mi.  Four barriers per path instead of
            // three per modulation.
            const int nw = OBS_ENV_THREADS >> 5;
            // k = 0..3, 6: sums; 4, 5: maxima.  Entries named in int_mask are small exact integers (counts, slot indices and
            // their squares): they take one warp-reduce instruction instead of five shuffle rounds in FP64.
            auto warp_part = [&](double (&v)[8], int slot, const unsigned int_mask) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    double x;
                    if ((int_mask >> k) & 1u) {
                        const int iv = v[k] < -1e299 ? -1 : (int)v[k];
                        const int r = k == 4 ? __reduce_max_sync(FULL, iv) : __reduce_add_sync(FULL, iv);
                        x = (k == 4 && r < 0) ? -1e300 : (double)r;
                    } else {
                        x = v[k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double y = __shfl_xor_sync(FULL, x, o);
                            x = (k == 4 || k == 5) ? fmax(x, y) : x + y;
                        }
                    }
                    if (lane == 0) sm->part[slot][warp][k] = x;
                }
            };
            // free-block statistics of the path availability (qrmsa.pyx:631-646), thread per slot
            {
                double bs[8] = {0, 0, 0, 0, -1e300, -1e300, 0, 0};
                for (int s = tid; s < S; s += OBS_ENV_THREADS) {
                    const bool free_s = (sm->av[s >> 5] >> (s & 31)) & 1u;
                    const bool free_n = (s + 1 < S) && ((sm->av[(s + 1) >> 5] >> ((s + 1) & 31)) & 1u);
                    if (free_s) bs[0] += 1.0;                       // total available slots
                    if (free_s && !free_n) {                        // a run ends here: walk back to its start
                        int b = s;
                        while (b > 0 && ((sm->av[(b - 1) >> 5] >> ((b - 1) & 31)) & 1u)) --b;
                        const double len = (double)(s - b + 1);
                        bs[1] += 1.0; bs[2] += len; bs[3] += len * len;
                    }
                }
                warp_part(bs, 7, 0x0fu);   // free slots, blocks, sum of lengths, sum of squared lengths
            }
            env_sync();   // X[], validM[] complete
            // per modulation: GSNR per valid start, mask, partial statistics
            double g_cached[OBS_MAX_IT];
#pragma unroll
            for (int it = 0; it < OBS_MAX_IT; ++it) g_cached[it] = 0.0;
            int n_cached = -1;
            for (int mi = 0; mi < M; ++mi) {
                const int m = (M - 1) - mi;
                const int n = t.need(rate * M + m), ncls = t.cls(rate * M + m);
                const double th = p.mod_thr_nomargin[m];
                // count, sum s, sum s^2, sum norm, max s, max norm, sum norm^2 (positions are small integers and the
                // normalised GSNR is O(1), so both variances come from one pass in FP64)
                double v[8] = {0, 0, 0, 0, -1e300, -1e300, 0, 0};
                const bool same_n = n == n_cached;   // same slot count as the previous modulation: same starts, same GSNR
#pragma unroll
                for (int it = 0; it < OBS_MAX_IT; ++it) {     // S <= 960 = OBS_MAX_IT * OBS_ENV_THREADS
                    const int s = tid + it * OBS_ENV_THREADS;
                    if (s >= S) break;
                    const bool ok = (sm->validM[mi][s >> 5] >> (s & 31)) & 1u;
                    uint8_t bit = 0;
                    if (ok) {
                        double g;
                        if (same_n) {
                            g = g_cached[it];
                        } else {
                            const double acc = gn_base(p, t, path, s, n, ncls).with(X[2 * s + n]);
                            g = 10.0 * log10(1.0 / acc);
                            g_cached[it] = g;
                        }
                        const double nrm = rint(((g - th) / fabs(th)) * 1e10) / 1e10;     // np.round(x, 10), osnr.pyx:366
                        bit = nrm >= 0.0 ? 1 : 0;
                        v[0] += 1.0; v[1] += (double)s; v[2] += (double)s * (double)s; v[3] += nrm;
                        v[4] = fmax(v[4], (double)s); v[5] = fmax(v[5], nrm);
                        v[6] = fma(nrm, nrm, v[6]);
                    }
                    mask[(size_t)pi * M * S + (size_t)mi * S + s] = bit;
                }
                warp_part(v, mi, 0x17u);   // count, sum s, sum s^2, max s
                n_cached = n;
            }
            env_sync();   // partial statistics complete
            for (int mi = warp; mi < M; mi += nw) {   // warp w finishes modulations w, w + nw, ...
                const int m = (M - 1) - mi;
                const int n = t.need(rate * M + m);
                double f[8], b4[4];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    double x = lane < nw ? sm->part[mi][lane][k] : ((k == 4 || k == 5) ? -1e300 : 0.0);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double y = __shfl_xor_sync(FULL, x, o);
                        x = (k == 4 || k == 5) ? fmax(x, y) : x + y;
                    }
                    f[k] = x;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double x = lane < nw ? sm->part[7][lane][k] : 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
                    b4[k] = x;
                }
                if (lane == 0) {
                    const double total_av = b4[0], nb = b4[1];
                    double mean_block = 0.0, std_block = 0.0;
                    if (nb > 0.0) {
                        const double mb = b4[2] / nb;
                        mean_block = ((mb - 4.0) / 4.0) / 100.0;
                        std_block = sqrt(fmax(b4[3] / nb - mb * mb, 0.0)) / 100.0;
                    }
                    const double cntv = f[0];
                    double f_avg = 0, f_std = 0, f_max = 0, best = 0, omean = 0, ovar = 0;
                    if (cntv > 0.0) {
                        f_avg = f[1] / cntv; omean = f[3] / cntv; f_max = f[4]; best = fmax(f[5], 0.0);
                        f_std = sqrt(fmax(f[2] / cntv - f_avg * f_avg, 0.0));
                        ovar = fmax(f[6] / cntv - omean * omean, 0.0);
                    }
                    float *o = obs + 3 + K + (pi * M + mi) * 12;
                    o[0] = (float)(cntv / (double)S);
                    o[1] = (float)(f_avg / (double)(S - 1));
                    o[2] = (float)(f_std / (double)(S - 1));
                    o[3] = (float)fmax(((double)n - 5.5) / 3.5, 0.0);
                    o[4] = (float)(2.0 * (total_av - 0.5 * (double)S) / (double)S);
                    o[5] = (float)mean_block;
                    o[6] = (float)std_block;
                    o[7] = (float)best;
                    o[8] = (float)omean;
                    o[9] = (float)ovar;
                    o[10] = (float)(2.0 * ((total_av / (double)S) - 0.5));
                    o[11] = (float)(f_max / (double)(S - 1));
                }
            }
        }
        env_sync();
    }
}

// --------------------------------------------------------------------------------------------------------
// k_observation_links: the same observation + mask for spectra up to 320 slots (D = 2S <= 4 * OBS_ENV_THREADS), built
// LINK-major.  The k paths of a node pair share links (NSFNET: 20.4 hop visits but 12.1 distinct links per request,
// nobel-eu: 25.1 / 13.3), and the neighbour sum of a link does not depend on the path it is reached by, so the sums
// are made once per DISTINCT link and added into X[p][c2] of every path p that crosses it.  Thread t owns the four
// centres c2 = t, t + 160, t + 320, t + 480 for the whole request: one channel record is fetched (broadcast) and
// decoded once for four table lookups, and the sums of a link stay in registers.  The kernel is bound by the
// shared-memory bandwidth of the G / INV lookups (16 bytes per term); this layout takes it from 6 to 4.25 wavefronts per
// 32 terms and cuts the terms by the link sharing.  Afterwards warp w owns path w (round robin): free-block
// statistics, GSNR per valid start, mask bytes and the 12 features per modulation need warp shuffles only -- three
// env-wide barriers per request instead of seven per path.
// --------------------------------------------------------------------------------------------------------
template <int V> struct IntC { static constexpr int value = V; };
constexpr int OBS2_NJ = 4;                          // centres per thread
constexpr int OBS2_MAX_D = OBS2_NJ * OBS_ENV_THREADS;
constexpr int OBS2_STAGE = 16;                      // channel records decoded and staged per warp and pass of a link's list

__host__ __device__ inline int obs2_env_smem(int K, int D, int W, int E) {
    const int VW = W + 1, NW = (D + 31) >> 5, S = D / 2, NWARP = OBS_ENV_THREADS / 32;
    // scratch after the fixed arrays: phase 2 keeps the compacted valid starts of each warp's unit there (u16 [NWARP][S]);
    // phase 1 the compacted needed centres (u16 [D]) and, 16-byte aligned, each warp's 16 staged records of 16 bytes
    const int scratch2 = 2 * NWARP * S, scratch1 = 2 * D + 16 + NWARP * OBS2_STAGE * 16;
    const int scratch = (scratch1 > scratch2 ? scratch1 : scratch2);
    const int words = K * 8 * VW + K * NW + K * VW + K * 32 + K + E + 4 + (scratch + 3) / 4;
    return ((8 * K * D + 8 * 3 * K + 4 * words) + 15) / 16 * 16;
}

__device__ __forceinline__ uint32_t spread16(uint32_t x) {   // bit i -> bit 2i
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// One env's shared-memory area of the link-major kernels (k_observation_links, k_step_highest_snr_links), after the table blob
struct LinkArea {
    double *X, *pstat;                   // [K][D] neighbour sum per path and centre; [K][3] free slots, mean / std of the free blocks
    uint32_t *validM, *need, *avs;       // [K][8][VW] valid starts per modulation; [K][NW] usable centres; [K][VW] availability
    int *plink, *phops;                  // [K][32] link ids; [K] hop counts
    uint32_t *lmask;                     // [E] paths crossing link l (0: none)
    int *tick, *smax;                    // work ticket of the per-path phase; max_modulation_idx of the open request
    unsigned char *scr;                  // scratch (obs2_env_smem): compacted centres + staged records, later compacted starts
    __device__ __forceinline__ void carve(unsigned char *base, int K, int D, int VW, int NW, int E) {
        X = reinterpret_cast<double *>(base);
        pstat = X + (size_t)K * D;
        validM = reinterpret_cast<uint32_t *>(pstat + 3 * K);
        need = validM + K * 8 * VW;
        avs = need + K * NW;
        plink = reinterpret_cast<int *>(avs + K * VW);
        phops = plink + K * 32;
        lmask = reinterpret_cast<uint32_t *>(phops + K);
        tick = reinterpret_cast<int *>(lmask + E);
        smax = tick + 1;
        scr = reinterpret_cast<unsigned char *>(tick + 4);
    }
};

// Phase 0 of a request, warp per path: links, availability, valid starts of every modulation, usable centres; with STATS
// (observation) the path-length feature, the free-block statistics and the -1 features of a missing path.
template <bool STATS>
__device__ __forceinline__ void links_open_paths(const KParams &p, const Tab &t, const LinkArea &ar, const uint32_t *bm, const int pbase,
                                                 const int rate, const int warp, const int lane, float *obs,
                                                 const double *__restrict__ path_len_norm) {
    const Dim<0, 0, 0> dm(p);
    const int S = p.S, W = p.W, M = p.M, K = p.K, D = p.D;
    const int VW = W + 1, NW = (D + 31) >> 5;
    constexpr int nw = OBS_ENV_THREADS >> 5;
    double *pstat = ar.pstat;
    uint32_t *validM = ar.validM, *need = ar.need, *avs = ar.avs, *lmask = ar.lmask;
    int *plink = ar.plink, *phops = ar.phops;
    for (int pi = warp; pi < K; pi += nw) {
        const int path = pbase + pi;
        const int hops = __ldg(p.path_hops + path) & 0x7f;
        if (lane == 0) {
            phops[pi] = hops;
            if (STATS) obs[3 + pi] = hops ? (float)path_len_norm[path] : 0.f;
        }
        for (int w = lane; w < NW; w += 32) need[pi * NW + w] = 0u;
        if (hops == 0) {   // fewer than k paths for this pair: features stay -1, no valid action (qrmsa.pyx:697)
            if (STATS)
                for (int i = lane; i < p.Mc * 12; i += 32) obs[3 + K + pi * p.Mc * 12 + i] = -1.f;
            continue;
        }
        const int l = lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
        plink[pi * 32 + lane] = l;
        if (lane < hops) atomicOr(&lmask[l], 1u << pi);
        const uint32_t a = path_available(dm, bm, hops, l, lane);
        if (lane < VW) avs[pi * VW + lane] = a;
        uint32_t r = a;
        int aa = 1;
        for (int mi = 0; mi < M; ++mi) {
            const int L = t.need(rate * M + (M - 1) - mi) + 1;
            if (L < aa) { r = a; aa = 1; }
            while (aa < L) { const int b = min(aa, L - aa); r &= shr_multi(r, b); aa += b; }
            if (lane < VW) validM[(pi * 8 + mi) * VW + lane] = r;
        }
        __syncwarp();
        // need[c2] = some modulation has a valid start s with 2 s + n == c2: for the 32 centres of word w and a
        // modulation of n slots these are the 16 starts from 16 w - (n >> 1), spread to every other bit
        for (int w = lane; w < NW; w += 32) {
            uint32_t bits = 0u;
            for (int mi = 0; mi < M; ++mi) {
                const int n = t.need(rate * M + (M - 1) - mi);
                int s0 = 16 * w - (n >> 1), sh = 0;
                if (s0 < 0) { sh = -s0; s0 = 0; }
                if (sh >= 16) continue;
                const uint32_t *row = validM + (pi * 8 + mi) * VW;
                const int q = s0 >> 5;
                const uint32_t lo = q < VW ? row[q] : 0u, hi = q + 1 < VW ? row[q + 1] : 0u;
                const uint32_t v16 = ((__funnelshift_r(lo, hi, s0 & 31) & 0xffffu) << sh) & 0xffffu;
                bits |= spread16(v16) << (n & 1);
            }
            if (w == NW - 1 && (D & 31)) bits &= (1u << (D & 31)) - 1u;
            need[pi * NW + w] = bits;
        }
        if (STATS) {
            // free blocks of the path availability (qrmsa.pyx:631-646): lane j looks at word j; a block is counted in the
            // word that holds its last slot.  A block that reaches bit 0 of its word continues the free run that ends at the
            // top of the word below (`carry`, handed up through all-free words)
            int b_free = 0, b_n = 0, b_len = 0, b_len2 = 0;
            {
                uint32_t wv = 0u;   // word `lane` of the availability without the virtual slot at S
                if (lane < W) {
                    wv = avs[pi * VW + lane];
                    if (lane == (S >> 5)) wv &= (1u << (S & 31)) - 1u;
                }
                b_free = __popc(wv);
                const int top = __clz((int)~wv);   // free slots at the top of the word (32: the whole word)
                int carry = 0, run = 0;
                for (int w = 0; w < W; ++w) {
                    if (lane == w) carry = run;
                    const int tl = __shfl_sync(FULL, top, w);
                    run = tl == 32 ? run + 32 : tl;
                }
                const uint32_t up = __shfl_down_sync(FULL, wv, 1);   // (lanes at and past W hold 0)
                uint32_t ends = wv & ~((wv >> 1) | (up << 31));
                while (ends) {
                    const int eb = __ffs(ends) - 1;
                    ends &= ends - 1u;
                    const uint32_t occ = ~wv & ((eb == 31) ? 0xffffffffu : ((2u << eb) - 1u));   // occupied slots below the end
                    const int len = occ ? eb - 31 + __clz((int)occ) : eb + 1 + carry;
                    b_n += 1; b_len += len; b_len2 += len * len;
                }
            }
            b_free = __reduce_add_sync(FULL, b_free); b_n = __reduce_add_sync(FULL, b_n);
            b_len = __reduce_add_sync(FULL, b_len); b_len2 = __reduce_add_sync(FULL, b_len2);
            if (lane == 0) {
                double mean_block = 0.0, std_block = 0.0;
                if (b_n > 0) {
                    const double nb = (double)b_n, mb = (double)b_len / nb;
                    mean_block = ((mb - 4.0) / 4.0) / 100.0;
                    std_block = sqrt(fmax((double)b_len2 / nb - mb * mb, 0.0)) / 100.0;
                }
                pstat[pi * 3 + 0] = (double)b_free; pstat[pi * 3 + 1] = mean_block; pstat[pi * 3 + 2] = std_block;
            }
        }
    }
}

// Phase 1 of a request: neighbour sums link by link (core/osnr.pyx:64-94, table form) into X[p][c2].  Returns the number
// of (record, centre) terms this thread summed.
__device__ __forceinline__ uint32_t links_neighbour_sums(const KParams &p, const Tab &t, const LinkArea &ar, const uint32_t *bm,
                                                         const uint32_t *lists, const int tid) {
    const int K = p.K, D = p.D, CAP = p.CAP, E = p.E, NW = (D + 31) >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    double *X = ar.X;
    const uint32_t *need = ar.need, *lmask = ar.lmask;
    uint32_t n_terms = 0u;
    // ---- phase 1: neighbour sums link by link (core/osnr.pyx:64-94, table form).  Only centres that some path can use
    // are summed: their union over the paths is compacted (every warp builds the same list -- identical stores, no
    // barrier) and thread t owns compacted centres t, t + 160, ... for the whole request, so the sums of a link stay in
    // registers and X[p][c2] has one writer.  nj = ceil(needed / 160) centre groups are live (a template parameter of
    // the inner loop: no per-term test).  A link's records are decoded once -- centre, G-row address, phi * bandwidth --
    // 16 per pass, lane per record, into the warp's staging area, and the loop over them is one 16-byte broadcast
    // LDS per record and, per centre, |d|, two addresses, two LDS.64, DADD, DFMA.
    {
        unsigned char *scr = ar.scr;
        uint16_t *clist = reinterpret_cast<uint16_t *>(scr);
        const uint32_t stage = (smem_u32(scr + 2 * D) + 15u) / 16u * 16u + (uint32_t)(warp * OBS2_STAGE * 16);
        uint32_t uw = 0u;
        if (lane < NW)
            for (int pi = 0; pi < K; ++pi) uw |= need[pi * NW + lane];
        int incl = __popc(uw);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int P = __shfl_sync(FULL, incl, 31);
        const int excl = incl - __popc(uw);
        for (int w = 0; w < NW; ++w) {
            const uint32_t word = __shfl_sync(FULL, uw, w);
            const int base = __shfl_sync(FULL, excl, w);
            if ((word >> lane) & 1u) clist[base + __popc(word & ((1u << lane) - 1u))] = (uint16_t)(w * 32 + lane);
        }
        __syncwarp();
        const int nj = (P + OBS_ENV_THREADS - 1) / OBS_ENV_THREADS;
        int c2j8[OBS2_NJ];            // 8 * centre (byte offset into a table row)
        uint32_t myneed[OBS2_NJ];     // bit pi: path pi uses this centre
#pragma unroll
        for (int j = 0; j < OBS2_NJ; ++j) {
            const int idx = tid + j * OBS_ENV_THREADS;
            int c2 = 0;
            uint32_t mset = 0u;
            if (idx < P) {
                c2 = clist[idx];
                for (int pi = 0; pi < K; ++pi) mset |= ((need[pi * NW + (c2 >> 5)] >> (c2 & 31)) & 1u) << pi;
            }
            c2j8[j] = 8 * c2;
            myneed[j] = mset;
        }
        const uint32_t d8 = 8u * (uint32_t)D;
#pragma unroll 1
        for (int l = 0; l < E && nj > 0; ++l) {
            const uint32_t pm = lmask[l];
            if (!pm) continue;      // no path of this request crosses the link
            // groups of centres with a taker in this warp for this link: all of them -> the nj-wide loop, some -> one
            // single-group loop per taker (the staged records are read once per group then), none -> next link
            uint32_t wmask = 0u;
#pragma unroll
            for (int j = 0; j < OBS2_NJ; ++j)
                if (__any_sync(FULL, (myneed[j] & pm) != 0u)) wmask |= 1u << j;
            if (!wmask) continue;
            const bool all_groups = wmask == (1u << nj) - 1u;
            const int cnt = (int)bm[(unsigned)(l * p.RW + p.RW - 1)];
            const uint32_t *lst = lists + (unsigned)(l * CAP);
            double s1[OBS2_NJ], s2[OBS2_NJ];
#pragma unroll
            for (int j = 0; j < OBS2_NJ; ++j) s1[j] = s2[j] = 0.0;
#pragma unroll 1
            for (int q0 = 0; q0 < cnt; q0 += OBS2_STAGE) {
                __syncwarp();   // the previous pass has been read
                if (lane < OBS2_STAGE) {   // (entries past the count hold the zero-contribution filler record; CAP is a multiple of 32)
                    const uint32_t rec = lst[q0 + lane];
                    const double phin = t.PHIN(rec >> 20);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stage + 16u * (uint32_t)lane),
                                 "r"(8u * (rec & 0xfffu)), "r"(t.sb + ((rec >> 23) + 1u) * d8),
                                 "r"((uint32_t)__double2loint(phin)), "r"((uint32_t)__double2hiint(phin)) : "memory");
                }
                __syncwarp();
                const int nq = min(OBS2_STAGE, cnt - q0);
                n_terms += (uint32_t)(nq * __popc(wmask));
                auto sum_pass = [&](auto njc) {
                    constexpr int NJ = decltype(njc)::value;
#pragma unroll 2
                    for (int q = 0; q < nq; ++q) {
                        uint32_t c2r8, gb, plo, phi;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c2r8), "=r"(gb), "=r"(plo), "=r"(phi)
                                     : "r"(stage + 16u * (uint32_t)q));
                        const double phin = __hiloint2double((int)phi, (int)plo);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const uint32_t dd = (uint32_t)abs((int)c2r8 - c2j8[j]);
                            double g, inv;
                            asm("ld.shared.f64 %0, [%1+%2];" : "=d"(g) : "r"(gb + dd), "n"(lay::INV));
                            asm("ld.shared.f64 %0, [%1+%2];" : "=d"(inv) : "r"(t.sb + dd), "n"(lay::INV));
                            s1[j] += g;
                            s2[j] = fma(phin, inv, s2[j]);
                        }
                    }
                };
                auto sum_group = [&](auto jc) {
                    constexpr int J = decltype(jc)::value;
#pragma unroll 2
                    for (int q = 0; q < nq; ++q) {
                        uint32_t c2r8, gb, plo, phi;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c2r8), "=r"(gb), "=r"(plo), "=r"(phi)
                                     : "r"(stage + 16u * (uint32_t)q));
                        const double phin = __hiloint2double((int)phi, (int)plo);
                        const uint32_t dd = (uint32_t)abs((int)c2r8 - c2j8[J]);
                        double g, inv;
                        asm("ld.shared.f64 %0, [%1+%2];" : "=d"(g) : "r"(gb + dd), "n"(lay::INV));
                        asm("ld.shared.f64 %0, [%1+%2];" : "=d"(inv) : "r"(t.sb + dd), "n"(lay::INV));
                        s1[J] += g;
                        s2[J] = fma(phin, inv, s2[J]);
                    }
                };
                if (all_groups) {
                    if (nj >= 4) sum_pass(IntC<4>()); else if (nj == 3) sum_pass(IntC<3>());
                    else if (nj == 2) sum_pass(IntC<2>()); else sum_pass(IntC<1>());
                } else {
                    if (wmask & 1u) sum_group(IntC<0>());
                    if (wmask & 2u) sum_group(IntC<1>());
                    if (wmask & 4u) sum_group(IntC<2>());
                    if (wmask & 8u) sum_group(IntC<3>());
                }
            }
            const double w1 = t.W1(l), w2 = t.W2(l);   // W2 is stored negated
#pragma unroll
            for (int j = 0; j < OBS2_NJ; ++j) {
                uint32_t ps = myneed[j] & pm;
                if (!ps) continue;
                const double y = fma(w2, s2[j], w1 * s1[j]);
                while (ps) {
                    const int pi = __ffs(ps) - 1;
                    ps &= ps - 1u;
                    X[pi * D + (c2j8[j] >> 3)] += y;                         // this thread owns the centre: no race
                }
            }
        }
    }
    return n_terms;
}

__global__ void __launch_bounds__(OBS_ENV_THREADS * OBS_MAX_EPC, 1)
    k_observation_links(const KParams p, const double *__restrict__ path_len_norm, const double inv_max_rate,
                        float *__restrict__ obs_out, uint8_t *__restrict__ mask_out, const int obs_dim, const int n_actions,
                        const int epc, const int env_smem) {
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<0, 0, 0> dm(p);
    const int S = p.S, W = p.W, M = p.M, Mc = p.Mc, K = p.K, D = p.D, CAP = p.CAP, E = p.E;
    const int VW = W + 1, NW = (D + 31) >> 5;
    const int slot = threadIdx.x / OBS_ENV_THREADS;
    auto env_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(OBS_ENV_THREADS) : "memory"); };
    unsigned char *base = qsmem + p.blob_bytes + (size_t)slot * env_smem;
    LinkArea ar;
    ar.carve(base, K, D, VW, NW, E);
    double *X = ar.X, *pstat = ar.pstat;
    uint32_t *validM = ar.validM, *lmask = ar.lmask;
    int *phops = ar.phops, *tick = ar.tick, *smax = ar.smax;
    const int tid = threadIdx.x - slot * OBS_ENV_THREADS, lane = tid & 31, warp = tid >> 5;
    uint16_t *slist = reinterpret_cast<uint16_t *>(ar.scr) + warp * S;   // [nw][S]  valid starts of the open unit, compacted
    const double inv_S = 1.0 / (double)S, inv_S1 = 1.0 / (double)(S - 1);

    for (int i = tid; i < E; i += OBS_ENV_THREADS) lmask[i] = 0u;
    env_sync();
    for (int env = blockIdx.x * epc + slot; env < p.n_envs; env += gridDim.x * epc) {
        const int4 st = p.estate[env];
        float *obs = obs_out + (size_t)env * obs_dim;
        uint8_t *mask = mask_out + (size_t)env * n_actions;
        const int cur = st.x < p.n_req ? st.x : p.n_req - 1;
        const uint4 rq = p.trace[(size_t)env * p.T + cur];
        const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
        const uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
        const uint32_t *lists = p.lists + (size_t)env * E * CAP;
        const int pbase = (src * p.N + dst) * K;
        if (tid == 0) {
            obs[0] = (float)((double)t.rate(rate) * inv_max_rate);                       // qrmsa.pyx:654-665
            obs[1] = (float)(p.N > 1 ? (double)src / (double)(p.N - 1) : 0.0);
            obs[2] = (float)(p.N > 1 ? (double)dst / (double)(p.N - 1) : 0.0);
            mask[n_actions - 1] = 1;                                                        // qrmsa.pyx:766
            *tick = 0;
        }
        for (int i = tid; i < K * D; i += OBS_ENV_THREADS) X[i] = 0.0;   // (its readers passed the barrier that ends the loop)
        {   // every mask entry is rewritten: zeros here, 16 bytes per store (n_actions is odd, so the env's row starts at any
            // byte: the unaligned head and tail go byte by byte); the valid candidates are set to 1 in phase 2, two barriers on
            const int total = K * Mc * S;
            const int head = min((int)((16u - ((uint32_t)(uintptr_t)mask & 15u)) & 15u), total);
            const int nvec = (total - head) >> 4;
            if (tid < head) mask[tid] = 0;
            uint4 *mv = reinterpret_cast<uint4 *>(mask + head);
            for (int i = tid; i < nvec; i += OBS_ENV_THREADS) mv[i] = make_uint4(0u, 0u, 0u, 0u);
            for (int i = head + (nvec << 4) + tid; i < total; i += OBS_ENV_THREADS) mask[i] = 0;
        }

        // ---- phase 0, warp per path: links, availability, free blocks, valid starts of every modulation, usable centres
        links_open_paths<true>(p, t, ar, bm, pbase, rate, warp, lane, obs, path_len_norm);
        env_sync();
        // ---- phase 1: neighbour sums of every usable centre, link by link
        links_neighbour_sums(p, t, ar, bm, lists, tid);
        env_sync();   // X complete
        for (int i = tid; i < E; i += OBS_ENV_THREADS) lmask[i] = 0u;   // (read above, set again after the closing barrier)
        // ---- modulations_to_consider < n_mods: get_max_modulation_index (qrmsa.pyx:543-581) -- paths in order, modulations
        // from the most efficient one, any valid start whose GSNR meets minimum_osnr + margin; never below Mc - 1.  The Mc
        // blocks of a path then stand for modulations max_idx, max_idx - 1, ... (qrmsa.pyx:716-719)
        int maxidx = M - 1;
        if (Mc < M) {
            if (warp == 0) {
                int found = -1;
                for (int pi = 0; pi < K && found < 0; ++pi) {
                    if (phops[pi] == 0) continue;
                    const double2 pg = __ldg(p.path_gn + pbase + pi);
                    for (int mi = 0; mi < M && found < 0; ++mi) {
                        const int m = (M - 1) - mi;
                        const int n = t.need(rate * M + m), ncls = t.cls(rate * M + m);
                        const uint32_t *vrow = validM + (pi * 8 + mi) * VW;
                        bool ok = false;
                        for (int it = 0; it * 32 < S; ++it) {
                            const int s = lane + it * 32;
                            if (s < S && ((vrow[it] >> lane) & 1u))
                                ok |= gn_base(p, t, pg, s, n, ncls).with(X[pi * D + 2 * s + n]) <= t.ACCT(m);
                        }
                        if (__any_sync(FULL, ok)) found = m;
                    }
                }
                if (lane == 0) {
                    *smax = max(found, Mc - 1);
                    p.maxmod[env] = (uint8_t)max(found, Mc - 1);
                }
            }
            env_sync();
            maxidx = *smax;
        }
        // block j of a path is modulation mod_of(j) (allowed_mods[j], qrmsa.pyx:821-829)
        auto mod_of = [&](int j) { return maxidx > 1 ? maxidx - j : (Mc - 1) - j; };
        // ---- phase 2: units = (path, run of modulations that need the same number of slots), handed to the warps by a
        // ticket.  The unit's valid starts are compacted so that every lane holds one; GSNR once per start, then mask
        // bytes and the 12 features per modulation of the unit (qrmsa.pyx:583-781).
        // the runs are the same on every path (slots needed depend on bit rate and modulation only): lane j < Mc holds the
        // slot count of block j, a ballot marks the heads of the runs
        const int my_n = lane < Mc ? t.need(rate * M + mod_of(lane)) : -1;
        const int prev_n = __shfl_up_sync(FULL, my_n, 1);
        const uint32_t heads = __ballot_sync(FULL, lane < Mc && (lane == 0 || prev_n != my_n));
        const int R = __popc(heads);
        const uint32_t rcpR = (65536u + (uint32_t)R - 1u) / (uint32_t)R;   // u / R for the few tickets of a request
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(tick, 1);
            u = __shfl_sync(FULL, u, 0);
            if (u >= K * R) break;
            const int pi = (int)(((uint32_t)u * rcpR) >> 16);
            if (phops[pi] == 0) continue;
            uint32_t h = heads;
            for (int i = u - pi * R; i > 0; --i) h &= h - 1u;
            const int mi0 = __ffs(h) - 1;
            h &= h - 1u;
            const int mi1 = h ? __ffs(h) - 1 : Mc;
            const int n = __shfl_sync(FULL, my_n, mi0);
            const int ncls = t.cls(rate * M + mod_of(mi0));
            const int path = pbase + pi;
            const double2 pg = __ldg(p.path_gn + path);
            const uint32_t *vrow = validM + (pi * 8 + (M - 1) - mod_of(mi0)) * VW;   // bitmaps are kept per modulation, best first
            const double *Xp = X + pi * D + n;
            // compaction: a word of the bitmap is its own ballot
            int cnt = 0;
            for (int it = 0; it * 32 < S; ++it) {
                const uint32_t w = vrow[it];
                if ((w >> lane) & 1u) slist[cnt + __popc(w & ((1u << lane) - 1u))] = (uint16_t)(lane + it * 32);
                cnt += __popc(w);
            }
            __syncwarp();
            // The modulations of a run see the same valid starts with the same GSNR g and differ only in their threshold:
            // the integer statistics (count, sum s, sum s^2, max s) and the moments of y = g - th0 (th0: the run's first
            // threshold) are taken ONCE per unit; a modulation's normalised margin (g - th) / |th| = (y + (th0 - th)) / |th|
            // follows from them in closed form.  What stays per (start, modulation) is the mask bit.  np.round(x, 10) >= 0
            // (osnr.pyx:366; rint(x * 1e10) >= 0, -0.0 included) <=> x * 1e10 >= -0.5 <=> g >= th - 0.5e-10 |th|; the
            // rounding itself (<= 5e-11) is far below the float32 features and is not applied to the moments.
            const double th0 = p.mod_thr_nomargin[mod_of(mi0)];
            uint8_t *mrow0 = mask + (size_t)(pi * Mc + mi0) * S;
            int c_s = 0, c_max = -1;
            uint32_t c_s2 = 0u;   // sum of s^2 over <= 960 slots < 2^32
            double v_sum = 0.0, v_max = -1e300, v_sq = 0.0;
#pragma unroll 1
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int idx = c0 + lane;
                if (idx < cnt) {
                    const int s = slist[idx];
                    const double acc = gn_base(p, t, pg, s, n, ncls).with(Xp[2 * s]);
                    const double g = -10.0 * log10(acc);
                    const double y = g - th0;
                    c_s += s; c_s2 += (uint32_t)(s * s); c_max = s;   // starts are held in ascending order
                    v_sum += y; v_max = fmax(v_max, y); v_sq = fma(y, y, v_sq);
                    for (int mi = mi0; mi < mi1; ++mi) {
                        const double th = p.mod_thr_nomargin[mod_of(mi)];
                        if (g >= fma(-0.5e-10, fabs(th), th)) mrow0[(mi - mi0) * S + s] = 1;
                    }
                }
            }
            c_s = __reduce_add_sync(FULL, c_s); c_s2 = __reduce_add_sync(FULL, c_s2); c_max = __reduce_max_sync(FULL, c_max);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v_sum += __shfl_xor_sync(FULL, v_sum, o);
                v_sq += __shfl_xor_sync(FULL, v_sq, o);
                v_max = fmax(v_max, __shfl_xor_sync(FULL, v_max, o));
            }
            // lane j finishes the features of the unit's j-th modulation
            if (lane < mi1 - mi0) {
                const double cntv = (double)cnt, total_av = pstat[pi * 3 + 0];
                double f_avg = 0, f_std = 0, f_max = 0, best = 0, omean = 0, ovar = 0;
                if (cnt > 0) {
                    const double th = p.mod_thr_nomargin[mod_of(mi0 + lane)], ia = 1.0 / fabs(th), dl = th0 - th;
                    const double icnt = 1.0 / cntv;   // (the features are float32: one reciprocal serves the four means)
                    const double k_sum = fma(cntv, dl, v_sum) * ia;
                    const double k_sq = fma(dl, fma(cntv, dl, 2.0 * v_sum), v_sq) * (ia * ia);
                    f_avg = (double)c_s * icnt; omean = k_sum * icnt; f_max = (double)c_max; best = fmax((v_max + dl) * ia, 0.0);
                    f_std = sqrt(fmax((double)c_s2 * icnt - f_avg * f_avg, 0.0));
                    ovar = fmax(k_sq * icnt - omean * omean, 0.0);
                }
                float *o = obs + 3 + K + (pi * Mc + mi0 + lane) * 12;
                o[0] = (float)(cntv * inv_S);
                o[1] = (float)(f_avg * inv_S1);
                o[2] = (float)(f_std * inv_S1);
                o[3] = (float)fmax(((double)n - 5.5) * (1.0 / 3.5), 0.0);
                o[4] = (float)(2.0 * (total_av - 0.5 * (double)S) * inv_S);
                o[5] = (float)pstat[pi * 3 + 1];
                o[6] = (float)pstat[pi * 3 + 2];
                o[7] = (float)best;
                o[8] = (float)omean;
                o[9] = (float)ovar;
                o[10] = (float)(2.0 * ((total_av * inv_S) - 0.5));
                o[11] = (float)(f_max * inv_S1);
            }
            __syncwarp();   // slist is rewritten by the next unit
        }
        env_sync();   // the next request rewrites X, need, lmask and the path scratch
    }
}

// --------------------------------------------------------------------------------------------------------
// heuristic_highest_snr (heuristics.py:272-328, benchmark heuristic #2) + env.step, n_steps requests per env.
// The reference QoT-checks EVERY valid start of every (path, modulation) and takes the acceptable candidate with
// the highest GSNR.  One CTA per env: per path the neighbour sum X[c2] is built once for every centre frequency
// (as in k_observation), after which a candidate is one lookup; the winner is the smallest acc = 1/GSNR, found by
// block reductions, and warp 0 commits it with the warp-level commit / release code of the first-fit kernel.
// Order of the search = order of the reference's loops (path, modulation descending, start ascending); the first
// candidate wins a tie.  The reference compares GSNR in dB, several 1/GSNR values share one dB value, so a
// runner-up within 1e-6 dB of the winner raises QRMSA_FLAG_NEAR_TIE (reported like the near-threshold flag).
// --------------------------------------------------------------------------------------------------------
// minimum over the OBS_ENV_THREADS threads of one env (named barrier `bar`); red has one entry per warp of the env
__device__ __forceinline__ void env_bar(int bar) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(OBS_ENV_THREADS) : "memory"); }

__device__ __forceinline__ unsigned long long env_min_u64(unsigned long long v, unsigned long long *red, unsigned long long *bc,
                                                          int lane, int warp, int bar) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL, v, o);
        v = y < v ? y : v;
    }
    if (lane == 0) red[warp] = v;
    env_bar(bar);
    if (warp == 0) {
        unsigned long long x = lane < OBS_ENV_THREADS / 32 ? red[lane] : ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(FULL, x, o);
            x = y < x ? y : x;
        }
        if (lane == 0) *bc = x;
    }
    env_bar(bar);
    const unsigned long long r = *bc;
    env_bar(bar);
    return r;
}

__global__ void __launch_bounds__(OBS_ENV_THREADS * OBS_MAX_EPC, 1)
    k_step_highest_snr(const KParams p, const int n_steps, const int epc, const int env_smem) {
    __shared__ uint64_t mbar;
    __shared__ unsigned long long red64_[OBS_MAX_EPC][OBS_ENV_THREADS / 32], bc64_[OBS_MAX_EPC];
    __shared__ unsigned int s_checks_[OBS_MAX_EPC], s_terms_[OBS_MAX_EPC], s_osnr_[OBS_MAX_EPC];
    __shared__ int s_cur_[OBS_MAX_EPC], s_err_[OBS_MAX_EPC];
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<0, 0, 0> dm(p);
    const int S = p.S, M = p.M, K = p.K, D = p.D, CAP = p.CAP;
    // epc envs per CTA, OBS_ENV_THREADS threads each, each env on its own named barrier (see k_observation)
    const int slot = threadIdx.x / OBS_ENV_THREADS, bar = slot + 1;
    unsigned long long *red64 = red64_[slot], &bc64 = bc64_[slot];
    unsigned int &s_checks = s_checks_[slot], &s_terms = s_terms_[slot], &s_osnr = s_osnr_[slot];
    int &s_cur = s_cur_[slot], &s_err = s_err_[slot];
    unsigned char *extra = qsmem + p.blob_bytes + (size_t)slot * env_smem;
    ObsSmem *sm = reinterpret_cast<ObsSmem *>(extra);
    double *X = reinterpret_cast<double *>(extra + sizeof(ObsSmem));          // [D]
    uint32_t *rec = reinterpret_cast<uint32_t *>(X + D + p.S);                  // [Hmax][CAP]  (same carve-up as k_observation)
    const int tid = threadIdx.x - slot * OBS_ENV_THREADS, lane = tid & 31, warp = tid >> 5;
    const int reject = K * M * S;
    const double TIE = 1.0000002302585359;   // 10^(1e-6 / 10): 1e-6 dB on the linear value

    for (int env = blockIdx.x * epc + slot; env < p.n_envs; env += gridDim.x * epc) {
        int4 st = p.estate[env];
        if (st.w != ENV_OK) continue;   // (uniform over the CTA)
        int cur = st.x, rel_ptr = st.y, err = 0;
        uint4 *tr = p.trace + (size_t)env * p.T;
        const unsigned long long *perm = p.perm + (size_t)env * p.T;
        uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
        uint32_t *lists = p.lists + (size_t)env * p.E * CAP;
        uint8_t *pos = p.pos + (size_t)env * p.pos_stride;
        Head head = load_head(p, tr, perm, rel_ptr);
        unsigned long long *cglob = p.counters + (size_t)(env / p.group_size) * QRMSA_N_COUNTERS;
        if (tid == 0) { s_checks = 0u; s_terms = 0u; s_osnr = 0u; }

#pragma unroll 1
        for (int step = 0; step < n_steps && cur + 1 < p.n_req && !err; ++step) {
            const uint4 rq = tr[cur];
            const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
            const int pbase = (src * p.N + dst) * K;
            // per-thread best / runner-up over the candidates this thread evaluates (acc > 0: its bit pattern orders it)
            unsigned long long best = ~0ull, second = ~0ull, near = ~0ull;   // near: smallest acc within 1e-3 dB of its threshold
            int best_idx = 0x7fffffff, best_phys = -1;   // phys: (path, start, slots) -- the same channel under another modulation
            int any_res = 0, any_osnr = 0;
            uint32_t n_checks = 0, n_terms = 0;

            for (int pi = 0; pi < K; ++pi) {
                const int path = pbase + pi;
                const int hops = __ldg(p.path_hops + path) & 0x7f;
                if (hops == 0) continue;
                env_bar(bar);
                if (tid < 32) {
                    const int l = tid < hops ? __ldg(p.path_links + path * p.Hmax + tid) : 0;
                    sm->link[tid] = l;
                    sm->cnt[tid] = tid < hops ? (int)bm[(unsigned)(l * p.RW + p.RW - 1)] : 0;
                    sm->w1[tid] = t.W1(l);
                    sm->w2[tid] = t.W2(l);
                    const uint32_t a = path_available(dm, bm, hops, l, lane);
                    sm->av[tid] = a;
                    valid_starts_all(t, a, rate, M, sm->validM, lane);
                }
                env_bar(bar);
                for (int i = 0; i < hops; ++i) {
                    const int c = sm->cnt[i];
                    const uint32_t *lst = lists + (unsigned)(sm->link[i] * CAP);
                    for (int q = tid; q < c; q += OBS_ENV_THREADS) rec[i * CAP + q] = lst[q];
                }
                env_bar(bar);
                for (int c0 = 0; c0 < D; c0 += OBS_ENV_THREADS) {   // X[c2]: neighbour sum for a candidate centred at c2, where needed
                    const int c2 = c0 + tid;
                    const bool need = c2 < D && centre_needed(t, sm->validM, rate, M, S, c2);
                    if (!__any_sync(FULL, need)) continue;
                    if (!need) continue;
                    double x = 0.0;
                    for (int i = 0; i < hops; ++i) {
                        const int c = sm->cnt[i];
                        const uint32_t *r = rec + i * CAP;
                        double s1 = 0.0, s2 = 0.0;
                        for (int q = 0; q < c; ++q) gn_term(t, D, r[q], c2, s1, s2);
                        n_terms += (uint32_t)c;
                        x = fma(sm->w1[i], s1, x);
                        x = fma(sm->w2[i], s2, x);
                    }
                    X[c2] = x;
                }
                env_bar(bar);   // X complete
                for (int mi = 0; mi < M; ++mi) {
                    const int m = (M - 1) - mi;
                    const int n = t.need(rate * M + m), ncls = t.cls(rate * M + m);
                    if (!sm->validM[mi][31]) any_res = 1;             // no valid start: blocked_resources (heuristics.py:295-297)
                    for (int s = tid; s < S; s += OBS_ENV_THREADS) {
                        if (!((sm->validM[mi][s >> 5] >> (s & 31)) & 1u)) continue;
                        const double acc = gn_base(p, t, path, s, n, ncls).with(X[2 * s + n]);
                        n_checks += 1;
                        const unsigned long long k = (unsigned long long)__double_as_longlong(acc);
                        if (acc > t.ACCLO(m) && acc < t.ACCHI(m) && k < near) near = k;
                        if (acc <= t.ACCT(m)) {                       // gsnr >= threshold (heuristics.py:312-313)
                            const int idx = (pi * M + mi) * S + s;     // = the action index
                            const int phys = (pi << 20) | (s << 8) | n;
                            if (k == best && phys == best_phys) {
                                // identical channel (same path, slots) met again under a lower modulation: same GSNR by
                                // construction, the first one keeps the tie (heuristics.py:314) and it is no runner-up
                            } else if (k < best) { second = best; best = k; best_idx = idx; best_phys = phys; }
                            else if (k < second) second = k;
                        } else {
                            any_osnr = 1;
                        }
                    }
                }
            }
            // ---- winner: smallest acc, first in search order among equals
            const unsigned long long gbest = env_min_u64(best, red64, &bc64, lane, warp, bar);
            const int widx = (int)env_min_u64(best == gbest ? (unsigned long long)(unsigned)best_idx : ~0ull, red64, &bc64, lane, warp, bar);
            const unsigned long long gsecond = env_min_u64((best == gbest && best_idx == widx) ? second : best, red64, &bc64, lane, warp, bar);
            const unsigned long long gnear = env_min_u64(near, red64, &bc64, lane, warp, bar);
            if (__any_sync(FULL, any_osnr) && lane == 0) atomicOr(&s_osnr, 1u);
            env_bar(bar);
            const int g_osnr = (int)s_osnr;
            const bool found = gbest != ~0ull;
            uint32_t flags = QRMSA_FLAG_DECIDED;
            // a check within 1e-3 dB of its threshold matters when flipping it could change the outcome: no acceptable
            // candidate at all, or its GSNR is at least the winner's
            if (gnear != ~0ull && (!found || __longlong_as_double((long long)gnear) <= __longlong_as_double((long long)gbest) * TIE))
                flags |= QRMSA_FLAG_NEAR_THRESHOLD;
            int action = reject;
            if (found) {
                flags |= QRMSA_FLAG_ACCEPTED;
                action = widx;
                if (gsecond != ~0ull && __longlong_as_double((long long)gsecond) <= __longlong_as_double((long long)gbest) * TIE)
                    flags |= QRMSA_FLAG_NEAR_TIE;
            } else {
                if (any_res && !g_osnr) flags |= QRMSA_FLAG_BLOCKED_RESOURCES;   // heuristics.py:324-326
                if (g_osnr) flags |= QRMSA_FLAG_BLOCKED_OSNR;
            }
            // work counters: block totals
            {
                uint32_t v0 = n_checks, v1 = n_terms;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(FULL, v0, o); v1 += __shfl_xor_sync(FULL, v1, o); }
                if (lane == 0) { atomicAdd(&s_checks, v0); atomicAdd(&s_terms, v1); }
            }
            env_bar(bar);
            if (warp == 0) {
                if (found) {
                    const int s = widx % S, mi = (widx / S) % M, pi = widx / (S * M), m = (M - 1) - mi;
                    const int path = pbase + pi;
                    const int hops = __ldg(p.path_hops + path) & 0x7f;
                    const int mylink = lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
                    const int mycnt = lane < hops ? (int)*cnt_word(bm, mylink, p.RW) : 0;
                    const int n = t.need(rate * M + m), ncls = t.cls(rate * M + m);
                    const uint32_t rec_w = (uint32_t)(2 * s + n) | ((uint32_t)n << 12) | ((uint32_t)m << 20) | ((uint32_t)ncls << 23);
                    if (commit(dm, p, bm, lists, pos, hops, mylink, mycnt, s, n, rec_w, lane)) err = ENV_ERR_LIST_OVERFLOW;
                    if (lane == 0 && p.gsnr_log) {
                        const double acc = __longlong_as_double((long long)gbest), ase = gn_base(p, t, path, s, n, ncls).ase;
                        double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                        gl[0] = -10.0 * log10(acc); gl[1] = -10.0 * log10(ase); gl[2] = -10.0 * log10(acc - ase);
                    }
                } else if (lane == 0 && p.gsnr_log) {
                    double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                    gl[0] = gl[1] = gl[2] = 0.0;
                }
                if (lane == 0) tr[cur].w = (uint32_t)action | flags;
                __syncwarp();
                uint32_t n_rel = 0;
                if (advance_and_release(dm, p, t, tr, perm, bm, lists, pos, cur, rel_ptr, head, lane, n_rel))
                    err = ENV_ERR_RELEASE_NOT_FOUND;
                if (lane == 0) {
                    atomicAdd(cglob + QRMSA_CNT_GN_EVALS, (unsigned long long)s_checks);
                    atomicAdd(cglob + QRMSA_CNT_GN_TERMS, (unsigned long long)s_terms);
                    atomicAdd(cglob + QRMSA_CNT_PATHS_TRIED, (unsigned long long)K);
                    if (n_rel) atomicAdd(cglob + QRMSA_CNT_RELEASES, (unsigned long long)n_rel);
                    s_checks = 0u; s_terms = 0u; s_osnr = 0u;
                    s_cur = cur; s_err = err;
                }
            }
            env_bar(bar);
            cur = s_cur;
            err = s_err;
            env_bar(bar);
        }
        if (tid == 0) p.estate[env] = make_int4(cur, rel_ptr, st.z, err);
        env_bar(bar);
    }
}

// --------------------------------------------------------------------------------------------------------
// heuristic_highest_snr + env.step, link-major (spectra up to 320 slots, at most 8 paths per pair): the first two phases
// are the observation kernel's -- warp per path (availability, valid starts of every modulation, usable centres), then the
// neighbour sums once per DISTINCT link of the request into X[p][c2] -- and the candidate scan runs as ticketed units
// (path, run of modulations with the same slot count) on compacted valid starts: one GN base + one lookup per start, one
// threshold test per modulation of the run.  Per request: four env-wide barriers (the round-1 kernel: five per path plus a
// dozen for its four block reductions) and the neighbour sums of shared links are made once instead of once per path.
// Winner = smallest 1/GSNR, the smallest action index among equals; runner-up and near-threshold rules as in
// k_step_highest_snr (heuristics.py:272-328).
// --------------------------------------------------------------------------------------------------------
struct HsnrPick {   // per-lane, per-warp and per-env summary of the candidates seen
    unsigned long long best, second, near;
    int idx, any;   // any: bit 0 some modulation had no valid start, bit 1 some check failed
};
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL, v, o);
        v = y < v ? y : v;
    }
    return v;
}
// min over the lanes: the winner (smallest acc, then smallest action index) and the best of everything else
__device__ __forceinline__ HsnrPick warp_pick(const HsnrPick a) {
    HsnrPick r;
    r.best = warp_min_u64(a.best);
    r.idx = (int)__reduce_min_sync(FULL, a.best == r.best ? (unsigned)a.idx : 0xffffffffu);
    r.second = warp_min_u64((a.best == r.best && a.idx == r.idx) ? a.second : a.best);
    r.near = warp_min_u64(a.near);
    r.any = (int)__reduce_or_sync(FULL, (unsigned)a.any);
    return r;
}

__global__ void __launch_bounds__(OBS_ENV_THREADS * OBS_MAX_EPC, 1)
    k_step_highest_snr_links(const KParams p, const int n_steps, const int epc, const int env_smem) {
    constexpr int nw = OBS_ENV_THREADS >> 5;
    __shared__ uint64_t mbar;
    __shared__ unsigned long long s_red_[OBS_MAX_EPC][nw][3];
    __shared__ int s_redi_[OBS_MAX_EPC][nw][4];
    __shared__ int s_cur_[OBS_MAX_EPC], s_err_[OBS_MAX_EPC];
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<0, 0, 0> dm(p);
    const int S = p.S, W = p.W, M = p.M, K = p.K, D = p.D, CAP = p.CAP, E = p.E;
    const int VW = W + 1, NW = (D + 31) >> 5;
    const int slot = threadIdx.x / OBS_ENV_THREADS;
    auto env_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(OBS_ENV_THREADS) : "memory"); };
    LinkArea ar;
    ar.carve(qsmem + p.blob_bytes + (size_t)slot * env_smem, K, D, VW, NW, E);
    const int tid = threadIdx.x - slot * OBS_ENV_THREADS, lane = tid & 31, warp = tid >> 5;
    uint16_t *slist = reinterpret_cast<uint16_t *>(ar.scr) + warp * S;
    unsigned long long (*s_red)[3] = s_red_[slot];
    int (*s_redi)[4] = s_redi_[slot];
    int &s_cur = s_cur_[slot], &s_err = s_err_[slot];
    const int reject = K * M * S;
    const double TIE = 1.0000002302585359;   // 10^(1e-6 / 10): 1e-6 dB on the linear value

    for (int i = tid; i < E; i += OBS_ENV_THREADS) ar.lmask[i] = 0u;
    env_sync();
    for (int env = blockIdx.x * epc + slot; env < p.n_envs; env += gridDim.x * epc) {
        const int4 st = p.estate[env];
        if (st.w != ENV_OK) continue;   // (uniform over the env's threads)
        int cur = st.x, rel_ptr = st.y, err = 0;
        uint4 *tr = p.trace + (size_t)env * p.T;
        const unsigned long long *perm = p.perm + (size_t)env * p.T;
        uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
        uint32_t *lists = p.lists + (size_t)env * E * CAP;
        uint8_t *pos = p.pos + (size_t)env * p.pos_stride;
        Head head;
        head.id = -1; head.rel = 0.f;
        if (warp == 0) head = load_head(p, tr, perm, rel_ptr);
        unsigned long long tot_checks = 0ull, tot_terms = 0ull;   // (warp 0)
        uint32_t tot_rel = 0u, tot_steps = 0u;

#pragma unroll 1
        for (int step = 0; step < n_steps && cur + 1 < p.n_req && !err; ++step) {
            const uint4 rq = tr[cur];
            const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff, rate = (rq.z >> 16) & 0xff;
            const int pbase = (src * p.N + dst) * K;
            if (tid == 0) *ar.tick = 0;
            for (int i = tid; i < K * D; i += OBS_ENV_THREADS) ar.X[i] = 0.0;
            links_open_paths<false>(p, t, ar, bm, pbase, rate, warp, lane, nullptr, nullptr);
            env_sync();
            uint32_t n_terms = links_neighbour_sums(p, t, ar, bm, lists, tid);
            env_sync();   // X complete
            for (int i = tid; i < E; i += OBS_ENV_THREADS) ar.lmask[i] = 0u;   // (set again after the barrier that closes the step)
            // ---- candidates: lane j < M holds the slot count of modulation M-1-j, a ballot marks the heads of the runs
            const int my_n = lane < M ? t.need(rate * M + (M - 1) - lane) : -1;
            const int prev_n = __shfl_up_sync(FULL, my_n, 1);
            const uint32_t heads = __ballot_sync(FULL, lane < M && (lane == 0 || prev_n != my_n));
            const int R = __popc(heads);
            const uint32_t rcpR = (65536u + (uint32_t)R - 1u) / (uint32_t)R;
            HsnrPick pk;
            pk.best = pk.second = pk.near = ~0ull;
            pk.idx = 0x7fffffff; pk.any = 0;
            int best_phys = -1;   // (path, start, slots): the same channel under another modulation
            uint32_t n_checks = 0u;
            for (;;) {
                int u = 0;
                if (lane == 0) u = atomicAdd(ar.tick, 1);
                u = __shfl_sync(FULL, u, 0);
                if (u >= K * R) break;
                const int pi = (int)(((uint32_t)u * rcpR) >> 16);
                if (ar.phops[pi] == 0) continue;
                uint32_t h = heads;
                for (int i = u - pi * R; i > 0; --i) h &= h - 1u;
                const int mi0 = __ffs(h) - 1;
                h &= h - 1u;
                const int mi1 = h ? __ffs(h) - 1 : M;
                const int n = __shfl_sync(FULL, my_n, mi0);
                const int ncls = t.cls(rate * M + (M - 1) - mi0);
                const double2 pg = __ldg(p.path_gn + pbase + pi);
                const uint32_t *vrow = ar.validM + (pi * 8 + mi0) * VW;
                const double *Xp = ar.X + pi * D + n;
                int cnt = 0;
                for (int it = 0; it * 32 < S; ++it) {
                    const uint32_t w = vrow[it];
                    if ((w >> lane) & 1u) slist[cnt + __popc(w & ((1u << lane) - 1u))] = (uint16_t)(lane + it * 32);
                    cnt += __popc(w);
                }
                __syncwarp();
                if (cnt == 0) pk.any |= 1;   // no valid start: blocked_resources (heuristics.py:295-297)
#pragma unroll 1
                for (int c0 = 0; c0 < cnt; c0 += 32) {
                    const int idx = c0 + lane;
                    if (idx < cnt) {
                        const int s = slist[idx];
                        const double acc = gn_base(p, t, pg, s, n, ncls).with(Xp[2 * s]);
                        const unsigned long long k = (unsigned long long)__double_as_longlong(acc);   // acc > 0: its bits order it
                        const int phys = (pi << 20) | (s << 8) | n;
                        for (int mi = mi0; mi < mi1; ++mi) {
                            const int m = (M - 1) - mi;
                            n_checks += 1u;
                            if (acc > t.ACCLO(m) && acc < t.ACCHI(m) && k < pk.near) pk.near = k;
                            if (acc <= t.ACCT(m)) {                       // gsnr >= threshold (heuristics.py:312-313)
                                const int a = (pi * M + mi) * S + s;       // = the action index
                                if (k == pk.best && phys == best_phys) {
                                    // the identical channel under a lower modulation: same GSNR by construction, the first
                                    // one keeps the tie (heuristics.py:314) and it is no runner-up
                                    if (a < pk.idx) pk.idx = a;
                                } else if (k < pk.best || (k == pk.best && a < pk.idx)) {
                                    pk.second = pk.best; pk.best = k; pk.idx = a; best_phys = phys;
                                } else if (k < pk.second) {
                                    pk.second = k;
                                }
                            } else {
                                pk.any |= 2;
                            }
                        }
                    }
                }
                __syncwarp();   // slist is rewritten by the next unit
            }
            // ---- winner: per warp by shuffles, then warp 0 over the env's warps
            {
                const HsnrPick wp = warp_pick(pk);
                n_checks = __reduce_add_sync(FULL, n_checks);
                n_terms = __reduce_add_sync(FULL, n_terms);
                if (lane == 0) {
                    s_red[warp][0] = wp.best; s_red[warp][1] = wp.second; s_red[warp][2] = wp.near;
                    s_redi[warp][0] = wp.idx; s_redi[warp][1] = wp.any; s_redi[warp][2] = (int)n_checks; s_redi[warp][3] = (int)n_terms;
                }
            }
            env_sync();
            if (warp == 0) {
                HsnrPick e;
                e.best = e.second = e.near = ~0ull;
                e.idx = 0x7fffffff; e.any = 0;
                uint32_t c_checks = 0u, c_terms = 0u;
                if (lane < nw) {
                    e.best = s_red[lane][0]; e.second = s_red[lane][1]; e.near = s_red[lane][2];
                    e.idx = s_redi[lane][0]; e.any = s_redi[lane][1];
                    c_checks = (uint32_t)s_redi[lane][2]; c_terms = (uint32_t)s_redi[lane][3];
                }
                const HsnrPick g = warp_pick(e);
                tot_checks += __reduce_add_sync(FULL, c_checks);
                tot_terms += __reduce_add_sync(FULL, c_terms);
                const bool found = g.best != ~0ull;
                uint32_t flags = QRMSA_FLAG_DECIDED;
                // a check within 1e-3 dB of its threshold matters when flipping it could change the outcome: no acceptable
                // candidate at all, or its GSNR is at least the winner's
                if (g.near != ~0ull && (!found || __longlong_as_double((long long)g.near) <= __longlong_as_double((long long)g.best) * TIE))
                    flags |= QRMSA_FLAG_NEAR_THRESHOLD;
                int action = reject;
                if (found) {
                    flags |= QRMSA_FLAG_ACCEPTED;
                    action = g.idx;
                    if (g.second != ~0ull && __longlong_as_double((long long)g.second) <= __longlong_as_double((long long)g.best) * TIE)
                        flags |= QRMSA_FLAG_NEAR_TIE;
                    const int s = g.idx % S, mi = (g.idx / S) % M, pi = g.idx / (S * M), m = (M - 1) - mi;
                    const int path = pbase + pi;
                    const int hops = __ldg(p.path_hops + path) & 0x7f;
                    const int mylink = lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
                    const int mycnt = lane < hops ? (int)*cnt_word(bm, mylink, p.RW) : 0;
                    const int n = t.need(rate * M + m), ncls = t.cls(rate * M + m);
                    const uint32_t rec_w = (uint32_t)(2 * s + n) | ((uint32_t)n << 12) | ((uint32_t)m << 20) | ((uint32_t)ncls << 23);
                    if (commit(dm, p, bm, lists, pos, hops, mylink, mycnt, s, n, rec_w, lane)) err = ENV_ERR_LIST_OVERFLOW;
                    if (lane == 0 && p.gsnr_log) {
                        const double acc = __longlong_as_double((long long)g.best), ase = gn_base(p, t, path, s, n, ncls).ase;
                        double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                        gl[0] = -10.0 * log10(acc); gl[1] = -10.0 * log10(ase); gl[2] = -10.0 * log10(acc - ase);
                    }
                } else {
                    if ((g.any & 1) && !(g.any & 2)) flags |= QRMSA_FLAG_BLOCKED_RESOURCES;   // heuristics.py:324-326
                    if (g.any & 2) flags |= QRMSA_FLAG_BLOCKED_OSNR;
                    if (lane == 0 && p.gsnr_log) {
                        double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                        gl[0] = gl[1] = gl[2] = 0.0;
                    }
                }
                if (lane == 0) tr[cur].w = (uint32_t)action | flags;
                __syncwarp();
                uint32_t n_rel = 0;
                if (advance_and_release(dm, p, t, tr, perm, bm, lists, pos, cur, rel_ptr, head, lane, n_rel))
                    err = ENV_ERR_RELEASE_NOT_FOUND;
                tot_rel += n_rel;
                tot_steps += 1u;
                if (lane == 0) { s_cur = cur; s_err = err; }
            }
            env_sync();   // the step's commit and releases are visible to every warp of the env
            cur = s_cur;
            err = s_err;
        }
        if (tid == 0) {
            p.estate[env] = make_int4(cur, rel_ptr, st.z, err);
            unsigned long long *cglob = p.counters + (size_t)(env / p.group_size) * QRMSA_N_COUNTERS;
            atomicAdd(cglob + QRMSA_CNT_GN_EVALS, tot_checks);
            atomicAdd(cglob + QRMSA_CNT_GN_TERMS, tot_terms);
            atomicAdd(cglob + QRMSA_CNT_PATHS_TRIED, (unsigned long long)tot_steps * (unsigned long long)K);
            if (tot_rel) atomicAdd(cglob + QRMSA_CNT_RELEASES, (unsigned long long)tot_rel);
        }
        env_sync();
    }
}

// --------------------------------------------------------------------------------------------------------
// reset (qrmsa.pyx:427-504): every slot free, lists empty, release pointer / request index / episode counters 0
// --------------------------------------------------------------------------------------------------------
__global__ void k_reset(const KParams p) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t words = (size_t)p.n_envs * p.bm_stride;
    for (size_t i = tid; i < words; i += nthreads) {
        const int j = (int)(i % p.RW);   // word of the link row; the count word (RW-1) and the padding start at 0
        uint32_t v = 0u;
        if (j < p.W) {
            const int left = p.S - (j << 5);
            v = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
        }
        p.bm[i] = v;
    }
    const size_t nrec = (size_t)p.n_envs * p.E * p.CAP;
    for (size_t i = tid; i < nrec; i += nthreads) p.lists[i] = p.sentinel;
    for (size_t i = tid; i < (size_t)p.n_envs; i += nthreads) {
        p.estate[i] = make_int4(0, 0, 0, 0);
        p.counted[i] = 0u;
        p.maxmod[i] = (uint8_t)(p.M - 1);
    }
}

// Request-major SoA [n_req][n_envs] -> per-env AoS records, through a shared-memory tile so that both the
// reads (consecutive envs) and the writes (consecutive requests of one env) are coalesced.
__global__ void k_ingest_trace(const KParams p, const uint8_t *__restrict__ src, const uint8_t *__restrict__ dst,
                               const uint8_t *__restrict__ rate, const float *__restrict__ arrival,
                               const float *__restrict__ holding, const int n_req) {
    __shared__ uint4 tile[32][33];
    const int e0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, e = e0 + tx;
        if (r < n_req && e < p.n_envs) {
            const size_t o = (size_t)r * p.n_envs + e;
            uint4 v;
            v.x = __float_as_uint(arrival[o]);
            v.y = __float_as_uint(holding[o]);
            v.z = (uint32_t)src[o] | ((uint32_t)dst[o] << 8) | ((uint32_t)rate[o] << 16);
            v.w = 0u;
            tile[rr][tx] = v;
        }
    }
    __syncthreads();
    for (int ee = ty; ee < 32; ee += 8) {
        const int e = e0 + ee, r = r0 + tx;
        if (r < n_req && e < p.n_envs) p.trace[(size_t)e * p.T + r] = tile[tx][ee];
    }
}

// --------------------------------------------------------------------------------------------------------
// On-device request generation: the draws of QRMSAEnv._next_service / _get_node_pair
// (qrmsa.pyx:1079-1099, :1134-1148) in the same order and arithmetic -- exponential inter-arrival and holding
// times rounded to float32, the float32 clock, source / destination / bit rate by bisecting cumulative weights --
// but from a counter-based Philox4x32-10 stream keyed by (seed, global env index, request index) instead of
// CPython's MT19937.  Streams are reproducible and independent of how envs are sharded over GPUs; they are NOT
// the reference's streams (use the host generator + qrmsa_load_trace for replay parity).
// --------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// 53-bit uniform in [0, 1) from two words, as CPython's random() builds it
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ int bisect_right_dev(const double *__restrict__ a, double x, int hi) {
    int lo = 0;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x < a[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One warp per env: the draws of 32 consecutive requests are made in parallel (they depend only on the request
// index), then the float32 clock -- at_k = float32(at_{k-1} + x_k), a rounding chain that cannot be reassociated --
// is walked through the 32 inter-arrival times; the records leave as one coalesced 512-byte store.
__global__ void k_generate_trace(const KParams p, const unsigned long long seed, const unsigned long long pos0,
                                 const long long env_offset, const double *__restrict__ load, const double mean_holding,
                                 const double *__restrict__ src_cum, const double *__restrict__ dst_cum,
                                 const double *__restrict__ rate_cum, float *__restrict__ clock, const int n_req) {
    const int lane = threadIdx.x & 31;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= p.n_envs) return;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long ge = (unsigned long long)(env_offset + e);
    const double mean_iat = 1.0 / (load[e] / mean_holding);   // qrmsa.pyx:1130 (holding time as float32, set by the host)
    const double lam = 1.0 / mean_iat, lam_hold = 1.0 / mean_holding;
    const int N = p.N, R = p.R;
    double now = (double)clock[e];
    uint4 *tr = p.trace + (size_t)e * p.T;
    for (int k0 = 0; k0 < n_req; k0 += 32) {
        const int k = k0 + lane;
        const unsigned long long c = pos0 + (unsigned long long)k;
        const uint4 a = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)ge, (uint32_t)(ge >> 32) * 2u), key);
        const uint4 b = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)ge, (uint32_t)(ge >> 32) * 2u + 1u), key);
        const double x = -log(1.0 - u53(a.x, a.y)) / lam;                        // qrmsa.pyx:1079-1081
        const float ht = (float)(-log(1.0 - u53(a.z, a.w)) / lam_hold);          // qrmsa.pyx:1083-1084
        const int src = bisect_right_dev(src_cum, ((double)b.x * (1.0 / 4294967296.0)) * src_cum[N - 1], N - 1);
        const double *dc = dst_cum + (size_t)src * N;
        const int dst = bisect_right_dev(dc, ((double)b.y * (1.0 / 4294967296.0)) * dc[N - 1], N - 1);
        const int rate = bisect_right_dev(rate_cum, ((double)b.z * (1.0 / 4294967296.0)) * rate_cum[R - 1], R - 1);
        float at = 0.f;
        const int m = min(32, n_req - k0);
        for (int j = 0; j < m; ++j) {   // the clock chain, same on every lane
            const float aj = (float)(now + __shfl_sync(FULL, x, j));
            now = (double)aj;
            if (lane == j) at = aj;
        }
        if (k < n_req)
            tr[k] = make_uint4(__float_as_uint(at), __float_as_uint(ht), (uint32_t)src | ((uint32_t)dst << 8) | ((uint32_t)rate << 16), 0u);
    }
    if (lane == 0) clock[e] = (float)now;
}

// request records [first, first+count) -> request-major SoA [count][n_envs] (the inverse of k_ingest_trace)
__global__ void k_gather_trace(const KParams p, const int first, const int count, uint8_t *__restrict__ src,
                               uint8_t *__restrict__ dst, uint8_t *__restrict__ rate, float *__restrict__ arrival,
                               float *__restrict__ holding) {
    __shared__ uint4 tile[32][33];
    const int e0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int ee = ty; ee < 32; ee += 8) {
        const int e = e0 + ee, r = r0 + tx;
        if (r < count && e < p.n_envs) tile[ee][tx] = p.trace[(size_t)e * p.T + first + r];
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, e = e0 + tx;
        if (r < count && e < p.n_envs) {
            const uint4 v = tile[tx][rr];
            const size_t o = (size_t)r * p.n_envs + e;
            arrival[o] = __uint_as_float(v.x);
            holding[o] = __uint_as_float(v.y);
            src[o] = (uint8_t)(v.z & 0xff); dst[o] = (uint8_t)((v.z >> 8) & 0xff); rate[o] = (uint8_t)((v.z >> 16) & 0xff);
        }
    }
}

// Release schedule: per env, request ids sorted by (float32(arrival + holding), id) -- the key of the
// reference's heap (qrmsa.pyx:1327-1330).  One CTA per env, bitonic sort of 64-bit keys in shared memory.
__global__ void __launch_bounds__(1024) k_build_schedule(const KParams p, const int n_req, const int n_pad) {
    extern __shared__ __align__(16) unsigned long long keys[];
    for (int env = blockIdx.x; env < p.n_envs; env += gridDim.x) {
        const uint4 *tr = p.trace + (size_t)env * p.T;
        for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
            unsigned long long k = ~0ull;
            if (i < n_req) {
                const uint4 r = tr[i];
                const float rel = __fadd_rn(__uint_as_float(r.x), __uint_as_float(r.y));
                k = ((unsigned long long)__float_as_uint(rel) << 32) | (unsigned)i;  // times are >= 0
            }
            keys[i] = k;
        }
        __syncthreads();
        for (int k = 2; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = keys[i], b = keys[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        unsigned long long *perm = p.perm + (size_t)env * p.T;
        for (int i = threadIdx.x; i < n_req; i += blockDim.x) perm[i] = keys[i];
        __syncthreads();
    }
}

// action words of requests [first, first+count) -> request-major int32 [count][n_envs]
__global__ void k_gather_actions(const KParams p, const int first, const int count, int32_t *__restrict__ out) {
    __shared__ uint32_t tile[32][33];
    const int e0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int ee = ty; ee < 32; ee += 8) {
        const int e = e0 + ee, r = r0 + tx;
        if (r < count && e < p.n_envs) tile[ee][tx] = p.trace[(size_t)e * p.T + first + r].w;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, e = e0 + tx;
        if (r < count && e < p.n_envs) out[(size_t)r * p.n_envs + e] = (int32_t)tile[tx][rr];
    }
}

__global__ void k_gather_gsnr(const KParams p, const int first, const int count, const int comp,
                              double *__restrict__ out) {
    const size_t n = (size_t)count * p.n_envs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / p.n_envs), e = (int)(i % p.n_envs);
        out[i] = p.gsnr_log[((size_t)e * p.T + first + r) * 3 + comp];
    }
}

// Counters that follow from the decision log (decided / accepted / rejected / bit rates / hops / modulation
// histogram / near-threshold and blocked-by flags), summed after a step launch over the requests each env
// decided since the last count; also maintains the env's accepted total (estate.z).  One warp per env at a time,
// coalesced reads of the 16-byte request records.
__global__ void k_count_decisions(const KParams p) {
    constexpr int NL = 20;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int per = (p.n_envs + nwarps - 1) / nwarps;
    const int e0 = warp * per, e1 = min(e0 + per, p.n_envs);
    const int *rate_tab = reinterpret_cast<const int *>(p.blob + lay::RATE);
    const int MS = p.M * p.S;
    // c[3] / c[4] (bit rate x 1000, up to 1e6 per request) stay unused: the two rate sums are kept and reduced in
    // 64 bits (rate_req / rate_prov) -- a warp covers ceil(n_envs / warps) envs x every step since the last count,
    // which overflows 32 bits from a few thousand decisions per lane on
    uint32_t c[NL];
    unsigned long long rate_req = 0ull, rate_prov = 0ull;
    int grp = -1;
    auto flush = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rate_req += __shfl_xor_sync(FULL, rate_req, o);
            rate_prov += __shfl_xor_sync(FULL, rate_prov, o);
        }
        if (lane == 0 && grp >= 0) {
            if (rate_req) atomicAdd(p.counters + (size_t)grp * QRMSA_N_COUNTERS + QRMSA_CNT_RATE_REQUESTED, rate_req);
            if (rate_prov) atomicAdd(p.counters + (size_t)grp * QRMSA_N_COUNTERS + QRMSA_CNT_RATE_PROVISIONED, rate_prov);
        }
        rate_req = rate_prov = 0ull;
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            uint32_t v = c[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            const int slot = k == 0 ? QRMSA_CNT_DECIDED : k == 1 ? QRMSA_CNT_ACCEPTED : k == 2 ? QRMSA_CNT_REJECTED :
                             k == 3 ? -1 : k == 4 ? -1 :
                             k == 5 ? QRMSA_CNT_HOPS_ACCEPTED : k == 6 ? QRMSA_CNT_NEAR_THRESHOLD :
                             k == 7 ? QRMSA_CNT_BLOCKED_RESOURCES : k == 8 ? QRMSA_CNT_BLOCKED_OSNR :
                             k == 9 ? QRMSA_CNT_ERRORS : k == 10 ? -1 : k == 11 ? -1 : QRMSA_CNT_MOD_HIST + (k - 12);
            if (lane == 0 && v && slot >= 0 && grp >= 0)
                atomicAdd(p.counters + (size_t)grp * QRMSA_N_COUNTERS + slot, (unsigned long long)v);
            c[k] = 0u;
        }
    };
#pragma unroll
    for (int k = 0; k < NL; ++k) c[k] = 0u;
    for (int env = e0; env < e1; ++env) {
        const int g = env / p.group_size;
        if (g != grp) {
            flush();
            grp = g;
        }
        const int4 es = p.estate[env];
        const int first = (int)p.counted[env], last = es.x;
        const uint4 *tr = p.trace + (size_t)env * p.T;
        uint32_t acc_here = 0u;
        for (int i = first + lane; i < last; i += 32) {
            const uint4 rq = tr[i];
            if (!(rq.w & QRMSA_FLAG_DECIDED)) continue;
            const int rate = rate_tab[(rq.z >> 16) & 0xff];
            c[0] += 1u;
            rate_req += (unsigned long long)rate;
            if (rq.w & QRMSA_FLAG_NEAR_THRESHOLD) c[6] += 1u;
            if (rq.w & QRMSA_FLAG_ACCEPTED) {
                const uint32_t a = rq.w & QRMSA_ACTION_MASK;
                const int pi = a / MS, m = (p.M - 1) - (int)((a / p.S) % p.M);
                const int path = ((rq.z & 0xff) * p.N + ((rq.z >> 8) & 0xff)) * p.K + pi;
                c[1] += 1u;
                acc_here += 1u;
                rate_prov += (unsigned long long)rate;
                c[5] += (uint32_t)(__ldg(p.path_hops + path) & 0x7f);
#pragma unroll
                for (int k = 0; k < 8; ++k) c[12 + k] += (m == k) ? 1u : 0u;
            } else {
                c[2] += 1u;
                if (rq.w & QRMSA_FLAG_BLOCKED_RESOURCES) c[7] += 1u;
                if (rq.w & QRMSA_FLAG_BLOCKED_OSNR) c[8] += 1u;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc_here += __shfl_xor_sync(FULL, acc_here, o);
        if (lane == 0) {
            if (es.w != ENV_OK && first < last) c[9] += 1u;   // an env that stopped on an error during this span
            if (acc_here) p.estate[env].z = es.z + (int)acc_here;
            p.counted[env] = (uint32_t)last;
        }
    }
    flush();
}

// reset(options={"only_episode_counters": True}) empties the release heap (qrmsa.pyx:433): every accepted service whose
// schedule entry is still ahead of the release pointer stays in the network for good.  One warp per env.
__global__ void k_cancel_releases(const KParams p) {
    const int lane = threadIdx.x & 31;
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= p.n_envs) return;
    const int4 st = p.estate[env];
    uint4 *tr = p.trace + (size_t)env * p.T;
    const unsigned long long *perm = p.perm + (size_t)env * p.T;
    for (int i = st.y + lane; i < p.n_req; i += 32) {
        const int id = (int)(unsigned)perm[i];
        if (id < st.x && (tr[id].w & QRMSA_FLAG_ACCEPTED)) tr[id].w |= QRMSA_FLAG_RELEASE_CANCELLED;
    }
}

// calculate_osnr for a hypothetical channel on one env (core/osnr.pyx:21-142); one warp.
__global__ void k_probe_gsnr(const KParams p, const int env, const int src, const int dst, const int pi, const int s,
                             const int n, double *out) {
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    const int path = (src * p.N + dst) * p.K + pi;
    const int hops = __ldg(p.path_hops + path) & 0x7f;
    const uint32_t *lists = p.lists + (size_t)env * p.E * p.CAP;
    const uint32_t *bm = p.bm + (size_t)env * p.bm_stride;
    const int mylink = lane < hops ? __ldg(p.path_links + path * p.Hmax + lane) : 0;
    const int mycnt = lane < hops ? (int)*cnt_word(bm, mylink, p.RW) : 0;
    // class of n: search the class table through NEED/CLS
    int ncls = -1;
    for (int i = 0; i < p.R * p.M; ++i)
        if (t.need(i) == n) ncls = t.cls(i);
    double g = nan(""), g_ase = nan(""), g_nli = nan("");
    if (ncls >= 0 && hops > 0) {
        uint32_t terms = 0;
        const GnBase gb = gn_base(p, t, path, s, n, ncls);
        const double acc = gb.with(gn_neighbours(Dim<0, 0, 0>(p), t, lists, hops, mylink, mycnt, 2 * s + n, lane, terms));
        g = -10.0 * log10(acc);                  // total, ASE-only and NLI-only figures (osnr.pyx:133-140)
        g_ase = -10.0 * log10(gb.ase);
        g_nli = -10.0 * log10(acc - gb.ase);
    }
    if (lane == 0) { out[0] = g; out[1] = g_ase; out[2] = g_nli; }
}

}  // namespace qrmsa
