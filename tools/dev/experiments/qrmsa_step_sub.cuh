// qrmsa_step_sub.cuh -- fused heuristic + env.step, LPE LANES PER ENVIRONMENT (4 up to 479 slots, 8 up to 991).
//
// Why not a warp per env: at 320 slots a link row is 10 words and a link carries ~28 channels, so most of a
// 32-lane warp idles through the control flow of one env.  Here a warp carries 32/LPE envs; lane `sub` of an env
// holds words 4*sub..4*sub+3 of every link row (one 16-byte load per lane, one 64/128-byte line per env), and the
// envs of a warp run DECOUPLED request streams through one warp-synchronous state machine:
//
//     while any env is alive:
//        SEARCH   every env without a candidate advances stage by stage -- finish the previous request (log the
//                 decision, next request, releases due) -> load the request -> open path pi (AND of the link rows)
//                 -> probe modulation m (run of n+1 free slots, empty-network pruning) -- until it holds a candidate
//                 (path, modulation, slot) that needs a QoT check, or has used up its steps
//        GN       all candidates are evaluated together: one flattened loop over (link, 4*LPE-record chunk)
//        DECIDE   accept -> commit -> FINISH; refuse -> next modulation
//
// Every branch is warp-uniform (taken on a vote), per-env activity is a predicate, so all shuffles/votes run
// converged with the full mask and width LPE.  The expensive stages (GN sum, release, commit) therefore execute
// with most lanes busy instead of once per env.
//
// Channel lists are padded with a filler record (class NC: G row of zeros, PHIN 0) so a chunk needs no bounds
// test: entries at or past the link's count contribute exactly +0.0.
#pragma once
#include "qrmsa_kernels.cuh"

namespace qrmsa {

#ifndef QRMSA_SUB_THREADS
#define QRMSA_SUB_THREADS 896
#endif

enum SubState : int { ST_DONE = 0, ST_FINISH = 1, ST_REQ = 2, ST_PATH = 3, ST_PROBE = 4 };

// link i of a path; link ids are bytes, 4 per word, word j in lane j of the env (LPE = 4: words 4..7 in lw1)
template <int LPE>
__device__ __forceinline__ int link_at(uint32_t lw0, uint32_t lw1, int i) {
    uint32_t w;
    if (LPE == 4) w = __shfl_sync(FULL, i < 16 ? lw0 : lw1, (i >> 2) & 3, 4);
    else w = __shfl_sync(FULL, lw0, (i >> 2) & 7, 8);
    return (int)((w >> ((i & 3) << 3)) & 0xffu);
}

// bits [s, e) of the row that fall into this lane's four words
__device__ __forceinline__ uint4 range_mask4(int s, int e, int sub) {
    const int j = sub << 2;
    return make_uint4(range_mask(s, e, j), range_mask(s, e, j + 1), range_mask(s, e, j + 2), range_mask(s, e, j + 3));
}

template <int LPE>
__device__ __forceinline__ int quad_min(int v) {
    v = min(v, __shfl_xor_sync(FULL, v, 1));
    v = min(v, __shfl_xor_sync(FULL, v, 2));
    if (LPE == 8) v = min(v, __shfl_xor_sync(FULL, v, 4));
    return v;
}

template <int LPE>
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    if (LPE == 8) v += __shfl_xor_sync(FULL, v, 4);
    return v;
}

// the empty-network part of 1/GSNR for a candidate (see GnBase), path constants from the 64-byte path record
__device__ __forceinline__ GnBase gn_base_rec(const KParams &p, const Tab &t, int path, int s, int n, int ncls) {
    const double2 pg = __ldg(reinterpret_cast<const double2 *>(p.prec + (size_t)path * 4 + 2));
    const double fc = p.f0 + (p.sb * (double)s) + (p.sb * ((double)n / 2.0));  // heuristics.py:948-951
    GnBase b;
    b.ase = t.ASEC(ncls) * fc * pg.x;
    b.cn = t.CN(ncls);
    b.selfpb = t.SELF(ncls) * pg.y;
    return b;
}

constexpr int SUB_HCAP = 16;   // hops per path the per-env link scratch holds (qrmsa_create routes longer paths to k_step_policy)

template <int LPE, int S_, int M_, int K_>
__global__ void __launch_bounds__(QRMSA_SUB_THREADS, 1) k_step_sub(const KParams p, const int n_steps) {
    __shared__ uint64_t mbar;
    stage_tables(p, &mbar);
    Tab t;
    t.init();
    const Dim<S_, M_, K_> dm(p);
    const int S = dm.S(), M = dm.M(), K = dm.K(), D = dm.D(), CAP = dm.CAP();
    constexpr int RW = 4 * LPE, EPW = 32 / LPE, NCR = 8 / LPE, NONE = 0x7fff;
    const int lane = threadIdx.x & 31, sub = lane & (LPE - 1);
    const int reject = K * M * S;
    const int n_groups = (p.n_envs + EPW - 1) / EPW;
    const int vs_word = S >> 5;
    const uint32_t vs_bit = 1u << (S & 31);
    // per-env scratch in shared memory: {link id, channel count} of every hop of the open path, written when the path
    // is opened (the counts cannot change before this env's own commit), read by the GN sum and the commit
    uint2 *sc = reinterpret_cast<uint2 *>(qsmem + p.blob_bytes) + ((threadIdx.x >> 5) * EPW + (lane / LPE)) * SUB_HCAP;

    // Work counters kept in registers: local slot c lives in lane (c % LPE), register (c / LPE).  Everything that
    // follows from the decision log is counted by k_count_decisions after the launch.
    enum { LC_LINKS_READ = 0, LC_RECORDS_READ, LC_GN_TERMS, LC_GN_EVALS, LC_GN_PRUNED, LC_PATHS_TRIED, LC_RELEASES, LC_N };
#define QC(slot, v) cr[(slot) / LPE] += (sub == ((slot) % LPE)) ? (uint32_t)(v) : 0u

    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(p.work, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const int env = min(g * EPW + lane / LPE, p.n_envs - 1);
        const bool real = g * EPW + lane / LPE < p.n_envs;

        // per-env arrays (addresses are cheap to form again; they are not carried through the loop)
#define TR (p.trace + (size_t)env * p.T)
#define PERM (p.perm + (size_t)env * p.T)
#define BM (p.bm + (size_t)env * p.bm_stride)
#define LISTS (p.lists + (size_t)env * p.E * CAP)
#define POS (p.pos + (size_t)env * p.pos_stride)

        // ---- env state
        int st = ST_DONE, cur = 0, cur_end = 0, rel_ptr = 0, err = 0;
        Head head;
        head.id = -1;
        head.rel = 0.f;
        if (real) {
            const int4 es = p.estate[env];
            cur = es.x; rel_ptr = es.y; err = es.w;
            cur_end = min(cur + n_steps, p.n_req - 1);   // requests [cur, cur_end) are decided by this launch
            if (err == ENV_OK && cur < cur_end) {
                st = ST_REQ;
                head = load_head(p, TR, PERM, rel_ptr);
            }
        }
        const bool was_alive = st != ST_DONE;
        uint32_t cr[NCR];
#pragma unroll
        for (int j = 0; j < NCR; ++j) cr[j] = 0u;

        // ---- request / path / candidate registers
        int rate = 0, pbase = 0, pi = 0, m = 0, action = reject;
        uint32_t flags = 0u;   // QRMSA_FLAG_* of the request being decided
        uint32_t pm = 0u;      // per (path, bit rate): modulations always / never refused on the empty-network bound
        bool cand = false, prunable = false, counted = false;
        int hops = 0, a = 1;
        uint32_t lw0 = 0u, lw1 = 0u;
        uint4 av = make_uint4(0u, 0u, 0u, 0u), r = av;
        int cs = 0, cn = 1, ccls = 0;
#ifdef QRMSA_SUB_STATS
        uint32_t dbg_rounds = 0, dbg_cands = 0, dbg_search = 0, dbg_gn = 0, dbg_alive = 0;
#endif

        // One round = every env advances to its next QoT check (or finds that its open path has none) and runs it.
        while (__any_sync(FULL, st != ST_DONE)) {
#ifdef QRMSA_SUB_STATS
            dbg_rounds += 1; dbg_alive += __popc(__ballot_sync(FULL, st != ST_DONE)) / LPE;
#endif
            // ---- FINISH: log the decision (qrmsa.pyx:996-1063), take the next request and release what is due
            //      (qrmsa.pyx:1067-1122)
            if (__any_sync(FULL, st == ST_FINISH)) {
                const bool fin = st == ST_FINISH;
                float now = 0.f;
                if (fin) {
                    if (sub == 0) {
                        TR[cur].w = (uint32_t)action | flags;
                        if (p.gsnr_log && !(flags & QRMSA_FLAG_ACCEPTED)) {
                            double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                            gl[0] = gl[1] = gl[2] = 0.0;
                        }
                    }
                    cur += 1;
                    now = __uint_as_float(TR[cur].x);
                }
                __syncwarp();
                uint32_t n_rel = 0;
                while (__any_sync(FULL, fin && head.id >= 0 && head.id < cur && head.rel <= now)) {
                    const bool due = fin && head.id >= 0 && head.id < cur && head.rel <= now;
                    uint4 rq = make_uint4(0u, 0u, 0u, 0u);
                    if (due) rq = TR[head.id];
                    const bool ract = due && (rq.w & QRMSA_FLAG_ACCEPTED);
                    if (__any_sync(FULL, ract)) {
                        // qrmsa.pyx:1332-1350: free [s, s+n+1) (clamped at S) on every link, drop the channel record
                        const uint32_t aw = rq.w & QRMSA_ACTION_MASK;
                        const int rs = aw % S, rm = (M - 1) - (int)((aw / S) % M), rpi = aw / (S * M);
                        const int rrate = (rq.z >> 16) & 0xff;
                        int rn = 1, rcls = 0, rhops = 0;
                        uint32_t rl0 = 0u, rl1 = 0u;
                        if (ract) {
                            rn = t.need(rrate * M + rm);
                            rcls = t.cls(rrate * M + rm);
                            const uint32_t *pr = reinterpret_cast<const uint32_t *>(
                                p.prec + (size_t)((((rq.z & 0xff) * p.N + ((rq.z >> 8) & 0xff)) * K + rpi)) * 4);
                            rhops = __ldg(pr + 12) & 0x7f;
                            rl0 = __ldg(pr + sub);
                            if (LPE == 4) rl1 = __ldg(pr + 4 + sub);
                        }
                        const uint32_t target = (uint32_t)(2 * rs + rn) | ((uint32_t)rn << 12) | ((uint32_t)rm << 20) |
                                                ((uint32_t)rcls << 23);
                        const uint4 mk = range_mask4(rs, min(rs + rn + 1, S), sub);
                        const int rmax = __reduce_max_sync(FULL, ract ? rhops : 0);
                        int bad = 0;
#pragma unroll 1
                        for (int i = 0; i < rmax; ++i) {
                            const int l = link_at<LPE>(rl0, rl1, i);
                            const bool on = ract && i < rhops;
                            int cw = 0;
                            if (on) {
                                uint4 *row = reinterpret_cast<uint4 *>(BM + (unsigned)(l * RW)) + sub;
                                uint4 v = *row;
                                v.x |= mk.x; v.y |= mk.y; v.z |= mk.z; v.w |= mk.w;
                                if (sub == LPE - 1) { cw = (int)v.w; v.w = (uint32_t)(cw > 0 ? cw - 1 : 0); }
                                *row = v;
                            }
                            const int c = __shfl_sync(FULL, cw, LPE - 1, LPE);
                            if (on && sub == 0) {
                                // the record sits where the position table says (kept by every commit and every move)
                                uint32_t *lst = LISTS + (unsigned)(l * CAP);
                                const unsigned pidx = pos_index(p, l, rs >> 1);
                                const int fpos = LPE == 4 ? (int)POS[pidx] : (int)reinterpret_cast<const uint16_t *>(POS)[pidx];
                                if (fpos >= c || lst[fpos] != target) {
                                    bad = 1;
                                } else {
                                    const uint32_t last = lst[c - 1];
                                    lst[fpos] = last;
                                    lst[c - 1] = p.sentinel;
                                    const unsigned midx = pos_index(p, l, rec_pair(last));
                                    if (LPE == 4) POS[midx] = (uint8_t)fpos;
                                    else reinterpret_cast<uint16_t *>(POS)[midx] = (uint16_t)fpos;
                                }
                            }
                        }
                        if (__shfl_sync(FULL, bad, 0, LPE)) err = ENV_ERR_RELEASE_NOT_FOUND;
                        __syncwarp();
                    }
                    if (due) {
                        n_rel += ract ? 1u : 0u;
                        rel_ptr += 1;
                        head = load_head(p, TR, PERM, rel_ptr);
                    }
                }
                if (fin) {
                    QC(LC_RELEASES, n_rel);
                    st = (cur < cur_end && !err) ? ST_REQ : ST_DONE;
                }
            }

            // ---- REQ: the current request (qrmsa.pyx:1079-1099 replayed from the trace)
            if (st == ST_REQ) {
                const uint4 rq = TR[cur];
                const int src = rq.z & 0xff, dst = (rq.z >> 8) & 0xff;
                rate = (rq.z >> 16) & 0xff;
                pbase = (src * p.N + dst) * K;
                flags = QRMSA_FLAG_DECIDED;
                action = reject;
                pi = 0;
                st = ST_PATH;
            }

            // ---- PATH: open path pi (qrmsa.pyx:1482-1512: AND of the link rows), or reject after the k-th
            if (__any_sync(FULL, st == ST_PATH)) {
                if (st == ST_PATH && pi >= K) st = ST_FINISH;   // heuristics.py:966: no path/modulation -> reject action
                bool open = st == ST_PATH;
                if (open) {
                    const uint32_t *pr = reinterpret_cast<const uint32_t *>(p.prec + (size_t)(pbase + pi) * 4);
                    const int hf = __ldg(pr + 12);   // hops | 0x80 if every neighbour term on the path is >= 0
                    hops = hf & 0x7f;
                    prunable = (hf & 0x80) != 0;
                    lw0 = __ldg(pr + sub);
                    if (LPE == 4) lw1 = __ldg(pr + 4 + sub);
                    pm = rate < 6 ? (__ldg(pr + 13 + (rate >> 1)) >> ((rate & 1) << 4)) & 0xffffu : 0u;
                    if (hops == 0) { pi += 1; open = false; }
                }
                const int hmax = __reduce_max_sync(FULL, open ? hops : 0);
                uint4 x = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
#pragma unroll 1
                for (int i = 0; i < hmax; ++i) {
                    const int l = link_at<LPE>(lw0, lw1, i);
                    if (open && i < hops) {
                        const uint4 v = *(reinterpret_cast<const uint4 *>(BM + (unsigned)(l * RW)) + sub);
                        x.x &= v.x; x.y &= v.y; x.z &= v.z; x.w &= v.w;
                        if (sub == LPE - 1) sc[i] = make_uint2((uint32_t)l, v.w);
                    }
                }
                if (open) {
                    if (sub == LPE - 1) x.w = 0u;   // the count word is not spectrum
                    // one virtual free slot at index S: "n slots if the run touches the spectrum end, else n+1"
                    // (qrmsa.pyx:529-540) becomes "n+1 consecutive free slots"
                    if (sub == (vs_word >> 2)) {
                        x.x |= (vs_word & 3) == 0 ? vs_bit : 0u;
                        x.y |= (vs_word & 3) == 1 ? vs_bit : 0u;
                        x.z |= (vs_word & 3) == 2 ? vs_bit : 0u;
                        x.w |= (vs_word & 3) == 3 ? vs_bit : 0u;
                    }
                    av = x;
                    r = x;
                    a = 1;
                    m = M - 1;
                    counted = false;
                    QC(LC_LINKS_READ, hops);
                    QC(LC_PATHS_TRIED, 1);
                    st = ST_PROBE;
                }
                __syncwarp();
            }

            // ---- PROBE: modulations m, m-1, ... on the open path (heuristics.py:931-958) until one has a block of n+1
            //      free slots that is not refused on the empty-network bound, or the path is used up
            {
                bool pr = st == ST_PROBE;
                while (__any_sync(FULL, pr)) {
#ifdef QRMSA_SUB_STATS
                    dbg_search += 1;
#endif
                    int n = 1, ncls = 0;
                    if (pr) {
                        n = t.need(rate * M + m);
                        ncls = t.cls(rate * M + m);
                    }
                    const int L = n + 1;
                    if (pr && L < a) { r = av; a = 1; }
                    // r = starts of runs of >= a free slots; doubling (b <= a keeps the cover exact, b <= 31 keeps the
                    // shift inside one neighbour word)
                    while (__any_sync(FULL, pr && a < L)) {
                        const uint32_t nx = __shfl_down_sync(FULL, r.x, 1, LPE);   // last lane: its words are 0 anyway
                        if (pr && a < L) {
                            const int b = min(min(a, L - a), 31);
                            r.x &= __funnelshift_r(r.x, r.y, b);
                            r.y &= __funnelshift_r(r.y, r.z, b);
                            r.z &= __funnelshift_r(r.z, r.w, b);
                            r.w &= __funnelshift_r(r.w, nx, b);
                            a += b;
                        }
                    }
                    uint32_t w0 = r.x;
                    int j = 0;
                    if (!w0) { w0 = r.y; j = 32; }
                    if (!w0) { w0 = r.z; j = 64; }
                    if (!w0) { w0 = r.w; j = 96; }
                    const int s = quad_min<LPE>((pr && w0) ? (sub << 7) + j + __ffs(w0) - 1 : NONE);
                    if (pr) {
                        bool next_mod = false;
                        if (s == NONE) {
                            // no block of n+1 slots: blocked_due_to_resources, next modulation (heuristics.py:938-940);
                            // when the remaining ones all need >= n slots none of them can fit either
                            flags |= QRMSA_FLAG_BLOCKED_RESOURCES;
                            if (p.need_monotone) { pi += 1; st = ST_PATH; pr = false; }
                            else next_mod = true;
                        } else {
                            // hopeless even in an empty network?  Decided from the per-path table when the answer is the
                            // same for every slot, else from the candidate's own centre frequency.
                            bool pruned = (pm >> m) & 1u;
                            if (!pruned && !((pm >> (8 + m)) & 1u)) {
                                const GnBase gb = gn_base_rec(p, t, pbase + pi, s, n, ncls);
                                pruned = prunable && gb.empty() >= t.ACCHI(m);
                            }
                            if (pruned) {
                                QC(LC_GN_PRUNED, 1);
                                flags = (flags | QRMSA_FLAG_BLOCKED_OSNR) & ~QRMSA_FLAG_BLOCKED_RESOURCES;
                                next_mod = true;
                            } else {
                                cand = true;
                                pr = false;
                                cs = s; cn = n; ccls = ncls;
                            }
                        }
                        if (next_mod) {
                            m -= 1;
                            if (m < 0) { pi += 1; st = ST_PATH; pr = false; }
                        }
                    }
                }
            }

            // ---- GN: core/osnr.pyx:21-142 in table form, x = sum over the path's links and every channel on them;
            //      one flattened loop over (link, chunk of 4*LPE records)
            if (__any_sync(FULL, cand)) {
#ifdef QRMSA_SUB_STATS
                dbg_cands += __popc(__ballot_sync(FULL, cand)) / LPE;
#endif
                const int c2 = 2 * cs + cn;
                bool gact = cand;
                int gi = 0, gq = 0, gc = 0;
                uint32_t terms = 0u;
                double s1 = 0.0, s2 = 0.0, x = 0.0, w1 = 0.0, w2 = 0.0;
                const uint32_t *lp = p.lists;
                if (gact) {
                    const uint2 e = sc[0];
                    gc = (int)e.y;
                    w1 = t.W1(e.x);
                    w2 = t.W2(e.x);
                    lp = LISTS + (unsigned)(e.x * CAP) + 4 * sub;
                    terms = e.y;
                }
                while (__any_sync(FULL, gact)) {
#ifdef QRMSA_SUB_STATS
                    dbg_gn += 1;
#endif
                    if (gact) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(lp + gq);
                        gn_term(t, D, v.x, c2, s1, s2);
                        gn_term(t, D, v.y, c2, s1, s2);
                        gn_term(t, D, v.z, c2, s1, s2);
                        gn_term(t, D, v.w, c2, s1, s2);
                        gq += 4 * LPE;
                        if (gq >= gc) {   // link done
                            x = fma(w1, s1, x);
                            x = fma(w2, s2, x);   // W2 is stored negated
                            s1 = s2 = 0.0;
                            gi += 1;
                            gq = 0;
                            if (gi >= hops) {
                                gact = false;
                            } else {
                                const uint2 e = sc[gi];
                                gc = (int)e.y;
                                w1 = t.W1(e.x);
                                w2 = t.W2(e.x);
                                lp = LISTS + (unsigned)(e.x * CAP) + 4 * sub;
                                terms += e.y;
                            }
                        }
                    }
                }
                x = quad_sum<LPE>(x);

                // ---- DECIDE
                bool ok = false;
                if (cand) {
                    const GnBase gb = gn_base_rec(p, t, pbase + pi, cs, cn, ccls);
                    const double acc = gb.with(x);
                    QC(LC_GN_EVALS, 1);
                    QC(LC_GN_TERMS, terms);
                    if (!counted) { QC(LC_RECORDS_READ, terms); counted = true; }
                    ok = qot_ok(t, m, acc, flags);   // heuristics.py:957-958
                    if (ok) {
                        if (p.gsnr_log && sub == 0) {   // 10*log10(1/acc): total, ASE-only, NLI-only (osnr.pyx:133-140)
                            double *gl = p.gsnr_log + ((size_t)env * p.T + cur) * 3;
                            gl[0] = -10.0 * log10(acc);
                            gl[1] = -10.0 * log10(gb.ase);
                            gl[2] = -10.0 * log10(acc - gb.ase);
                        }
                    } else {
                        flags = (flags | QRMSA_FLAG_BLOCKED_OSNR) & ~QRMSA_FLAG_BLOCKED_RESOURCES;
                        m -= 1;
                        if (m < 0) { pi += 1; st = ST_PATH; }
                    }
                }
                // ---- COMMIT (qrmsa.pyx:1288-1325): occupy [s, s+n) plus one guard slot unless the block ends at S
                if (__any_sync(FULL, ok)) {
                    int e = cs + cn;
                    if (e < S) e += 1;
                    const uint4 mk = range_mask4(cs, e, sub);
                    const uint32_t rec = (uint32_t)c2 | ((uint32_t)cn << 12) | ((uint32_t)m << 20) | ((uint32_t)ccls << 23);
                    const int hmax = __reduce_max_sync(FULL, ok ? hops : 0);
                    int ovf = 0;
#pragma unroll 1
                    for (int i = 0; i < hmax; ++i) {
                        if (ok && i < hops) {
                            const int l = (int)sc[i].x;
                            uint4 *row = reinterpret_cast<uint4 *>(BM + (unsigned)(l * RW)) + sub;
                            uint4 v = *row;
                            v.x &= ~mk.x; v.y &= ~mk.y; v.z &= ~mk.z; v.w &= ~mk.w;
                            if (sub == LPE - 1) {
                                const int c = (int)v.w;
                                if (c >= CAP) {
                                    ovf = 1;
                                } else {
                                    LISTS[(unsigned)(l * CAP + c)] = rec;
                                    const unsigned pidx = pos_index(p, l, cs >> 1);
                                    if (LPE == 4) POS[pidx] = (uint8_t)c;
                                    else reinterpret_cast<uint16_t *>(POS)[pidx] = (uint16_t)c;
                                    v.w = (uint32_t)(c + 1);
                                }
                            }
                            *row = v;
                        }
                    }
                    if (__shfl_sync(FULL, ovf, LPE - 1, LPE)) err = ENV_ERR_LIST_OVERFLOW;
                    if (ok) {
                        flags = (flags | QRMSA_FLAG_ACCEPTED) & ~(QRMSA_FLAG_BLOCKED_RESOURCES | QRMSA_FLAG_BLOCKED_OSNR);
                        action = pi * M * S + ((M - 1) - m) * S + cs;
                        st = ST_FINISH;
                    }
                    __syncwarp();
                }
                cand = false;
            }
        }

#ifdef QRMSA_SUB_STATS
        if (lane == 0) {
            atomicAdd(p.counters + 25, (unsigned long long)dbg_rounds); atomicAdd(p.counters + 26, (unsigned long long)dbg_cands);
            atomicAdd(p.counters + 27, (unsigned long long)dbg_search); atomicAdd(p.counters + 28, (unsigned long long)dbg_gn);
            atomicAdd(p.counters + 29, (unsigned long long)dbg_alive);
        }
#endif
        // ---- write back
        if (real && was_alive) {
            if (sub == 0) {
                int4 *es = p.estate + env;   // .z (accepted total) is maintained by k_count_decisions
                es->x = cur; es->y = rel_ptr; es->w = err;
            }
            unsigned long long *cbase = p.counters + (size_t)(env / p.group_size) * QRMSA_N_COUNTERS;
#pragma unroll
            for (int j = 0; j < NCR; ++j) {
                const int lc = j * LPE + sub;
                const int slot = lc == LC_LINKS_READ ? QRMSA_CNT_LINKS_READ : lc == LC_RECORDS_READ ? QRMSA_CNT_RECORDS_READ :
                                 lc == LC_GN_TERMS ? QRMSA_CNT_GN_TERMS : lc == LC_GN_EVALS ? QRMSA_CNT_GN_EVALS :
                                 lc == LC_GN_PRUNED ? QRMSA_CNT_GN_PRUNED : lc == LC_PATHS_TRIED ? QRMSA_CNT_PATHS_TRIED :
                                 QRMSA_CNT_RELEASES;
                if (cr[j] && lc < LC_N) atomicAdd(cbase + slot, (unsigned long long)cr[j]);
            }
        }
#undef TR
#undef PERM
#undef BM
#undef LISTS
#undef POS
    }
#undef QC
}

}  // namespace qrmsa
