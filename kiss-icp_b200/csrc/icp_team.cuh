// Registration::AlignPointsToMap (core/Registration.cpp:138-167) on a TEAM of CTAs.
//
// Why: the loop is a dependent chain of ~18 iterations per scan. In op_icp (device_ops.cuh) every
// iteration crosses the whole 148-CTA grid three times (partials -> group leaders -> coordinator ->
// everybody, ~1.2 us per L2 hop) for ~1 us of useful work, and 15 of 16 warps of every SM wait. Here
//   * ONE pass over the map per scan (icp_fill_pass, every CTA of the launch, one warp per source point)
//     runs the 27-voxel search of GetClosestNeighbor (core/VoxelHashMap.cpp:46-70) and leaves, per point, the
//     short list of map points that can still become its nearest neighbour while it moves by <= R
//     (exactness argument: QCache in device_ops.cuh);
//   * the iterations then run on a small team (T = ceil(n / 128) CTAs, 22 for a KITTI scan): FOUR lanes per
//     source point walk its list (8 points per warp, 128 per CTA and pass), the 16 accumulators of J^T w J /
//     J^T w r are split over those four lanes, and the T partial systems meet in ONE all-gather of epoch-tagged
//     16-byte chunks: every team CTA polls all T partials, adds them in the same fixed order and solves the
//     6x6 itself, so there is no coordinator, no broadcast hop and one L2 round trip per iteration;
//   * the rest of the launch (the other ~126 SMs) is free meanwhile: k_register_frame runs the NEXT scan's
//     preprocessing and voxel downsampling on it.
// A point whose list is no longer valid (moved > R, or left its voxel) is searched again by its whole warp
// inside the iteration (nn_search_list), exactly like the first time.
#pragma once

#include "device_ops.cuh"

namespace kb {

constexpr int TEAM_MAX = 128;  // CTAs of an ICP team (one tagged chunk per value and member; a lane gathers 4 members)
constexpr int TQ_LANES = 4;    // lanes that share one source point in the list walk
constexpr int TQ_PER_WARP = 32 / TQ_LANES;
constexpr int TQ_PER_PASS = BLOCK / TQ_LANES;  // source points one CTA handles per pass

// one source point of the ICP loop: candidate list + state. Lives in global memory after the fill pass and in
// the owning team CTA's shared memory during the iterations. The list holds POINT INDICES (slot * cap + k) in the
// reference's visiting order (voxel_shifts order, then insertion order), so the first strict minimum over the
// list is the reference's answer, ties included; coordinates are read through L1 (the map is immutable
// during AlignPointsToMap).
struct QList {
    int count;      // >= 0: idx[0..count) valid; -1: not cacheable (too many candidates / max_points_per_voxel > 32)
    int full;       // points of the whole 27-voxel neighbourhood (bookkeeping of algorithmic bytes)
    int any_voxel;  // list valid whatever voxel the point is in (d* + 3R < voxel_size)
    int vx, vy, vz; // voxel at fill time
    int direct;     // nn[] holds this iteration's answer (written by a re-search inside the iteration)
    int pad;
    double pf[3];   // position at fill time
    double p[3];    // current position (TransformPoints is applied in place, Registration.cpp:55-58,160)
    double nn[4];   // nearest neighbour and distance of a re-search
    int idx[QC_MAX];
};
static_assert(sizeof(QList) % 16 == 0, "QList is copied as int4");
constexpr int TQ_CAP = static_cast<int>(QC_BYTES / sizeof(QList));  // source points per team CTA (same dynamic smem as op_icp)
static_assert(TQ_CAP >= TQ_PER_PASS, "one pass of source points must fit in shared memory");

struct TeamScratch {
    uint4 *ll;        // [2][NPART][TEAM_MAX] epoch-tagged partial systems, ping-pong by iteration parity
    double *out;      // [16] result record: pose(7) iters cand_total query_total cache_stats(3)
    QList *qrec;      // [n] per source point, written by the fill pass
};

// team size for n source points: q_per_cta points per CTA (one pass of four-lane groups by default)
__device__ __forceinline__ int icp_team_size(int n, int q_per_cta, int grid) {
    int T = (n + q_per_cta - 1) / q_per_cta;
    T = max(T, 1);
    T = min(T, min(grid, TEAM_MAX));
    return T;
}
__device__ __forceinline__ bool icp_team_fits(int n, int T) { return (n + T - 1) / T <= TQ_CAP; }

// GetClosestNeighbor for one point by one warp (like nn_search_warp) + its candidate list into *out
// (generic pointer: global memory in the fill pass, shared memory inside an iteration).
__device__ __noinline__ NNResult nn_search_list(const MapView &m, const V3 &q, int lane, WarpNN &w, QList *out,
                                                double cache_radius) {
    const int3 v = point_to_voxel(q.x, q.y, q.z, m.vdiv);
    int cnt = 0, slot = -1;
    if (lane < 27) {
        slot = map_find(m, v.x + c_shifts[lane][0], v.y + c_shifts[lane][1], v.z + c_shifts[lane][2], &cnt);
        if (slot < 0) cnt = 0;
    }
    const int cap = m.cap;
    NNResult r;
    int count = -1;
    if (cap > NN_FLAT_CAP) {
        // general path (max_points_per_voxel > 32): no list, the point is searched again every iteration
        double best = DBL_MAX, best_d2 = DBL_MAX;
        int bseq = INT_MAX;
        V3 bp{0, 0, 0};
        unsigned occ = __ballot_sync(FULL, cnt > 0);
        int total = 0;
        while (occ) {
            const int vi = __ffs(occ) - 1;
            occ &= occ - 1;
            const int c = __shfl_sync(FULL, cnt, vi);
            const int s = __shfl_sync(FULL, slot, vi);
            total += c;
            const double *blk = m.points + static_cast<size_t>(s) * cap * 3;
            for (int k = lane; k < c; k += 32)
                nn_consider(V3{blk[3 * k], blk[3 * k + 1], blk[3 * k + 2]}, q, vi * 1024 + k, best, best_d2, bseq, bp);
        }
        nn_reduce(best, bseq, bp);
        r = NNResult{best, bp, total};
    } else {
        r = nn_flat_search(m, q, lane, w, cnt, slot);
        const int total = r.candidates;
        // second pass (L1-hot): the candidates within d* + 2R of the point, in reference order
        if (r.d < DBL_MAX) {
            const double thr = r.d + 2.0 * cache_radius;
            const double thr2 = thr * thr * (1.0 + 1e-12);
            count = 0;
            constexpr int U2 = 4;
            for (int base = 0; base < total; base += 32 * U2) {
#pragma unroll
                for (int u = 0; u < U2; ++u) {
                    const int j = base + u * 32 + lane;
                    bool keep = false;
                    int gi = 0;
                    if (j < total) {
                        const int vi = w.owner[j];
                        gi = w.slot[vi] * cap + (j - w.start[vi]);
                        const V3 c = ld_point24(m.points + static_cast<size_t>(gi) * 3);
                        keep = sqnorm(c - q) <= thr2;
                    }
                    const unsigned mask = __ballot_sync(FULL, keep);
                    const int pos = count + __popc(mask & ((1u << lane) - 1u));
                    if (keep && pos < QC_MAX) out->idx[pos] = gi;
                    count += __popc(mask);
                }
            }
            if (count > QC_MAX) count = -1;
        } else {
            count = 0;  // empty neighbourhood: stays empty while the point stays in its voxel
        }
    }
    if (lane == 0) {
        out->count = count;
        out->full = r.candidates;
        out->any_voxel = (r.d < DBL_MAX && r.d + 3.0 * cache_radius < m.voxel_size) ? 1 : 0;
        out->vx = v.x;
        out->vy = v.y;
        out->vz = v.z;
        out->pf[0] = q.x;
        out->pf[1] = q.y;
        out->pf[2] = q.z;
    }
    __syncwarp();
    return r;
}

// every CTA of `g`, one warp per source point: source = initial_guess * source (Registration.cpp:146-147),
// 27-voxel search, candidate list -> qrec[point]
__device__ __noinline__ void icp_fill_pass(const Grid &g, Shared &sh, const MapView &m, const double *src, int n,
                                           const SE3 &guess, QList *qrec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double radius = 0.2 * m.voxel_size;
    for (int qi = g.rank + g.size * warp; qi < n; qi += g.size * NWARPS) {
        const V3 p = se3_act(guess, V3{src[3 * qi], src[3 * qi + 1], src[3 * qi + 2]});
        QList *rec = &qrec[qi];
        nn_search_list(m, p, lane, sh.wnn[warp], rec, radius);
        if (lane == 0) {
            rec->p[0] = p.x;
            rec->p[1] = p.y;
            rec->p[2] = p.z;
            rec->direct = 0;
        }
    }
}

// the 16 distinct entries of one correspondence's J^T w J / J^T w r (icp_term's formulas); lane g4 of the
// point's four lanes keeps entries 4 g4 .. 4 g4 + 3
__device__ __forceinline__ void icp_term4(int g4, const V3 &s, const V3 &t, double kscale, double acc[4]) {
    const V3 r = s - t;
    const double r2 = sqnorm(r);
    const double w = (kscale * kscale) * fast_rcp((kscale + r2) * (kscale + r2));
    const double xw = s.x * w, yw = s.y * w, zw = s.z * w;
    double a0, a1, a2, a3;
    if (g4 == 0) {
        a0 = w;
        a1 = xw;
        a2 = yw;
        a3 = zw;
    } else if (g4 == 1) {
        a0 = zw * s.z + yw * s.y;  // (3,3)
        a1 = -(xw * s.y);          // (4,3)
        a2 = zw * s.z + xw * s.x;  // (4,4)
        a3 = -(xw * s.z);          // (5,3)
    } else if (g4 == 2) {
        a0 = -(yw * s.z);          // (5,4)
        a1 = yw * s.y + xw * s.x;  // (5,5)
        a2 = w * r.x;
        a3 = w * r.y;
    } else {
        a0 = w * r.z;
        a1 = -(zw * r.y) + yw * r.z;
        a2 = zw * r.x - xw * r.z;
        a3 = -(yw * r.x) + xw * r.y;
    }
    acc[0] += a0;
    acc[1] += a1;
    acc[2] += a2;
    acc[3] += a3;
}

// merge two (minimum, second minimum, position, point) records of a list walk; symmetric, so both sides of a
// butterfly end with the same record. Equal squares keep the smaller list position (= reference order).
__device__ __forceinline__ void tq_merge(double &b2, double &s2, int &bk, V3 &bp, int o) {
    const double ob2 = __shfl_xor_sync(FULL, b2, o), os2 = __shfl_xor_sync(FULL, s2, o);
    const int ok = __shfl_xor_sync(FULL, bk, o);
    const V3 obp{__shfl_xor_sync(FULL, bp.x, o), __shfl_xor_sync(FULL, bp.y, o), __shfl_xor_sync(FULL, bp.z, o)};
    const double hi = fmax(b2, ob2), lo = fmin(b2, ob2);
    double ns2 = fmin(s2, os2);
    if (hi > lo) ns2 = fmin(ns2, hi);
    if ((ob2 < b2) || (ob2 == b2 && ok < bk)) {
        bk = ok;
        bp = obp;
    }
    b2 = lo;
    s2 = ns2;
}

// DataAssociation + BuildLinearSystem (Registration.cpp:60-121) for this CTA's source points, iteration j:
// the partial system goes out as tagged chunks ll[parity][value][member]
__device__ __noinline__ void team_queries(const TeamScratch &ts, Shared &sh, const MapView &m, QList *tq, int nq, int j,
                                          double max_dist, double kscale, int member, unsigned tag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g4 = lane & (TQ_LANES - 1), grp = lane / TQ_LANES;
    const double radius = 0.2 * m.voxel_size, r2max = radius * radius;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int corr = 0, n_hit = 0, n_fill = 0, n_over = 0;
    double cand = 0.0;
    for (int base = warp * TQ_PER_WARP; base < nq; base += NWARPS * TQ_PER_WARP) {  // warp-uniform trip count
        const int li = base + grp;
        const bool have = li < nq;
        QList &t = tq[have ? li : base];
        V3 p{0.0, 0.0, 0.0};
        bool valid = false;
        if (have) {
            p = V3{t.p[0], t.p[1], t.p[2]};
            if (j > 0) p = se3_act(sh.pending, p);  // TransformPoints(estimation, source)  Registration.cpp:160
            const V3 moved = p - V3{t.pf[0], t.pf[1], t.pf[2]};
            const int3 v = point_to_voxel(p.x, p.y, p.z, m.vdiv);
            valid = t.count >= 0 && sqnorm(moved) <= r2max && (t.any_voxel || (t.vx == v.x && t.vy == v.y && t.vz == v.z));
        }
        __syncwarp();  // all four lanes have read t.p
        if (have && g4 == 0 && j > 0) {
            t.p[0] = p.x;
            t.p[1] = p.y;
            t.p[2] = p.z;
        }
        // points whose list is stale are searched again by the whole warp, one after the other
        unsigned need = __ballot_sync(FULL, have && !valid && g4 == 0);
        const unsigned hits = __ballot_sync(FULL, have && valid && g4 == 0);
        if (lane == 0) {
            n_fill += __popc(need);
            n_hit += __popc(hits);
        }
        while (need) {
            const int b = __ffs(need) - 1;
            need &= need - 1;
            const V3 pq{__shfl_sync(FULL, p.x, b), __shfl_sync(FULL, p.y, b), __shfl_sync(FULL, p.z, b)};
            QList *tr = &tq[base + b / TQ_LANES];
            const NNResult r = nn_search_list(m, pq, lane, sh.wnn[warp], tr, radius);
            if (lane == 0) {
                tr->nn[0] = r.p.x;
                tr->nn[1] = r.p.y;
                tr->nn[2] = r.p.z;
                tr->nn[3] = r.d;
                tr->direct = 1;
                n_over += (tr->count < 0 && r.d < DBL_MAX) ? 1 : 0;
            }
            __syncwarp();
        }
        // nearest neighbour from the list: four lanes stride it, first strict minimum of the squared distance
        const bool direct = have && t.direct != 0;
        const int cnt = (have && !direct) ? t.count : 0;
        double b2 = DBL_MAX, s2 = DBL_MAX;
        int bk = INT_MAX;
        V3 bp{0.0, 0.0, 0.0};
        for (int k0 = g4; k0 < cnt; k0 += 4 * TQ_LANES) {
            V3 c[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // four independent loads in flight per lane
                const int k = k0 + u * TQ_LANES;
                ok[u] = k < cnt;
                if (ok[u]) c[u] = ld_point24(m.points + static_cast<size_t>(t.idx[k]) * 3);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u]) {
                    const double d2 = sqnorm(c[u] - p);
                    if (d2 < b2) {
                        s2 = b2;
                        b2 = d2;
                        bk = k0 + u * TQ_LANES;
                        bp = c[u];
                    } else if (d2 > b2 && d2 < s2) {
                        s2 = d2;
                    }
                }
        }
        tq_merge(b2, s2, bk, bp, 1);
        tq_merge(b2, s2, bk, bp, 2);
        double d = (cnt > 0) ? sqrt(b2) : DBL_MAX;
        V3 np = bp;
        // two squares within a few ulps could round to the same root: then compare rounded roots in reference
        // order like GetClosestNeighbor does (in practice never)
        const bool near = cnt > 0 && s2 <= b2 * (1.0 + 8.8817841970012523e-16);
        if (__any_sync(FULL, near)) {
            double best = DBL_MAX;
            int ek = INT_MAX;
            V3 ep{0.0, 0.0, 0.0};
            if (near)
                for (int k = g4; k < cnt; k += TQ_LANES) {
                    const V3 c = ld_point24(m.points + static_cast<size_t>(t.idx[k]) * 3);
                    const double dd = norm(c - p);
                    if (dd < best) {
                        best = dd;
                        ek = k;
                        ep = c;
                    }
                }
#pragma unroll
            for (int o = 1; o < TQ_LANES; o <<= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int ok2 = __shfl_xor_sync(FULL, ek, o);
                const V3 op{__shfl_xor_sync(FULL, ep.x, o), __shfl_xor_sync(FULL, ep.y, o), __shfl_xor_sync(FULL, ep.z, o)};
                if ((ob < best) || (ob == best && ok2 < ek)) {
                    best = ob;
                    ek = ok2;
                    ep = op;
                }
            }
            if (near) {
                d = best;
                np = ep;
            }
        }
        if (direct) {
            d = t.nn[3];
            np = V3{t.nn[0], t.nn[1], t.nn[2]};
        }
        __syncwarp();  // all four lanes have read the direct answer
        if (direct && g4 == 0) t.direct = 0;
        if (have && g4 == 0) cand += t.full;
        if (have && d < max_dist) {  // DataAssociation's gate, Registration.cpp:72
            icp_term4(g4, p, np, kscale, acc);
            corr += (g4 == 0) ? 1 : 0;
        }
    }
    // eight points of a warp -> one partial per accumulator (lane g4 holds entries 4 g4 ..)
#pragma unroll
    for (int o = TQ_LANES; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
        corr += __shfl_xor_sync(FULL, corr, o);
        cand += __shfl_xor_sync(FULL, cand, o);
    }
    if (lane < TQ_LANES) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sh.warp_d[warp][4 * lane + i] = acc[i];
    }
    if (lane == 0) {
        sh.warp_d[warp][NACC] = static_cast<double>(corr);
        sh.warp_d[warp][NACC + 1] = cand;
        sh.warp_d[warp][NACC + 2] = static_cast<double>(n_hit);
        sh.warp_d[warp][NACC + 3] = static_cast<double>(n_fill);
        sh.warp_d[warp][NACC + 4] = static_cast<double>(n_over);
    }
    __syncthreads();
    // 16-warp tree per value: thread t -> value t / 16, warp t % 16
    if (threadIdx.x < ((NPART * NWARPS + 31) / 32) * 32) {
        const int val = min(static_cast<int>(threadIdx.x) / NWARPS, NPART - 1);
        double v = sh.warp_d[threadIdx.x & (NWARPS - 1)][val];
#pragma unroll
        for (int o = NWARPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if ((threadIdx.x & (NWARPS - 1)) == 0 && threadIdx.x < NPART * NWARPS)
            ll_store(&ts.ll[(static_cast<size_t>(tag & 1u) * NPART + val) * TEAM_MAX + member], v, tag);
    }
}

// all-gather of the T tagged partial systems: warp w sums values w and w + 16 over the members (a lane polls
// members lane, lane + 32, lane + 64, lane + 96), fixed order -> the same bits in every team CTA
__device__ __forceinline__ void team_gather(const TeamScratch &ts, Shared &sh, int T, unsigned tag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 *base = ts.ll + static_cast<size_t>(tag & 1u) * NPART * TEAM_MAX;
    for (int v = warp; v < NPART; v += NWARPS) {
        double x[TEAM_MAX / 32];
        bool ok[TEAM_MAX / 32];
#pragma unroll
        for (int u = 0; u < TEAM_MAX / 32; ++u) {
            x[u] = 0.0;
            ok[u] = lane + 32 * u >= T;
        }
        bool all = false;
        while (!all) {
            all = true;
#pragma unroll
            for (int u = 0; u < TEAM_MAX / 32; ++u) {
                if (!ok[u]) ok[u] = ll_load(&base[static_cast<size_t>(v) * TEAM_MAX + lane + 32 * u], tag, &x[u]);
                all = all && ok[u];
            }
            all = __all_sync(FULL, all);
        }
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < TEAM_MAX / 32; ++u) s += (lane + 32 * u < T) ? x[u] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) sh.red[v] = s;
    }
    __syncthreads();
}

// the iterations, on CTAs [0, T) of the launch. Precondition: icp_fill_pass + a grid barrier, map not empty,
// max_iter > 0, icp_team_fits(n, T). Output in sh.result / sh.iters / sh.cand_total / ... of every team CTA.
__device__ __noinline__ void op_icp_team(const TeamScratch &ts, const Scratch &sc, Shared &sh, const MapView &m, int n,
                                         const SE3 &guess, double max_dist, double kscale, int max_iter, double conv,
                                         QList *tq, int T, unsigned tag_base) {
    const int member = static_cast<int>(blockIdx.x);
    const int nq = member < n ? (n - member + T - 1) / T : 0;
    {
        constexpr int W = static_cast<int>(sizeof(QList) / 16);
        for (int i = threadIdx.x; i < nq * W; i += BLOCK) {
            const int li = i / W, w = i - li * W;
            reinterpret_cast<int4 *>(&tq[li])[w] = __ldcg(reinterpret_cast<const int4 *>(&ts.qrec[member + T * li]) + w);
        }
    }
    if (threadIdx.x == 0) {
        sh.t_icp = se3_identity();
        sh.cand_total = 0.0;
        sh.cache_stats[0] = sh.cache_stats[1] = sh.cache_stats[2] = 0.0;
    }
    __syncthreads();
    int j = 0;
    for (;; ++j) {
        if (sc.profile && member == 0 && threadIdx.x == 0 && j < 20) sc.dbg[41 + j] = globaltimer_ns();
        const unsigned tag = tag_base + static_cast<unsigned>(j) + 1u;
        team_queries(ts, sh, m, tq, nq, j, max_dist, kscale, member, tag);
        team_gather(ts, sh, T, tag);
        if (threadIdx.x == 0) {
            double sys[NACC];
#pragma unroll
            for (int i = 0; i < NACC; ++i) sys[i] = sh.red[i];
            double dx[6];
            if (!icp_solve_schur(sys, dx)) {      // Registration.cpp:156 — structured 3x3 Schur solve; degenerate
                double JTJ[36], JTr[6], rhs[6];   // systems take the pivoted LDL^T with Eigen's zero-pivot rule
                icp_expand(sys, JTJ, JTr);
#pragma unroll
                for (int i = 0; i < 6; ++i) rhs[i] = -JTr[i];
                ldlt6_solve_fast(JTJ, rhs, dx);
            }
            const SE3 est = se3_exp_fast(dx);  // :157
            double n2 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) n2 += dx[i] * dx[i];
            sh.pending = est;
            sh.flag = ((sqrt(n2) < conv) || (j + 1 >= max_iter)) ? 1 : 0;  // :163 / :151
            sh.t_icp = se3_mul_fast(est, sh.t_icp);                        // :161
            sh.cand_total += sh.red[NACC + 1];
            for (int i = 0; i < 3; ++i) sh.cache_stats[i] += sh.red[NACC + 2 + i];
        }
        __syncthreads();
        if (sh.flag) break;
    }
    if (threadIdx.x == 0) {
        sh.result = se3_mul(sh.t_icp, guess);  // :166
        sh.iters = j + 1;
        sh.query_total = static_cast<double>(n) * (j + 1);
    }
    __syncthreads();
}

// result record <-> shared memory (member 0 publishes before a grid barrier, everybody reads after it)
__device__ __forceinline__ void team_publish(const TeamScratch &ts, const Shared &sh) {
    double *o = ts.out;
    o[0] = sh.result.q.x;
    o[1] = sh.result.q.y;
    o[2] = sh.result.q.z;
    o[3] = sh.result.q.w;
    o[4] = sh.result.t.x;
    o[5] = sh.result.t.y;
    o[6] = sh.result.t.z;
    o[7] = static_cast<double>(sh.iters);
    o[8] = sh.cand_total;
    o[9] = sh.query_total;
    o[10] = sh.cache_stats[0];
    o[11] = sh.cache_stats[1];
    o[12] = sh.cache_stats[2];
}
__device__ __forceinline__ void team_collect(const TeamScratch &ts, Shared &sh) {
    const double *o = ts.out;
    sh.result = SE3{{__ldcg(o + 0), __ldcg(o + 1), __ldcg(o + 2), __ldcg(o + 3)}, {__ldcg(o + 4), __ldcg(o + 5), __ldcg(o + 6)}};
    sh.iters = static_cast<int>(__ldcg(o + 7));
    sh.cand_total = __ldcg(o + 8);
    sh.query_total = __ldcg(o + 9);
    sh.cache_stats[0] = __ldcg(o + 10);
    sh.cache_stats[1] = __ldcg(o + 11);
    sh.cache_stats[2] = __ldcg(o + 12);
}

}  // namespace kb
