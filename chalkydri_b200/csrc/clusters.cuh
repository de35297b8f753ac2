// clusters.cuh -- row A4 of SURVEY.md 8a: gradient_clusters().
//
// Upstream (apriltag_quad_thresh.c do_gradient_clusters): every pixel (1..w-2, 1..h-2) whose component has
// >= 25 pixels looks at the neighbours (1,0) (0,1) (-1,1) (1,1); where the two pixels are black/white opposites
// and the neighbour's component is also >= 25 pixels, the half-way point {2x+dx, 2y+dy, gx, gy} is appended to
// the cluster keyed by the ordered pair of component representatives.  The (-1,1) probe is skipped when the
// previous pixel's (1,1) probe fired (`connected_last`), a purely local rule that is re-evaluated here.
//
// B200 mapping (atomic / latency bound): the size gate is folded into the `mark` map, so one pass over the
// pixels only reads that map plus two labels per emitted point.  Pass "count" inserts the 64-bit pair key into
// a per-frame open-addressing table and counts points per cluster (warp-aggregated with match.any); a per-frame
// CTA then selects the clusters fit_quad() would accept (24 <= n <= 3(2w+2h)) and prefix-sums their offsets;
// pass "scatter" re-walks the pixels and writes only the points of selected clusters.  The huge clusters that
// dominate the point count on noisy frames (and that upstream discards after building them) are never stored.
#pragma once
#include "common.cuh"

namespace cb {

constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ uint32_t hash_key(unsigned long long k)
{
    // two 32-bit multiplies (the 64-bit finaliser this replaces was ~10 % of the count pass' instructions)
    uint32_t h = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return h;
}

// find-or-insert; returns slot index inside the frame's sub-table or 0xffffffff when the table is full
__device__ __forceinline__ uint32_t slot_insert(ClusterSlot *tab, uint32_t nslots, unsigned long long key)
{
    uint32_t s = hash_key(key) & (nslots - 1);
    for (uint32_t probe = 0; probe < nslots; probe++) {
        unsigned long long cur = *((volatile unsigned long long *)&tab[s].key);
        if (cur == key) return s;
        if (cur == EMPTY_KEY) {
            unsigned long long old = atomicCAS(&tab[s].key, EMPTY_KEY, key);
            if (old == EMPTY_KEY || old == key) return s;
        }
        s = (s + 1) & (nslots - 1);
    }
    return 0xffffffffu;
}

__device__ __forceinline__ uint32_t slot_find(const ClusterSlot *tab, uint32_t nslots, unsigned long long key)
{
    uint32_t s = hash_key(key) & (nslots - 1);
    for (uint32_t probe = 0; probe < nslots; probe++) {
        unsigned long long cur = tab[s].key;
        if (cur == key) return s;
        if (cur == EMPTY_KEY) return 0xffffffffu;
        s = (s + 1) & (nslots - 1);
    }
    return 0xffffffffu;
}

// One CTA per 64x16-pixel tile, one thread per pixel and step (4 steps).  The points of one cluster inside a tile (a stretch of boundary, typically 30..100 points) are first
// aggregated in a 256-entry shared-memory table -- lanes with the same key elect a leader (match.any), the leader bumps the
// tile-local count -- so the global table sees one probe and one atomic per (tile, cluster) instead of one per warp row and
// probe direction.  In the scatter pass the tile-local count doubles as the rank of the point inside the tile's share of
// the cluster, and one atomicAdd on the cluster's cursor reserves that share.
constexpr int CL_TW = 64, CL_TH = 16, CL_THREADS = 256, CL_PER = CL_TW * CL_TH / CL_THREADS, CL_CAP = 256, CL_PROBES = 24;
struct ClShared {
    unsigned long long key[CL_CAP];
    uint32_t cnt[CL_CAP];
};

// find-or-insert in the tile table; -1 when no free entry is found within CL_PROBES probes (the caller goes to the global table)
__device__ __forceinline__ int tile_insert(ClShared &S, unsigned long long key, uint32_t h, int max_probes)
{
    uint32_t s = h % (CL_CAP - 2);      // entries 0xfe / 0xff stay unused: they are the "overflow" / "no entry" marks of the per-point word
    for (int probe = 0; probe < max_probes; probe++) {
        const unsigned long long cur = *((volatile unsigned long long *)&S.key[s]);
        if (cur == key) return (int)s;
        if (cur == EMPTY_KEY) {
            const unsigned long long old = atomicCAS(&S.key[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY || old == key) return (int)s;
        }
        s = s + 1 == CL_CAP - 2 ? 0 : s + 1;
    }
    return -1;
}

// the same on a bare key array (the band pass keeps one table per warp)
__device__ __forceinline__ int tile_insert_n(unsigned long long *keys, unsigned long long key, uint32_t h, int max_probes)
{
    uint32_t s = h % (CL_CAP - 2);
    for (int probe = 0; probe < max_probes; probe++) {
        const unsigned long long cur = *((volatile unsigned long long *)&keys[s]);
        if (cur == key) return (int)s;
        if (cur == EMPTY_KEY) {
            const unsigned long long old = atomicCAS(&keys[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY || old == key) return (int)s;
        }
        s = s + 1 == CL_CAP - 2 ? 0 : s + 1;
    }
    return -1;
}

// The count pass also saves what the scatter pass would otherwise recompute: the per-point words (16 bytes per pixel) and
// the tile's table (keys, counts), so the scatter pass (cluster_scatter_kernel) only reserves the tile's share of each
// selected cluster and streams the points out.  The step loop is kept rolled: unrolled four times the kernel was 66 KB of
// code and stalled on instruction fetch.
constexpr uint32_t CL_ENT_OVERFLOW = 0xfeu;     // point that did not get a tile entry: the scatter pass looks its cluster up itself
__global__ void __launch_bounds__(CL_THREADS)
cluster_count_kernel(const uint8_t *__restrict__ mark, const uint32_t *__restrict__ labels, ClusterSlot *__restrict__ table,
                     uint32_t *__restrict__ errflag, uint4 *__restrict__ ent_out, unsigned long long *__restrict__ tile_keys,
                     uint32_t *__restrict__ tile_cnt, Geom g, Caps caps)
{
    __shared__ ClShared S;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * CL_TW, y0 = 1 + blockIdx.y * CL_TH;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    const uint8_t *m = mark + (size_t)b * g.h * g.tp;
    const uint32_t *lab = labels + (size_t)b * g.npix;
    ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    for (int e = threadIdx.x; e < CL_CAP; e += CL_THREADS) { S.key[e] = EMPTY_KEY; S.cnt[e] = 0; }
    __syncthreads();
    const int dxs[4] = {1, 0, -1, 1}, dys[4] = {0, 1, 1, 1};
    const int x = x0 + (wid & 1) * 32 + lane;
#pragma unroll 1
    for (int it = 0; it < CL_PER; it++) {
        const int y = y0 + it * (CL_THREADS / CL_TW) + (wid >> 1);
        const bool inside = x >= 1 && x <= g.w - 2 && y <= g.h - 2;
        uint32_t v0 = 127;
        if (inside) v0 = m[(size_t)y * g.tp + x];
        uint32_t vn[4] = {127, 127, 127, 127};
        bool emit[4] = {false, false, false, false};
        if (v0 != 127) {
            const uint8_t *r0 = m + (size_t)y * g.tp, *r1 = r0 + g.tp;
            vn[0] = r0[x + 1]; vn[1] = r1[x]; vn[2] = r1[x - 1]; vn[3] = r1[x + 1];
            const uint32_t vprev = r0[x - 1];
            const bool connected_last = (x - 1 >= 1) && vprev != 127 && (vprev + vn[1] == 255);
            emit[0] = v0 + vn[0] == 255;
            emit[1] = v0 + vn[1] == 255;
            emit[2] = (v0 + vn[2] == 255) && !connected_last;
            emit[3] = v0 + vn[3] == 255;
        }
        uint32_t rep0 = 0;
        if (emit[0] | emit[1] | emit[2] | emit[3]) rep0 = lab[(size_t)y * g.w + x];
        // per probe: tile entry (8 bits, 0xff = none, 0xfe = overflow) | rank inside the tile entry << 8 | gradient sign << 31
        uint32_t ent[4];
#pragma unroll
        for (int d = 0; d < 4; d++) {
            ent[d] = 0xffu;
            unsigned long long key = EMPTY_KEY;
            if (emit[d]) {
                const uint32_t rep1 = lab[(size_t)(y + dys[d]) * g.w + x + dxs[d]];
                key = rep0 < rep1 ? ((unsigned long long)rep1 << 32) | rep0 : ((unsigned long long)rep0 << 32) | rep1;
            }
            if (__ballot_sync(full, emit[d]) == 0) continue;     // warp-uniform: nobody probes this direction
            const uint32_t peers = __match_any_sync(full, key);
            if (!emit[d]) continue;
            const int leader = __ffs(peers) - 1;
            const uint32_t rank = __popc(peers & ((1u << lane) - 1)), npeers = __popc(peers);
            const uint32_t sign = vn[d] > v0 ? 1u : 0u;
            int e = -1;
            uint32_t r0 = 0;
            if (lane == leader) {
                e = tile_insert(S, key, hash_key(key), (int)caps.tile_probes);
                if (e >= 0) r0 = atomicAdd(&S.cnt[e], npeers);
            }
            e = __shfl_sync(peers, e, leader);
            r0 = __shfl_sync(peers, r0, leader);
            if (e >= 0) {
                ent[d] = (uint32_t)e | ((r0 + rank) << 8) | (sign << 31);
                continue;
            }
            // tile table full (a tile crossed by > ~200 clusters): straight to the global table
            ent[d] = CL_ENT_OVERFLOW | (sign << 31);
            if (lane == leader) {
                const uint32_t s = slot_insert(tab, caps.slots_per_frame, key);
                if (s == 0xffffffffu) atomicOr(errflag, ERR_HASH_FULL);
                else atomicAdd(&tab[s].count, npeers);
            }
        }
        if (x < g.w && y <= g.h - 2) ent_out[(size_t)b * g.npix + (size_t)y * g.w + x] = make_uint4(ent[0], ent[1], ent[2], ent[3]);
    }
    __syncthreads();
    // one global probe + atomic per (tile, cluster)
    const size_t tile = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    for (int e = threadIdx.x; e < CL_CAP; e += CL_THREADS) {
        const unsigned long long key = S.key[e];
        tile_keys[tile * CL_CAP + e] = key;
        tile_cnt[tile * CL_CAP + e] = S.cnt[e];
        if (key == EMPTY_KEY) continue;
        const uint32_t s = slot_insert(tab, caps.slots_per_frame, key);
        if (s == 0xffffffffu) atomicOr(errflag, ERR_HASH_FULL);
        else atomicAdd(&tab[s].count, S.cnt[e]);
    }
}

// Scatter pass on the saved results of the count pass: per tile, one global probe + one atomicAdd on the cluster cursor per
// table entry, then every point goes to (share of its entry) + (its rank inside the tile).
__global__ void __launch_bounds__(CL_THREADS)
cluster_scatter_kernel(const uint8_t *__restrict__ mark, const uint32_t *__restrict__ labels, const ClusterSlot *__restrict__ table,
                       ClusterRec *__restrict__ clusters, uint32_t *__restrict__ scankey, const uint4 *__restrict__ ent_in,
                       const unsigned long long *__restrict__ tile_keys, const uint32_t *__restrict__ tile_cnt, Geom g, Caps caps)
{
    __shared__ uint32_t s_base[CL_CAP];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * CL_TW, y0 = 1 + blockIdx.y * CL_TH;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    const size_t tile = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    // issue the per-point loads before the table work
    uint4 en[CL_PER];
#pragma unroll
    for (int it = 0; it < CL_PER; it++) {
        const int x = x0 + (wid & 1) * 32 + lane, y = y0 + it * (CL_THREADS / CL_TW) + (wid >> 1);
        en[it] = make_uint4(0xffu, 0xffu, 0xffu, 0xffu);
        if (x >= 1 && x <= g.w - 2 && y <= g.h - 2) en[it] = ent_in[(size_t)b * g.npix + (size_t)y * g.w + x];
    }
    for (int e = threadIdx.x; e < CL_CAP; e += CL_THREADS) {
        const unsigned long long key = tile_keys[tile * CL_CAP + e];
        uint32_t base = 0xffffffffu;
        if (key != EMPTY_KEY) {
            const uint32_t s = slot_find(tab, caps.slots_per_frame, key);
            if (s != 0xffffffffu) {
                const uint32_t c = tab[s].cluster;
                if (c != 0xffffffffu) {
                    ClusterRec *cr = clusters + (size_t)b * caps.clusters_per_frame + c;
                    base = cr->offset + atomicAdd(&cr->cursor, tile_cnt[tile * CL_CAP + e]);
                }
            }
        }
        s_base[e] = base;
    }
    __syncthreads();
    const int dxs[4] = {1, 0, -1, 1}, dys[4] = {0, 1, 1, 1};
#pragma unroll
    for (int it = 0; it < CL_PER; it++) {
        const int x = x0 + (wid & 1) * 32 + lane, y = y0 + it * (CL_THREADS / CL_TW) + (wid >> 1);
        const uint32_t w4[4] = {en[it].x, en[it].y, en[it].z, en[it].w};
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const uint32_t w = w4[d], e = w & 0xffu;
            if (e == 0xffu) continue;
            const uint32_t word = ((uint32_t)(y * g.w + x) << 3) | ((uint32_t)d << 1) | (w >> 31);
            if (e != CL_ENT_OVERFLOW) {
                const uint32_t base = s_base[e];
                if (base != 0xffffffffu) scankey[(size_t)b * caps.points_per_frame + base + ((w >> 8) & 0x7fffffu)] = word;
                continue;
            }
            // the tile table was full when this point was counted: look its cluster up directly
            const uint32_t *lab = labels + (size_t)b * g.npix;
            const uint32_t rep0 = lab[(size_t)y * g.w + x], rep1 = lab[(size_t)(y + dys[d]) * g.w + x + dxs[d]];
            const unsigned long long key = rep0 < rep1 ? ((unsigned long long)rep1 << 32) | rep0 : ((unsigned long long)rep0 << 32) | rep1;
            const uint32_t s = slot_find(tab, caps.slots_per_frame, key);
            if (s == 0xffffffffu) continue;
            const uint32_t c = tab[s].cluster;
            if (c == 0xffffffffu) continue;
            ClusterRec *cr = clusters + (size_t)b * caps.clusters_per_frame + c;
            scankey[(size_t)b * caps.points_per_frame + cr->offset + atomicAdd(&cr->cursor, 1u)] = word;
        }
    }
}

// ---- round 2: band-ordered count / scatter -----------------------------------------------------------------------------------
// fit_quad() needs every cluster's points in upstream's append order, which is scan order (y, x, probe).  The tile passes above
// hand out positions with atomics, so a cluster's points came out in arbitrary order and a whole sort per cluster (five kernels)
// put them back.  Here the order is never lost:
//   * count pass: ONE WARP owns a band of `rows` full image rows and walks it in scan order.  Every boundary point is appended
//     to the band's staging list (4 bytes: table slot, row in the band, probe, gradient sign, x) at a running position -- scan
//     order by construction, all clusters interleaved -- and counted in the warp's private 256-slot table (shared memory, no
//     other warp touches it).  At the end of the band the used slots go out as (key, count) records and one atomicAdd per
//     (band, cluster) updates the frame's hash table.  A table that fills up closes the band early ("sub-band": the records go
//     out, a chained record area takes over), so there is no overflow path that could break the order;
//   * select: unchanged (clusters fit_quad() can accept get a range of the point buffer);
//   * resolve: every record looks its cluster up once (parallel);
//   * prefix: per frame the bands are walked in order and every record receives the position of its band's share inside the
//     cluster (running cursors in shared memory: the only sequential step, a few microseconds per frame);
//   * scatter: one warp per (sub-)band streams its staging list, 32 points per step, one match.any on the slot number per
//     step, and writes every point of a selected cluster straight to its scan-order position.
// Per point the stage moves 4 B (staging) out and in and 4 B (scan key) out, instead of 16 B per PIXEL out and in; the scan-order
// sort is gone.
constexpr int CLB_CAP = 256, CLB_WARPS = 8;
constexpr uint32_t CLB_NONE = 0xffffffffu;
struct ClbRec { unsigned long long key; uint32_t cnt; uint32_t slot; };     // resolve: key <- cluster index; prefix: cnt <- first position
struct ClbArea {
    uint32_t n_used, st_start, st_end, next;     // records; staging range of this (sub-)band, relative to the band's list; next area of the band
    uint32_t band, pad[3];                       // band id = frame * nbands + band index
    ClbRec rec[CLB_CAP];
};
struct BandPlan { int rows, nbands; uint32_t pool_cap; };    // rows per band, bands per frame, chained areas available to the whole batch

// staging word: slot << 24 | row in band << 21 | probe << 19 | sign << 18 | x
__device__ __forceinline__ size_t band_stage_base(const Geom &g, int b, int y0) { return ((size_t)b * g.npix + (size_t)(y0 - 1) * g.w) * 4; }

__global__ void __launch_bounds__(CLB_WARPS * 32)
cluster_band_count_kernel(const uint8_t *__restrict__ mark, const uint32_t *__restrict__ labels, ClusterSlot *__restrict__ table,
                          uint32_t *__restrict__ errflag, uint32_t *__restrict__ stage, ClbArea *__restrict__ areas,
                          uint32_t *__restrict__ pool_counter, Geom g, Caps caps, BandPlan bp)
{
    __shared__ unsigned long long s_key[CLB_WARPS][CLB_CAP];
    __shared__ uint32_t s_cnt[CLB_WARPS][CLB_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    const uint32_t job = blockIdx.x * CLB_WARPS + wid;
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    if (job >= nfirst) return;                                  // (no block-wide barrier in this kernel)
    const int b = job / bp.nbands, band = job % bp.nbands;
    const int y0 = 1 + band * bp.rows, y1 = min(y0 + bp.rows, g.h - 1);       // rows y0 .. y1 - 1 (upstream scans y = 1 .. h - 2)
    unsigned long long *K = s_key[wid];
    uint32_t *Cn = s_cnt[wid];
    for (int e = lane; e < CLB_CAP; e += 32) { K[e] = EMPTY_KEY; Cn[e] = 0; }
    __syncwarp();
    const uint8_t *m = mark + (size_t)b * g.h * g.tp;
    const uint32_t *lab = labels + (size_t)b * g.npix;
    ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    uint32_t *st = stage + band_stage_base(g, b, y0);
    uint32_t npts = 0, sub_start = 0, area_idx = job;
    bool dead = false;

    // close the current (sub-)band: records out, frame table updated, private table cleared; `more`: another sub-band follows
    auto flush = [&](bool more) {
        __syncwarp();
        ClbArea *A = areas + area_idx;
        uint32_t n_used = 0;
        for (int s0 = 0; s0 < CLB_CAP; s0 += 32) {
            const int s = s0 + lane;
            const unsigned long long key = K[s];
            const uint32_t cnt = Cn[s];
            const bool used = key != EMPTY_KEY && cnt > 0;
            const uint32_t bu = __ballot_sync(full, used);
            if (used) {
                ClbRec r;
                r.key = key; r.cnt = cnt; r.slot = (uint32_t)s;
                A->rec[n_used + __popc(bu & lt)] = r;
                const uint32_t gs = slot_insert(tab, caps.slots_per_frame, key);
                if (gs == 0xffffffffu) atomicOr(errflag, ERR_HASH_FULL);
                else atomicAdd(&tab[gs].count, cnt);
            }
            n_used += __popc(bu);
            if (key != EMPTY_KEY) { K[s] = EMPTY_KEY; Cn[s] = 0; }
        }
        uint32_t next = CLB_NONE;
        if (more) {
            if (lane == 0) next = atomicAdd(pool_counter, 1u);
            next = __shfl_sync(full, next, 0);
            if (next >= bp.pool_cap) { if (lane == 0) atomicOr(errflag, ERR_HASH_FULL); next = CLB_NONE; dead = true; }
            else next += nfirst;
        }
        if (lane == 0) { A->n_used = n_used; A->st_start = sub_start; A->st_end = npts; A->next = next; A->band = job; }
        sub_start = npts;
        if (next != CLB_NONE) area_idx = next;
        __syncwarp();
    };

    const int nseg = (g.w + 31) / 32;
    for (int y = y0; y < y1 && !dead; y++) {
        const uint8_t *r0 = m + (size_t)y * g.tp, *r1 = r0 + g.tp;
        const uint32_t *l0 = lab + (size_t)y * g.w, *l1 = l0 + g.w;
        uint32_t p0 = 127, p1 = 127;                        // values left of lane 0: the previous segment's last pixels
        uint32_t c0 = lane < g.w ? r0[lane] : 127u, c1 = lane < g.w ? r1[lane] : 127u;
        for (int seg = 0; seg < nseg && !dead; seg++) {
            const int x = seg * 32 + lane, xn = x + 32;
            uint32_t n0 = 127, n1 = 127;                    // the next segment's pixels are in flight while this one is processed
            if (xn < g.w) { n0 = r0[xn]; n1 = r1[xn]; }
            uint32_t vl = __shfl_up_sync(full, c0, 1), vr = __shfl_down_sync(full, c0, 1);
            uint32_t bl = __shfl_up_sync(full, c1, 1), br = __shfl_down_sync(full, c1, 1);
            const uint32_t nf0 = __shfl_sync(full, n0, 0), nf1 = __shfl_sync(full, n1, 0);
            if (lane == 0) { vl = p0; bl = p1; }
            if (lane == 31) { vr = nf0; br = nf1; }
            const uint32_t v0 = c0, b0 = c1;
            p0 = __shfl_sync(full, c0, 31); p1 = __shfl_sync(full, c1, 31);
            c0 = n0; c1 = n1;
            uint32_t em = 0;
            if (x >= 1 && x <= g.w - 2 && v0 != 127u) {
                const bool connected_last = (x - 1 >= 1) && vl != 127u && (vl + b0 == 255u);
                em = (v0 + vr == 255u ? 1u : 0u) | (v0 + b0 == 255u ? 2u : 0u) | ((v0 + bl == 255u && !connected_last) ? 4u : 0u) | (v0 + br == 255u ? 8u : 0u);
            }
            const uint32_t bal0 = __ballot_sync(full, em & 1u), bal1 = __ballot_sync(full, em & 2u), bal2 = __ballot_sync(full, em & 4u),
                           bal3 = __ballot_sync(full, em & 8u);
            if ((bal0 | bal1 | bal2 | bal3) == 0) continue;
            const uint32_t bals[4] = {bal0, bal1, bal2, bal3};
            unsigned long long key[4];
            uint32_t sgn = 0;
            {
                const uint32_t rep0 = em ? l0[x] : 0u;
                const uint32_t vn[4] = {vr, b0, bl, br};
#pragma unroll
                for (int d = 0; d < 4; d++) {
                    key[d] = EMPTY_KEY;
                    if ((em >> d) & 1u) {
                        const uint32_t rep1 = d == 0 ? l0[x + 1] : (d == 1 ? l1[x] : (d == 2 ? l1[x - 1] : l1[x + 1]));
                        key[d] = rep0 < rep1 ? ((unsigned long long)rep1 << 32) | rep0 : ((unsigned long long)rep0 << 32) | rep1;
                        sgn |= (vn[d] > v0 ? 1u : 0u) << d;
                    }
                }
            }
            // phase 1: every key gets a slot of the warp's table; a full table closes the sub-band first (the retry probes the
            // whole, now empty, table: a step has at most 128 distinct keys)
            int e[4];
            uint32_t peers[4];
            for (int attempt = 0; attempt < 2; attempt++) {
                bool ovf = false;
                const int probes = attempt ? CLB_CAP : (int)caps.tile_probes;
#pragma unroll
                for (int d = 0; d < 4; d++) {
                    e[d] = -1; peers[d] = 0;
                    if (bals[d] == 0) continue;                                   // warp-uniform
                    peers[d] = __match_any_sync(full, key[d]);
                    if (!((em >> d) & 1u)) continue;
                    const int leader = __ffs(peers[d]) - 1;
                    int ee = -1;
                    if (lane == leader) ee = tile_insert_n(K, key[d], hash_key(key[d]), probes);
                    ee = __shfl_sync(peers[d], ee, leader);
                    e[d] = ee;
                    ovf |= ee < 0;
                }
                if (!__any_sync(full, ovf)) break;
                flush(true);
                if (dead) break;
            }
            if (dead) break;
            // phase 2: counts, and the points themselves at their scan-order position (lane-major, then probe)
            uint32_t off = npts + __popc(bal0 & lt) + __popc(bal1 & lt) + __popc(bal2 & lt) + __popc(bal3 & lt);
#pragma unroll
            for (int d = 0; d < 4; d++) {
                if (!((em >> d) & 1u)) continue;
                if (lane == __ffs(peers[d]) - 1) atomicAdd(&Cn[e[d]], (uint32_t)__popc(peers[d]));
                st[off++] = ((uint32_t)e[d] << 24) | ((uint32_t)(y - y0) << 21) | ((uint32_t)d << 19) | (((sgn >> d) & 1u) << 18) | (uint32_t)x;
            }
            npts += __popc(bal0) + __popc(bal1) + __popc(bal2) + __popc(bal3);
        }
    }
    if (!dead) flush(false);
}

// resolve: every record learns its cluster (index into the frame's selected-cluster list, CLB_NONE = not selected).  One warp per area.
__global__ void __launch_bounds__(CLB_WARPS * 32)
cluster_band_resolve_kernel(const ClusterSlot *__restrict__ table, ClbArea *__restrict__ areas, const uint32_t *__restrict__ pool_counter,
                            Geom g, Caps caps, BandPlan bp)
{
    const int lane = threadIdx.x & 31;
    const uint32_t a = blockIdx.x * CLB_WARPS + (threadIdx.x >> 5);
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    if (a >= nfirst + min(*pool_counter, bp.pool_cap)) return;
    ClbArea *A = areas + a;
    const uint32_t n_used = A->n_used;
    const int b = A->band / bp.nbands;
    const ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    for (uint32_t i = lane; i < n_used; i += 32) {
        const uint32_t s = slot_find(tab, caps.slots_per_frame, A->rec[i].key);
        A->rec[i].key = s == 0xffffffffu ? CLB_NONE : tab[s].cluster;
    }
}

// prefix: one CTA per frame walks the frame's bands (and their chained sub-bands) in scan order; record t of an area is thread t's.
// cursors: clusters_per_frame words in shared memory when they fit, otherwise the caller passes a global array.
__global__ void __launch_bounds__(CLB_CAP)
cluster_band_prefix_kernel(ClbArea *__restrict__ areas, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ nclusters,
                           uint32_t *__restrict__ global_cursors, Caps caps, BandPlan bp)
{
    extern __shared__ uint32_t s_cursor[];
    const int b = blockIdx.x, t = threadIdx.x;
    uint32_t *cur = global_cursors ? global_cursors + (size_t)b * caps.clusters_per_frame : s_cursor;
    const uint32_t ncl = min(nclusters[b], caps.clusters_per_frame);
    for (uint32_t c = t; c < ncl; c += CLB_CAP) cur[c] = clusters[(size_t)b * caps.clusters_per_frame + c].offset;
    __syncthreads();
    // the first area of band i + 1 is loaded while band i is processed (its address does not depend on anything)
    uint32_t a = (uint32_t)b * bp.nbands;
    uint32_t n_used = areas[a].n_used, next = areas[a].next;
    ClbRec r = areas[a].rec[t];
    for (int band = 0; band < bp.nbands; band++) {
        uint32_t pn_used = 0, pnext = CLB_NONE;
        ClbRec pr;
        pr.key = 0; pr.cnt = 0; pr.slot = 0;
        const uint32_t a_next_band = (uint32_t)b * bp.nbands + band + 1;
        if (band + 1 < bp.nbands) { pn_used = areas[a_next_band].n_used; pnext = areas[a_next_band].next; pr = areas[a_next_band].rec[t]; }
        for (;;) {
            if ((uint32_t)t < n_used) {
                const uint32_t c = (uint32_t)r.key;
                uint32_t base = CLB_NONE;
                if (c < ncl) { base = cur[c]; cur[c] = base + r.cnt; }      // the records of one area name distinct clusters
                areas[a].rec[t].cnt = base;
            }
            __syncthreads();
            if (next == CLB_NONE) break;
            a = next;                                                       // chained sub-band (rare): loaded on demand
            n_used = areas[a].n_used; next = areas[a].next; r = areas[a].rec[t];
        }
        a = a_next_band; n_used = pn_used; next = pnext; r = pr;
    }
}

// scatter: one warp per (sub-)band
__global__ void __launch_bounds__(CLB_WARPS * 32)
cluster_band_scatter_kernel(const uint32_t *__restrict__ stage, const ClbArea *__restrict__ areas, const uint32_t *__restrict__ pool_counter,
                            uint32_t *__restrict__ scankey, Geom g, Caps caps, BandPlan bp)
{
    __shared__ uint32_t s_cur[CLB_WARPS][CLB_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    const uint32_t a = blockIdx.x * CLB_WARPS + wid;
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    if (a >= nfirst + min(*pool_counter, bp.pool_cap)) return;
    const ClbArea *A = areas + a;
    const uint32_t n_used = A->n_used, st_start = A->st_start, st_end = A->st_end, job = A->band;
    if (st_end <= st_start) return;
    const int b = job / bp.nbands, band = job % bp.nbands, y0 = 1 + band * bp.rows;
    uint32_t *cur = s_cur[wid];
    for (int s = lane; s < CLB_CAP; s += 32) cur[s] = CLB_NONE;
    __syncwarp();
    bool any = false;
    for (uint32_t i = lane; i < n_used; i += 32) {
        const ClbRec r = A->rec[i];
        cur[r.slot] = r.cnt;
        any |= r.cnt != CLB_NONE;
    }
    if (!__any_sync(full, any)) return;                         // no selected cluster crosses this band
    __syncwarp();
    const uint32_t *st = stage + band_stage_base(g, b, y0);
    uint32_t *out = scankey + (size_t)b * caps.points_per_frame;
    uint32_t word = st_start + lane < st_end ? st[st_start + lane] : 0xffffffffu;
    for (uint32_t i0 = st_start; i0 < st_end; i0 += 32) {
        const uint32_t w = word;
        if (i0 + 32 + lane < st_end) word = st[i0 + 32 + lane]; else word = 0xffffffffu;
        const uint32_t slot = w >> 24;                          // idle lanes: slot 255, which is never used
        const uint32_t peers = __match_any_sync(full, slot);
        const int leader = __ffs(peers) - 1;
        uint32_t base = CLB_NONE;
        if (lane == leader) { base = cur[slot]; if (base != CLB_NONE) cur[slot] = base + __popc(peers); }
        base = __shfl_sync(full, base, leader);
        if (base != CLB_NONE) {
            const uint32_t x = w & 0x3ffffu, row = (w >> 21) & 7u, ds = (w >> 18) & 7u;      // ds = probe << 1 | sign
            out[base + __popc(peers & lt)] = (((uint32_t)(y0 + row) * (uint32_t)g.w + x) << 3) | ds;
        }
        __syncwarp();
    }
}

// One thread per hash slot: pick the clusters fit_quad() can accept (24 <= n <= 3(2w+2h)), give each a slot in the
// frame's cluster list and a contiguous range of the frame's point buffer (atomic bump allocation: the layout is
// arbitrary, which is fine because every later stage is order independent), and append it to the batch-wide work list
// of its tier (one warp per small cluster, one CTA per large one).
__global__ void __launch_bounds__(256)
cluster_select_kernel(ClusterSlot *__restrict__ table, ClusterRec *__restrict__ clusters, uint32_t *__restrict__ nclusters,
                      uint32_t *__restrict__ npoints, uint32_t *__restrict__ worklists, size_t list_stride, uint32_t *__restrict__ nwork,
                      int nwork_stride, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t *__restrict__ errflag, Geom g, Caps caps,
                      int min_cluster_pixels)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    // (slots_per_frame is a power of two >= 1024, so whole warps are in range together)
    ClusterSlot *slot = table + (size_t)b * caps.slots_per_frame + min(s, caps.slots_per_frame - 1);
    const uint32_t maxsz = 3u * (2u * g.w + 2u * g.h);
    const uint32_t minsz = (uint32_t)max(24, min_cluster_pixels);
    unsigned long long key = EMPTY_KEY;
    uint32_t cnt = 0;
    if (s < caps.slots_per_frame) { key = slot->key; if (key != EMPTY_KEY) cnt = slot->count; }
    const bool sel = key != EMPTY_KEY && cnt >= minsz && cnt <= maxsz;          // otherwise slot->cluster stays 0xffffffff
    // the three allocations are warp-aggregated: the batch-wide tier counters would otherwise take one atomic per cluster
    const uint32_t m = __ballot_sync(full, sel);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    const uint32_t padded = sel ? (cnt + 7u) & ~7u : 0u;        // 8-aligned: one checkpoint slot per 8 points (quads.cuh LF_CP)
    uint32_t scan = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(full, scan, o); if (lane >= o) scan += v; }
    const uint32_t total = __shfl_sync(full, scan, 31);
    uint32_t ci0 = 0, off0 = 0;
    if (lane == leader) { ci0 = atomicAdd(&nclusters[b], (uint32_t)__popc(m)); off0 = atomicAdd(&npoints[b], total); }
    ci0 = __shfl_sync(full, ci0, leader); off0 = __shfl_sync(full, off0, leader);
    const uint32_t ci = ci0 + __popc(m & lt), off = off0 + scan - padded;
    const int tier = cnt <= t0 ? 0 : (cnt <= t1 ? 1 : (cnt <= t2 ? 2 : 3));     // work list of the quad-fitting tier
    bool ok = sel;
    if (ok && ci >= caps.clusters_per_frame) { atomicOr(errflag, ERR_CLUSTERS_FULL); ok = false; }
    if (ok && off + cnt > caps.points_per_frame) { atomicOr(errflag, ERR_POINTS_FULL); ok = false; }
    uint32_t pos = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const uint32_t mt = __ballot_sync(full, ok && tier == t);
        if (mt == 0) continue;
        const int ld = __ffs(mt) - 1;
        uint32_t p0 = 0;
        if (lane == ld) p0 = atomicAdd(&nwork[t * nwork_stride], (uint32_t)__popc(mt));
        p0 = __shfl_sync(full, p0, ld);
        if (ok && tier == t) pos = p0 + __popc(mt & lt);
    }
    if (!ok) return;
    ClusterRec r;
    r.key = key; r.offset = off; r.count = cnt; r.cursor = 0; r.pad = 0;
    clusters[(size_t)b * caps.clusters_per_frame + ci] = r;
    slot->cluster = ci;
    worklists[(size_t)tier * list_stride + pos] = (uint32_t)b * caps.clusters_per_frame + ci;
}

}  // namespace cb
