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
                if (s == 0xffffffffu) flag_overflow(errflag, blockIdx.z, ERR_HASH_FULL);
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
        if (s == 0xffffffffu) flag_overflow(errflag, blockIdx.z, ERR_HASH_FULL);
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

// find-or-insert in a warp's private 256-slot table (power-of-two mask; the table is closed long before it is full, so the
// probe sequence always ends); *fresh = the key was not there before
__device__ __forceinline__ int band_insert(unsigned long long *keys, unsigned long long key, uint32_t h, bool *fresh)
{
    uint32_t s = h & (CLB_CAP - 1);
    for (;;) {
        const unsigned long long cur = *((volatile unsigned long long *)&keys[s]);
        if (cur == key) { *fresh = false; return (int)s; }
        if (cur == EMPTY_KEY) {
            const unsigned long long old = atomicCAS(&keys[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY) { *fresh = true; return (int)s; }
            if (old == key) { *fresh = false; return (int)s; }
        }
        s = (s + 1) & (CLB_CAP - 1);
    }
}

// (One warp per CTA: the band, its row range and every loop bound then depend on blockIdx alone, which the compiler knows to be
// warp-uniform.)  The walk has a producer and a consumer half.  Producer, per 32-pixel step: pixel values (loaded three steps
// ahead) -> which probes emit -> component labels (loaded one step ahead) -> pair keys, appended in scan order to a small
// queue in shared memory.  Consumer, whenever 32 points are queued: ONE match.any on the keys, one table probe per distinct
// key, one coalesced 128-byte store of staging words -- every lane busy, where probing the four directions one after the
// other kept about a quarter of the lanes busy and cost four match rounds per step.
__global__ void __launch_bounds__(32, 32)
cluster_band_count_kernel(const uint8_t *__restrict__ mark, const uint32_t *__restrict__ labels, ClusterSlot *__restrict__ table,
                          uint32_t *__restrict__ errflag, uint32_t *__restrict__ stage, ClbArea *__restrict__ areas,
                          uint32_t *__restrict__ pool_counter, Geom g, Caps caps, BandPlan bp)
{
    __shared__ unsigned long long K[CLB_CAP];       // the band's table: pair key -> slot
    __shared__ uint32_t Cn[CLB_CAP];                //                   points per slot
    __shared__ unsigned long long Qk[256];          // queue of points waiting for a full round: key
    __shared__ uint32_t Qm[256];                    //                                           staging word without the slot
    const int lane = threadIdx.x;
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    const uint32_t job = blockIdx.x;
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    const int b = job / bp.nbands, band = job % bp.nbands;
    const int y0 = 1 + band * bp.rows, y1 = min(y0 + bp.rows, g.h - 1);       // rows y0 .. y1 - 1 (upstream scans y = 1 .. h - 2)
    for (int e = lane; e < CLB_CAP; e += 32) { K[e] = EMPTY_KEY; Cn[e] = 0; }
    __syncwarp();
    const uint8_t *m = mark + (size_t)b * g.h * g.tp;
    const uint32_t *lab = labels + (size_t)b * g.npix;
    ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    uint32_t *st = stage + band_stage_base(g, b, y0);
    uint32_t npts = 0, sub_start = 0, area_idx = job, n_slots = 0, q_head = 0, q_tail = 0;
    // a sub-band is closed before its table is half full (CB_TILE_PROBES lowers the limit so that tests reach the chained areas)
    const uint32_t slot_limit = caps.tile_probes >= CL_PROBES ? 120u : caps.tile_probes * 48u;

    // close the current (sub-)band: records out, frame table updated, private table cleared; `more`: another sub-band follows
    auto flush = [&](bool more) {
        __syncwarp();
        ClbArea *A = areas + area_idx;
        uint32_t n_used = 0;
        for (int s0 = 0; s0 < CLB_CAP; s0 += 32) {
            const int s = s0 + lane;
            const unsigned long long key = K[s];
            const uint32_t cnt = Cn[s];
            const bool used = key != EMPTY_KEY;
            const uint32_t bu = __ballot_sync(full, used);
            if (used) {
                ClbRec r;
                r.key = key; r.cnt = cnt; r.slot = (uint32_t)s;
                A->rec[n_used + __popc(bu & lt)] = r;
                const uint32_t gs = slot_insert(tab, caps.slots_per_frame, key);
                if (gs == 0xffffffffu) flag_overflow(errflag, b, ERR_HASH_FULL);
                else atomicAdd(&tab[gs].count, cnt);
                K[s] = EMPTY_KEY; Cn[s] = 0;
            }
            n_used += __popc(bu);
        }
        uint32_t next = CLB_NONE;
        if (more) {
            if (lane == 0) next = atomicAdd(pool_counter, 1u);
            next = __shfl_sync(full, next, 0);
            // no chained area left: the call fails with ERR_HASH_FULL; this area is simply reused so that the walk stays in bounds
            if (next >= bp.pool_cap) { if (lane == 0) flag_overflow(errflag, b, ERR_HASH_FULL); next = CLB_NONE; }
            else next += nfirst;
        }
        if (lane == 0) { A->n_used = n_used; A->st_start = sub_start; A->st_end = npts; A->next = next; A->band = job; }
        sub_start = npts; n_slots = 0;
        if (next != CLB_NONE) area_idx = next;
        __syncwarp();
    };

    // consumer: the n (<= 32) oldest queued points get their slot, are counted and leave for the staging list
    auto consume = [&](uint32_t n) {
        if (n_slots > 0 && n_slots + n > slot_limit) flush(true);
        const bool active = (uint32_t)lane < n;
        const uint32_t qi = (q_head + lane) & 255u;
        const unsigned long long key = active ? Qk[qi] : EMPTY_KEY;
        const uint32_t meta = Qm[qi];
        const uint32_t peers = __match_any_sync(full, key);
        const int leader = __ffs(peers) - 1;
        bool fresh = false;
        int e = 0;
        if (active && lane == leader) {
            // (two multiplies and a fold: the table has 256 slots and at most half of them are ever used)
            const uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
            e = band_insert(K, key, h >> 24, &fresh);
            atomicAdd(&Cn[e], (uint32_t)__popc(peers));
        }
        n_slots += __popc(__ballot_sync(full, fresh));
        e = __shfl_sync(full, e, leader);
        if (active) st[npts + lane] = ((uint32_t)e << 24) | meta;
        npts += n; q_head += n;
        __syncwarp();
    };

    // The band is walked as one sequence of 32-pixel steps (row-major).
    const int nseg = (g.w + 31) / 32, nrows = y1 - y0, nsteps = nrows * nseg;
    struct Pos { int seg, y; };                                  // a step's segment and row, advanced without divisions
    auto advance = [&](Pos &q) { if (++q.seg == nseg) { q.seg = 0; q.y++; } };
    auto load_px = [&](const Pos &q, uint32_t &a0, uint32_t &a1) {
        a0 = 127u; a1 = 127u;
        const int x = q.seg * 32 + lane;
        if (q.y < y1 && x < g.w) { const uint8_t *r0 = m + (size_t)q.y * g.tp; a0 = r0[x]; a1 = r0[g.tp + x]; }
    };
    // emit mask and labels of a step whose pixel values are (c0, c1), left neighbours (p0, p1), next segment (n0, n1)
    struct Pending { uint32_t em, sgn, rep0, rep[4], meta; };
    auto prepare = [&](const Pos &q, uint32_t c0, uint32_t c1, uint32_t p0, uint32_t p1, uint32_t n0, uint32_t n1, Pending &P) {
        const int seg = q.seg, x = seg * 32 + lane;
        P.meta = ((uint32_t)(q.y - y0) << 21) | (uint32_t)x;
        uint32_t vl = __shfl_up_sync(full, c0, 1), vr = __shfl_down_sync(full, c0, 1);
        uint32_t bl = __shfl_up_sync(full, c1, 1), br = __shfl_down_sync(full, c1, 1);
        const uint32_t nf0 = __shfl_sync(full, n0, 0), nf1 = __shfl_sync(full, n1, 0);
        if (lane == 0) { vl = seg == 0 ? 127u : p0; bl = seg == 0 ? 127u : p1; }
        if (lane == 31) { vr = seg + 1 < nseg ? nf0 : 127u; br = seg + 1 < nseg ? nf1 : 127u; }
        P.em = 0; P.sgn = 0; P.rep0 = 0;
        if (x >= 1 && x <= g.w - 2 && c0 != 127u) {
            const bool connected_last = (x - 1 >= 1) && vl != 127u && (vl + c1 == 255u);
            P.em = (c0 + vr == 255u ? 1u : 0u) | (c0 + c1 == 255u ? 2u : 0u) | ((c0 + bl == 255u && !connected_last) ? 4u : 0u) | (c0 + br == 255u ? 8u : 0u);
            P.sgn = (vr > c0 ? 1u : 0u) | (c1 > c0 ? 2u : 0u) | (bl > c0 ? 4u : 0u) | (br > c0 ? 8u : 0u);
        }
        if (P.em) {
            const uint32_t *l0 = lab + (size_t)q.y * g.w + x, *l1 = l0 + g.w;
            P.rep0 = l0[0];
            if (P.em & 1u) P.rep[0] = l0[1];
            if (P.em & 2u) P.rep[1] = l1[0];
            if (P.em & 4u) P.rep[2] = l1[-1];
            if (P.em & 8u) P.rep[3] = l1[1];
        }
    };

    uint32_t c0, c1, n0, n1, nn0, nn1, p0 = 127u, p1 = 127u;
    Pos q_prep{0, y0}, q_load{0, y0};
    load_px(q_load, c0, c1); advance(q_load);
    load_px(q_load, n0, n1); advance(q_load);
    load_px(q_load, nn0, nn1); advance(q_load);
    Pending cur, nxt;
    prepare(q_prep, c0, c1, p0, p1, n0, n1, cur); advance(q_prep);
    for (int step = 0; step < nsteps; step++) {
        // producer, stage A for step + 1 (its labels fly while this step is queued), pixel values for step + 3
        p0 = __shfl_sync(full, c0, 31); p1 = __shfl_sync(full, c1, 31);
        c0 = n0; c1 = n1; n0 = nn0; n1 = nn1;
        load_px(q_load, nn0, nn1); advance(q_load);
        nxt.em = 0;
        if (step + 1 < nsteps) { prepare(q_prep, c0, c1, p0, p1, n0, n1, nxt); advance(q_prep); }
        // producer, stage B for this step: keys into the queue in scan order (lane-major, then probe)
        const uint32_t em = cur.em;
        const uint32_t bal0 = __ballot_sync(full, em & 1u), bal1 = __ballot_sync(full, em & 2u), bal2 = __ballot_sync(full, em & 4u),
                       bal3 = __ballot_sync(full, em & 8u);
        if ((bal0 | bal1 | bal2 | bal3) != 0) {
            uint32_t off = q_tail + __popc(bal0 & lt) + __popc(bal1 & lt) + __popc(bal2 & lt) + __popc(bal3 & lt);
#pragma unroll
            for (int d = 0; d < 4; d++) {
                if (!((em >> d) & 1u)) continue;
                const uint32_t r0 = cur.rep0, r1 = cur.rep[d];
                Qk[off & 255u] = r0 < r1 ? ((unsigned long long)r1 << 32) | r0 : ((unsigned long long)r0 << 32) | r1;
                Qm[off & 255u] = cur.meta | ((uint32_t)d << 19) | (((cur.sgn >> d) & 1u) << 18);
                off++;
            }
            q_tail += __popc(bal0) + __popc(bal1) + __popc(bal2) + __popc(bal3);
            __syncwarp();
            while (q_tail - q_head >= 32u) consume(32u);
        }
        cur = nxt;
    }
    if (q_tail != q_head) consume(q_tail - q_head);
    flush(false);
}

// resolve: every record learns its cluster (index into the frame's selected-cluster list, CLB_NONE = not selected).  One warp per area.
__global__ void __launch_bounds__(CLB_WARPS * 32)
cluster_band_resolve_kernel(const ClusterSlot *__restrict__ table, ClbArea *__restrict__ areas, const uint32_t *__restrict__ pool_counter,
                            Geom g, Caps caps, BandPlan bp, uint32_t *__restrict__ dense)
{
    const int lane = threadIdx.x & 31;
    const uint32_t a = blockIdx.x * CLB_WARPS + (threadIdx.x >> 5);
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    if (a >= nfirst + min(*pool_counter, bp.pool_cap)) return;
    ClbArea *A = areas + a;
    const uint32_t n_used = A->n_used;
    const int b = A->band / bp.nbands;
    const ClusterSlot *tab = table + (size_t)b * caps.slots_per_frame;
    // dense form of the prefix (small batches, no chained sub-band anywhere): the record's count goes to cell (band, cluster)
    const bool dense_on = dense != nullptr && *pool_counter == 0;
    for (uint32_t i = lane; i < n_used; i += 32) {
        const uint32_t s = slot_find(tab, caps.slots_per_frame, A->rec[i].key);
        const uint32_t ci = s == 0xffffffffu ? CLB_NONE : tab[s].cluster;
        A->rec[i].key = ci;
        if (dense_on && ci < caps.clusters_per_frame) dense[(size_t)A->band * caps.clusters_per_frame + ci] = A->rec[i].cnt;
    }
}

// Dense form of the prefix pass, for small batches: walking a frame's bands one after the other (the kernel below) is a chain of
// `nbands` dependent steps -- 107 us for the 358 one-row bands of a single 1280x720 frame, a sixth of that frame's latency.  With
// the counts in a (band x cluster) matrix the running position of every cluster is a column scan: one thread per cluster, the loads
// of a column independent of each other (only the stored values form a chain), coalesced across clusters.  Only when no band of
// the batch needed a chained sub-band (pool_counter == 0; otherwise the walk below does the work) and the matrix fits a fixed
// budget (api.cu); the throughput configuration (256 frames, four-row bands) keeps the walk.
// (128 threads = 32 clusters x 4 quarters of the band range: every thread sums its quarter, the quarters' sums give its start, then it
// writes the running positions of its quarter -- twice the loads, a quarter of the chain.)
__global__ void __launch_bounds__(128)
cluster_band_prefix_dense_kernel(uint32_t *__restrict__ dense, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ nclusters,
                                 const uint32_t *__restrict__ pool_counter, Caps caps, BandPlan bp)
{
    __shared__ uint32_t part[4][32];
    if (*pool_counter != 0) return;
    const int b = blockIdx.y, lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * 32 + lane;
    const bool on = c < min(nclusters[b], caps.clusters_per_frame);
    const int per = (bp.nbands + 3) / 4, j_lo = min(q * per, bp.nbands), j_hi = min(j_lo + per, bp.nbands);
    uint32_t *col = dense + (size_t)b * bp.nbands * caps.clusters_per_frame + c;
    constexpr int U = 16;
    uint32_t sum = 0;
    if (on)
        for (int j0 = j_lo; j0 < j_hi; j0 += U) {
            uint32_t v[U];
#pragma unroll
            for (int k = 0; k < U; k++) v[k] = j0 + k < j_hi ? col[(size_t)(j0 + k) * caps.clusters_per_frame] : 0u;
#pragma unroll
            for (int k = 0; k < U; k++) sum += v[k];
        }
    part[q][lane] = sum;
    __syncthreads();
    if (!on) return;
    uint32_t run = clusters[(size_t)b * caps.clusters_per_frame + c].offset;
    for (int k = 0; k < q; k++) run += part[k][lane];
    for (int j0 = j_lo; j0 < j_hi; j0 += U) {
        uint32_t v[U];
#pragma unroll
        for (int k = 0; k < U; k++) v[k] = j0 + k < j_hi ? col[(size_t)(j0 + k) * caps.clusters_per_frame] : 0u;
#pragma unroll
        for (int k = 0; k < U; k++)
            if (j0 + k < j_hi) { col[(size_t)(j0 + k) * caps.clusters_per_frame] = run; run += v[k]; }
    }
}

// prefix: one CTA per frame walks the frame's bands (and their chained sub-bands) in scan order; record t of an area is thread t's.
// cursors: clusters_per_frame words in shared memory when they fit, otherwise the caller passes a global array.
__global__ void __launch_bounds__(CLB_CAP)
cluster_band_prefix_kernel(ClbArea *__restrict__ areas, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ nclusters,
                           uint32_t *__restrict__ global_cursors, Caps caps, BandPlan bp, const uint32_t *__restrict__ dense_done_unless)
{
    extern __shared__ uint32_t s_cursor[];
    const int b = blockIdx.x, t = threadIdx.x;
    if (dense_done_unless != nullptr && *dense_done_unless == 0) return;      // the dense form above did the work (no chained sub-band)
    uint32_t *cur = global_cursors ? global_cursors + (size_t)b * caps.clusters_per_frame : s_cursor;
    const uint32_t ncl = min(nclusters[b], caps.clusters_per_frame);
    for (uint32_t c = t; c < ncl; c += CLB_CAP) cur[c] = clusters[(size_t)b * caps.clusters_per_frame + c].offset;
    __syncthreads();
    // The first areas of the next PF bands are loaded together (their addresses do not depend on anything), so the walk pays one
    // memory round trip per PF bands instead of one per band; chained sub-bands (rare) are loaded on demand.
    constexpr int PF = 8;
    for (int band0 = 0; band0 < bp.nbands; band0 += PF) {
        uint32_t n_used[PF], next[PF];
        ClbRec r[PF];
#pragma unroll
        for (int k = 0; k < PF; k++) {
            n_used[k] = 0; next[k] = CLB_NONE; r[k].key = CLB_NONE; r[k].cnt = 0; r[k].slot = 0;
            if (band0 + k < bp.nbands) {
                const ClbArea *A = areas + (size_t)b * bp.nbands + band0 + k;
                n_used[k] = A->n_used; next[k] = A->next; r[k] = A->rec[t];
            }
        }
#pragma unroll
        for (int k = 0; k < PF; k++) {
            if (band0 + k >= bp.nbands) break;
            uint32_t a = (uint32_t)b * bp.nbands + band0 + k, nu = n_used[k], nx = next[k];
            ClbRec rr = r[k];
            for (;;) {
                if ((uint32_t)t < nu) {
                    const uint32_t c = (uint32_t)rr.key;
                    uint32_t base = CLB_NONE;
                    if (c < ncl) { base = cur[c]; cur[c] = base + rr.cnt; }      // the records of one area name distinct clusters
                    areas[a].rec[t].cnt = base;
                }
                __syncthreads();
                if (nx == CLB_NONE) break;
                a = nx;
                nu = areas[a].n_used; nx = areas[a].next; rr = areas[a].rec[t];
            }
        }
    }
}

// scatter: one warp per (sub-)band, CLB_SC_WARPS independent warps per CTA (an SM holds at most 32 CTAs, so one-warp CTAs cap it at half
// of its warps, and this pass waits on memory)
constexpr int CLB_SC_WARPS = 4;
__global__ void __launch_bounds__(CLB_SC_WARPS * 32)
cluster_band_scatter_kernel(const uint32_t *__restrict__ stage, const ClbArea *__restrict__ areas, const uint32_t *__restrict__ pool_counter,
                            uint32_t *__restrict__ scankey, Geom g, Caps caps, BandPlan bp, const uint32_t *__restrict__ dense)
{
    __shared__ uint32_t cur_all[CLB_SC_WARPS][CLB_CAP];
    uint32_t *cur = cur_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t full = 0xffffffffu, lt = (1u << lane) - 1u;
    const uint32_t a = blockIdx.x * CLB_SC_WARPS + (threadIdx.x >> 5);
    const uint32_t nfirst = (uint32_t)g.batch * bp.nbands;
    if (a >= nfirst + min(*pool_counter, bp.pool_cap)) return;
    const ClbArea *A = areas + a;
    const uint32_t n_used = A->n_used, st_start = A->st_start, st_end = A->st_end, job = A->band;
    if (st_end <= st_start) return;
    const int b = job / bp.nbands, band = job % bp.nbands, y0 = 1 + band * bp.rows;
    for (int s = lane; s < CLB_CAP; s += 32) cur[s] = CLB_NONE;
    __syncwarp();
    bool any = false;
    const bool dense_on = dense != nullptr && *pool_counter == 0;
    for (uint32_t i = lane; i < n_used; i += 32) {
        const ClbRec r = A->rec[i];
        uint32_t first = r.cnt;                                 // the walk's result, or (dense form) cell (band, cluster) of the scanned matrix
        if (dense_on) first = (uint32_t)r.key < caps.clusters_per_frame ? dense[(size_t)job * caps.clusters_per_frame + (uint32_t)r.key] : CLB_NONE;
        cur[r.slot] = first;
        any |= first != CLB_NONE;
    }
    if (!__any_sync(full, any)) return;                         // no selected cluster crosses this band
    __syncwarp();
    const uint32_t *st = stage + band_stage_base(g, b, y0);
    uint32_t *out = scankey + (size_t)b * caps.points_per_frame;
    uint32_t word = st_start + lane < st_end ? st[st_start + lane] : 0u;
    for (uint32_t i0 = st_start; i0 < st_end; i0 += 32) {
        const uint32_t w = word;
        const bool active = i0 + lane < st_end;
        word = i0 + 32 + lane < st_end ? st[i0 + 32 + lane] : 0u;
        const uint32_t slot = active ? (w >> 24) : 256u + (uint32_t)lane;          // idle lanes: a group of their own
        const uint32_t peers = __match_any_sync(full, slot);
        const int leader = __ffs(peers) - 1;
        uint32_t base = CLB_NONE;
        if (active && lane == leader) { base = cur[slot]; if (base != CLB_NONE) cur[slot] = base + __popc(peers); }
        base = __shfl_sync(full, base, leader);
        if (active && base != CLB_NONE) {
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
    if (ok && ci >= caps.clusters_per_frame) { flag_overflow(errflag, b, ERR_CLUSTERS_FULL); ok = false; }
    if (ok && off + cnt > caps.points_per_frame) { flag_overflow(errflag, b, ERR_POINTS_FULL); ok = false; }
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
