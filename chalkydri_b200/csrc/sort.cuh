// sort.cuh -- the two sorts of fit_quad() (row A5 of SURVEY.md 8a; upstream apriltag_quad_thresh.c fit_quad / ptsort).
//
// Upstream sorts every cluster by slope with ptsort(): a top-down merge sort split at sz/2 whose leaves (<= 5 points) are
// sorting networks that swap only when strictly greater, and whose merges take from the SECOND half unless the first is
// strictly smaller.  Slopes are floats with 7..8 fractional bits in three of the four quadrants, so equal keys are common
// and the order of equal keys -- which decides the order of the sequential line-fit sums -- is whatever that particular
// tree produces from the order the points were appended in (scan order).
//
// That order has a closed form.  After the leaf networks ran, let q be the position of a point in the array, L(q) the
// first position of its leaf.  By induction over the tree every node's output is its points ordered by
//     (slope ascending, L descending, q ascending):
// a merge emits an element of the second half before any element of the first half that is not strictly smaller.
// So ptsort() == leaf networks (emulated exactly, one thread per leaf) followed by ANY sort on the unique 64-bit key
//     slope << 32 | (0xffffff - L) << 3 | (q - L),
// and the scan-order sort before it has unique keys anyway.  Both therefore run on one generic block sort built for the
// SM instead of for upstream's tree:
//   * every thread sorts E consecutive elements in registers (bitonic network, compile-time indices);
//   * log2(n / E) merge passes over shared memory.  Per pass a thread owns E consecutive outputs: ONE merge-path binary
//     search finds where they start in the two runs, then it loads the next E of each run (2E independent loads),
//     takes x[i] = min(a[i], b[E-1-i]) -- a bitonic sequence holding exactly the next E outputs -- and finishes with a
//     log2(E)-stage in-register bitonic merge.  No dependent load -> compare -> load chain, no second search;
//   * arrays are padded by one element per E so the blocked accesses (stride E) are bank-conflict free.
#pragma once
#include "common.cuh"

namespace cb {

template <int NT>
struct Grp {
    static __device__ __forceinline__ int tid() { return NT == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
    static __device__ __forceinline__ void sync() { if (NT == 32) __syncwarp(); else __syncthreads(); }
    template <typename T, typename Op>
    static __device__ __forceinline__ T reduce(T v, Op op, T *scratch)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (NT == 32) return v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) scratch[wid] = v;
        __syncthreads();
        T r = scratch[0];
#pragma unroll
        for (int k = 1; k < NT / 32; k++) r = op(r, scratch[k]);
        return r;
    }
};

__device__ __forceinline__ uint32_t float_orderable(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// scan key layout (clusters.cuh): (pixel index << 3) | (probe << 1) | (v1 > v0)
__device__ __forceinline__ void decode_point(uint32_t key, int w, int &px, int &py, int &gx, int &gy)
{
    const uint32_t pix = key >> 3;
    const int d = (key >> 1) & 3, s = key & 1;
    const int x = pix % w, y = pix / w;
    const int dx = d == 2 ? -1 : (d == 1 ? 0 : 1), dy = d == 0 ? 0 : 1;
    const int dv = s ? 255 : -255;
    px = 2 * x + dx; py = 2 * y + dy; gx = dx * dv; gy = dy * dv;
}

// ---- generic block sort of unique keys ------------------------------------------------------------------------------
__device__ __forceinline__ void sort_cas(uint32_t &a, uint32_t &b)
{
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
__device__ __forceinline__ void sort_cas(unsigned long long &a, unsigned long long &b)
{
    const bool sw = b < a;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}
__device__ __forceinline__ uint32_t sort_min(uint32_t a, uint32_t b) { return min(a, b); }
__device__ __forceinline__ unsigned long long sort_min(unsigned long long a, unsigned long long b) { return b < a ? b : a; }

template <int E, bool PAD>
__device__ __forceinline__ int sidx(int p) { return PAD ? p + (int)((unsigned)p / (unsigned)E) : p; }

template <int E, typename T>
__device__ __forceinline__ void bitonic_merge_regs(T (&x)[E])
{
#pragma unroll
    for (int s = E / 2; s >= 1; s >>= 1)
#pragma unroll
        for (int i = 0; i < E; i++)
            if ((i & s) == 0) sort_cas(x[i], x[i + s]);
}

template <int E, typename T>
__device__ __forceinline__ void bitonic_sort_regs(T (&x)[E])
{
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
        for (int i = 0; i < E; i++) {
            const int j = i ^ (k - 1);
            if (j > i) sort_cas(x[i], x[j]);
        }
#pragma unroll
        for (int s = k >> 2; s >= 1; s >>= 1)
#pragma unroll
            for (int i = 0; i < E; i++)
                if ((i & s) == 0) sort_cas(x[i], x[i + s]);
    }
}

// Sorts the n unique keys of src (layout sidx<E, PAD>) ascending; src and dst are swapped per pass, the result is in src.
// All ones is reserved as the padding value.
template <int NT, int E, bool PAD, typename T>
__device__ __forceinline__ void block_sort(T *&src, T *&dst, int n, int tid)
{
    const T SENT = ~(T)0;
    const int nchunks = (n + E - 1) / E;
    for (int c = tid; c < nchunks; c += NT) {
        T x[E];
        const int base = sidx<E, PAD>(c * E);
#pragma unroll
        for (int i = 0; i < E; i++) x[i] = c * E + i < n ? src[base + i] : SENT;
        bitonic_sort_regs<E>(x);
#pragma unroll
        for (int i = 0; i < E; i++)
            if (c * E + i < n) src[base + i] = x[i];
    }
    Grp<NT>::sync();
    const int npad = nchunks * E;
    for (int run = E; run < npad; run <<= 1) {
        for (int c = tid; c < nchunks; c += NT) {
            const int p = c * E;
            const int lo = p & ~(2 * run - 1);
            const int mid = min(lo + run, npad), hi = min(lo + 2 * run, npad);
            const int an = mid - lo, bn = hi - mid, k = p - lo;
            const int obase = sidx<E, PAD>(p);
            T x[E];
            if (bn <= 0) {                      // unpaired run at the end: copy
#pragma unroll
                for (int i = 0; i < E; i++) x[i] = p + i < n ? src[obase + i] : SENT;
            } else {
                // merge path: l = how many of the first k outputs come from the first run
                int l = max(0, k - bn), h = min(k, an);
                while (l < h) {
                    const int m = (l + h) >> 1;
                    const int ia = lo + m, ib = mid + k - m - 1;
                    const T av = ia < n ? src[sidx<E, PAD>(ia)] : SENT;
                    const T bv = ib < n ? src[sidx<E, PAD>(ib)] : SENT;
                    if (av < bv) l = m + 1; else h = m;
                }
                const int a0 = lo + l, b0 = mid + (k - l);
#pragma unroll
                for (int i = 0; i < E; i++) {
                    const int ia = a0 + i, ib = b0 + (E - 1 - i);
                    const T av = (ia < mid && ia < n) ? src[sidx<E, PAD>(ia)] : SENT;
                    const T bv = (ib < hi && ib < n) ? src[sidx<E, PAD>(ib)] : SENT;
                    x[i] = sort_min(av, bv);
                }
                bitonic_merge_regs<E>(x);
            }
#pragma unroll
            for (int i = 0; i < E; i++)
                if (p + i < n) dst[obase + i] = x[i];
        }
        Grp<NT>::sync();
        T *t = src; src = dst; dst = t;
    }
}

struct SortScratch {      // per group (warp or CTA)
    double red_d[16];
    int red_i[16];
    float red_f[16];
    int work;
};

// ---- sort #1: bounding box, border polarity, scan-order sort ----------------------------------------------------------
// Rejected clusters (too small a box, reversed border -- tag36h11 has a normal border only) get cursor = 0xffffffff and
// are skipped by every later kernel.  The sorted scan keys replace the unsorted ones in K.
template <int NT, int E, bool PAD>
__device__ __forceinline__ void sort_scan_cluster(uint32_t *__restrict__ K, int n, uint32_t *A, uint32_t *B, SortScratch &S,
                                                  ClusterRec *__restrict__ rec_global, const Geom &g, const DetParams &prm)
{
    typedef Grp<NT> G;
    const int tid = G::tid();
    int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
    for (int i = tid; i < n; i += NT) {
        const uint32_t key = K[i];
        A[sidx<E, PAD>(i)] = key;
        int px, py, gx, gy;
        decode_point(key, g.w, px, py, gx, gy);
        xmin = min(xmin, px); xmax = max(xmax, px); ymin = min(ymin, py); ymax = max(ymax, py);
    }
    xmin = G::reduce(xmin, [](int a, int c) { return min(a, c); }, S.red_i);
    xmax = G::reduce(xmax, [](int a, int c) { return max(a, c); }, S.red_i);
    ymin = G::reduce(ymin, [](int a, int c) { return min(a, c); }, S.red_i);
    ymax = G::reduce(ymax, [](int a, int c) { return max(a, c); }, S.red_i);
    const float cx = (float)((xmin + xmax) * 0.5 + 0.05118);
    const float cy = (float)((ymin + ymax) * 0.5 + -0.028581);
    G::sync();
    if ((xmax - xmin) * (ymax - ymin) < prm.min_tag_width) { if (tid == 0) rec_global->cursor = 0xffffffffu; return; }
    // border polarity (upstream: float dot = sum over the points IN SCAN ORDER of dx*gx + dy*gy; reversed_border = dot < 0).
    // The per-point terms are independent and evaluated exactly as upstream does; only the float accumulation is order
    // dependent.  Its result differs from the exact sum S of the terms by at most (n - 1) * 2^-24 * sum|t|, so when |S|
    // clears twice that bound the sign is known before any sorting (the usual case, and half of all clusters end here).
    // Otherwise (thin, nearly symmetric clusters whose terms cancel) the chain is replayed in scan order after the sort.
    double sum = 0, sum_abs = 0;
    for (int i = tid; i < n; i += NT) {
        int px, py, gx, gy;
        decode_point(A[sidx<E, PAD>(i)], g.w, px, py, gx, gy);
        const float dx = (float)px - cx, dy = (float)py - cy;
        const float t = dx * (float)gx + dy * (float)gy;
        sum += (double)t; sum_abs += fabs((double)t);
    }
    sum = G::reduce(sum, [](double a, double c) { return a + c; }, S.red_d);
    sum_abs = G::reduce(sum_abs, [](double a, double c) { return a + c; }, S.red_d);
    const bool certain = fabs(sum) > 2.0 * (double)n * 5.9604644775390625e-8 * sum_abs;
    if (certain && sum < 0) { if (tid == 0) rec_global->cursor = 0xffffffffu; return; }
    G::sync();
    uint32_t *src = A, *dst = B;
    block_sort<NT, E, PAD>(src, dst, n, tid);
    if (!certain) {
        if (tid == 0) {
            float dot = 0.f;
            for (int i = 0; i < n; i++) {
                int px, py, gx, gy;
                decode_point(src[sidx<E, PAD>(i)], g.w, px, py, gx, gy);
                const float dx = (float)px - cx, dy = (float)py - cy;
                dot += dx * (float)gx + dy * (float)gy;
            }
            S.work = dot < 0.f ? 1 : 0;
            if (dot < 0.f) rec_global->cursor = 0xffffffffu;
        }
        G::sync();
        if (S.work) return;
    }
    for (int j = tid; j < n; j += NT) K[j] = src[sidx<E, PAD>(j)];
}

// ---- bucket form of the slope sort -------------------------------------------------------------------------------------------
// The composite keys are unique, so ANY correct sort gives upstream's order -- and the slope is an angle, spread roughly evenly
// over a closed boundary.  So instead of O(n log n) merge passes: a monotone map from the float slope to one of ~n/4 buckets
// (same order as the keys: every step of it is a monotone function evaluated without contraction), a shared-memory histogram,
// a prefix sum, a scatter, and the exact position of a point inside its bucket by counting the smaller keys among the handful
// of points that share it.  ~100 instructions per point where the merge sort spent well over a thousand.  Clusters whose
// points crowd into a few directions (a bucket above SORT_BUCKET_LIMIT: long thin shapes seen from inside) take the merge sort.
constexpr int SORT_BUCKET_LIMIT = 96;

// monotone non-decreasing in the float `slope` (= quadrant + dy/dx, quadrant in {-65536, 0, 65536, 131072}, dy/dx >= 0)
__device__ __forceinline__ int slope_bucket(uint32_t orderable, int nb4, float scale)
{
    const uint32_t bits = (orderable & 0x80000000u) ? (orderable & 0x7fffffffu) : ~orderable;     // inverse of float_orderable
    const float s = __uint_as_float(bits);
    int q;
    float t;
    if (s >= 131072.f) { q = 3; t = s - 131072.f; }
    else if (s >= 65536.f) { q = 2; t = s - 65536.f; }
    else if (s >= 0.f) { q = 1; t = s; }
    else { q = 0; t = s + 65536.f; }                        // (the differences are exact: Sterbenz)
    const float u = t < 1.f ? 0.5f * t : 1.f - 0.5f / t;    // [0, 1): tan -> something close to the angle; each branch monotone, equal at 1
    int k = (int)(u * scale);
    k = max(0, min(k, nb4 - 1));
    return q * nb4 + k;
}

// exclusive prefix sum of cnt[0 .. nb) into base[0 .. nb) by a group of NT threads (nb a power of two >= 32); returns the largest count
template <int NT>
__device__ __forceinline__ uint32_t bucket_scan(const uint32_t *cnt, uint32_t *base, int nb, int tid, SortScratch &S)
{
    typedef Grp<NT> G;
    const int per = nb >= NT ? nb / NT : 1;
    const bool on = tid * per < nb;
    uint32_t sum = 0, mx = 0;
    if (on)
        for (int k = 0; k < per; k++) { const uint32_t c = cnt[tid * per + k]; sum += c; mx = max(mx, c); }
    // inclusive scan of `sum` across the group
    const int lane = tid & 31;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    uint32_t warp_off = 0;
    if (NT > 32) {
        const int wid = tid >> 5;
        __syncthreads();
        if (lane == 31) S.red_i[wid] = (int)incl;
        __syncthreads();
        for (int k = 0; k < wid; k++) warp_off += (uint32_t)S.red_i[k];
    }
    uint32_t run = warp_off + incl - sum;
    if (on)
        for (int k = 0; k < per; k++) { const uint32_t c = cnt[tid * per + k]; base[tid * per + k] = run; run += c; }
    mx = G::reduce(mx, [](uint32_t a, uint32_t c) { return max(a, c); }, reinterpret_cast<uint32_t *>(S.red_f));
    return mx;
}

// ---- sort #2: slopes in scan order, ptsort() ------------------------------------------------------------------------
// K holds the scan-ordered keys on entry and the sorted points (px | py << 16) on return.
// CHECK: K arrives in scan order straight from the band scatter pass (clusters.cuh), and the two tests sort #1 used to make --
// bounding box against min_tag_width, border polarity -- happen here before any sorting; a rejected cluster gets
// cursor = 0xffffffff.  Returns false for a rejected cluster.
template <int NT, int E, bool PAD, bool CHECK>
__device__ __forceinline__ bool sort_slope_cluster(uint32_t *__restrict__ K, int n, unsigned long long *A, unsigned long long *B,
                                                   SortScratch &S, const Geom &g, ClusterRec *__restrict__ rec_global, const DetParams &prm,
                                                   uint32_t *hist, int hist_buckets)
{
    typedef Grp<NT> G;
    const int tid = G::tid();
    int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
    for (int i = tid; i < n; i += NT) {
        int px, py, gx, gy;
        decode_point(K[i], g.w, px, py, gx, gy);
        xmin = min(xmin, px); xmax = max(xmax, px); ymin = min(ymin, py); ymax = max(ymax, py);
    }
    xmin = G::reduce(xmin, [](int a, int c) { return min(a, c); }, S.red_i);
    xmax = G::reduce(xmax, [](int a, int c) { return max(a, c); }, S.red_i);
    ymin = G::reduce(ymin, [](int a, int c) { return min(a, c); }, S.red_i);
    ymax = G::reduce(ymax, [](int a, int c) { return max(a, c); }, S.red_i);
    const float cx = (float)((xmin + xmax) * 0.5 + 0.05118);
    const float cy = (float)((ymin + ymax) * 0.5 + -0.028581);
    if (CHECK) {
        if ((xmax - xmin) * (ymax - ymin) < prm.min_tag_width) { if (tid == 0) rec_global->cursor = 0xffffffffu; return false; }
        // border polarity: see sort_scan_cluster -- the sign of upstream's float chain is known from the exact sum unless the
        // terms nearly cancel; then the chain is replayed, and the points already are in the order upstream adds them in
        double sum = 0, sum_abs = 0;
        for (int i = tid; i < n; i += NT) {
            int px, py, gx, gy;
            decode_point(K[i], g.w, px, py, gx, gy);
            const float dx = (float)px - cx, dy = (float)py - cy;
            const float t = dx * (float)gx + dy * (float)gy;
            sum += (double)t; sum_abs += fabs((double)t);
        }
        sum = G::reduce(sum, [](double a, double c) { return a + c; }, S.red_d);
        sum_abs = G::reduce(sum_abs, [](double a, double c) { return a + c; }, S.red_d);
        const bool certain = fabs(sum) > 2.0 * (double)n * 5.9604644775390625e-8 * sum_abs;
        if (certain && sum < 0) { if (tid == 0) rec_global->cursor = 0xffffffffu; return false; }
        if (!certain) {
            G::sync();
            if (tid == 0) {
                float dot = 0.f;
                for (int i = 0; i < n; i++) {
                    int px, py, gx, gy;
                    decode_point(K[i], g.w, px, py, gx, gy);
                    const float dx = (float)px - cx, dy = (float)py - cy;
                    dot += dx * (float)gx + dy * (float)gy;
                }
                S.work = dot < 0.f ? 1 : 0;
                if (dot < 0.f) rec_global->cursor = 0xffffffffu;
            }
            G::sync();
            if (S.work) return false;
        }
    }
    // slopes in scan order (upstream fit_quad step 1)
    for (int j = tid; j < n; j += NT) {
        int px, py, gx, gy;
        decode_point(K[j], g.w, px, py, gx, gy);
        float dx = (float)px - cx, dy = (float)py - cy;
        float quadrant;
        if (dy > 0) quadrant = dx > 0 ? 65536.f : 131072.f; else quadrant = dx > 0 ? 0.f : -65536.f;
        if (dy < 0) { dy = -dy; dx = -dx; }
        if (dx < 0) { const float t = dx; dx = dy; dy = -t; }
        const float slope = quadrant + dy / dx;
        B[j] = ((unsigned long long)float_orderable(slope) << 32) | (uint32_t)px | ((uint32_t)py << 16);
    }
    G::sync();
    // leaves of ptsort()'s recursion tree, one thread each: upstream's networks on the slope, then the composite key
    for (int i = tid; i < n; i += NT) {
        int lo = 0, hi = n;
        while (hi - lo > 5) { const int mid = lo + (hi - lo) / 2; if (i < mid) hi = mid; else lo = mid; }
        if (i != lo) continue;
        const int sz = hi - lo;
        unsigned long long *a = B + lo;
#define QF_SWAP(x, y) if ((uint32_t)(a[x] >> 32) > (uint32_t)(a[y] >> 32)) { const unsigned long long t = a[x]; a[x] = a[y]; a[y] = t; }
        if (sz == 2) { QF_SWAP(0, 1); }
        else if (sz == 3) { QF_SWAP(0, 1); QF_SWAP(1, 2); QF_SWAP(0, 1); }
        else if (sz == 4) { QF_SWAP(0, 1); QF_SWAP(2, 3); QF_SWAP(0, 2); QF_SWAP(1, 3); QF_SWAP(1, 2); }
        else if (sz == 5) { QF_SWAP(0, 1); QF_SWAP(3, 4); QF_SWAP(2, 4); QF_SWAP(2, 3); QF_SWAP(0, 3); QF_SWAP(0, 2); QF_SWAP(1, 4); QF_SWAP(1, 3); QF_SWAP(1, 2); }
#undef QF_SWAP
        const uint32_t tie = (0xffffffu - (uint32_t)lo) << 3;
        for (int j = 0; j < sz; j++) {
            const unsigned long long v = a[j];
            A[sidx<E, PAD>(lo + j)] = (v & 0xffffffff00000000ull) | tie | (uint32_t)j;
            K[lo + j] = (uint32_t)v;
        }
    }
    G::sync();
    if (hist != nullptr) {
        // bucket form: histogram -> prefix -> scatter -> exact rank inside the bucket
        int nb = 32;
        while (nb < hist_buckets && nb * 4 < n) nb <<= 1;
        const int nb4 = nb >> 2;
        const float scale = (float)nb4;
        uint32_t *cnt = hist, *base = hist + hist_buckets;
        for (int k = tid; k < nb; k += NT) cnt[k] = 0;
        G::sync();
        for (int i = tid; i < n; i += NT) atomicAdd(&cnt[slope_bucket((uint32_t)(A[sidx<E, PAD>(i)] >> 32), nb4, scale)], 1u);
        G::sync();
        const uint32_t biggest = bucket_scan<NT>(cnt, base, nb, tid, S);
        G::sync();
        if (biggest <= (uint32_t)SORT_BUCKET_LIMIT) {
            for (int k = tid; k < nb; k += NT) cnt[k] = base[k];                   // cursors
            G::sync();
            for (int i = tid; i < n; i += NT) {
                const unsigned long long key = A[sidx<E, PAD>(i)];
                B[atomicAdd(&cnt[slope_bucket((uint32_t)(key >> 32), nb4, scale)], 1u)] = key;
            }
            G::sync();
            uint32_t *stage = reinterpret_cast<uint32_t *>(A);               // the keys now live in B
            for (int p = tid; p < n; p += NT) {
                const unsigned long long key = B[p];
                const int bk = slope_bucket((uint32_t)(key >> 32), nb4, scale);
                const uint32_t s0 = base[bk], s1 = cnt[bk];                     // the cursor ended at the bucket's end
                uint32_t r = s0;
                for (uint32_t q = s0; q < s1; q++) r += B[q] < key ? 1u : 0u;
                const uint32_t tie = (uint32_t)key;
                stage[r] = K[(0xffffffu - (tie >> 3)) + (tie & 7u)];
            }
            G::sync();
            for (int j = tid; j < n; j += NT) K[j] = stage[j];
            return true;
        }
    }
    unsigned long long *src = A, *dst = B;
    block_sort<NT, E, PAD>(src, dst, n, tid);
    uint32_t *stage = reinterpret_cast<uint32_t *>(dst);
    for (int j = tid; j < n; j += NT) {
        const uint32_t tie = (uint32_t)src[sidx<E, PAD>(j)];
        stage[j] = K[(0xffffffu - (tie >> 3)) + (tie & 7u)];
    }
    G::sync();
    for (int j = tid; j < n; j += NT) K[j] = stage[j];
    return true;
}

// ---- kernels: persistent groups pulling (frame, cluster) items from a range of the tier work lists ----------------------
// item of position wi in the concatenation of lists T_HI, T_HI - 1, ..., T_LO (largest clusters first)
template <int T_LO, int T_HI>
__device__ __forceinline__ bool tier_item(uint32_t wi, const uint32_t *__restrict__ nwork /* stride 2 */,
                                          const uint32_t *__restrict__ worklists, size_t wl_stride, uint32_t &item)
{
#pragma unroll
    for (int t = T_HI; t >= T_LO; t--) {
        const uint32_t c = nwork[2 * t];
        if (wi < c) { item = worklists[(size_t)t * wl_stride + wi]; return true; }
        wi -= c;
    }
    return false;
}

template <int E> constexpr int sort_padded(int n) { return n + n / E + 1; }

// Configuration of one sort kernel.  WHICH = 1: sort_scan_cluster on u32, WHICH = 2: sort_slope_cluster on u64, WHICH = 3:
// sort_slope_cluster with sort #1's box / polarity tests in front (points arrive in scan order, no sort #1 ran).
// NT = 32: SORT_WARPS clusters per CTA, one per warp; otherwise one cluster per CTA of NT threads.  The kernel pulls from
// work lists T_LO..T_HI and processes the clusters of NMIN..NMAX points (several kernels can share one list); above MAXN
// points the work arrays live in the global scratch area.
constexpr int SORT_WARPS = 8;
template <int NT_, int E_, int MAXN_, int WHICH_, int T_LO_, int T_HI_, int NMIN_, int NMAX_>
struct SortCfg {
    static constexpr int NT = NT_, E = E_, MAXN = MAXN_, WHICH = WHICH_, T_LO = T_LO_, T_HI = T_HI_, NMIN = NMIN_, NMAX = NMAX_;
    static constexpr int GROUPS = NT == 32 ? SORT_WARPS : 1;
    static constexpr int THREADS = NT * GROUPS;
    static constexpr int ELEM = WHICH == 1 ? 4 : 8;
    static constexpr size_t ARRAY_BYTES = ((size_t)sort_padded<E>(MAXN) * ELEM + 15) / 16 * 16;
    static constexpr int HIST_BUCKETS = WHICH == 1 ? 0 : MAXN / 4;               // bucket form of the slope sort: counts + starts
    static constexpr size_t SCRATCH_BYTES = (sizeof(SortScratch) + 15) / 16 * 16;
    static constexpr size_t GROUP_BYTES = 2 * ARRAY_BYTES + SCRATCH_BYTES + (size_t)2 * HIST_BUCKETS * sizeof(uint32_t);
    static constexpr size_t BYTES = GROUPS * GROUP_BYTES;
};

template <typename C>
__global__ void __launch_bounds__(C::THREADS)
sort_clusters_kernel(uint32_t *__restrict__ scankey, ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ worklists,
                     size_t wl_stride, const uint32_t *__restrict__ nwork, uint32_t *__restrict__ work_counter,
                     unsigned long long *__restrict__ scratch, Geom g, Caps caps, DetParams prm, int bucket_form)
{
    typedef C SH;
    constexpr int NT = C::NT, E = C::E, MAXN = C::MAXN, WHICH = C::WHICH, T_LO = C::T_LO, T_HI = C::T_HI, NMIN = C::NMIN, NMAX = C::NMAX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = NT == 32 ? (int)(threadIdx.x >> 5) : 0;
    unsigned char *base = smem_raw + (size_t)grp * SH::GROUP_BYTES;
    SortScratch &S = *reinterpret_cast<SortScratch *>(base + 2 * SH::ARRAY_BYTES);
    const int tid = Grp<NT>::tid();
    for (;;) {
        uint32_t wi;
        if (NT == 32) {
            wi = 0;
            if (tid == 0) wi = atomicAdd(work_counter, 1u);
            wi = __shfl_sync(0xffffffffu, wi, 0);
        } else {
            __syncthreads();
            if (tid == 0) S.work = (int)atomicAdd(work_counter, 1u);
            __syncthreads();
            wi = (uint32_t)S.work;
        }
        uint32_t item;
        if (!tier_item<T_LO, T_HI>(wi, nwork, worklists, wl_stride, item)) return;
        const int b = item / caps.clusters_per_frame;
        const ClusterRec rec = clusters[item];
        const int n = (int)rec.count;
        if (n < 24 || n < NMIN || n > NMAX) continue;
        if (WHICH == 2 && rec.cursor == 0xffffffffu) continue;
        constexpr bool CHECK = WHICH == 3;
        const size_t pbase = (size_t)b * caps.points_per_frame + rec.offset;
        // NMAX <= MAXN: the global-memory variant is never needed (and not compiled into the kernel).
        // (A half-E variant for the small clusters of the one-warp tiers measured slower: twice the code in the kernel.)
        constexpr bool NEVER_GLOBAL = NMAX <= MAXN;
        constexpr int EH = E;
        if (WHICH == 1) {
            uint32_t *A = reinterpret_cast<uint32_t *>(base), *B = reinterpret_cast<uint32_t *>(base + SH::ARRAY_BYTES);
            if (EH != E && n <= MAXN / 2) sort_scan_cluster<NT, EH, true>(scankey + pbase, n, A, B, S, clusters + item, g, prm);
            else if (NEVER_GLOBAL || n <= MAXN) sort_scan_cluster<NT, E, true>(scankey + pbase, n, A, B, S, clusters + item, g, prm);
            else {
                uint32_t *GA = reinterpret_cast<uint32_t *>(scratch + pbase * 2);
                sort_scan_cluster<NT, E, false>(scankey + pbase, n, GA, GA + n, S, clusters + item, g, prm);
            }
        } else {
            unsigned long long *A = reinterpret_cast<unsigned long long *>(base), *B = reinterpret_cast<unsigned long long *>(base + SH::ARRAY_BYTES);
            uint32_t *hist = bucket_form ? reinterpret_cast<uint32_t *>(base + 2 * SH::ARRAY_BYTES + SH::SCRATCH_BYTES) : nullptr;
            if (EH != E && n <= MAXN / 2) sort_slope_cluster<NT, EH, true, CHECK>(scankey + pbase, n, A, B, S, g, clusters + item, prm, hist, SH::HIST_BUCKETS);
            else if (NEVER_GLOBAL || n <= MAXN) sort_slope_cluster<NT, E, true, CHECK>(scankey + pbase, n, A, B, S, g, clusters + item, prm, hist, SH::HIST_BUCKETS);
            else {
                unsigned long long *GA = scratch + pbase * 2;
                sort_slope_cluster<NT, E, false, CHECK>(scankey + pbase, n, GA, GA + n, S, g, clusters + item, prm, nullptr, 0);      // above the shared-memory tiers: merge sort in global memory
            }
        }
        if (NT == 32) __syncwarp();
    }
}

}  // namespace cb
