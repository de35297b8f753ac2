// api.cu -- context management and the C ABI of libchalkydri_b200.so (include/chalkydri_b200.h).
//
// Host-side orchestration only: every stage of the hot path is a CUDA kernel from the headers below; there is
// no CPU implementation behind any entry point (a missing / non-sm_100 device makes cb_create fail).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "threshold.cuh"
#include "ccl.cuh"
#include "clusters.cuh"
#include "quads.cuh"
#include "decode.cuh"
#include "sqpnp.cuh"
#include "cat.cuh"

using namespace cb;

static const unsigned long long kHostCodes[kNumCodes] = {
#include "tag36h11_codes.inc"
};
static const int kHostBitX[36] = {1,2,3,4,5,2,3,4,3, 6,6,6,6,6,5,5,5,4, 6,5,4,3,2,5,4,3,4, 1,1,1,1,1,2,2,2,3};
static const int kHostBitY[36] = {1,1,1,1,1,2,2,2,3, 1,2,3,4,5,2,3,4,3, 6,6,6,6,6,5,5,5,4, 6,5,4,3,2,5,4,3,4};

static thread_local std::string g_create_error;

struct ThrChoice { int variant, ysegs; };
typedef std::tuple<int, int, int, int> ThrKey;      // W, H, stride, batch

enum Stage { ST_THRESH = 1, ST_LABELS = 2, ST_QUADS = 3, ST_FULL = 4, ST_CLUSTERS = 10 /* stop after the cluster passes (tap) */ };

struct cb_ctx {
    int device = 0;
    int max_w = 0, max_h = 0, max_batch = 0, max_dets = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;           // H2D of the next chunk overlaps the kernels of the current one
    cudaEvent_t ev_copied[2]{}, ev_consumed[2]{};
    cudaStream_t tier_stream[4]{};                 // the size tiers of the two cluster sorts run side by side
    cudaEvent_t ev_fork = nullptr, ev_tier[4]{};
    uint32_t *h_chunk_err = nullptr;              // pinned: error flag of every chunk of a pipelined call
    // streaming form (cb_detect_gray_submit / cb_detect_gray_collect): up to two batches in flight, each with its own pinned
    // staging, so the H2D copy of batch k+1 runs under the kernels of batch k
    struct StreamSlot {
        cb_detection *h_dets = nullptr;
        int32_t *h_counts = nullptr;
        uint32_t *h_err = nullptr;               // error flag per chunk
        uint32_t *h_ferr = nullptr;              // overflow word per frame
        cudaEvent_t start = nullptr, done = nullptr;
        int batch = 0, nchunks = 0, launches = 0, thr_launches = 0;
        // fused detect -> pose batches (cb_detect_pose_gray_submit): gyro staging in, poses out
        bool has_pose = false;
        double *h_gyro = nullptr;
        cb_pose *h_poses = nullptr;
        uint8_t *h_ok = nullptr;
        int32_t *h_ntags = nullptr;
        cudaEvent_t pose_done = nullptr;
    } ss[2];
    int ss_head = 0, ss_pending = 0;
    unsigned ss_chunk = 0;                       // batches submitted so far: parity picks d_in or d_in2
    uint8_t *d_in2 = nullptr;                    // second input buffer of the streaming form (allocated on the first submit)
    std::string err;
    bool family_set = false;
    DetParams prm{};
    DecodeConst dc{};
    int sq_max_iter = 15;
    bool sq_attr_set = false;
    double *d_sq_scratch = nullptr;      // intermediates of the three-phase solver (api_solver.inc)
    size_t sq_scratch_cap = 0;
    double sq_tol_sq = 1e-16;
    // fused detect -> pose (cb_detect_pose_gray): field layout, camera, per-frame SQPnP problems
    int32_t *d_field_ids = nullptr;
    cb_iso3 *d_field_poses = nullptr;
    int n_field = 0;
    double *d_cam9 = nullptr;            // 9 intrinsics, then robot_to_cam (cb_iso3) behind them
    bool camera_set = false;
    uint8_t *d_pose_buf = nullptr;       // [tags | bearings | n_tags | gyro | poses | ok] for pose_cap frames
    int pose_cap = 0;
    bool external_map = false;           // cb_cat_detect_tags: d_thresh was filled from CAT's colour map, run_pipeline skips A1+A2
    bool pose_active = false;            // run_pipeline appends the chunk's problems at frame pose_frame_base ...
    int pose_frame_base = 0;
    double pose_sign_change_error = 0;   // ... and solves them on pose_stream, under the kernels of the next chunk
    cudaStream_t pose_stream = nullptr;
    cudaEvent_t ev_pose_ready = nullptr, ev_pose_done = nullptr;
    // chunks of <= 4 frames (the reference's one-frame call): the solve is joined and its poses are read back inside the chunk, so the
    // whole detect -> pose call is one capturable sequence (queue_chunk); larger chunks leave the solve running under the next chunk
    bool pose_join = false;
    cb_pose *h_pose_stage = nullptr; uint8_t *h_pose_ok_stage = nullptr; int32_t *h_pose_nt_stage = nullptr;      // pinned, 4 frames
    cb_pose *pose_dst = nullptr; uint8_t *pose_ok_dst = nullptr; int32_t *pose_nt_dst = nullptr;                 // the caller's arrays
    int num_sms = 148;

    // device buffers (sized for max_batch frames of max_w x max_h at decimation >= 1)
    uint8_t *d_in = nullptr;      size_t in_bytes = 0;
    uint8_t *d_gray = nullptr;    size_t gray_bytes = 0;
    uint8_t *d_thresh = nullptr, *d_mark = nullptr;  size_t map_bytes = 0;
    uint8_t *d_tmin = nullptr, *d_tmax = nullptr;    size_t tile_bytes = 0;
    uint32_t *d_labels = nullptr, *d_sizes = nullptr; size_t label_bytes = 0;
    ClusterSlot *d_table = nullptr;
    uint8_t *d_cat = nullptr, *d_cat_tot = nullptr;   // CAT process_frame scratch (grow-only)
    size_t cat_bytes = 0, cat_tot_bytes = 0;
    int sort_bucket_form = 1;                  // slope sort: bucket form (round 2) or merge sort only (CB_SORT=merge, A/B hook)
    int cluster_mode = 0;                      // 0: band-ordered count / scatter (round 2), 1: tile passes + scan-order sort (round 1; CB_CLUSTERS=tiles)
    ClbArea *d_areas = nullptr;                // band tables of the count pass (first areas, then the pool of chained sub-band areas)
    uint32_t areas_first_cap = 0, areas_pool_cap = 0;
    uint32_t *d_band_dense = nullptr;          // (band x cluster) matrix of the dense prefix form (small batches; allocated on first use)
    size_t band_dense_cells = 0;
    uint32_t *d_cursors = nullptr;             // prefix-pass cursors when clusters_per_frame does not fit shared memory
    uint4 *d_ent = nullptr;                    // tiles: per-point words of the count pass, 16 B per decimated pixel; bands: the staging lists (same size)
    unsigned long long *d_tile_keys = nullptr;  // per-tile tables of the count pass
    uint32_t *d_tile_cnt = nullptr;
    ClusterRec *d_clusters = nullptr;
    uint32_t *d_worklist = nullptr;
    uint32_t *d_scankey = nullptr;
    double *d_errs = nullptr;       // window errors, 8 B per point
    double *d_cp = nullptr;         // prefix-moment checkpoints, 48 B per LF_CP points
    unsigned long long *d_scratch = nullptr;
    QuadRec *d_quads = nullptr;
    RawDet *d_raw = nullptr;
    cb_detection *d_dets = nullptr;
    int32_t *d_counts = nullptr;
    uint32_t *d_small = nullptr;   // [nclusters B][npoints B][nquads B][nraw B][misc 16]
    Caps caps{};
    size_t max_npix = 0;           // per frame, at decimation 1 (worst case)
    // pinned staging for outputs
    cb_detection *h_dets = nullptr;
    int32_t *h_counts = nullptr;
    uint32_t *h_small = nullptr;
    size_t out_block_bytes = 0;                   // d_dets | d_counts | d_small (and h_*) are one allocation of this size

    cudaEvent_t ev[10]{};
    cb_timing timing{};
    std::vector<uint32_t> frame_flags;            // overflow bits per frame of the last completed detection call (cb_frame_flags)
    uint32_t *h_frame_err = nullptr; size_t h_frame_err_cap = 0;      // pinned: the same for the chunks of a pipelined call
    std::map<ThrKey, ThrChoice> thr_plans;          // threshold kernel shape per frame geometry (threshold_plan)
    // small batches: the pipeline of one geometry captured as a CUDA graph on its second use (detect_device_chunk)
    struct GraphSlot { int seen = 0; bool failed = false; cudaGraphExec_t exec = nullptr; int launches = 0, thr_launches = 0; };
    std::map<std::tuple<int, int, int, int, const void *, int, double>, GraphSlot> graphs;      // W, H, stride, batch, device frames, 1 + first pose frame (0: no poses), sign_change_error
    bool graph_off = false;
};

static int fail(cb_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(ctx, CB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// the synchronous entry points share buffers with batches still in flight on the streaming form
#define CB_NOT_STREAMING(ctx)                                                                                                            \
    do {                                                                                                                                 \
        if ((ctx)->ss_pending) return fail(ctx, CB_ERR_STATE, "%d submitted batch(es) not collected yet: call cb_detect_gray_collect first", (ctx)->ss_pending); \
    } while (0)

static void drop_graphs(cb_ctx *ctx)
{
    for (auto &kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
}

static uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

__global__ void table_init_kernel(ClusterSlot *t, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { t[i].key = EMPTY_KEY; t[i].count = 0; t[i].cluster = 0xffffffffu; }
}

static void set_default_params(cb_ctx *ctx)
{
    DetParams &p = ctx->prm;
    p.quad_decimate = 2.0f; p.refine_edges = 1; p.decode_sharpening = 0.25; p.min_cluster_pixels = 5; p.max_nmaxima = 10;
    p.critical_rad = (float)(10 * M_PI / 180); p.max_line_fit_mse = 10.0f; p.min_white_black_diff = 5; p.bits_corrected = 3;
    p.cos_critical_rad = cos(p.critical_rad);
    for (int i = 0; i < 7; i++) { int j = i - 3; p.smooth_f[i] = (float)exp(-j * j / (2 * 1.0 * 1.0)); }
    int mtw = (int)(8 / p.quad_decimate); if (mtw < 3) mtw = 3;
    p.min_tag_width = mtw;
    for (int k = 0; k < 4; k++) { double th = k * M_PI / 2.0; ctx->dc.rot_c[k] = cos(th); ctx->dc.rot_s[k] = sin(th); }
}

extern "C" {

const char *cb_version(void) { return "chalkydri_b200 0.1 (sm_100a)"; }

int cb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *cb_last_error(const cb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void cb_destroy(cb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);      // batches submitted and never collected
    if (ctx->pose_stream) cudaStreamSynchronize(ctx->pose_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void *ptrs[] = {ctx->d_in, ctx->d_gray, ctx->d_thresh, ctx->d_mark, ctx->d_tmin, ctx->d_tmax, ctx->d_labels, ctx->d_sizes,
                    ctx->d_table, ctx->d_areas, ctx->d_cursors, ctx->d_cat, ctx->d_cat_tot, ctx->d_ent, ctx->d_tile_keys, ctx->d_tile_cnt, ctx->d_clusters, ctx->d_worklist, ctx->d_scankey, ctx->d_errs, ctx->d_cp, ctx->d_scratch,
                    ctx->d_quads, ctx->d_raw, ctx->d_dets /* the read-back block: lists, counts, counters */};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (ctx->d_in2) cudaFree(ctx->d_in2);
    for (auto &kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (ctx->d_sq_scratch) cudaFree(ctx->d_sq_scratch);
    if (ctx->d_band_dense) cudaFree(ctx->d_band_dense);
    for (void *p : {(void *)ctx->d_field_ids, (void *)ctx->d_field_poses, (void *)ctx->d_cam9, (void *)ctx->d_pose_buf}) if (p) cudaFree(p);
    if (ctx->h_dets) cudaFreeHost(ctx->h_dets);          // (h_counts, h_small live in the same block)
    if (ctx->h_chunk_err) cudaFreeHost(ctx->h_chunk_err);
    if (ctx->h_frame_err) cudaFreeHost(ctx->h_frame_err);
    if (ctx->h_pose_stage) cudaFreeHost(ctx->h_pose_stage);
    if (ctx->h_pose_ok_stage) cudaFreeHost(ctx->h_pose_ok_stage);
    if (ctx->h_pose_nt_stage) cudaFreeHost(ctx->h_pose_nt_stage);
    for (auto &sl : ctx->ss) {
        if (sl.h_dets) cudaFreeHost(sl.h_dets);
        if (sl.h_counts) cudaFreeHost(sl.h_counts);
        if (sl.h_err) cudaFreeHost(sl.h_err);
        if (sl.h_ferr) cudaFreeHost(sl.h_ferr);
        if (sl.start) cudaEventDestroy(sl.start);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.h_gyro) cudaFreeHost(sl.h_gyro);
        if (sl.h_poses) cudaFreeHost(sl.h_poses);
        if (sl.h_ok) cudaFreeHost(sl.h_ok);
        if (sl.h_ntags) cudaFreeHost(sl.h_ntags);
        if (sl.pose_done) cudaEventDestroy(sl.pose_done);
    }
    for (int i = 0; i < 2; i++) { if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]); if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pose_stream) cudaStreamDestroy(ctx->pose_stream);
    if (ctx->ev_pose_ready) cudaEventDestroy(ctx->ev_pose_ready);
    if (ctx->ev_pose_done) cudaEventDestroy(ctx->ev_pose_done);
    for (int i = 0; i < 4; i++) { if (ctx->tier_stream[i]) cudaStreamDestroy(ctx->tier_stream[i]); if (ctx->ev_tier[i]) cudaEventDestroy(ctx->ev_tier[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

cb_ctx *cb_create(int device, int max_width, int max_height, int max_batch, int max_dets_per_frame)
{
    cb_ctx *ctx = nullptr;
    if (max_width < 8 || max_height < 8 || max_batch < 1 || max_dets_per_frame < 1 || max_width >= 65536 || max_height >= 65536) {
        fail(nullptr, CB_ERR_ARG, "cb_create: bad sizes %dx%d batch %d dets %d", max_width, max_height, max_batch, max_dets_per_frame);
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fail(nullptr, CB_ERR_CUDA, "cb_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(nullptr, CB_ERR_ARG, "cb_create: device %d out of range (0..%d)", device, ndev - 1); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        fail(nullptr, CB_ERR_CUDA, "cb_create: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { fail(nullptr, CB_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    ctx = new cb_ctx();
    ctx->device = device; ctx->max_w = max_width; ctx->max_h = max_height; ctx->max_batch = max_batch; ctx->max_dets = max_dets_per_frame;
    ctx->num_sms = prop.multiProcessorCount;
    set_default_params(ctx);
    const size_t B = (size_t)max_batch;
    const size_t npix = (size_t)max_width * max_height;   // worst case: decimation 1
    ctx->max_npix = npix;
    if (B * npix >= 0xffffffffull) { fail(nullptr, CB_ERR_ARG, "cb_create: batch*pixels must stay below 2^32 (labels are 32-bit)"); delete ctx; return nullptr; }
    // capacities are per frame and scale with the decimated frame (default decimation 2 => npix/4), computed for the
    // worst case decimation 1 only when the caller asks for it through cb_set_params; allocate for decimation >= 2 by
    // default and grow lazily in ensure_capacity().
    auto alloc = [&](void **p, size_t bytes) -> bool {
        cudaError_t er = cudaMalloc(p, bytes);
        if (er != cudaSuccess) { fail(nullptr, CB_ERR_CUDA, "cb_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(er)); return false; }
        return true;
    };
    bool ok = true;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->pose_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_pose_ready, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ctx->ev_pose_done, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4; i++) ok = ok && cudaStreamCreateWithFlags(&ctx->tier_stream[i], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&ctx->ev_tier[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2; i++) ok = ok && cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&ctx->h_chunk_err, 4096 * sizeof(uint32_t)) == cudaSuccess;
    for (auto &ev : ctx->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    // decimated worst case handled by this allocation: ceil(W/2) x ceil(H/2); decimation 1 re-allocates on demand
    const size_t dw = (max_width + 1) / 2, dh = (max_height + 1) / 2;
    const size_t dpix = dw * dh;
    const size_t tp = (dw + 15) / 16 * 16;
    ctx->in_bytes = B * (npix + 16);
    ctx->map_bytes = B * dh * tp;
    ctx->tile_bytes = B * ((dw / 4) + 1) * ((dh / 4) + 1);
    ctx->label_bytes = B * dpix * sizeof(uint32_t);
    Caps &c = ctx->caps;
    c.slots_per_frame = next_pow2((uint32_t)std::max<size_t>(1024, dpix / 8));
    c.clusters_per_frame = (uint32_t)std::max<size_t>(2048, dpix / 128);
    c.tile_probes = CL_PROBES;
    if (const char *e = getenv("CB_TILE_PROBES")) c.tile_probes = (uint32_t)std::max(1, std::min(CL_CAP, atoi(e)));   // test hook
    c.points_per_frame = (uint32_t)((std::max<size_t>(65536, dpix + dpix / 2) + 7) / 8 * 8);   // noisy frames emit ~0.85 points / pixel
    c.quads_per_frame = (uint32_t)std::max<size_t>(2048, dpix / 32);   // pure-noise frames produce thousands of candidate quads
    c.dets_per_frame = (uint32_t)max_dets_per_frame;
    if (const char *e = getenv("CB_TEST_QUADS_PER_FRAME")) c.quads_per_frame = (uint32_t)std::max(1, atoi(e));      // test hook: reach the overflow path
    ok = ok && alloc((void **)&ctx->d_in, ctx->in_bytes + 64);
    ok = ok && alloc((void **)&ctx->d_thresh, ctx->map_bytes) && alloc((void **)&ctx->d_mark, ctx->map_bytes);
    ok = ok && alloc((void **)&ctx->d_tmin, ctx->tile_bytes) && alloc((void **)&ctx->d_tmax, ctx->tile_bytes);
    ok = ok && alloc((void **)&ctx->d_labels, ctx->label_bytes) && alloc((void **)&ctx->d_sizes, ctx->label_bytes);
    ok = ok && alloc((void **)&ctx->d_table, B * c.slots_per_frame * sizeof(ClusterSlot));
    {
        const size_t dw = (size_t)ctx->max_w / 2 + 1, dh = (size_t)ctx->max_h / 2 + 1;      // decimated size bound (quad_decimate = 2)
        const size_t tiles = ((dw + CL_TW - 1) / CL_TW) * ((dh + CL_TH - 1) / CL_TH);
        if (const char *e = getenv("CB_CLUSTERS")) ctx->cluster_mode = strcmp(e, "tiles") == 0 ? 1 : 0;      // A/B hook
        if (const char *e = getenv("CB_SORT")) ctx->sort_bucket_form = strcmp(e, "merge") == 0 ? 0 : 1;
        ok = ok && alloc((void **)&ctx->d_ent, B * dw * dh * sizeof(uint4));
        if (ctx->cluster_mode == 1) {
            ok = ok && alloc((void **)&ctx->d_tile_keys, B * tiles * CL_CAP * sizeof(unsigned long long));
            ok = ok && alloc((void **)&ctx->d_tile_cnt, B * tiles * CL_CAP * sizeof(uint32_t));
        } else {
            // bands: at most ~num_sms * 128 first areas when bands are short, B * ceil(h / 4) when they have 4 rows; as many
            // again for chained sub-bands (a band whose private table fills up)
            const size_t first = std::max<size_t>((size_t)ctx->num_sms * 128 + 2 * B, B * (dh / 4 + 2));
            ctx->areas_first_cap = (uint32_t)first; ctx->areas_pool_cap = (uint32_t)first;
            ok = ok && alloc((void **)&ctx->d_areas, 2 * first * sizeof(ClbArea));
            if ((size_t)c.clusters_per_frame * sizeof(uint32_t) > 160 * 1024) ok = ok && alloc((void **)&ctx->d_cursors, B * c.clusters_per_frame * sizeof(uint32_t));
        }
    }
    ok = ok && alloc((void **)&ctx->d_clusters, B * c.clusters_per_frame * sizeof(ClusterRec));
    ok = ok && alloc((void **)&ctx->d_worklist, 4 * B * c.clusters_per_frame * sizeof(uint32_t));
    ok = ok && alloc((void **)&ctx->d_scankey, B * c.points_per_frame * sizeof(uint32_t));
    ok = ok && alloc((void **)&ctx->d_errs, B * c.points_per_frame * sizeof(double));
    ok = ok && alloc((void **)&ctx->d_cp, B * c.points_per_frame / LF_CP * 6 * sizeof(double));
    ok = ok && alloc((void **)&ctx->d_scratch, B * c.points_per_frame * 2 * sizeof(unsigned long long));
    ok = ok && alloc((void **)&ctx->d_quads, B * c.quads_per_frame * sizeof(QuadRec));
    ok = ok && alloc((void **)&ctx->d_raw, B * c.quads_per_frame * sizeof(RawDet));
    {
        // what a call reads back -- detection lists, counts, counters and flags -- is ONE block on either side, so a small context
        // (the one-frame call) fetches it with one copy instead of three (queue_chunk)
        const size_t b_dets = (B * c.dets_per_frame * sizeof(cb_detection) + 255) / 256 * 256, b_counts = (B * sizeof(int32_t) + 255) / 256 * 256;
        ctx->out_block_bytes = b_dets + b_counts + (5 * B + 48) * sizeof(uint32_t);
        uint8_t *d_blk = nullptr, *h_blk = nullptr;
        ok = ok && alloc((void **)&d_blk, ctx->out_block_bytes);
        ok = ok && cudaMallocHost((void **)&h_blk, ctx->out_block_bytes) == cudaSuccess;
        if (ok) {
            ctx->d_dets = (cb_detection *)d_blk; ctx->d_counts = (int32_t *)(d_blk + b_dets); ctx->d_small = (uint32_t *)(d_blk + b_dets + b_counts);
            ctx->h_dets = (cb_detection *)h_blk; ctx->h_counts = (int32_t *)(h_blk + b_dets); ctx->h_small = (uint32_t *)(h_blk + b_dets + b_counts);
        } else {
            if (d_blk) cudaFree(d_blk);
            if (h_blk) cudaFreeHost(h_blk);
        }
    }
    if (ok) {
        ok = ok && cudaMemcpyToSymbol(c_codes, kHostCodes, sizeof(kHostCodes)) == cudaSuccess;
        ok = ok && cudaMemcpyToSymbol(c_bit_x, kHostBitX, sizeof(kHostBitX)) == cudaSuccess;
        ok = ok && cudaMemcpyToSymbol(c_bit_y, kHostBitY, sizeof(kHostBitY)) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(threshold_f2_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, thr_tma_warp_bytes(THR_TMA_ROWB) * THR_TMA_WARPS) == cudaSuccess;
#define CB_SORT_ATTR(CFG) ok = ok && cudaFuncSetAttribute(sort_clusters_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CFG::BYTES) == cudaSuccess;
        CB_SORT_ATTR(SortS8<1>) CB_SORT_ATTR(SortS16<1>) CB_SORT_ATTR(SortM<1>) CB_SORT_ATTR(SortL1<1>) CB_SORT_ATTR(SortL2<1>)
        CB_SORT_ATTR(SortS8<2>) CB_SORT_ATTR(SortS16<2>) CB_SORT_ATTR(SortM<2>) CB_SORT_ATTR(SortL1<2>) CB_SORT_ATTR(SortL2<2>)
        CB_SORT_ATTR(SortS8<3>) CB_SORT_ATTR(SortS16<3>) CB_SORT_ATTR(SortM<3>) CB_SORT_ATTR(SortL1<3>) CB_SORT_ATTR(SortL2<3>)
        // (function attributes are shared by every context of the device: always the same value, the most a context may ask for)
        ok = ok && cudaFuncSetAttribute(cluster_band_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(lfps_big_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LFB_SMEM) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(lfps_big_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, LFB_SMEM) == cudaSuccess;
#undef CB_SORT_ATTR
        {   // 4-subsets of {0..9} in colex order (subsets of {0..k-1} first), packed m0<<12|m1<<8|m2<<4|m3
            uint16_t combos[210];
            int nc = 0;
            for (int m3 = 3; m3 < 10; m3++)
                for (int m2 = 2; m2 < m3; m2++)
                    for (int m1 = 1; m1 < m2; m1++)
                        for (int m0 = 0; m0 < m1; m0++) combos[nc++] = (uint16_t)((m0 << 12) | (m1 << 8) | (m2 << 4) | m3);
            ok = ok && nc == 210 && cudaMemcpyToSymbol(c_combos, combos, sizeof(combos)) == cudaSuccess;
        }
        if (!ok && g_create_error.empty()) fail(nullptr, CB_ERR_CUDA, "cb_create: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!ok) {
        if (g_create_error.empty()) fail(nullptr, CB_ERR_CUDA, "cb_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cb_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

int cb_set_family_tag36h11(cb_ctx *ctx, int bits_corrected)
{
    if (!ctx) return CB_ERR_ARG;
    if (bits_corrected < 0 || bits_corrected > 3) return fail(ctx, CB_ERR_UNSUPPORTED, "bits_corrected %d: upstream's quick-decode table supports 0..3", bits_corrected);
    ctx->prm.bits_corrected = bits_corrected;
    ctx->family_set = true;
    drop_graphs(ctx);               // kernel parameters are baked into the captured graphs
    return CB_OK;
}

int cb_set_params(cb_ctx *ctx, float quad_decimate, float quad_sigma, int refine_edges, double decode_sharpening,
                  int min_cluster_pixels, int max_nmaxima, float critical_rad, float max_line_fit_mse, int min_white_black_diff)
{
    if (!ctx) return CB_ERR_ARG;
    if (quad_sigma != 0.0f) return fail(ctx, CB_ERR_UNSUPPORTED, "quad_sigma %g: blur is not implemented (the reference leaves it at 0)", quad_sigma);
    // upstream: factor 1.5 averages 3x3 blocks, every other factor f > 1 point-samples at (int)f; f <= 1 works on the frame itself
    if (!(quad_decimate >= 1.0f && quad_decimate <= 16.0f) || (float)(int)quad_decimate != quad_decimate)
        return fail(ctx, CB_ERR_UNSUPPORTED, "quad_decimate %g: integer factors 1..16 are implemented (2 is the reference's setting and the fused fast path; 1.5 is not offered)", quad_decimate);
    CB_NOT_STREAMING(ctx);
    if (max_nmaxima < 4 || max_nmaxima > 10) return fail(ctx, CB_ERR_UNSUPPORTED, "max_nmaxima %d outside 4..10", max_nmaxima);
    DetParams &p = ctx->prm;
    p.quad_decimate = quad_decimate; p.refine_edges = refine_edges; p.decode_sharpening = decode_sharpening;
    p.min_cluster_pixels = min_cluster_pixels; p.max_nmaxima = max_nmaxima; p.critical_rad = critical_rad;
    p.max_line_fit_mse = max_line_fit_mse; p.min_white_black_diff = min_white_black_diff;
    p.cos_critical_rad = cos(p.critical_rad);
    int mtw = (int)(8 / p.quad_decimate); if (mtw < 3) mtw = 3;
    p.min_tag_width = mtw;
    drop_graphs(ctx);
    return CB_OK;
}

int cb_decimated_size(const cb_ctx *ctx, int width, int height, int *w, int *h)
{
    if (!ctx || !w || !h) return CB_ERR_ARG;
    const int f = (int)ctx->prm.quad_decimate;
    *w = 1 + (width - 1) / f; *h = 1 + (height - 1) / f;
    return CB_OK;
}

int cb_frame_flags(const cb_ctx *ctx, uint32_t *flags, int n)
{
    if (!ctx || n < 0 || (n > 0 && !flags)) return CB_ERR_ARG;
    int bad = 0;
    for (int i = 0; i < n; i++) {
        flags[i] = i < (int)ctx->frame_flags.size() ? ctx->frame_flags[i] : 0u;
        bad += flags[i] != 0;
    }
    return bad;
}

int cb_get_timing(const cb_ctx *ctx, cb_timing *t)
{
    if (!ctx || !t) return CB_ERR_ARG;
    *t = ctx->timing;
    return CB_OK;
}

}  // extern "C"

static int make_geom(cb_ctx *ctx, int width, int height, int stride, size_t frame_stride, int batch, Geom &g)
{
    if (width < 8 || height < 8 || width > ctx->max_w || height > ctx->max_h)
        return fail(ctx, CB_ERR_ARG, "frame %dx%d outside the context's capacity %dx%d", width, height, ctx->max_w, ctx->max_h);
    if (stride < width || frame_stride < (size_t)stride * (height - 1) + width) return fail(ctx, CB_ERR_ARG, "bad stride");
    if (batch < 1 || batch > ctx->max_batch) return fail(ctx, CB_ERR_ARG, "batch %d outside 1..%d", batch, ctx->max_batch);
    g.W = width; g.H = height; g.stride = stride; g.frame_stride = frame_stride;
    g.f = (int)ctx->prm.quad_decimate;
    g.w = 1 + (width - 1) / g.f; g.h = 1 + (height - 1) / g.f;
    // the per-pixel buffers hold ceil(max_w / 2) x ceil(max_h / 2) decimated pixels (cb_create): quad_decimate = 1 needs a
    // context created with twice the frame size
    if (g.w > (ctx->max_w + 1) / 2 || g.h > (ctx->max_h + 1) / 2)
        return fail(ctx, CB_ERR_ARG, "quad_decimate %d: the detector's working image %dx%d exceeds the context's %dx%d (create the context with max_width / max_height of twice the frame size)",
                    g.f, g.w, g.h, (ctx->max_w + 1) / 2, (ctx->max_h + 1) / 2);
    g.tp = (g.w + 15) / 16 * 16;
    g.tw = g.w / 4; g.th = g.h / 4;
    g.batch = batch;
    g.npix = (uint32_t)g.w * g.h;
    return CB_OK;
}

// Runs the device pipeline on `d_frames` (device memory) up to `stage`.  Events: ev[1] start, ev[2] after threshold,
// ev[3] after ccl, ev[4] after clusters, ev[5] after quads, ev[6] after decode+reconcile.
static int sqpnp_launch(cb_ctx *ctx, const cb_iso3 *d_tags, const double *d_bear, const int32_t *d_nt, int max_tags, const cb_iso3 *d_r2c,
                        const double *d_gyro, double sign_change_error, int64_t n, cb_pose *d_out, uint8_t *d_ok, cudaStream_t stream = nullptr);
// layout of d_pose_buf for pose_cap frames
struct PoseBufs { cb_iso3 *tags; double *bearings; int32_t *n_tags; double *gyro; cb_pose *poses; uint8_t *ok; };
static PoseBufs pose_bufs(cb_ctx *ctx)
{
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t n = (size_t)ctx->pose_cap;
    uint8_t *p = ctx->d_pose_buf;
    PoseBufs b;
    b.tags = (cb_iso3 *)p; p += up(n * SQ_MAX_TAGS * sizeof(cb_iso3));
    b.bearings = (double *)p; p += up(n * SQ_MAX_TAGS * 12 * sizeof(double));
    b.n_tags = (int32_t *)p; p += up(n * sizeof(int32_t));
    b.gyro = (double *)p; p += up(n * sizeof(double));
    b.poses = (cb_pose *)p; p += up(n * sizeof(cb_pose));
    b.ok = p;
    return b;
}
static size_t pose_buf_bytes(size_t n)
{
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    return up(n * SQ_MAX_TAGS * sizeof(cb_iso3)) + up(n * SQ_MAX_TAGS * 12 * sizeof(double)) + up(n * sizeof(int32_t)) + up(n * sizeof(double)) +
           up(n * sizeof(cb_pose)) + up(n);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libcuda is not linked)
typedef CUresult (*cb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cb_encode_tiled_fn tensor_map_encoder()
{
    static cb_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (cb_encode_tiled_fn)p;
    }();
    return fn;
}

// Launch of the tensor-map threshold kernel with `ysegs` row segments per frame (<= 0: the wave-count model below).
// Returns false when the tensor map cannot be built (the caller then takes the 1-D TMA kernel).
template <class C>
static bool launch_threshold_tm(cb_ctx *ctx, const uint8_t *d_frames, const Geom &g, int min_diff, cudaStream_t st, int ysegs)
{
    constexpr int T = 2 * C::P, TM_WARPS = C::WARPS, TM_STAGES = C::STAGES;
    if (C::SMEM > 48 * 1024 && cudaFuncSetAttribute(threshold_tm_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) return false;
    cb_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    // 3-D tensor of 8-byte elements: x = tile column (8 input bytes = 4 decimated pixels), y = EVEN input row (row stride doubled),
    // z = frame.  Box = the 32 T tiles of a warp x the 4 rows of a tile row.
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)(g.stride / 8), (cuuint64_t)g.h, (cuuint64_t)g.batch};
    const cuuint64_t strides[2] = {(cuuint64_t)2 * g.stride, (cuuint64_t)g.frame_stride};
    const cuuint32_t box[3] = {32u * T, 4u, 1u}, estr[3] = {1u, 1u, 1u};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t *>(d_frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    TmPlan plan;
    plan.strips = (g.tw + C::MAX_IW - 1) / C::MAX_IW;
    plan.iw = ((g.tw + plan.strips - 1) / plan.strips + 3) / 4 * 4;      // multiple of 4 tiles: 16-byte aligned strip starts (bulk stores)
    if (ysegs <= 0) {   // equal waves: every warp does seg_rows + 2 tile rows plus the fill of its ring
        const int ctas_per_sm = std::max(1, std::min((int)C::MIN_CTAS, (int)((227 * 1024) / (C::SMEM + 1024))));
        const long long resident = (long long)ctx->num_sms * ctas_per_sm * TM_WARPS;
        const int max_segs = std::max(1, g.th / 6);
        long long best_cost = -1;
        ysegs = 1;
        for (int ys = 1; ys <= max_segs; ys++) {
            const int rows = (g.th + ys - 1) / ys, segs = (g.th + rows - 1) / rows;
            if (segs != ys) continue;
            const long long warps = (long long)plan.strips * segs * g.batch;
            const long long cost = ((warps + resident - 1) / resident) * (rows + 2 + TM_STAGES);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; ysegs = ys; }
        }
    }
    plan.seg_rows = (g.th + ysegs - 1) / ysegs;
    plan.ysegs = (g.th + plan.seg_rows - 1) / plan.seg_rows;
    const long long warps = (long long)plan.strips * plan.ysegs * g.batch;
    const int wt = (g.w % 4 || g.h % 4) ? 1 : 0;
    threshold_tm_kernel<C><<<(unsigned)((warps + TM_WARPS - 1) / TM_WARPS), TM_WARPS * 32, C::SMEM, st>>>(map, ctx->d_thresh, ctx->d_tmin, ctx->d_tmax, g,
                                                                                                     std::max(-1000, std::min(1000, min_diff)), plan, wt);
    return true;
}

// The shapes of the tensor-map kernel: tiles per lane x ring depth x CTAs per SM x store form.  One warp per CTA throughout (a CTA
// that ends frees its slot at once; 4-warp CTAs waited for their slowest warp).  Which one is fastest depends on the frame width
// (lanes busy with T = 6 or 4), on how the segments fill the SMs' slots and on the L2 state the stores meet, none of which a
// static model predicted in the micro-benchmark (spread 0.59 .. 0.71 of the HBM peak on 256 x 1280x720, tools/cuda/thr_bench.cu); the
// default is the shape that measured best inside the pipeline, timing the shapes on the batch is opt-in (threshold_plan).
typedef bool (*thr_launch_fn)(cb_ctx *, const uint8_t *, const Geom &, int, cudaStream_t, int);
static const thr_launch_fn kThrVariants[] = {
    launch_threshold_tm<TmCfg<6, 3, 1, 9, 1>>,      // 0: bulk stores, 3-deep ring
    launch_threshold_tm<TmCfg<6, 2, 1, 12, 1>>,     // 1: bulk stores, 2-deep ring, 12 warps per SM
    launch_threshold_tm<TmCfg<6, 3, 1, 12, 0>>,     // 2: direct row-major stores
    launch_threshold_tm<TmCfg<4, 3, 1, 12, 1>>,     // 3: T = 4, bulk stores
    launch_threshold_tm<TmCfg<4, 3, 1, 16, 0>>,     // 4: T = 4, direct stores
};
static const int kThrMaxIw[] = {TmCfg<6>::MAX_IW, TmCfg<6>::MAX_IW, TmCfg<6>::MAX_IW, TmCfg<4>::MAX_IW, TmCfg<4>::MAX_IW};
constexpr int kNumThrVariants = 5;

static ThrChoice threshold_plan(cb_ctx *ctx, const uint8_t *d_frames, const Geom &g, int min_diff, cudaStream_t st)
{
    const ThrKey key{g.W, g.H, g.stride, g.batch};
    auto it = ctx->thr_plans.find(key);
    if (it != ctx->thr_plans.end()) return it->second;
    // static choice: the tiles per lane that keep most lanes busy (one strip of T = 6 spans up to 188 tiles = 1504 input pixels)
    const int s6 = (g.tw + kThrMaxIw[0] - 1) / kThrMaxIw[0], s4 = (g.tw + kThrMaxIw[3] - 1) / kThrMaxIw[3];
    // T = 6 shapes: bulk stores with a 2-deep ring and 12 one-warp CTAs per SM measured best inside the pipeline (tools/run_thr_static.sh:
    // 0.80 of the HBM peak on 256 x 1280x720 with the wave model's segment count, 0.70 for the 3-deep ring, 0.74 for direct stores)
    ThrChoice best{(double)g.tw / (s6 * 32 * 6) >= (double)g.tw / (s4 * 32 * 4) ? 1 : 3, 0};
    if (const char *e = getenv("CB_THR_T")) best.variant = atoi(e) == 6 ? 0 : 3;                      // experiment hooks
    if (const char *e = getenv("CB_THR_CFG")) best.variant = std::max(0, std::min(kNumThrVariants - 1, atoi(e)));
    if (const char *e = getenv("CB_THR_YSEGS")) best.ysegs = std::max(1, std::min(std::max(1, g.th / 6), atoi(e)));
    const char *tune_env = getenv("CB_THR_TUNE");
    // Timing the shapes on the batch itself is opt-in (CB_THR_TUNE=1): stand-alone launches meet another L2 than the pipeline's launch
    // does (back to back they find their input resident, behind a flush they compete with its write-back), so the winner of the timing
    // was not reliably the fastest shape in the pipeline (0.60 .. 0.81 over repeated runs against 0.80 for the static choice).
    const bool tune = tune_env ? atoi(tune_env) != 0 : false;
    if (tune && (size_t)g.batch * g.W * g.H >= ((size_t)32 << 20)) {
        // time every shape x a ladder of segment heights on this batch (about 40 ms, once per geometry and context).  The first launches
        // after an idle period run at ramping clocks, so the GPU is warmed up first and every candidate is timed in two separate rounds
        // (the better of the two counts): a shape must not win or lose by where in the sequence it was measured.
        cudaEvent_t e0, e1;
        if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
            const int rows_ladder[] = {5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 18, 20, 23, 26, 30, 36, 45, 60, 90};
            struct Cand { int v, ys; float ms; };
            std::vector<Cand> cands;
            for (int v = 0; v < kNumThrVariants; v++) {
                int last_ys = -1;
                for (int rows : rows_ladder) {
                    if (rows > g.th) break;
                    const int ys = (g.th + rows - 1) / rows;
                    if (ys == last_ys || ys > std::max(1, g.th / 5)) continue;
                    last_ys = ys;
                    cands.push_back(Cand{v, ys, 1e30f});
                }
            }
            bool usable = true;
            for (int r = 0; r < 40 && usable; r++) usable = kThrVariants[best.variant](ctx, d_frames, g, min_diff, st, best.ysegs);      // warm-up
            for (int round = 0; round < 2 && usable; round++)
                for (Cand &c : cands) {
                    float sum = 0;
                    if (!kThrVariants[c.v](ctx, d_frames, g, min_diff, st, c.ys)) { usable = false; break; }
                    cudaEventRecord(e0, st);
                    for (int r = 0; r < 3; r++) kThrVariants[c.v](ctx, d_frames, g, min_diff, st, c.ys);
                    cudaEventRecord(e1, st);
                    if (cudaEventSynchronize(e1) != cudaSuccess) { usable = false; break; }
                    cudaEventElapsedTime(&sum, e0, e1);
                    c.ms = std::min(c.ms, sum);
                }
            if (usable) {
                float best_ms = 1e30f;
                for (const Cand &c : cands) if (c.ms < best_ms) { best_ms = c.ms; best = ThrChoice{c.v, c.ys}; }
            }
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
    }
    if (getenv("CB_THR_VERBOSE")) fprintf(stderr, "chalkydri_b200: threshold plan %dx%d x %d: CB_THR_CFG=%d CB_THR_YSEGS=%d\n", g.W, g.H, g.batch, best.variant, best.ysegs);
    ctx->thr_plans[key] = best;
    return best;
}

// packed RGB -> gray for `n` frames: the persistent ring kernel (8 CTAs of 4 warps per SM, measured best of 2..16: 0.92 of the HBM
// copy peak on 256 x 1456x1088 against 0.71 for the CTA-refill kernel, tools/cuda/rgb_bench.cu) when every chunk start is 16-byte
// aligned, else the round-1 kernel, which checks alignment per frame and converts ragged frames lane by lane.
static void launch_rgb_to_gray(cb_ctx *ctx, const uint8_t *d_rgb, uint8_t *d_gray, size_t npix, size_t in_stride, size_t out_stride, int n, dim3 grid_refill)
{
    const bool aligned = ((uintptr_t)d_rgb % 16 == 0) && ((uintptr_t)d_gray % 16 == 0) && in_stride % 16 == 0 && out_stride % 16 == 0 && npix >= 512;
    static const bool force_refill = getenv("CB_RGB") && strcmp(getenv("CB_RGB"), "refill") == 0;      // A/B hook
    if (aligned && !force_refill)
        rgb_to_gray_ring_kernel<<<ctx->num_sms * 8, RGB_RING_WARPS * 32, 0, ctx->stream>>>(d_rgb, d_gray, npix, in_stride, out_stride, n);
    else
        rgb_to_gray_kernel<<<grid_refill, RGB_THREADS, 0, ctx->stream>>>(d_rgb, d_gray, npix, in_stride, out_stride);
}

static int run_pipeline(cb_ctx *ctx, const uint8_t *d_frames, const Geom &g, int stage)
{
    cudaStream_t st = ctx->stream;
    const Caps &caps = ctx->caps;
    const DetParams &prm = ctx->prm;
    const int B = g.batch;
    int launches = 0, thr_launches = 0;
    const size_t wl_stride = (size_t)ctx->max_batch * caps.clusters_per_frame;       // four tier work lists, back to back
    uint32_t *d_ncl = ctx->d_small, *d_npt = ctx->d_small + ctx->max_batch, *d_nq = ctx->d_small + 2 * ctx->max_batch,
             *d_nraw = ctx->d_small + 3 * ctx->max_batch, *d_misc = ctx->d_small + 4 * ctx->max_batch;
    // misc: [0] errflag, [1] nwork, [2] work_counter, [3] nquads_total, [4] decode counter; behind misc: one overflow word per frame
    CK(cudaMemsetAsync(ctx->d_small, 0, (5 * (size_t)ctx->max_batch + 48) * sizeof(uint32_t), st));
    CK(cudaEventRecord(ctx->ev[1], st));
    // ---- A1+A2 threshold ----
    const bool fast = g.f == 2 && (g.stride % 16 == 0) && (g.frame_stride % 16 == 0) && ((uintptr_t)d_frames % 16 == 0) && g.tw > 0 && g.th > 0;
    if (ctx->external_map) {
        // the ternary map is already in d_thresh
    } else if (fast) {
        // variant switch for A/B profiling: CB_THRESHOLD = tmap (default: tensor-map TMA, 6 or 4 tiles per lane) | tma (round 1: four
        // 1-D bulk copies per tile row) | tiled (first version, CTA tiles with block barriers)
        static const char *variant_env = getenv("CB_THRESHOLD");
        static const int variant = variant_env == nullptr ? 0 : (strcmp(variant_env, "tiled") == 0 ? 2 : (strcmp(variant_env, "tma") == 0 ? 1 : 0));
        bool done = false;
        if (variant == 0) {
            const ThrChoice ch = threshold_plan(ctx, d_frames, g, prm.min_white_black_diff, st);
            done = kThrVariants[ch.variant](ctx, d_frames, g, prm.min_white_black_diff, st, ch.ysegs);
        }
        if (done) {
        } else if (variant == 2) {
            dim3 grid((g.tw + THR_IW - 1) / THR_IW, (g.th + THR_IH - 1) / THR_IH, B), block(THR_TX, THR_TY);
            threshold_f2_kernel<<<grid, block, 0, st>>>(d_frames, ctx->d_thresh, ctx->d_tmin, ctx->d_tmax, g, prm.min_white_black_diff);
        } else {
            RollPlan plan;
            plan.strips = (g.tw + THR_ROLL_MAXIW - 1) / THR_ROLL_MAXIW;
            plan.iw = ((g.tw + plan.strips - 1) / plan.strips + 3) / 4 * 4;
            // Every warp does the same work (seg_rows + 2 halo tile rows, plus the fill of its TMA ring), and an SM holds 3 CTAs
            // of 4 warps (66 KB of ring each), so the launch runs in ceil(warps / resident) equal waves: pick the segment
            // count that minimises waves x (rows per warp).  (A fixed "24 warps per SM" target gave 2.02 waves -> 3 rounds.)
            {
                const int nlanes_max = (plan.iw + 3 + 3) / 4;                                 // lanes that hold data: iw inner tiles + 3 halo, 4 tiles per lane
                plan.rowb = std::min(THR_TMA_ROWB, (nlanes_max * 32 + 127) / 128 * 128);
                plan.warp_bytes = thr_tma_warp_bytes(plan.rowb);
                const int ctas_per_sm = std::max(1, std::min(5, (int)((227 * 1024) / (plan.warp_bytes * THR_TMA_WARPS + 1024))));   // 5: register limit (87 regs)
                const long long resident = (long long)ctx->num_sms * ctas_per_sm * THR_TMA_WARPS;
                const int max_segs = std::max(1, g.th / 8);
                long long best_cost = -1;
                int best = 1;
                for (int ys = 1; ys <= max_segs; ys++) {
                    const int rows = (g.th + ys - 1) / ys, segs = (g.th + rows - 1) / rows;
                    if (segs != ys) continue;
                    const long long warps = (long long)plan.strips * segs * B;
                    const long long cost = ((warps + resident - 1) / resident) * (rows + 2 + THR_TMA_STAGES);
                    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ys; }
                }
                plan.seg_rows = (g.th + best - 1) / best;
                plan.ysegs = (g.th + plan.seg_rows - 1) / plan.seg_rows;
            }
            const long long warps = (long long)plan.strips * plan.ysegs * B;
            const int wt = (g.w % 4 || g.h % 4) ? 1 : 0;
            threshold_f2_tma_kernel<<<(unsigned)((warps + THR_TMA_WARPS - 1) / THR_TMA_WARPS), THR_TMA_WARPS * 32, (size_t)plan.warp_bytes * THR_TMA_WARPS, st>>>(
                    d_frames, ctx->d_thresh, ctx->d_tmin, ctx->d_tmax, g, prm.min_white_black_diff, plan, wt);
        }
        launches++; thr_launches++;
        if (g.w % 4 || g.h % 4) {
            dim3 gr((g.w + 127) / 128, g.h, B);
            threshold_generic_kernel<<<gr, 128, 0, st>>>(d_frames, ctx->d_tmin, ctx->d_tmax, ctx->d_thresh, g, prm.min_white_black_diff, 1);
            launches++;
        }
    } else {
        if (g.tw > 0 && g.th > 0) {
            dim3 gr((g.tw * g.th + 127) / 128, B);
            tile_minmax_generic_kernel<<<gr, 128, 0, st>>>(d_frames, ctx->d_tmin, ctx->d_tmax, g);
            launches++;
        }
        dim3 gr((g.w + 127) / 128, g.h, B);
        threshold_generic_kernel<<<gr, 128, 0, st>>>(d_frames, ctx->d_tmin, ctx->d_tmax, ctx->d_thresh, g, prm.min_white_black_diff, 0);
        launches++; thr_launches++;
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    if (stage >= ST_LABELS) {
        // ---- A3 connected components ----
        dim3 grid((g.w + CCL_TW - 1) / CCL_TW, (g.h + CCL_TH - 1) / CCL_TH, B);
        // local roots go to a list (in the staging area of the cluster passes, free until they start) instead of a dense sizes plane
        static const bool dense_sizes = getenv("CB_CCL") && strcmp(getenv("CB_CCL"), "dense") == 0;      // A/B hook: round 1's form
        uint32_t *d_roots = dense_sizes ? nullptr : reinterpret_cast<uint32_t *>(ctx->d_ent);
        ccl_local_kernel<0><<<grid, CCL_THREADS, 0, st>>>(ctx->d_thresh, ctx->d_labels, ctx->d_sizes, g, d_roots, d_misc + 7);
        launches++;
        const int nrows = (g.h - 1) / CCL_TH, ncols = (g.w - 1) / CCL_TW + 1;   // borders: rows k*TH (k>=1); columns k*TW and k*TW-1
        if (nrows > 0) {
            dim3 gm((g.w + 127) / 128, nrows, B);
            ccl_merge_kernel<0><<<gm, 128, 0, st>>>(ctx->d_thresh, ctx->d_labels, g, 0);
            launches++;
        }
        {
            dim3 gm((g.h + 127) / 128, ncols, B);
            ccl_merge_kernel<0><<<gm, 128, 0, st>>>(ctx->d_thresh, ctx->d_labels, g, 1);
            launches++;
        }
        const uint32_t total = (uint32_t)B * g.npix;
        if (d_roots) ccl_flatten_list_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(ctx->d_labels, ctx->d_sizes, d_roots, d_misc + 7);
        else ccl_flatten_roots_kernel<<<(total + 256 * ROOTS_PER - 1) / (256 * ROOTS_PER), 256, 0, st>>>(ctx->d_labels, ctx->d_sizes, total);
        dim3 gfin((g.w + 256 * MARK_PER - 1) / (256 * MARK_PER), g.h, B);
        ccl_finish_kernel<<<gfin, 256, 0, st>>>(ctx->d_thresh, ctx->d_labels, ctx->d_sizes, ctx->d_mark, g);
        launches += 2;
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    if (stage >= ST_QUADS) {
        // ---- A4 gradient clusters ----
        const size_t nslots = (size_t)B * caps.slots_per_frame;
        table_init_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(ctx->d_table, nslots);
        launches += 1;
        if (g.h > 2 && g.w > 2 && ctx->cluster_mode == 0) {
            // band-ordered passes (clusters.cuh): rows per band so that the launch has ~128 warps per SM to balance
            BandPlan bp;
            const long long target = (long long)ctx->num_sms * 128;
            bp.rows = (int)std::max<long long>(1, std::min<long long>(4, ((long long)B * (g.h - 2) + target - 1) / target));      // <= target + B bands when rows < 4
            bp.nbands = (g.h - 2 + bp.rows - 1) / bp.rows;
            if ((uint32_t)B * bp.nbands > ctx->areas_first_cap) return fail(ctx, CB_ERR_ARG, "internal: band tables too small (%d bands)", B * bp.nbands);
            bp.pool_cap = ctx->areas_pool_cap;
            const uint32_t nfirst = (uint32_t)B * bp.nbands, nall = nfirst + bp.pool_cap;
            uint32_t *d_pool = d_misc + 6;
            uint32_t *d_stage = reinterpret_cast<uint32_t *>(ctx->d_ent);
            cluster_band_count_kernel<<<nfirst, 32, 0, st>>>(ctx->d_mark, ctx->d_labels, ctx->d_table, d_misc, d_stage,
                                                                                                   ctx->d_areas, d_pool, g, caps, bp);
            cluster_select_kernel<<<dim3((caps.slots_per_frame + 255) / 256, B), 256, 0, st>>>(ctx->d_table, ctx->d_clusters, d_ncl, d_npt, ctx->d_worklist, wl_stride,
                                                                                            d_misc + 8, 2, (uint32_t)QT0, (uint32_t)QT1,
                                                                                            (uint32_t)QT2, d_misc, g, caps, prm.min_cluster_pixels);
            // dense form of the prefix for small batches (clusters.cuh): (band x cluster) matrix within a fixed budget
            uint32_t *d_dense = nullptr;
            {
                const size_t cells = (size_t)nfirst * caps.clusters_per_frame;
                static const bool dense_off = getenv("CB_BAND_PREFIX") && strcmp(getenv("CB_BAND_PREFIX"), "walk") == 0;      // A/B hook
                if (!dense_off && cells * sizeof(uint32_t) <= ((size_t)48 << 20)) {
                    if (ctx->band_dense_cells < cells) {
                        if (ctx->d_band_dense) {
                            CK(cudaStreamSynchronize(st));
                            cudaFree(ctx->d_band_dense); ctx->d_band_dense = nullptr; ctx->band_dense_cells = 0;
                            drop_graphs(ctx);       // graphs captured for other geometries hold the old address
                        }
                        CK(cudaMalloc((void **)&ctx->d_band_dense, cells * sizeof(uint32_t)));
                        ctx->band_dense_cells = cells;
                    }
                    d_dense = ctx->d_band_dense;
                    CK(cudaMemsetAsync(d_dense, 0, cells * sizeof(uint32_t), st));
                }
            }
            cluster_band_resolve_kernel<<<(nall + CLB_WARPS - 1) / CLB_WARPS, CLB_WARPS * 32, 0, st>>>(ctx->d_table, ctx->d_areas, d_pool, g, caps, bp, d_dense);
            if (d_dense) {
                cluster_band_prefix_dense_kernel<<<dim3((caps.clusters_per_frame + 31) / 32, B), 128, 0, st>>>(d_dense, ctx->d_clusters, d_ncl, d_pool, caps, bp);
                launches++;
            }
            cluster_band_prefix_kernel<<<B, CLB_CAP, ctx->d_cursors ? 0 : caps.clusters_per_frame * sizeof(uint32_t), st>>>(ctx->d_areas, ctx->d_clusters, d_ncl,
                                                                                                                      ctx->d_cursors, caps, bp, d_dense ? d_pool : nullptr);
            cluster_band_scatter_kernel<<<(nall + CLB_SC_WARPS - 1) / CLB_SC_WARPS, CLB_SC_WARPS * 32, 0, st>>>(d_stage, ctx->d_areas, d_pool, ctx->d_scankey, g, caps, bp, d_dense);
            launches += 5;
        }
        if (g.h > 2 && g.w > 2 && ctx->cluster_mode == 1) {
            dim3 gc((g.w + CL_TW - 1) / CL_TW, (g.h - 2 + CL_TH - 1) / CL_TH, B);
            cluster_count_kernel<<<gc, CL_THREADS, 0, st>>>(ctx->d_mark, ctx->d_labels, ctx->d_table, d_misc, ctx->d_ent, ctx->d_tile_keys, ctx->d_tile_cnt, g, caps);
            // misc: [0] error flags, [3] quads total, [4] decode counter, [8 + 2t] items of tier t, [9 + 2t] its work counter
            cluster_select_kernel<<<dim3((caps.slots_per_frame + 255) / 256, B), 256, 0, st>>>(ctx->d_table, ctx->d_clusters, d_ncl, d_npt, ctx->d_worklist, wl_stride,
                                                                                            d_misc + 8, 2, (uint32_t)QT0, (uint32_t)QT1,
                                                                                            (uint32_t)QT2, d_misc, g, caps, prm.min_cluster_pixels);
            cluster_scatter_kernel<<<gc, CL_THREADS, 0, st>>>(ctx->d_mark, ctx->d_labels, ctx->d_table, ctx->d_clusters, ctx->d_scankey, ctx->d_ent, ctx->d_tile_keys,
                                                               ctx->d_tile_cnt, g, caps);
            launches += 3;
        }
        CK(cudaEventRecord(ctx->ev[4], st));
        if (stage == ST_CLUSTERS) {      // cb_clusters: the selected clusters with their points in scan order, before any sort touches them
            CK(cudaEventRecord(ctx->ev[5], st));
            CK(cudaEventRecord(ctx->ev[6], st));
            CK(cudaGetLastError());
            ctx->timing.kernel_launches = launches;
            ctx->timing.threshold_launches = thr_launches;
            return CB_OK;
        }
        // ---- A5 quad fitting ----
        // sort #1 | sort #2 | prefix moments | fit, largest tier first inside each (long jobs).
        // misc: [8 + 2t] items of tier t; work counters: [16..20] sort #1, [21..25] sort #2, [26] moments, [27] fit
        // Each tier's sort #2 only depends on the same tier's sort #1, so the five tiers run as five independent chains (the
        // main stream plus four side streams): a tier with few, large CTAs no longer leaves the rest of the SMs idle.
#define CB_LAUNCH_SORT1(CFG, GRID, CNT, STREAM)                                                                                                  \
        sort_clusters_kernel<CFG><<<ctx->num_sms * (GRID), CFG::THREADS, CFG::BYTES, STREAM>>>(ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride, d_misc + 8, \
                                                                                             d_misc + (CNT), ctx->d_scratch, g, caps, prm, ctx->sort_bucket_form);
        CK(cudaEventRecord(ctx->ev_fork, st));
        for (int i = 0; i < 4; i++) CK(cudaStreamWaitEvent(ctx->tier_stream[i], ctx->ev_fork, 0));
        if (ctx->cluster_mode == 0) {
            // the points arrive in scan order: one sort (by slope) per tier, with sort #1's box / polarity tests in front
            CB_LAUNCH_SORT1(SortL2<3>, 1, 25, st)
            CB_LAUNCH_SORT1(SortL1<3>, 3, 24, ctx->tier_stream[0])
            CB_LAUNCH_SORT1(SortM<3>, 6, 23, ctx->tier_stream[1])
            CB_LAUNCH_SORT1(SortS16<3>, 3, 22, ctx->tier_stream[2])
            CB_LAUNCH_SORT1(SortS8<3>, 4, 21, ctx->tier_stream[3])
            launches -= 5;
        } else {
        CB_LAUNCH_SORT1(SortL2<1>, 1, 20, st)
        CB_LAUNCH_SORT1(SortL2<2>, 1, 25, st)
        CB_LAUNCH_SORT1(SortL1<1>, 3, 19, ctx->tier_stream[0])
        CB_LAUNCH_SORT1(SortL1<2>, 3, 24, ctx->tier_stream[0])
        CB_LAUNCH_SORT1(SortM<1>, 6, 18, ctx->tier_stream[1])
        CB_LAUNCH_SORT1(SortM<2>, 6, 23, ctx->tier_stream[1])
        CB_LAUNCH_SORT1(SortS16<1>, 3, 17, ctx->tier_stream[2])
        CB_LAUNCH_SORT1(SortS16<2>, 3, 22, ctx->tier_stream[2])
        CB_LAUNCH_SORT1(SortS8<1>, 4, 16, ctx->tier_stream[3])
        CB_LAUNCH_SORT1(SortS8<2>, 4, 21, ctx->tier_stream[3])
        }
#undef CB_LAUNCH_SORT1
        for (int i = 0; i < 4; i++) {
            CK(cudaEventRecord(ctx->ev_tier[i], ctx->tier_stream[i]));
            CK(cudaStreamWaitEvent(st, ctx->ev_tier[i], 0));
        }
        // small batches: the largest clusters get a CTA each (lfps_big_kernel, a side stream) instead of a warp -- one frame's latency is
        // its largest cluster's chain; large batches keep one warp per cluster (same work, more clusters in flight)
        int n_big = 0x7fffffff;
        if (B <= 8) n_big = 768;
        if (const char *e = getenv("CB_LFPS_BIG")) n_big = atoi(e) > 0 ? std::max(600, atoi(e)) : 0x7fffffff;      // A/B hook (0 = off)
        if (n_big != 0x7fffffff) {
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->tier_stream[0], ctx->ev_fork, 0));
            if (n_big > QT2) lfps_big_kernel<3><<<ctx->num_sms * 2, LFB_THREADS, LFB_SMEM, ctx->tier_stream[0]>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride,
                                                                                                      d_misc + 8, d_misc + 28, ctx->d_errs, ctx->d_cp, g, caps, n_big);
            else lfps_big_kernel<2><<<ctx->num_sms * 2, LFB_THREADS, LFB_SMEM, ctx->tier_stream[0]>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride,
                                                                                                  d_misc + 8, d_misc + 28, ctx->d_errs, ctx->d_cp, g, caps, n_big);
            CK(cudaEventRecord(ctx->ev_tier[0], ctx->tier_stream[0]));
            launches++;
        }
        lfps_kernel<<<ctx->num_sms * 8, LF_WARPS * 32, 0, st>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride, d_misc + 8, d_misc + 26,
                                                                  ctx->d_errs, ctx->d_cp, g, caps, n_big);
        if (n_big != 0x7fffffff) CK(cudaStreamWaitEvent(st, ctx->ev_tier[0], 0));
        if (n_big != 0x7fffffff) {      // the same split for the corner search: the large clusters' scans on eight warps each, side stream
            CK(cudaEventRecord(ctx->ev_fork, st));
            CK(cudaStreamWaitEvent(ctx->tier_stream[0], ctx->ev_fork, 0));
            if (n_big > QT2) fit_quads_big_kernel<3><<<ctx->num_sms * 2, FQB_WARPS * 32, sizeof(FqBig), ctx->tier_stream[0]>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride, d_misc + 8, d_misc + 29,
                        ctx->d_errs, ctx->d_cp, ctx->d_quads, d_nq, d_misc + 3, d_misc, g, caps, prm, n_big);
            else fit_quads_big_kernel<2><<<ctx->num_sms * 2, FQB_WARPS * 32, sizeof(FqBig), ctx->tier_stream[0]>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride, d_misc + 8, d_misc + 29,
                        ctx->d_errs, ctx->d_cp, ctx->d_quads, d_nq, d_misc + 3, d_misc, g, caps, prm, n_big);
            CK(cudaEventRecord(ctx->ev_tier[0], ctx->tier_stream[0]));
            launches++;
        }
        fit_quads_kernel<<<ctx->num_sms * 8, FQ_WARPS * 32, 0, st>>>(d_frames, ctx->d_scankey, ctx->d_clusters, ctx->d_worklist, wl_stride, d_misc + 8, d_misc + 27,
                                                                       ctx->d_errs, ctx->d_cp, ctx->d_quads, d_nq, d_misc + 3, d_misc, g, caps, prm, n_big);
        if (n_big != 0x7fffffff) CK(cudaStreamWaitEvent(st, ctx->ev_tier[0], 0));
        launches += 12;
        CK(cudaEventRecord(ctx->ev[5], st));
    } else {
        CK(cudaEventRecord(ctx->ev[4], st));
        CK(cudaEventRecord(ctx->ev[5], st));
    }
    if (stage >= ST_FULL) {
        // ---- A6-A9 refine, homography, decode, reconcile ----
        decode_quads_kernel<<<ctx->num_sms * 6, DEC_WARPS * 32, 0, st>>>(d_frames, ctx->d_quads, d_misc + 3, d_misc + 4, ctx->d_raw, d_nraw, g, caps, prm, ctx->dc);
        reconcile_kernel<<<(B + REC_WARPS - 1) / REC_WARPS, REC_WARPS * 32, 0, st>>>(ctx->d_raw, d_nraw, ctx->d_dets, ctx->d_counts, caps, B, d_misc + 48);
        launches += 2;
        if (ctx->pose_active) {
            PoseBufs pb = pose_bufs(ctx);
            assemble_pose_problems_kernel<<<(B + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, st>>>(ctx->d_dets, ctx->d_counts, (int)caps.dets_per_frame, ctx->d_field_ids, ctx->d_field_poses,
                                                                         ctx->n_field, ctx->d_cam9, pb.gyro, SQ_MAX_TAGS, pb.tags, pb.bearings, pb.n_tags,
                                                                         ctx->pose_frame_base, B);
            // a few hundred problems are one latency-bound wave of warps (~2 ms): solved on a side stream they disappear
            // behind the next chunk's detection kernels
            CK(cudaEventRecord(ctx->ev_pose_ready, st));
            CK(cudaStreamWaitEvent(ctx->pose_stream, ctx->ev_pose_ready, 0));
            const size_t o = (size_t)ctx->pose_frame_base;
            int prc = sqpnp_launch(ctx, pb.tags + o * SQ_MAX_TAGS, pb.bearings + o * SQ_MAX_TAGS * 12, pb.n_tags + o, SQ_MAX_TAGS,
                                   (const cb_iso3 *)(ctx->d_cam9 + 9), pb.gyro + o, ctx->pose_sign_change_error, B, pb.poses + o, pb.ok + o,
                                   ctx->pose_stream);
            if (prc) return prc;
            launches += 2;
        }
    }
    CK(cudaEventRecord(ctx->ev[6], st));
    CK(cudaGetLastError());
    ctx->timing.kernel_launches = launches;
    ctx->timing.threshold_launches = thr_launches;
    return CB_OK;
}

static int finish_timing(cb_ctx *ctx, bool h2d, bool d2h)
{
    cb_timing &t = ctx->timing;
    auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return ms; };
    t.h2d_ms = h2d ? el(0, 1) : 0.f;
    t.preprocess_ms = 0.f;
    t.threshold_ms = el(1, 2); t.ccl_ms = el(2, 3); t.cluster_ms = el(3, 4); t.quad_ms = el(4, 5); t.decode_ms = el(5, 6);
    t.d2h_ms = d2h ? el(6, 7) : 0.f;
    t.total_ms = el(h2d ? 0 : 1, d2h ? 7 : 6);
    return CB_OK;
}

// A frame that overflowed one of its fixed-size device tables (cluster hash, clusters, points, quads) reports an empty list and a flag
// word (cb_frame_flags); the other frames of the batch are complete, so the call itself succeeds -- upstream has no such tables, and
// one cluttered camera frame must not fail a batch.  The text stays available through cb_last_error.
static void note_overflow(cb_ctx *ctx, uint32_t flag, int frames)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "%d frame(s) overflowed a device table (flags 0x%x:%s%s%s%s) and report no detections; cb_frame_flags tells which", frames, flag,
             (flag & ERR_HASH_FULL) ? " cluster-hash" : "", (flag & ERR_CLUSTERS_FULL) ? " clusters" : "", (flag & ERR_POINTS_FULL) ? " points" : "",
             (flag & ERR_QUADS_FULL) ? " quads" : "");
    ctx->err = buf;
}
// (simple path and stage taps: everything is in h_small)
static int check_errflag(cb_ctx *ctx, int batch = 0)
{
    const uint32_t flag = ctx->h_small[4 * (size_t)ctx->max_batch];
    ctx->frame_flags.assign((size_t)batch, 0u);
    int bad = 0;
    for (int b = 0; b < batch; b++) { ctx->frame_flags[b] = ctx->h_small[4 * (size_t)ctx->max_batch + 48 + b]; bad += ctx->frame_flags[b] != 0; }
    if (flag) note_overflow(ctx, flag, bad);
    return CB_OK;
}

// The kernels of a chunk and the read-back of its lists, queued on ctx->stream
static int queue_chunk(cb_ctx *ctx, const uint8_t *d_frames, const Geom &g)
{
    int rc = run_pipeline(ctx, d_frames, g, ST_FULL);
    if (rc) return rc;
    const size_t B = g.batch;
    if (ctx->out_block_bytes <= ((size_t)64 << 10)) {
        // a small context: the whole read-back block in one copy (a copy costs ~6 us whatever its size up to here)
        CK(cudaMemcpyAsync(ctx->h_dets, ctx->d_dets, ctx->out_block_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CK(cudaMemcpyAsync(ctx->h_dets, ctx->d_dets, B * ctx->caps.dets_per_frame * sizeof(cb_detection), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, B * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_small, ctx->d_small, (5 * (size_t)ctx->max_batch + 48) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ctx->pose_active && ctx->pose_join) {
        // the chunk's solve (queued on pose_stream by run_pipeline) ran beside the read-back of the lists; join it and fetch the poses
        PoseBufs pb = pose_bufs(ctx);
        const size_t o = (size_t)ctx->pose_frame_base;
        CK(cudaEventRecord(ctx->ev_pose_done, ctx->pose_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_pose_done, 0));
        CK(cudaMemcpyAsync(ctx->h_pose_stage, pb.poses + o, B * sizeof(cb_pose), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pose_ok_stage, pb.ok + o, B, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pose_nt_stage, pb.n_tags + o, B * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return CB_OK;
}

// full detector on device frames; copies detections back to the caller's arrays.
// Small batches (the reference's one-frame-per-call shape, crates/apriltags/src/lib.rs:293-301) are launch bound on the host side: two
// dozen kernels on five streams, memsets, event records.  From the second call with the same geometry and frame buffer on, all of it
// is ONE cudaGraphLaunch (captured with the side streams' forks and joins; kernel parameters are baked in, so the graphs are dropped
// when a parameter changes).  The stage events live inside the graph then, so cb_get_timing reports the total only.
static int detect_device_chunk(cb_ctx *ctx, const uint8_t *d_frames, const Geom &g, cb_detection *out, int32_t *out_counts, bool timed_h2d)
{
    if (!ctx->family_set) return fail(ctx, CB_ERR_STATE, "no tag family set: call cb_set_family_tag36h11 first");
    static const bool graphs_env_off = getenv("CB_GRAPH") && atoi(getenv("CB_GRAPH")) == 0;       // A/B hook
    const bool with_poses = ctx->pose_active && ctx->pose_join;
    const bool eligible = g.batch <= 4 && (!ctx->pose_active || with_poses) && !ctx->external_map && !ctx->graph_off && !graphs_env_off;
    bool used_graph = false;
    const auto gkey = std::make_tuple(g.W, g.H, g.stride, g.batch, (const void *)d_frames, with_poses ? 1 + ctx->pose_frame_base : 0,
                                      with_poses ? ctx->pose_sign_change_error : 0.0);
    if (eligible && (ctx->graphs.size() < 32 || ctx->graphs.count(gkey))) {      // (a caller cycling through many device buffers: plain launches)
        cb_ctx::GraphSlot &slot = ctx->graphs[gkey];
        if (!slot.exec && !slot.failed && slot.seen >= 1) {
            // second use: capture (the first, plain run did every lazy allocation and timed nothing we would miss)
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = queue_chunk(ctx, d_frames, g);
                const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc == CB_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&slot.exec, graph, 0) == cudaSuccess) {
                    slot.launches = ctx->timing.kernel_launches; slot.thr_launches = ctx->timing.threshold_launches;
                } else {
                    slot.exec = nullptr; slot.failed = true;
                    cudaGetLastError();                        // a refused capture must not poison the plain path
                }
                if (graph) cudaGraphDestroy(graph);
            } else { slot.failed = true; cudaGetLastError(); }
        }
        slot.seen++;
        if (slot.exec) {
            if (!timed_h2d) CK(cudaEventRecord(ctx->ev[0], ctx->stream));
            CK(cudaGraphLaunch(slot.exec, ctx->stream));
            ctx->timing.kernel_launches = slot.launches; ctx->timing.threshold_launches = slot.thr_launches;
            used_graph = true;
        }
    }
    if (!used_graph) {
        int rc = queue_chunk(ctx, d_frames, g);
        if (rc) return rc;
    }
    const size_t B = g.batch;
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (used_graph) {
        const int kl = ctx->timing.kernel_launches, tl = ctx->timing.threshold_launches;
        ctx->timing = cb_timing{};
        cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[7]);
        ctx->timing.kernel_launches = kl; ctx->timing.threshold_launches = tl;
    } else {
        finish_timing(ctx, timed_h2d, true);
    }
    int rc = check_errflag(ctx, (int)B);
    if (rc) return rc;
    memcpy(out_counts, ctx->h_counts, B * sizeof(int32_t));
    for (size_t b = 0; b < B; b++)
        memcpy(out + b * ctx->caps.dets_per_frame, ctx->h_dets + b * ctx->caps.dets_per_frame, (size_t)ctx->h_counts[b] * sizeof(cb_detection));
    if (with_poses) {
        const size_t o = (size_t)ctx->pose_frame_base;
        memcpy(ctx->pose_dst + o, ctx->h_pose_stage, B * sizeof(cb_pose));
        memcpy(ctx->pose_ok_dst + o, ctx->h_pose_ok_stage, B);
        if (ctx->pose_nt_dst) memcpy(ctx->pose_nt_dst + o, ctx->h_pose_nt_stage, B * sizeof(int32_t));
    }
    return CB_OK;
}

static void accumulate_timing(cb_timing &acc, const cb_timing &t)
{
    acc.h2d_ms += t.h2d_ms; acc.preprocess_ms += t.preprocess_ms; acc.threshold_ms += t.threshold_ms; acc.ccl_ms += t.ccl_ms;
    acc.cluster_ms += t.cluster_ms; acc.quad_ms += t.quad_ms; acc.decode_ms += t.decode_ms; acc.d2h_ms += t.d2h_ms; acc.total_ms += t.total_ms;
    acc.threshold_launches += t.threshold_launches; acc.kernel_launches += t.kernel_launches;
}

// upload a chunk of host frames into d_in with pitch == stride
static int upload_chunk(cb_ctx *ctx, const uint8_t *frames, size_t frame_stride, int height, int stride, int n, size_t &dev_frame_stride)
{
    const size_t bytes = (size_t)stride * height;
    dev_frame_stride = (bytes + 15) / 16 * 16;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (frame_stride == dev_frame_stride) {
        CK(cudaMemcpyAsync(ctx->d_in, frames, (size_t)(n - 1) * frame_stride + bytes, cudaMemcpyHostToDevice, ctx->stream));   // never past the last frame's pixels
    } else {
        CK(cudaMemcpy2DAsync(ctx->d_in, dev_frame_stride, frames, frame_stride, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    return CB_OK;
}

extern "C" {

int cb_detect_gray_device(cb_ctx *ctx, const uint8_t *frames_dev, int width, int height, int stride, size_t frame_stride,
                          int batch, cb_detection *out, int32_t *out_counts)
{
    if (!ctx || !frames_dev || !out || !out_counts) return CB_ERR_ARG;
    CB_NOT_STREAMING(ctx);
    CK(cudaSetDevice(ctx->device));
    std::vector<uint32_t> all_flags;
    cb_timing acc{};
    for (int b0 = 0; b0 < batch; b0 += ctx->max_batch) {
        const int n = std::min(ctx->max_batch, batch - b0);
        Geom g;
        int rc = make_geom(ctx, width, height, stride, frame_stride, n, g);
        if (rc) return rc;
        rc = detect_device_chunk(ctx, frames_dev + (size_t)b0 * frame_stride, g, out + (size_t)b0 * ctx->caps.dets_per_frame, out_counts + b0, false);
        if (rc) return rc;
        all_flags.insert(all_flags.end(), ctx->frame_flags.begin(), ctx->frame_flags.end());
        for (int i = 0; i < n; i++)
            for (int k = 0; k < out_counts[b0 + i]; k++) out[(size_t)(b0 + i) * ctx->caps.dets_per_frame + k].frame = b0 + i;
        accumulate_timing(acc, ctx->timing);
    }
    ctx->frame_flags = all_flags;
    ctx->timing = acc;
    return CB_OK;
}

// Host frames in, detection lists out.  Large batches are cut into chunks whose H2D copy (copy stream, second half of the
// input buffer) overlaps the kernels of the previous chunk (compute stream): the PCIe transfer disappears behind the compute.
static int detect_gray_pipelined(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                                 cb_detection *out, int32_t *out_counts)
{
    const size_t D = ctx->caps.dets_per_frame;
    const size_t bytes = (size_t)stride * height;
    const size_t dfs = (bytes + 15) / 16 * 16;
    const int half = ctx->max_batch / 2;
    // The first copy cannot overlap anything and small chunks run the kernels less efficiently, so the chunks ramp up
    // steeply: batch/8, 3 batch/8, then half the context's capacity (measured best on the c2 workload: 32 | 96 | 128
    // frames, against 4 x 64 and 2 x 128).  End to end the call costs about one small copy plus all kernels.
    std::vector<int> cstart;
    {
        const int min_first = (int)std::max<size_t>(1, ((size_t)16 << 20) / std::max<size_t>(bytes, 1));   // a first copy of >= 16 MB
        const int s1 = std::min(half, std::max(std::min(8, std::max(min_first, 1)), batch / 8));      // never more than one buffer half
        int b0 = 0, k = 0;
        while (b0 < batch) {
            const int sz = k == 0 ? s1 : (k == 1 ? std::min(half, 3 * s1) : half);
            cstart.push_back(b0);
            b0 += std::min(sz, batch - b0);
            k++;
        }
    }
    if (const char *e = getenv("CB_E2E_CHUNKS")) {     // experiment hook: explicit chunk sizes "32,96,128" (each <= max_batch / 2)
        cstart.clear();
        int b0 = 0;
        for (const char *p = e; *p && b0 < batch;) {
            const int sz = std::max(1, std::min(half, atoi(p)));
            cstart.push_back(b0); b0 += std::min(sz, batch - b0);
            while (*p && *p != ',') p++;
            if (*p == ',') p++;
            if (!*p) while (b0 < batch) { cstart.push_back(b0); b0 += std::min(sz, batch - b0); }
        }
    }
    cstart.push_back(batch);
    const int nchunks = (int)cstart.size() - 1;
    if (nchunks > 4096) return fail(ctx, CB_ERR_ARG, "too many chunks");
    uint8_t *bufs[2] = {ctx->d_in, ctx->d_in + (size_t)half * dfs};
    cb_timing acc{};
    if (ctx->h_frame_err_cap < (size_t)batch) {
        if (ctx->h_frame_err) { cudaFreeHost(ctx->h_frame_err); ctx->h_frame_err = nullptr; ctx->h_frame_err_cap = 0; }
        CK(cudaMallocHost((void **)&ctx->h_frame_err, (size_t)batch * sizeof(uint32_t)));
        ctx->h_frame_err_cap = (size_t)batch;
    }
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    CK(cudaEventRecord(ctx->ev_consumed[0], ctx->stream));   // make the first waits trivially satisfied
    CK(cudaEventRecord(ctx->ev_consumed[1], ctx->stream));
    int done_base = 0;    // first frame of the group currently staged in h_dets (h_dets holds max_batch frames)
    for (int c = 0; c < nchunks; c++) {
        const int b0 = cstart[c], n = cstart[c + 1] - b0, bi = c & 1;
        // staging area full: drain what has been produced so far
        if (b0 + n - done_base > ctx->max_batch) {
            CK(cudaStreamSynchronize(ctx->stream));
            for (int b = done_base; b < b0; b++) {
                out_counts[b] = ctx->h_counts[b - done_base];
                memcpy(out + (size_t)b * D, ctx->h_dets + (size_t)(b - done_base) * D, (size_t)out_counts[b] * sizeof(cb_detection));
                for (int k = 0; k < out_counts[b]; k++) out[(size_t)b * D + k].frame = b;
            }
            done_base = b0;
        }
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[bi], 0));
        if (frame_stride == dfs) CK(cudaMemcpyAsync(bufs[bi], frames + (size_t)b0 * frame_stride, (size_t)(n - 1) * frame_stride + bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        else CK(cudaMemcpy2DAsync(bufs[bi], dfs, frames + (size_t)b0 * frame_stride, frame_stride, bytes, n, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copied[bi], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[bi], 0));
        Geom g;
        int rc = make_geom(ctx, width, height, stride, dfs, n, g);
        if (rc) return rc;
        ctx->pose_frame_base = b0;
        rc = run_pipeline(ctx, bufs[bi], g, ST_FULL);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_consumed[bi], ctx->stream));
        const int off = b0 - done_base;
        CK(cudaMemcpyAsync(ctx->h_dets + (size_t)off * D, ctx->d_dets, (size_t)n * D * sizeof(cb_detection), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_counts + off, ctx->d_counts, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_chunk_err + c, ctx->d_small + 4 * (size_t)ctx->max_batch, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_frame_err + b0, ctx->d_small + 4 * (size_t)ctx->max_batch + 48, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        acc.kernel_launches += ctx->timing.kernel_launches;
        acc.threshold_launches += ctx->timing.threshold_launches;
    }
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {
        uint32_t flag = 0;
        for (int c = 0; c < nchunks; c++) flag |= ctx->h_chunk_err[c];
        ctx->frame_flags.assign(ctx->h_frame_err, ctx->h_frame_err + batch);
        if (flag) { int bad = 0; for (int b = 0; b < batch; b++) bad += ctx->frame_flags[b] != 0; note_overflow(ctx, flag, bad); }
    }
    for (int b = done_base; b < batch; b++) {
        out_counts[b] = ctx->h_counts[b - done_base];
        memcpy(out + (size_t)b * D, ctx->h_dets + (size_t)(b - done_base) * D, (size_t)out_counts[b] * sizeof(cb_detection));
        for (int k = 0; k < out_counts[b]; k++) out[(size_t)b * D + k].frame = b;
    }
    cudaEventElapsedTime(&acc.total_ms, ctx->ev[0], ctx->ev[7]);
    ctx->timing = acc;
    return CB_OK;
}

int cb_detect_gray(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                   cb_detection *out, int32_t *out_counts)
{
    if (!ctx || !frames || !out || !out_counts) return CB_ERR_ARG;
    CB_NOT_STREAMING(ctx);
    CK(cudaSetDevice(ctx->device));
    if ((size_t)stride * height > ctx->max_npix) return fail(ctx, CB_ERR_ARG, "stride*height exceeds the context's frame capacity");
    if (!ctx->family_set) return fail(ctx, CB_ERR_STATE, "no tag family set: call cb_set_family_tag36h11 first");
    // pipelined (copy of chunk i+1 under the kernels of chunk i) when there is enough to overlap: many frames, or few large ones
    if (ctx->max_batch >= 8 && batch >= 8 && (batch >= 64 || (size_t)batch * stride * height >= ((size_t)128 << 20))) {
        Geom g;
        int rc = make_geom(ctx, width, height, stride, frame_stride, 1, g);      // argument validation
        if (rc) return rc;
        return detect_gray_pipelined(ctx, frames, width, height, stride, frame_stride, batch, out, out_counts);
    }
    cb_timing acc{};
    std::vector<uint32_t> all_flags;
    for (int b0 = 0; b0 < batch; b0 += ctx->max_batch) {
        const int n = std::min(ctx->max_batch, batch - b0);
        size_t dfs;
        int rc = upload_chunk(ctx, frames + (size_t)b0 * frame_stride, frame_stride, height, stride, n, dfs);
        if (rc) return rc;
        Geom g;
        rc = make_geom(ctx, width, height, stride, dfs, n, g);
        if (rc) return rc;
        ctx->pose_frame_base = b0;
        rc = detect_device_chunk(ctx, ctx->d_in, g, out + (size_t)b0 * ctx->caps.dets_per_frame, out_counts + b0, true);
        if (rc) return rc;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < out_counts[b0 + i]; k++) out[(size_t)(b0 + i) * ctx->caps.dets_per_frame + k].frame = b0 + i;
        accumulate_timing(acc, ctx->timing);
        all_flags.insert(all_flags.end(), ctx->frame_flags.begin(), ctx->frame_flags.end());
    }
    ctx->timing = acc;
    ctx->frame_flags = all_flags;
    return CB_OK;
}

// ---- streaming form: a continuous feed of batches (the reference's camera loop hands frames over from a 4-slot host pool,
//      crates/chalkydri/src/cameras/gst_to_cu.rs:66,72) ----
// submit() only enqueues: the H2D copy of a batch goes to the one of two whole-batch input buffers that the batch before the
// previous one has released, on the copy stream, so it runs under the kernels of the batch in front of it, and the kernels
// see the batch in one piece (no chunking: 180 GB of HBM hold a second input buffer easily).  collect() waits for the oldest batch.
}  // extern "C"

static int stream_submit(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                         const double *gyro, double sign_change_error)
{
    if (!ctx || !frames) return CB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->family_set) return fail(ctx, CB_ERR_STATE, "no tag family set: call cb_set_family_tag36h11 first");
    if (ctx->ss_pending >= 2) return fail(ctx, CB_ERR_STATE, "two batches are already in flight: call cb_detect_gray_collect first");
    if (batch < 1 || batch > ctx->max_batch) return fail(ctx, CB_ERR_ARG, "submit batch %d outside 1..max_batch (%d)", batch, ctx->max_batch);
    if ((size_t)stride * height > ctx->max_npix) return fail(ctx, CB_ERR_ARG, "stride*height exceeds the context's frame capacity");
    {
        Geom g;
        int rc = make_geom(ctx, width, height, stride, frame_stride, 1, g);      // argument validation
        if (rc) return rc;
    }
    const size_t D = ctx->caps.dets_per_frame;
    cb_ctx::StreamSlot &sl = ctx->ss[(ctx->ss_head + ctx->ss_pending) & 1];
    if (!sl.h_dets) {
        CK(cudaMallocHost((void **)&sl.h_dets, (size_t)ctx->max_batch * D * sizeof(cb_detection)));
        CK(cudaMallocHost((void **)&sl.h_counts, (size_t)ctx->max_batch * sizeof(int32_t)));
        CK(cudaMallocHost((void **)&sl.h_err, 8 * sizeof(uint32_t)));
        CK(cudaMallocHost((void **)&sl.h_ferr, (size_t)ctx->max_batch * sizeof(uint32_t)));
        CK(cudaEventCreate(&sl.start));
        CK(cudaEventCreate(&sl.done));
    }
    if (!ctx->d_in2) CK(cudaMalloc((void **)&ctx->d_in2, ctx->in_bytes + 64));
    const int slot = (ctx->ss_head + ctx->ss_pending) & 1;
    const int pose_base = slot * ctx->max_batch;          // each slot owns max_batch frames of the pose buffers
    sl.has_pose = gyro != nullptr;
    if (gyro) {
        if (!ctx->camera_set) return fail(ctx, CB_ERR_STATE, "no camera set: call cb_set_camera first");
        if (!sl.h_gyro) {
            CK(cudaMallocHost((void **)&sl.h_gyro, (size_t)ctx->max_batch * sizeof(double)));
            CK(cudaMallocHost((void **)&sl.h_poses, (size_t)ctx->max_batch * sizeof(cb_pose)));
            CK(cudaMallocHost((void **)&sl.h_ok, (size_t)ctx->max_batch));
            CK(cudaMallocHost((void **)&sl.h_ntags, (size_t)ctx->max_batch * sizeof(int32_t)));
            CK(cudaEventCreateWithFlags(&sl.pose_done, cudaEventDisableTiming));
        }
        if (ctx->pose_cap < 2 * ctx->max_batch) {
            // a pose batch in flight implies the buffer already has this size, so nothing is using the old one
            if (ctx->d_pose_buf) { CK(cudaStreamSynchronize(ctx->pose_stream)); cudaFree(ctx->d_pose_buf); ctx->d_pose_buf = nullptr; ctx->pose_cap = 0; }
            drop_graphs(ctx);
            CK(cudaMalloc((void **)&ctx->d_pose_buf, pose_buf_bytes((size_t)2 * ctx->max_batch)));
            ctx->pose_cap = 2 * ctx->max_batch;
        }
        memcpy(sl.h_gyro, gyro, (size_t)batch * sizeof(double));         // the caller's array is free again when submit returns
        PoseBufs pb = pose_bufs(ctx);
        CK(cudaMemcpyAsync(pb.gyro + pose_base, sl.h_gyro, (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(pb.poses + pose_base, 0, (size_t)batch * sizeof(cb_pose), ctx->stream));
        ctx->pose_sign_change_error = sign_change_error;
    }
    const size_t bytes = (size_t)stride * height;
    const size_t dfs = (bytes + 15) / 16 * 16;
    const int nbuf = 2;
    const int half = ctx->max_batch;             // one chunk per batch
    uint8_t *bufs[2] = {ctx->d_in, ctx->d_in2};
    sl.batch = batch; sl.nchunks = 0; sl.launches = 0; sl.thr_launches = 0;
    CK(cudaEventRecord(sl.start, ctx->copy_stream));
    for (int b0 = 0; b0 < batch; b0 += half) {
        const int n = std::min(half, batch - b0), bi = (int)(ctx->ss_chunk % nbuf), c = sl.nchunks;
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[bi], 0));
        if (frame_stride == dfs) CK(cudaMemcpyAsync(bufs[bi], frames + (size_t)b0 * frame_stride, (size_t)(n - 1) * frame_stride + bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        else CK(cudaMemcpy2DAsync(bufs[bi], dfs, frames + (size_t)b0 * frame_stride, frame_stride, bytes, n, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copied[bi], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[bi], 0));
        Geom g;
        int rc = make_geom(ctx, width, height, stride, dfs, n, g);
        if (rc) return rc;
        ctx->pose_active = gyro != nullptr;
        ctx->pose_frame_base = pose_base + b0;
        rc = run_pipeline(ctx, bufs[bi], g, ST_FULL);
        ctx->pose_active = false;
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_consumed[bi], ctx->stream));
        CK(cudaMemcpyAsync(sl.h_dets + (size_t)b0 * D, ctx->d_dets, (size_t)n * D * sizeof(cb_detection), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(sl.h_counts + b0, ctx->d_counts, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(sl.h_err + c, ctx->d_small + 4 * (size_t)ctx->max_batch, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(sl.h_ferr + b0, ctx->d_small + 4 * (size_t)ctx->max_batch + 48, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        sl.launches += ctx->timing.kernel_launches;
        sl.thr_launches += ctx->timing.threshold_launches;
        sl.nchunks++;
        ctx->ss_chunk++;
    }
    CK(cudaEventRecord(sl.done, ctx->stream));
    if (gyro) {
        // the batch's solve was queued on pose_stream by run_pipeline; its results are read back on the same stream, so
        // the detection kernels of the next batch (ctx->stream) never wait for a solve
        PoseBufs pb = pose_bufs(ctx);
        CK(cudaMemcpyAsync(sl.h_poses, pb.poses + pose_base, (size_t)batch * sizeof(cb_pose), cudaMemcpyDeviceToHost, ctx->pose_stream));
        CK(cudaMemcpyAsync(sl.h_ok, pb.ok + pose_base, (size_t)batch, cudaMemcpyDeviceToHost, ctx->pose_stream));
        CK(cudaMemcpyAsync(sl.h_ntags, pb.n_tags + pose_base, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->pose_stream));
        CK(cudaEventRecord(sl.pose_done, ctx->pose_stream));
    }
    ctx->ss_pending++;
    return CB_OK;
}

static int stream_collect(cb_ctx *ctx, cb_detection *out, int32_t *out_counts, cb_pose *poses, uint8_t *pose_ok, int32_t *pose_tags, bool want_pose)
{

    if (!ctx || !out || !out_counts || (want_pose && (!poses || !pose_ok))) return CB_ERR_ARG;
    if (ctx->ss_pending == 0) return fail(ctx, CB_ERR_STATE, "nothing to collect: no batch has been submitted");
    CK(cudaSetDevice(ctx->device));
    cb_ctx::StreamSlot &sl = ctx->ss[ctx->ss_head];
    if (want_pose && !sl.has_pose) return fail(ctx, CB_ERR_STATE, "the oldest batch in flight was submitted without poses: collect it with cb_detect_gray_collect");
    cudaError_t e = cudaEventSynchronize(sl.done);
    if (e == cudaSuccess && sl.has_pose) e = cudaEventSynchronize(sl.pose_done);
    ctx->ss_head ^= 1;            // the batch leaves the queue whatever its outcome
    ctx->ss_pending--;
    if (e != cudaSuccess) return fail(ctx, CB_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e));
    {
        uint32_t flag = 0;
        for (int c = 0; c < sl.nchunks; c++) flag |= sl.h_err[c];
        ctx->frame_flags.assign(sl.h_ferr, sl.h_ferr + sl.batch);
        if (flag) { int bad = 0; for (int b = 0; b < sl.batch; b++) bad += ctx->frame_flags[b] != 0; note_overflow(ctx, flag, bad); }
    }
    const size_t D = ctx->caps.dets_per_frame;
    for (int b = 0; b < sl.batch; b++) {
        out_counts[b] = sl.h_counts[b];
        memcpy(out + (size_t)b * D, sl.h_dets + (size_t)b * D, (size_t)out_counts[b] * sizeof(cb_detection));
        for (int k = 0; k < out_counts[b]; k++) out[(size_t)b * D + k].frame = b;
    }
    if (want_pose) {
        memcpy(poses, sl.h_poses, (size_t)sl.batch * sizeof(cb_pose));
        memcpy(pose_ok, sl.h_ok, (size_t)sl.batch);
        if (pose_tags) memcpy(pose_tags, sl.h_ntags, (size_t)sl.batch * sizeof(int32_t));
    }
    cb_timing t{};
    cudaEventElapsedTime(&t.total_ms, sl.start, sl.done);      // first copy queued -> lists on the host (includes waiting behind the batch in front)
    t.kernel_launches = sl.launches; t.threshold_launches = sl.thr_launches;
    ctx->timing = t;
    return CB_OK;
}

extern "C" {

int cb_detect_gray_submit(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch)
{
    return stream_submit(ctx, frames, width, height, stride, frame_stride, batch, nullptr, 0.0);
}

int cb_detect_gray_collect(cb_ctx *ctx, cb_detection *out, int32_t *out_counts)
{
    return stream_collect(ctx, out, out_counts, nullptr, nullptr, nullptr, false);
}

// AprilTags::process (crates/apriltags/src/lib.rs:293-379) for a continuous feed: cb_detect_pose_gray in the streaming form
int cb_detect_pose_gray_submit(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                               const double *gyro, double sign_change_error)
{
    if (!gyro) return CB_ERR_ARG;
    return stream_submit(ctx, frames, width, height, stride, frame_stride, batch, gyro, sign_change_error);
}

int cb_detect_pose_gray_collect(cb_ctx *ctx, cb_detection *out, int32_t *out_counts, cb_pose *poses, uint8_t *pose_ok, int32_t *pose_tags)
{
    return stream_collect(ctx, out, out_counts, poses, pose_ok, pose_tags, true);
}

int cb_detect_gray_pending(const cb_ctx *ctx) { return ctx ? ctx->ss_pending : CB_ERR_ARG; }

// pre-processing front ends: convert into a gray buffer on the device, then the gray pipeline
static int detect_converted(cb_ctx *ctx, const uint8_t *frames, int width, int height, int bytes_per_px, int batch, cb_detection *out,
                            int32_t *out_counts)
{
    if (!ctx || !frames || !out || !out_counts) return CB_ERR_ARG;
    CB_NOT_STREAMING(ctx);
    CK(cudaSetDevice(ctx->device));
    if (width > ctx->max_w || height > ctx->max_h) return fail(ctx, CB_ERR_ARG, "frame larger than the context's capacity");
    const size_t npix = (size_t)width * height;
    const size_t gfs = (npix + 15) / 16 * 16;
    const size_t need_gray = (size_t)ctx->max_batch * gfs, need_raw = (size_t)ctx->max_batch * npix * bytes_per_px;
    if (ctx->gray_bytes < need_gray + need_raw + 64) {
        if (ctx->d_gray) cudaFree(ctx->d_gray);
        ctx->d_gray = nullptr;
        CK(cudaMalloc((void **)&ctx->d_gray, need_gray + need_raw + 64));
        ctx->gray_bytes = need_gray + need_raw + 64;
    }
    uint8_t *d_rawin = ctx->d_gray + (need_gray + 15) / 16 * 16;
    cb_timing acc{};
    for (int b0 = 0; b0 < batch; b0 += ctx->max_batch) {
        const int n = std::min(ctx->max_batch, batch - b0);
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        CK(cudaMemcpyAsync(d_rawin, frames + (size_t)b0 * npix * bytes_per_px, (size_t)n * npix * bytes_per_px, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ev[8], ctx->stream));
        {   // one launch for the whole chunk: grid.y = frame; every warp converts 512 pixels
            const dim3 grid((unsigned)((npix + PRE_PX_PER_WARP * (PRE_THREADS / 32) - 1) / (PRE_PX_PER_WARP * (PRE_THREADS / 32))), (unsigned)n);
            const dim3 grid_rgb((unsigned)((npix + RGB_PX_PER_BLOCK - 1) / RGB_PX_PER_BLOCK), (unsigned)n);
            if (bytes_per_px == 3) launch_rgb_to_gray(ctx, d_rawin, ctx->d_gray, npix, npix * 3, gfs, n, grid_rgb);
            else yuyv_to_gray_kernel<<<grid, PRE_THREADS, 0, ctx->stream>>>(d_rawin, ctx->d_gray, npix, npix * 2, gfs);
        }
        CK(cudaEventRecord(ctx->ev[9], ctx->stream));
        Geom g;
        int rc = make_geom(ctx, width, height, width, gfs, n, g);
        if (rc) return rc;
        rc = detect_device_chunk(ctx, ctx->d_gray, g, out + (size_t)b0 * ctx->caps.dets_per_frame, out_counts + b0, true);
        if (rc) return rc;
        float pre = 0, h2d = 0;
        cudaEventElapsedTime(&pre, ctx->ev[8], ctx->ev[9]);
        cudaEventElapsedTime(&h2d, ctx->ev[0], ctx->ev[8]);
        ctx->timing.preprocess_ms = pre; ctx->timing.h2d_ms = h2d; ctx->timing.kernel_launches += 1;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < out_counts[b0 + i]; k++) out[(size_t)(b0 + i) * ctx->caps.dets_per_frame + k].frame = b0 + i;
        accumulate_timing(acc, ctx->timing);
    }
    ctx->timing = acc;
    return CB_OK;
}

// stage tap of the pre-processing kernels: host frames in, gray planes out (tests and tools/bench_preprocess.py)
static int convert_tap(cb_ctx *ctx, const uint8_t *frames, int width, int height, int bytes_per_px, int batch, uint8_t *gray_out)
{
    if (!ctx || !frames || !gray_out || width < 1 || height < 1 || batch < 1 || batch > 65535) return CB_ERR_ARG;
    CB_NOT_STREAMING(ctx);
    CK(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)width * height;
    uint8_t *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)batch * npix * (bytes_per_px + 1) + 32));
    uint8_t *d_out = d + ((size_t)batch * npix * bytes_per_px + 15) / 16 * 16;
    cudaError_t e = cudaMemcpyAsync(d, frames, (size_t)batch * npix * bytes_per_px, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[8], ctx->stream);
    if (e == cudaSuccess) {
        const dim3 grid((unsigned)((npix + PRE_PX_PER_WARP * (PRE_THREADS / 32) - 1) / (PRE_PX_PER_WARP * (PRE_THREADS / 32))), (unsigned)batch);
        const dim3 grid_rgb((unsigned)((npix + RGB_PX_PER_BLOCK - 1) / RGB_PX_PER_BLOCK), (unsigned)batch);
        if (bytes_per_px == 3) launch_rgb_to_gray(ctx, d, d_out, npix, npix * 3, npix, batch, grid_rgb);
        else yuyv_to_gray_kernel<<<grid, PRE_THREADS, 0, ctx->stream>>>(d, d_out, npix, npix * 2, npix);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[9], ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(gray_out, d_out, (size_t)batch * npix, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) {
        ctx->timing = cb_timing{};
        cudaEventElapsedTime(&ctx->timing.preprocess_ms, ctx->ev[8], ctx->ev[9]);
        ctx->timing.kernel_launches = 1;
    }
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, CB_ERR_CUDA, "pre-processing tap: %s", cudaGetErrorString(e));
    return CB_OK;
}

int cb_rgb_to_gray(cb_ctx *ctx, const uint8_t *frames_rgb, int width, int height, int batch, uint8_t *gray_out)
{
    return convert_tap(ctx, frames_rgb, width, height, 3, batch, gray_out);
}

int cb_yuyv_to_gray(cb_ctx *ctx, const uint8_t *frames_yuyv, int width, int height, int batch, uint8_t *gray_out)
{
    return convert_tap(ctx, frames_yuyv, width, height, 2, batch, gray_out);
}

int cb_detect_rgb(cb_ctx *ctx, const uint8_t *frames_rgb, int width, int height, int batch, cb_detection *out, int32_t *out_counts)
{
    return detect_converted(ctx, frames_rgb, width, height, 3, batch, out, out_counts);
}

int cb_detect_yuyv(cb_ctx *ctx, const uint8_t *frames_yuyv, int width, int height, int batch, cb_detection *out, int32_t *out_counts)
{
    if (width % 2) return fail(ctx, CB_ERR_ARG, "YUYV needs an even width");
    return detect_converted(ctx, frames_yuyv, width, height, 2, batch, out, out_counts);
}

// NV12 / NV21 / I420 / YV12 camera buffers (crates/chalkydri/src/cameras/gst_to_cu.rs:152-188): the first width*height bytes of
// each frame are the Y plane, which IS the gray image; the chroma planes behind it are never uploaded (a strided copy).
int cb_detect_yuv420(cb_ctx *ctx, const uint8_t *frames, int width, int height, int batch, cb_detection *out, int32_t *out_counts)
{
    if (!ctx) return CB_ERR_ARG;
    if ((width | height) & 1) return fail(ctx, CB_ERR_ARG, "4:2:0 formats need even width and height");
    return cb_detect_gray(ctx, frames, width, height, width, (size_t)width * height * 3 / 2, batch, out, out_counts);
}

// ---- stage taps ----
static int tap_common(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, int stage, Geom &g)
{
    if (!ctx || !frames) return CB_ERR_ARG;
    CB_NOT_STREAMING(ctx);
    CK(cudaSetDevice(ctx->device));
    if (batch > ctx->max_batch) return fail(ctx, CB_ERR_ARG, "tap batch %d exceeds max_batch %d", batch, ctx->max_batch);
    if ((size_t)stride * height > ctx->max_npix) return fail(ctx, CB_ERR_ARG, "stride*height exceeds the context's frame capacity");
    size_t dfs;
    int rc = upload_chunk(ctx, frames, frame_stride, height, stride, batch, dfs);
    if (rc) return rc;
    rc = make_geom(ctx, width, height, stride, dfs, batch, g);
    if (rc) return rc;
    rc = run_pipeline(ctx, ctx->d_in, g, stage);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_small, ctx->d_small, (5 * (size_t)ctx->max_batch + 48) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx, true, true);
    return check_errflag(ctx, batch);
}

int cb_threshold(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, uint8_t *out)
{
    Geom g;
    if (!out) return CB_ERR_ARG;
    int rc = tap_common(ctx, frames, width, height, stride, frame_stride, batch, ST_THRESH, g);
    if (rc) return rc;
    CK(cudaMemcpy2D(out, g.w, ctx->d_thresh, g.tp, g.w, (size_t)g.h * batch, cudaMemcpyDeviceToHost));
    return CB_OK;
}

int cb_labels(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, uint32_t *labels, uint32_t *sizes)
{
    Geom g;
    if (!labels) return CB_ERR_ARG;
    int rc = tap_common(ctx, frames, width, height, stride, frame_stride, batch, ST_LABELS, g);
    if (rc) return rc;
    const size_t n = (size_t)batch * g.npix;
    CK(cudaMemcpy(labels, ctx->d_labels, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // labels are batch-global indices; report them frame-local, sizes per pixel (size of its component)
    std::vector<uint32_t> sz;
    if (sizes) { sz.resize(n); CK(cudaMemcpy(sz.data(), ctx->d_sizes, n * sizeof(uint32_t), cudaMemcpyDeviceToHost)); }
    for (size_t i = 0; i < n; i++) {
        if (sizes) sizes[i] = sz[labels[i]];
        labels[i] -= (uint32_t)(i / g.npix) * g.npix;
    }
    return CB_OK;
}

int cb_quads(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, float *quads, int cap,
             int32_t *counts, int64_t *npoints_total)
{
    Geom g;
    if (!quads || !counts) return CB_ERR_ARG;
    int rc = tap_common(ctx, frames, width, height, stride, frame_stride, batch, ST_QUADS, g);
    if (rc) return rc;
    const uint32_t total = ctx->h_small[4 * (size_t)ctx->max_batch + 3];
    std::vector<QuadRec> q(total);
    if (total) CK(cudaMemcpy(q.data(), ctx->d_quads, total * sizeof(QuadRec), cudaMemcpyDeviceToHost));
    for (int b = 0; b < batch; b++) counts[b] = 0;
    for (uint32_t i = 0; i < total; i++) {
        const int b = q[i].frame;
        if (counts[b] < cap) memcpy(quads + ((size_t)b * cap + counts[b]) * 8, q[i].p, 8 * sizeof(float));
        counts[b]++;
    }
    if (npoints_total) {
        int64_t t = 0;
        for (int b = 0; b < batch; b++) t += ctx->h_small[ctx->max_batch + b];
        *npoints_total = t;
    }
    return CB_OK;
}

int cb_clusters(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, int16_t *pts,
                int32_t *cluster_of, int64_t cap, int64_t *npoints, int32_t *nclusters)
{
    Geom g;
    if (!pts || !cluster_of || !npoints || !nclusters || cap < 0) return CB_ERR_ARG;
    int rc = tap_common(ctx, frames, width, height, stride, frame_stride, batch, ST_CLUSTERS, g);
    if (rc) return rc;
    const Caps &caps = ctx->caps;
    int64_t k = 0;
    int32_t cluster_base = 0;
    for (int b = 0; b < batch; cluster_base += nclusters[b], b++) {
        const uint32_t ncl = std::min(ctx->h_small[b], caps.clusters_per_frame);          // d_ncl[b]
        nclusters[b] = (int32_t)ncl;
        std::vector<ClusterRec> recs(ncl);
        if (ncl) CK(cudaMemcpy(recs.data(), ctx->d_clusters + (size_t)b * caps.clusters_per_frame, ncl * sizeof(ClusterRec), cudaMemcpyDeviceToHost));
        std::vector<uint32_t> keys;
        for (uint32_t c = 0; c < ncl; c++) {
            keys.resize(recs[c].count);
            if (recs[c].count) CK(cudaMemcpy(keys.data(), ctx->d_scankey + (size_t)b * caps.points_per_frame + recs[c].offset, recs[c].count * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            for (uint32_t i = 0; i < recs[c].count; i++, k++) {
                if (k >= cap) continue;
                // scan key (clusters.cuh): (pixel index << 3) | (probe << 1) | (v1 > v0); the point upstream stores for it (sort.cuh decode_point)
                const uint32_t key = keys[i], pix = key >> 3;
                const int d = (key >> 1) & 3, sgn = key & 1;
                const int x = (int)(pix % (uint32_t)g.w), y = (int)(pix / (uint32_t)g.w);
                const int dx = d == 2 ? -1 : (d == 1 ? 0 : 1), dy = d == 0 ? 0 : 1, dv = sgn ? 255 : -255;
                pts[4 * k] = (int16_t)(2 * x + dx); pts[4 * k + 1] = (int16_t)(2 * y + dy); pts[4 * k + 2] = (int16_t)(dx * dv); pts[4 * k + 3] = (int16_t)(dy * dv);
                cluster_of[k] = cluster_base + (int32_t)c;
            }
        }
    }
    *npoints = k;
    return CB_OK;
}

// ---- plumbing ----
void *cb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void cb_host_free(void *p) { if (p) cudaFreeHost(p); }
void *cb_device_alloc(cb_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { fail(ctx, CB_ERR_CUDA, "cudaMalloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void cb_device_free(cb_ctx *ctx, void *p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }
int cb_memcpy_h2d(cb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!ctx) return CB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return CB_OK;
}
int cb_memcpy_d2h(cb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ctx) return CB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return CB_OK;
}

}  // extern "C"

#include "api_solver.inc"
#include "api_cat.inc"
#include "api_pool.inc"
