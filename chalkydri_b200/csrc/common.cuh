// common.cuh -- shared declarations of the B200 detector / solver kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chalkydri_b200.h"

namespace cb {

// Geometry of one batch launch.  Decimated image w x h; thresh / mark maps use pitch tp (multiple of 16),
// labels / sizes use pitch w so that a label value decodes to (frame, y, x) directly.
struct Geom {
    int W, H, stride;        // full-resolution input
    size_t frame_stride;     // bytes between frames of the input
    int f;                   // integer decimation factor (quad_decimate)
    int w, h, tp;            // decimated size and pitch of the u8 maps
    int tw, th;              // number of FULL 4x4 tiles (w/4, h/4)
    int batch;
    uint32_t npix;           // w*h
};

struct DetParams {
    float quad_decimate;
    int refine_edges;
    double decode_sharpening;
    int min_cluster_pixels;
    int max_nmaxima;
    float critical_rad;
    float max_line_fit_mse;
    int min_white_black_diff;
    int bits_corrected;
    double cos_critical_rad;   // host libm cos(critical_rad), like upstream's precomputed qtp.cos_critical_rad
    float smooth_f[7];         // host-computed low-pass taps (float)exp(-j*j/2), j=-3..3
    int min_tag_width;
};

// hash table entry for gradient clusters (one sub-table per frame)
struct ClusterSlot {
    unsigned long long key;   // (rep_hi << 32) | rep_lo, ~0 = empty
    uint32_t count;           // boundary points emitted for this pair
    uint32_t cluster;         // index into the frame's selected-cluster list, 0xffffffff = not selected
};

struct ClusterRec {
    unsigned long long key;
    uint32_t offset;          // first point (index inside the frame's point buffer)
    uint32_t count;
    uint32_t cursor;          // scatter cursor
    uint32_t pad;
};

struct QuadRec {
    float p[4][2];            // corners (decimated coordinates as fitted; rescaled by the decode kernel)
    int32_t reversed_border;
    int32_t npoints;
    unsigned long long key;
    int32_t frame;
    int32_t pad;
};

struct Caps {
    uint32_t slots_per_frame;     // power of two
    uint32_t clusters_per_frame;
    uint32_t points_per_frame;
    uint32_t quads_per_frame;
    uint32_t dets_per_frame;
    uint32_t tile_probes;         // probe limit of the per-tile cluster table (clusters.cuh); lowered by tests to force the overflow path
};

// error flag bits written by kernels
enum : uint32_t { ERR_HASH_FULL = 1, ERR_CLUSTERS_FULL = 2, ERR_POINTS_FULL = 4, ERR_QUADS_FULL = 8 };
// A fixed per-frame table overflowed in frame b: the batch-wide flag word and, 48 words behind it, the frame's own word (api.cu lays
// the small counters out as [.. | misc 48 | frame flags B]).  Every table is per frame, so the other frames of the batch are complete;
// the reconcile pass empties the list of a flagged frame.
__device__ __forceinline__ void flag_overflow(uint32_t *errflag, int b, uint32_t bit)
{
    atomicOr(errflag, bit);
    atomicOr(errflag + 48 + b, bit);
}

static constexpr int kNumCodes = 587;

}  // namespace cb
