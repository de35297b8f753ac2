/*
 * chalkydri_b200.h -- C ABI of libchalkydri_b200.so, the B200 (sm_100a) drop-in for Chalkydri's
 * AprilTag detection + SQPnP hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point cites the reference interface it replaces (paths relative to /root/reference).
 * Conventions: int return, 0 = ok, <0 = error (cb_last_error gives the text); nothing
 * unwinds across the boundary; the caller owns every buffer it passes in; outputs are caller-allocated
 * fixed-capacity arrays.  A context is single-threaded (like one `AprilTags` task instance,
 * crates/apriltags/src/lib.rs:166-182); several contexts (one per GPU / stream) may run concurrently.
 * There is no CPU fallback: every call fails with CB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CHALKYDRI_B200_H
#define CHALKYDRI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK 0
#define CB_ERR_ARG (-1)        /* bad argument (null pointer, size beyond the context's capacity, ...) */
#define CB_ERR_CUDA (-2)       /* CUDA runtime error; text in cb_last_error */
#define CB_ERR_UNSUPPORTED (-3)/* parameter value the kernels do not implement (quad_sigma != 0, decimate 1.5, ...) */
#define CB_ERR_OVERFLOW (-4)   /* an output list does not fit the capacity the caller gave (CAT lists, stage taps) */
#define CB_ERR_STATE (-5)      /* family not set etc. */

typedef struct cb_ctx cb_ctx;

/* mirrors apriltag_detection_t as exposed by `apriltag::Detection` (id(), hamming(), decision_margin(),
 * corners(), center(), homography(); used at crates/apriltags/src/lib.rs:306,310-314) plus the frame index */
typedef struct {
    int32_t frame;            /* index inside the batch */
    int32_t id;
    int32_t hamming;
    float   decision_margin;
    double  H[9];             /* row-major 3x3 */
    double  c[2];             /* centre */
    double  p[4][2];          /* corners, counter-clockwise starting at tag (-1, 1) like apriltag.c */
} cb_detection;

/* nalgebra Isometry3<f64> (chalkydri_sqpnp Iso3, crates/chalkydri_sqpnp/src/lib.rs:24) */
typedef struct {
    double t[3];
    double q[4];              /* unit quaternion w, x, y, z */
} cb_iso3;

/* Some((Rot3, Vec3 position, Vec3 std_devs)) of SqPnP::solve_robot_pose (lib.rs:297-304,376) */
typedef struct {
    double rot[9];            /* column-major 3x3 like nalgebra */
    double pos[3];
    double std_devs[3];
} cb_pose;

/* per-stage device time of the last detect call, CUDA events on the context's stream */
typedef struct {
    float h2d_ms, preprocess_ms, threshold_ms, ccl_ms, cluster_ms, quad_ms, decode_ms, d2h_ms, total_ms;
    int32_t threshold_launches;   /* launches of the fused decimate+threshold kernel inside the call */
    int32_t kernel_launches;      /* all kernel launches inside the call */
} cb_timing;

/* ---- lifetime: replaces apriltag_detector_create/destroy behind DetectorBuilder::build()
 *      (crates/apriltags/src/lib.rs:258-261) ---- */
cb_ctx *cb_create(int device, int max_width, int max_height, int max_batch, int max_dets_per_frame);
void cb_destroy(cb_ctx *ctx);
/* last error text of this context; ctx == NULL returns the text of the last failed cb_create on this thread */
const char *cb_last_error(const cb_ctx *ctx);

/* tag36h11_create + apriltag_detector_add_family_bits (DetectorBuilder::add_family_bits, lib.rs:259,280).
 * bits_corrected in [0,3]. */
int cb_set_family_tag36h11(cb_ctx *ctx, int bits_corrected);

/* apriltag_detector_t fields the reference leaves at their defaults.  quad_decimate must be an
 * integer in 1..16 (2, the reference's value, takes the fused fast path; upstream's special 1.5 is CB_ERR_UNSUPPORTED);
 * quad_decimate = 1 needs a context created with twice the frame size (the per-pixel buffers are sized for the decimated
 * frame).  quad_sigma must be 0; deglitch is not offered.  CB_ERR_STATE while batches are in flight. */
int cb_set_params(cb_ctx *ctx, float quad_decimate, float quad_sigma, int refine_edges, double decode_sharpening,
                  int min_cluster_pixels, int max_nmaxima, float critical_rad, float max_line_fit_mse,
                  int min_white_black_diff);

/* Detector::detect (crates/apriltags/src/lib.rs:301) on a batch of 8-bit gray frames in HOST memory
 * (image_u8_t{buf,width,height,stride}, lib.rs:197-213; frame b starts at frames + b*frame_stride bytes).
 * out holds batch*max_dets_per_frame records, frame b's detections start at out[b*max_dets_per_frame],
 * sorted by id; out_counts[b] is their number.  Includes H2D of the frames and D2H of the lists. */
int cb_detect_gray(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride,
                   int batch, cb_detection *out, int32_t *out_counts);
/* same with the frames already resident in device memory (frames_dev is a device pointer) */
int cb_detect_gray_device(cb_ctx *ctx, const uint8_t *frames_dev, int width, int height, int stride,
                          size_t frame_stride, int batch, cb_detection *out, int32_t *out_counts);
/* Streaming form of cb_detect_gray for a continuous feed of batches: the reference's camera loop hands frames to
 * AprilTags::process one after another from a 4-slot host pool (crates/chalkydri/src/cameras/gst_to_cu.rs:66,72;
 * crates/apriltags/src/lib.rs:293-301), so batch k+1 is already in host memory while batch k is being detected.
 *   cb_detect_gray_submit  enqueues one batch (1..max_batch frames, same argument meaning as cb_detect_gray) and returns
 *                          without waiting; at most two batches may be in flight (CB_ERR_STATE beyond that).  The frames
 *                          must stay valid and unchanged until the batch has been collected; pinned memory
 *                          (cb_host_alloc) lets the H2D copy of batch k+1 run under the kernels of batch k.
 *   cb_detect_gray_collect waits for the OLDEST submitted batch and writes its lists exactly like cb_detect_gray
 *                          (CB_ERR_STATE when nothing is in flight).  A failed batch still leaves the queue.
 *   cb_detect_gray_pending number of submitted, not yet collected batches (0..2).
 * While batches are in flight every other detection / tap entry point of the context returns CB_ERR_STATE. */
int cb_detect_gray_submit(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride,
                          int batch);
int cb_detect_gray_collect(cb_ctx *ctx, cb_detection *out, int32_t *out_counts);
int cb_detect_gray_pending(const cb_ctx *ctx);
/* packed-RGB input (CAT contract, crates/chalkydri-apriltags/src/lib.rs:265-267): gray = utils.rs:43 */
int cb_detect_rgb(cb_ctx *ctx, const uint8_t *frames_rgb, int width, int height, int batch, cb_detection *out,
                  int32_t *out_counts);
/* YUYV (YUY2) camera buffers (crates/chalkydri/src/cameras/gst_to_cu.rs:152-188 lists the formats): gray = Y */
int cb_detect_yuyv(cb_ctx *ctx, const uint8_t *frames_yuyv, int width, int height, int batch, cb_detection *out,
                   int32_t *out_counts);
/* planar / semi-planar 4:2:0 camera buffers -- NV12, NV21, I420, YV12 (gst_to_cu.rs:152-188): frame b starts at
 * frames + b*width*height*3/2 and its first width*height bytes (the Y plane) are the gray image.  Same as cb_detect_gray
 * with stride = width and frame_stride = width*height*3/2; only the Y planes are uploaded.  width, height even. */
int cb_detect_yuv420(cb_ctx *ctx, const uint8_t *frames, int width, int height, int batch, cb_detection *out,
                     int32_t *out_counts);

/* ---- stage taps (parity tests; run the pipeline up to that stage on HOST input) ---- */
/* pre-processing alone: packed RGB -> gray with CAT's formula (crates/chalkydri-apriltags/src/utils.rs:33-46), YUYV -> Y;
 * gray_out[batch][height][width].  Frame b starts at frames + b*width*height*{3,2}; planes that do not start on a 16-byte
 * boundary take a scalar path with the same results. */
int cb_rgb_to_gray(cb_ctx *ctx, const uint8_t *frames_rgb, int width, int height, int batch, uint8_t *gray_out);
int cb_yuyv_to_gray(cb_ctx *ctx, const uint8_t *frames_yuyv, int width, int height, int batch, uint8_t *gray_out);
/* decimated size the detector works on */
int cb_decimated_size(const cb_ctx *ctx, int width, int height, int *w, int *h);
/* threshold(): out[batch][h][w] in {0,127,255} */
int cb_threshold(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                 uint8_t *out);
/* connected_components(): labels[batch][h][w] = smallest pixel index (y*w+x) of the component; sizes = its size */
int cb_labels(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
              uint32_t *labels, uint32_t *sizes);
/* gradient_clusters(): the clusters fit_quad() can accept (24 <= n <= 3(2w+2h) boundary points), every cluster's points in upstream's
 * append order (scan order y, x, probe), before any sort.  pts[k] = (x, y, gx, gy) exactly as upstream stores a point (half-pixel
 * coordinates 2x+dx, 2y+dy; gradient +-255 along the probe), cluster_of[k] = running cluster number over the batch (frame 0's clusters first; nclusters[b] per frame).  npoints = points the frames hold (may exceed cap: then only the first cap were written), nclusters[batch]. */
int cb_clusters(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch, int16_t *pts,
                int32_t *cluster_of, int64_t cap, int64_t *npoints, int32_t *nclusters);
/* The detector's device tables (cluster hash, clusters, points, candidate quads) are sized per frame from the context's frame size.
 * A frame that overflows one of them -- upstream has no such limit; think of a frame of pure noise -- reports an EMPTY detection list
 * and a flag word; the other frames of the batch are complete and the call returns CB_OK (cb_last_error holds a note).  flags[i] for
 * frame i of the last completed detection call (blocking call: at return; streaming form: at collect): bit 0 cluster hash, 1 clusters,
 * 2 points, 3 quads.  Returns the number of flagged frames among the first n. */
int cb_frame_flags(const cb_ctx *ctx, uint32_t *flags, int n);

/* fit_quads(): quads[batch][cap] corners in decimated coordinates (float[4][2]) + counts */
int cb_quads(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
             float *quads, int cap, int32_t *counts, int64_t *npoints_total);

int cb_get_timing(const cb_ctx *ctx, cb_timing *t);

/* ---- solver: SqPnP::solve_robot_pose (crates/chalkydri_sqpnp/src/lib.rs:297-377), batched.
 * Problem i uses tags[i*max_tags .. +n_tags[i]) and bearings[(i*max_tags*4) .. ) (3 doubles each, 4 per tag in the
 * detector's corner order).  ok[i] = 1 for Some, 0 for None.  max_iter / tol mirror the builder (lib.rs:214-222). */
int cb_sqpnp_set(cb_ctx *ctx, int max_iter, double tolerance);
int cb_sqpnp_batch(cb_ctx *ctx, const cb_iso3 *tags, const double *bearings, const int32_t *n_tags, int max_tags,
                   const cb_iso3 *robot_to_cam, const double *gyro, double sign_change_error, int64_t n,
                   cb_pose *out, uint8_t *ok);
/* same with every array already in device memory */
int cb_sqpnp_batch_device(cb_ctx *ctx, const cb_iso3 *tags, const double *bearings, const int32_t *n_tags, int max_tags,
                          const cb_iso3 *robot_to_cam, const double *gyro, double sign_change_error, int64_t n,
                          cb_pose *out, uint8_t *ok);
/* ---- row B0, AprilTags::process on the device (crates/apriltags/src/lib.rs:293-379): detect -> field lookup -> un-project ->
 *      one multi-tag SQPnP per frame, without the detections leaving the device in between.
 *  cb_set_field: the tag layout of field.json (AprilTagFieldLayout, crates/apriltags/src/field_layout.rs:18-44); tags not in
 *      it are skipped like lib.rs:306-308.
 *  cb_set_camera: OpenCVModel5 intrinsics {fx, fy, cx, cy, k1, k2, p1, p2, k3} (lib.rs:232-233) and the robot->camera
 *      transform from cb_create_solver_camera_transform; NULL = identity (lib.rs:333).
 *  cb_detect_pose_gray: host frames in; detection lists like cb_detect_gray, plus per frame the Some((rot, pos, std_devs)) of
 *      solve_robot_pose in poses[b] with pose_ok[b] = 1, or pose_ok[b] = 0 for None (no detections, no usable tag, solver
 *      None) and for gyro[b] = NaN (comm.gyro_angle() == None, lib.rs:329).  pose_tags (optional): tags used per frame.
 *      At most 32 tags per frame enter the solver (the first 32 of the list, which is ordered by id). ---- */
int cb_set_field(cb_ctx *ctx, const int32_t *ids, const cb_iso3 *poses, int n);
int cb_set_camera(cb_ctx *ctx, const double *params9, const cb_iso3 *robot_to_cam);
int cb_detect_pose_gray(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride, int batch,
                        const double *gyro, double sign_change_error, cb_detection *out, int32_t *out_counts, cb_pose *poses,
                        uint8_t *pose_ok, int32_t *pose_tags);

/* Streaming form of cb_detect_pose_gray (same queue and rules as cb_detect_gray_submit / _collect; the two forms may be mixed,
 * batches are collected oldest first).  gyro[batch] is copied before submit returns.  A batch's SQPnP solve and the read-back
 * of its poses run on a side stream under the detection kernels of the next batch.  cb_detect_pose_gray_collect on a batch
 * that was submitted without poses returns CB_ERR_STATE and leaves it queued. */
int cb_detect_pose_gray_submit(cb_ctx *ctx, const uint8_t *frames, int width, int height, int stride, size_t frame_stride,
                               int batch, const double *gyro, double sign_change_error);
int cb_detect_pose_gray_collect(cb_ctx *ctx, cb_detection *out, int32_t *out_counts, cb_pose *poses, uint8_t *pose_ok,
                                int32_t *pose_tags);

/* ---- output contract of the task: the 64-byte record whacknet sends to the robot controller
 *      (struct VisionMeasurement, crates/whacknet/src/lib.rs:40-66; the reference's one test checks its size, :92-95) ---- */
typedef struct {
    double x, y, rot;             /* RobotPose (whacknet/src/lib.rs:17-26): position x, y and yaw */
    double std_x, std_y, std_rot; /* VisionUncertainty (:29-38) */
    uint64_t ts;                  /* microseconds */
    uint8_t camera_id;
    uint8_t tag_count;
    uint8_t reserved[6];
} cb_vision_measurement;
/* What AprilTags::process publishes for frame i (crates/apriltags/src/lib.rs:340-376): pose_ok[i] != 0 -> {x = pos[0],
 * y = pos[1], rot = Rotation3::euler_angles().2, the three std-devs, tag_count = min(det_counts[i], 255)}; otherwise the
 * heartbeat record (default pose and uncertainty, tag_count 0).  Host-side arithmetic only; ts_us[i] is copied through. */
int cb_pack_vision_measurements(const cb_pose *poses, const uint8_t *pose_ok, const int32_t *det_counts, const uint64_t *ts_us,
                                uint8_t camera_id, int n, cb_vision_measurement *out);

/* SqPnP::create_solver_camera_transform (lib.rs:430-461); host-side scalar helper */
int cb_create_solver_camera_transform(double fwd_m, double left_m, double up_m, double roll_deg, double pitch_deg,
                                      double yaw_deg, cb_iso3 *out);
/* GenericModel::unproject for OpenCVModel5 (crates/apriltags/src/lib.rs:316-321), batched on the device:
 * params = fx,fy,cx,cy,k1,k2,p1,p2,k3; px[n][2] -> bearings[n][3], ok[n] */
int cb_unproject_opencv5(cb_ctx *ctx, const double *params9, const double *px, int64_t n, double *bearings, uint8_t *ok);

/* ---- CAT stages (crates/chalkydri-apriltags/src/lib.rs), HOST buffers ---- */
/* Detector::calc_otsu (lib.rs:191-259): packed RGB -> Color map (0 Black, 1 White, 2 Other) */
int cb_cat_calc_otsu(cb_ctx *ctx, const uint8_t *rgb, int width, int height, uint8_t *color);
/* Detector::thresh (lib.rs:319-334) */
int cb_cat_thresh(cb_ctx *ctx, const uint8_t *rgb, int width, int height, uint8_t *color);
/* Detector::detect_corners (lib.rs:291-309,345-400): (x,y) pairs in the reference's scan order (x-major) */
int cb_cat_detect_corners(cb_ctx *ctx, const uint8_t *color, int width, int height, int32_t *xy, int64_t cap, int64_t *n);
/* Detector::check_edges (lib.rs:409-499): (x1,y1,x2,y2) in the reference's order */
int cb_cat_check_edges(cb_ctx *ctx, const uint8_t *color, int width, int height, const int32_t *xy, int64_t npts,
                       int32_t *lines, int64_t cap, int64_t *n);
/* Detector::process_frame (lib.rs:265-287) in ONE call: calc_otsu -> reset -> detect_corners -> check_edges.  The frame is uploaded
 * once; gray plane, colour map and corner list stay on the device between the stages (the reference's only benchmark times exactly
 * this call on a 703x905 frame, crates/chalkydri-apriltags/bench.rs:8-26).  xy / lines as above; color (optional, may be NULL)
 * receives the Color map.  CB_ERR_OVERFLOW when a list does not fit its capacity (n_points / n_lines still report the need).
 * cb_get_timing: h2d_ms, preprocess_ms (all CAT kernels), d2h_ms, total_ms. */
int cb_cat_process_frame(cb_ctx *ctx, const uint8_t *rgb, int width, int height, uint8_t *color, int32_t *xy, int64_t xy_cap,
                         int64_t *n_points, int32_t *lines, int64_t lines_cap, int64_t *n_lines);
/* CAT's decode intent -- book/src/maintenance/apriltags.md:58-60 ("Decoding tags is done pretty much the same way the C library does
 * it") and the commented-out cluster() over connected_components(), lib.rs:551-613: the frame is thresholded the CAT way
 * (use_otsu != 0: calc_otsu, lib.rs:191-259; 0: thresh, lib.rs:319-334) at full resolution, and from that map on the stages are the
 * detector's (connected components, gradient clusters, quad fit, refine, decode, reconcile) sampling CAT's gray plane
 * (utils.rs:33-46).  out[0 .. max_dets) / *out_count like cb_detect_gray with batch 1.  The context needs a tag family and, because
 * the stages run undecimated, a capacity of twice the frame size (cb_create(.., 2*width, 2*height, ..)). */
int cb_cat_detect_tags(cb_ctx *ctx, const uint8_t *rgb, int width, int height, int use_otsu, cb_detection *out, int32_t *out_count);
/* Detector::connected_components (lib.rs:501-549): min-index labels and component sizes */
int cb_cat_connected_components(cb_ctx *ctx, const uint8_t *color, int width, int height, uint32_t *labels, uint32_t *sizes);

/* ---- several GPUs of one box, one process: frames are independent, so a batch is sharded over the GPUs with no
 *      collective -- one context and one host thread per GPU, each GPU's lists written straight into its slice of the caller's
 *      one output array.  This is the batched form of running one AprilTags task per camera (crates/apriltags/src/lib.rs:166-182,
 *      three of them in chalkydri.ron:2-105) with camera streams spread over GPUs. ---- */
typedef struct cb_pool cb_pool;
typedef struct {
    float wall_ms;                /* host wall time of the last cb_pool_detect_gray call */
    float max_device_ms;          /* busiest / least busy worker thread (submit of its first batch .. collect of its last) */
    float min_device_ms;
    int32_t n_devices;
} cb_pool_timing;
/* devices[n_devices]: CUDA ordinals (one context each; an ordinal may repeat); devices == NULL or n_devices <= 0: every visible
 * GPU.  The other arguments are cb_create's, per GPU.  NULL on failure (cb_pool_last_error(NULL) has the text). */
cb_pool *cb_pool_create(const int *devices, int n_devices, int max_width, int max_height, int max_batch, int max_dets_per_frame);
void cb_pool_destroy(cb_pool *pool);
const char *cb_pool_last_error(const cb_pool *pool);
int cb_pool_size(const cb_pool *pool);
cb_ctx *cb_pool_context(cb_pool *pool, int i);       /* the i-th GPU's context, e.g. for cb_set_params; owned by the pool */
int cb_pool_set_family_tag36h11(cb_pool *pool, int bits_corrected);
/* cb_detect_gray over every GPU of the pool: GPU g takes the contiguous frames [g*n/N, (g+1)*n/N) in batches of at most
 * max_batch through the streaming form (two batches in flight per GPU); out / out_counts are indexed by the frame's position in
 * `frames` exactly like cb_detect_gray, and cb_detection.frame is that position.  Use pinned frames (cb_host_alloc). */
int cb_pool_detect_gray(cb_pool *pool, const uint8_t *frames, int width, int height, int stride, size_t frame_stride,
                        int n_frames, cb_detection *out, int32_t *out_counts);
int cb_pool_get_timing(const cb_pool *pool, cb_pool_timing *t);

/* ---- plumbing ---- */
void *cb_host_alloc(size_t bytes);        /* pinned host memory for frames / outputs */
void cb_host_free(void *p);
void *cb_device_alloc(cb_ctx *ctx, size_t bytes);
void cb_device_free(cb_ctx *ctx, void *p);
int cb_memcpy_h2d(cb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int cb_memcpy_d2h(cb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int cb_device_count(void);
const char *cb_version(void);

#ifdef __cplusplus
}
#endif
#endif
