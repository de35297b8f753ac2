/*
 * oracle/oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load liboracle.so.  The product (chalkydri_b200/) never links or calls it.
 *
 * What it restates (all citations relative to /root/reference):
 *   - detector:  the UMich AprilTag-3 pipeline that crates/apriltags/src/lib.rs:258-261,301
 *                calls through the un-vendored `apriltag-sys` git dependency
 *                (crates/apriltags/Cargo.toml:10-11, branch master, NO pinned revision).
 *                Restated from the library's published algorithm (AprilTag 3.4.x sources,
 *                apriltag.c / apriltag_quad_thresh.c / tag36h11.c / homography.c / g2d.c).
 *                PARITY UNPINNED: the reference holds no golden vector or test for this path
 *                (SURVEY.md 8c); pins available offline are (i) the tag36h11 code table known
 *                answers and (ii) agreement with cv2.aruco on ids / corners of synthetic frames.
 *   - solver:    crates/chalkydri_sqpnp/src/lib.rs:42-479 on nalgebra 0.34.1-equivalent linear
 *                algebra restated here (Jacobi eigen / Jacobi SVD / partial-pivot LU).
 *                PARITY UNPINNED (no tests in the reference); pinned offline against
 *                ground-truth synthetic poses and cv2.solvePnP(SOLVEPNP_SQPNP).
 *   - CAT:       crates/chalkydri-apriltags/src/lib.rs:42-113,191-259,291-549 and utils.rs:1-46
 *                on statrs 0.18.0 quantile semantics restated here. PARITY UNPINNED.
 */
#ifndef CHALKYDRI_ORACLE_H
#define CHALKYDRI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float quad_decimate;        /* default 2.0 */
    int   refine_edges;         /* default 1   */
    double decode_sharpening;   /* default 0.25 */
    int   min_cluster_pixels;   /* 5  */
    int   max_nmaxima;          /* 10 */
    float critical_rad;         /* 10 deg in rad */
    float max_line_fit_mse;     /* 10 */
    int   min_white_black_diff; /* 5  */
    int   bits_corrected;       /* crates/apriltags/src/lib.rs:230 default 3 (1 with no config, :280) */
} orc_params;

typedef struct {
    int32_t id;
    int32_t hamming;
    float   decision_margin;
    float   pad_;
    double  H[9];
    double  c[2];
    double  p[4][2];
} orc_detection;

/* quad as produced by fit_quads() (decimated coordinates, before rescale / refine) */
typedef struct {
    float p[4][2];
    int32_t reversed_border;
    int32_t npoints;      /* cluster size that produced it */
    uint64_t cluster_id;  /* (rep_hi<<32)|rep_lo with min-index representatives */
} orc_quad;

/* optional taps; every pointer may be NULL */
typedef struct {
    uint8_t  *thresh;      /* [w*h]   ternary map of the decimated image (0/127/255)          */
    uint32_t *labels;      /* [w*h]   component label = smallest pixel index of the component
                                      (pixels never touched by the union-find are singletons) */
    uint32_t *comp_size;   /* [w*h]   size of the component the pixel belongs to (1 if untouched) */
    orc_quad *quads;       /* [quads_cap] */
    int32_t   quads_cap;
    int32_t   nquads;      /* out */
    int32_t   nclusters;   /* out: clusters with >= min_cluster_pixels points */
    int64_t   npoints;     /* out: total boundary points emitted */
    int32_t   w, h;        /* out: decimated size */
    /* boundary point dump (x,y,gx,gy as 4 x int16 + 64-bit min-rep cluster id), sorted by (id,y,x,gx,gy) */
    int16_t  *pts;         /* [pts_cap*4] */
    uint64_t *pts_cluster; /* [pts_cap]   */
    int64_t   pts_cap;
} orc_taps;

void orc_default_params(orc_params *p);
const uint64_t *orc_tag36h11_codes(int *ncodes);

/* decimated size for an input of W x H */
void orc_decimated_size(int W, int H, float quad_decimate, int *w, int *h);

/* A1+A2 only */
int orc_threshold(const uint8_t *im, int W, int H, int stride, const orc_params *prm, uint8_t *out /* [w*h] */);

/* whole detector on one gray frame; returns number of detections written (<= cap), <0 on error */
int orc_detect(const uint8_t *im, int W, int H, int stride, const orc_params *prm,
               orc_detection *out, int cap, orc_taps *taps);
/* CAT decode intent (book/src/maintenance/apriltags.md:58-60): `map` = ternary map 0 / 127 / 255 of the (decimated) frame, pitch =
   its width, used in place of upstream's threshold(); all later stages are upstream's */
int orc_detect_with_map(const uint8_t *im, int W, int H, int stride, const uint8_t *map, const orc_params *prm, orc_detection *out, int cap);

/* one frame per worker thread; counts[b] detections written at out[b*cap ...] */
int orc_detect_batch(const uint8_t *frames, int W, int H, int stride, int64_t frame_stride, int batch,
                     const orc_params *prm, orc_detection *out, int cap, int32_t *counts, int nthreads);

/* ---------------- SQPnP (crates/chalkydri_sqpnp/src/lib.rs) ---------------- */
typedef struct {
    double t[3];
    double q[4];   /* unit quaternion w,x,y,z (nalgebra Quaternion::new(w,i,j,k) order, field_layout.rs:35-36) */
} orc_iso3;

typedef struct {
    double rot[9];      /* robot rotation, column-major like nalgebra Matrix3 */
    double pos[3];
    double std_devs[3];
} orc_robot_pose;

/* solve_robot_pose (lib.rs:297-377). Returns 1 = Some, 0 = None. */
int orc_sqpnp_solve_robot_pose(const orc_iso3 *tags, int n_tags, const double *bearings /* [4*n_tags*3] */,
                               int n_bearings, const orc_iso3 *robot_to_cam, double gyro, double sign_change_error,
                               int max_iter, double tol_sq, orc_robot_pose *out);

/* batch with fixed max_tags stride: tags[i*max_tags..], bearings[i*max_tags*12..], n_tags[i] */
int orc_sqpnp_batch(const orc_iso3 *tags, const double *bearings, const int32_t *n_tags, int max_tags,
                    const orc_iso3 *robot_to_cam, const double *gyro, double sign_change_error,
                    int64_t n, orc_robot_pose *out, uint8_t *ok, int nthreads);

/* create_solver_camera_transform (lib.rs:430-461) */
void orc_create_solver_camera_transform(double fwd, double left, double up, double roll_deg, double pitch_deg,
                                        double yaw_deg, orc_iso3 *out);

/* stage taps for tests */
void orc_sqpnp_omega(const double *pts3d /* centred, [n*3] */, const double *bearings, int n,
                     double *omega /* 81 col-major */, double *q_tt_inv /* 9 */, double *q_rt /* 27 col-major 9x3 */);
void orc_sym_eigen9(const double *a /* 81 */, double *evals /* 9 */, double *evecs /* 81 col-major */);
void orc_nearest_so3(const double *r9, double *out9);

/* OpenCV-5 un-projection used between detector and solver (crates/apriltags/src/lib.rs:316-321;
 * camera-intrinsic-model `OpenCVModel5`, un-vendored). params = fx,fy,cx,cy,k1,k2,p1,p2,k3.
 * Returns 1 and a bearing (x,y,1) on success, 0 when the undistortion does not converge. */
int orc_unproject_opencv5(const double *params9, double u, double v, double *bearing3);

/* ---------------- CAT (crates/chalkydri-apriltags) ---------------- */
/* utils.rs:33-46 */
uint8_t orc_cat_grayscale(uint8_t r, uint8_t g, uint8_t b);
/* lib.rs:191-259: packed RGB -> Color map (0 Black, 1 White, 2 Other) */
void orc_cat_calc_otsu(const uint8_t *rgb, int w, int h, uint8_t *color);
/* lib.rs:319-334 */
void orc_cat_thresh(const uint8_t *rgb, int w, int h, uint8_t *color);
/* lib.rs:291-309,345-400: returns number of corners; xy as (x,y) int32 pairs in reference scan order */
int64_t orc_cat_detect_corners(const uint8_t *color, int w, int h, int32_t *xy, int64_t cap);
/* lib.rs:409-499: lines as (x1,y1,x2,y2) in reference order; out-of-image samples count as Other */
int64_t orc_cat_check_edges(const uint8_t *color, int w, int h, const int32_t *xy, int64_t npts,
                            int32_t *lines, int64_t cap);
/* lib.rs:501-549: labels = min pixel index per component, sizes = component size per pixel */
void orc_cat_connected_components(const uint8_t *color, int w, int h, uint32_t *labels, uint32_t *sizes);

#ifdef __cplusplus
}
#endif
#endif
