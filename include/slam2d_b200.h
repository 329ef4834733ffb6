/*
 * slam2d_b200.h -- C ABI of the B200-native scan-match / FastSLAM hot path.
 *
 * Drop-in boundary.  The reference has no FFI: its hot path sits behind plain
 * Python methods.  Each entry point below is what a ctypes binding of that method would call; the
 * Python facade in slam-2d-lidar-scan_b200/ does exactly that.
 *
 *   slam_matcher_create / slam_match_scan   <- ScanMatcher.matchScan          Utils/ScanMatcher_OGBased.py:47-79
 *                                              (frameSearchSpace :20-39, generateProbSearchSpace :41-45,
 *                                               covertMeasureToXY :81-89, searchToMatch :91-151)
 *   slam_motion_priors                      <- heading prior of searchToMatch  Utils/ScanMatcher_OGBased.py:104-110
 *   slam_grid_init / slam_update_grid       <- OccupancyGrid.__init__ counts / updateOccupancyGrid
 *                                                                              Utils/OccupancyGrid.py:13-14,127-152
 *   slam_propose_poses / slam_finish_step   <- Particle.updateEstimatedPose / getMovingTheta / weight *= confidence
 *                                                                              Algorithm/FastSlam.py:77-135
 *   slam_normalize_weights                  <- ParticleFilter.normalizeWeights + weightUnbalanced  FastSlam.py:30-48
 *   slam_resample_indices                   <- np.random.choice(arange(N), N, p=w)                 FastSlam.py:59
 *   slam_gather_particles                   <- the deepcopy loop of ParticleFilter.resample        FastSlam.py:60-62
 *
 * Conventions: every pointer named d_* is DEVICE memory, h_* is HOST memory; no torch types.  All functions
 * return 0 on success or a non-zero code (CUDA error code, or SLAM_E_*); slam_last_error() describes the
 * last failure on the calling thread.  Kernels are enqueued on `stream` (a cudaStream_t passed as void*);
 * nothing synchronises unless stated.  Floating point on the path is IEEE float64 without FMA contraction.
 *
 * Map storage: one lattice per particle, row-major [N][G][pitch] cells of {float visited, float total}
 * (counts are integer-valued; float32 is exact to 2^24).  pitch = G rounded up to a multiple of 4 cells so
 * rows are 32-byte aligned (TMA needs 16-byte strides).
 */
#ifndef SLAM2D_B200_H
#define SLAM2D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAM_MAX_BLUR_RADIUS 8
#define SLAM_MAX_BEAMS 512

/* error codes (besides cudaError_t values, which are < 1000) */
#define SLAM_E_BADARG 1001
#define SLAM_E_UNSUPPORTED 1002
#define SLAM_E_NOMEM 1003

/* per-particle status bits written by slam_match_scan / slam_update_grid (0 = ok) */
#define SLAM_ST_WINDOW_OUTSIDE_MAP 1 /* search window left the pre-sized lattice (reference would expand)  */
#define SLAM_ST_INDEX_OUT_OF_FIELD 2 /* a gather fell outside probSP (numpy would raise IndexError / wrap) */
#define SLAM_ST_NAN_SCORE 4          /* NaN in the score volume (np.random.choice would raise ValueError)  */
#define SLAM_ST_SCAN_OUTSIDE_MAP 8   /* map update touched cells outside the lattice                       */
#define SLAM_ST_HEADING_MISSING 16   /* prevMatchedMovingTheta is None where the reference does None+float */

/* Lattice + lidar geometry, shared by all particles.  Built on the host with the reference's own
 * expressions (OccupancyGrid.py:8-45) and uploaded once. */
typedef struct {
  int32_t G;               /* lattice side = int(mapLength/unit)+1                           */
  int32_t pitch;           /* row pitch in cells (multiple of 4, >= G)                       */
  int32_t K;               /* numSamplesPerRev                                               */
  int32_t L;               /* lidar-local patch side = 2*int(maxRange/unit)+1                */
  int32_t numSpokes;       /* int(rint(2*pi/angularStep))                                    */
  int32_t spokesStartIdx;  /* OccupancyGrid.py:30                                            */
  double unit;             /* unitGridSize                                                   */
  double mapX0, mapX1;     /* mapXLim                                                        */
  double mapY0, mapY1;     /* mapYLim                                                        */
  double fovHalf;          /* lidarFOV / 2                                                   */
  double maxRange;         /* lidarMaxRange                                                  */
  double wallHalf;         /* wallThickness / 2                                              */
  const double* d_gridX;   /* [G]  OccupancyGridX[0, :]                                      */
  const double* d_gridY;   /* [G]  OccupancyGridY[:, 0]                                      */
  const int16_t* d_sector; /* [L*L] bearing sector of each local cell (OccupancyGrid.py:39-43) */
  const double* d_radius;  /* [L*L] sqrt(x^2+y^2) of each local cell (OccupancyGrid.py:44)   */
  const double* d_localAxis; /* [L] linspace(-maxRange, maxRange, L)                         */
} slam_geometry;

/* One stage (coarse or fine) of the correlative search.  All values are computed on the host with the
 * reference's Python expressions so that device constants are bit-identical to the oracle's. */
typedef struct {
  double unitLength;       /* coarseFactor*unit, or unit                                     */
  double logMiss;          /* math.log(missMatchProb of this stage)                          */
  int32_t blurRadius;      /* int(4*sigma + 0.5), <= SLAM_MAX_BLUR_RADIUS                    */
  double blurW[2 * SLAM_MAX_BLUR_RADIUS + 1]; /* scipy gaussian taps, blurW[blurRadius] = centre */
  int32_t nHalf;           /* int(searchRadius/unitLength): offsets -nHalf..nHalf            */
  int32_t nTheta;          /* len(thetaRange)                                                */
  const double* h_thetas;  /* [nTheta] np.arange(-h, h+angStep, angStep)                     */
  const double* h_cos;     /* [nTheta] np.cos(thetas)                                        */
  const double* h_sin;     /* [nTheta] np.sin(thetas)                                        */
} slam_stage_desc;

typedef struct {
  double windowRadius;     /* 1.1*lidarMaxRange + searchRadius (ScanMatcher_OGBased.py:21)   */
  slam_stage_desc coarse;
  slam_stage_desc fine;
} slam_matcher_desc;

typedef struct slam_matcher slam_matcher; /* opaque: device tables + memory plan */

const char* slam_last_error(void);
int slam_version(void);

/* Copies the geometry/stage tables it needs to the device (small allocations owned by the handle). */
int slam_matcher_create(const slam_geometry* geom, const slam_matcher_desc* desc, slam_matcher** out);
void slam_matcher_destroy(slam_matcher* m);
/* Device scratch the caller must provide to slam_match_scan (likelihood fields in flight; L2-resident). */
size_t slam_matcher_workspace_bytes(const slam_matcher* m);
/* The same for a call with N particles: one slot per CTA that gets a particle (a standalone ScanMatcher, N = 1, needs a
 * single slot instead of one per SM). */
size_t slam_matcher_workspace_bytes_n(const slam_matcher* m, int32_t N);
/* Side of the largest field / number of hypotheses per stage (stage 0 = coarse, 1 = fine). */
int slam_matcher_field_side(const slam_matcher* m, int stage);
int slam_matcher_num_poses(const slam_matcher* m, int stage);

/* Optional per-stage dumps for parity tests; any pointer may be NULL. */
typedef struct {
  double* d_prob[2];      /* [N][side*side]  probSP after the clamp, row pitch = side            */
  int32_t* d_probDims[2]; /* [N][2]          (rows, cols) actually used                          */
  double* d_vol[2];       /* [N][nPoses]     convTotal, C order (theta, dy, dx)                  */
} slam_match_debug;

/*
 * ScanMatcher.matchScan for a batch of particles, count >= 2 (ScanMatcher_OGBased.py:53-79), fused:
 * window threshold + scatter, separable blur, min/clamp, beam projection, per-theta rotate/index/unique,
 * correlation, priors, argmax or softmax-CDF sampling, confidence -- coarse then fine.
 *
 *  d_grid      [N][G][pitch][2] float  (visited, total)
 *  d_ranges    [K] double               the scan (shared by all particles)
 *  d_estPose   [N][3] double            (x, y, theta) proposals
 *  d_rv        [nOffC*nOffC] double     radial motion prior of the coarse stage (shared; host-built, :101-103)
 *  d_tw        [N][nOffC*nOffC] double  heading prior per particle, or NULL for zeros (:104-110)
 *  d_uniforms  [N] double or NULL       NULL -> matchMax=True (argmax); else the double np.random.choice draws
 *  d_outPose   [N][3] double            fine-stage matched pose
 *  d_outConf   [N] double               coarse confidence = sum(exp(convTotal))
 *  d_outIdx    [N][6] int32             (itheta, iy, ix) coarse then fine
 *  d_status    [N] int32                SLAM_ST_* bits, OR-ed into the existing words (never cleared here)
 */
int slam_match_scan(slam_matcher* m, const float* d_grid, int32_t N, const double* d_ranges,
                    const double* d_estPose, const double* d_rv, const double* d_tw,
                    const double* d_uniforms, double* d_outPose, double* d_outConf, int32_t* d_outIdx,
                    int32_t* d_status, void* d_workspace, size_t workspaceBytes,
                    const slam_match_debug* debug, void* stream);

/* The same with a slot table: particle p reads lattice d_slots[p] of the numLattices lattices at d_grid (NULL =
 * identity).  Everything else (poses, priors, outputs, status) stays indexed by particle. */
int slam_match_scan_slots(slam_matcher* m, const float* d_grid, const int32_t* d_slots, int32_t numLattices, int32_t N,
                          const double* d_ranges, const double* d_estPose, const double* d_rv, const double* d_tw,
                          const double* d_uniforms, double* d_outPose, double* d_outConf, int32_t* d_outIdx,
                          int32_t* d_status, void* d_workspace, size_t workspaceBytes, const slam_match_debug* debug,
                          void* stream);

/*
 * Stage-level entry points: the reference's public stage methods, for callers that drive the stages themselves
 * (slam_match_scan fuses them and keeps the fields on chip).
 *
 * slam_field_build   <- ScanMatcher.frameSearchSpace (:20-39) incl. generateProbSearchSpace (:41-45): the clamped
 *   likelihood field of `stage` (0 coarse, 1 fine) around d_centre[p] = (x, y, -); d_prob [N][side*side] with row pitch
 *   side = slam_matcher_field_side(stage), d_probDims [N][2] = (rows, cols) used.  xRangeList / yRangeList are
 *   centre -+ windowRadius (host arithmetic).
 * slam_correlate     <- ScanMatcher.searchToMatch (:91-151) against a caller-provided field of the same layout:
 *   d_centre [N][3] pose, d_origin [N][2] = (xRangeList[0], yRangeList[0]), d_rv / d_tw priors (NULL = zeros, the
 *   fineSearch=True case), d_uniforms NULL -> matchMax=True else the double np.random.choice draws;
 *   d_vol [N][nPoses] convTotal, d_outIdx [N][3] (itheta, iy, ix), d_outConf [N] = sum(exp(convTotal)).
 *   Workspace: min(N, SMs) * nPoses * 8 bytes (+512).
 * slam_blur_clamp    <- ScanMatcher.generateProbSearchSpace (:41-45) on an arbitrary float64 array: scipy's
 *   gaussian_filter (taps d_taps[2*radius+1], reflect borders, axis 0 then 1), min, clamp.  d_tmp: rows*cols doubles.
 */
int slam_field_build(slam_matcher* m, int32_t stage, const float* d_grid, int32_t N, const double* d_centre,
                     double* d_prob, int32_t* d_probDims, int32_t* d_status, void* d_workspace, size_t workspaceBytes,
                     void* stream);
int slam_correlate(slam_matcher* m, int32_t stage, int32_t N, const double* d_prob, const int32_t* d_probDims,
                   const double* d_ranges, const double* d_centre, const double* d_origin, const double* d_rv,
                   const double* d_tw, const double* d_uniforms, double* d_vol, int32_t* d_outIdx, double* d_outConf,
                   int32_t* d_status, void* d_workspace, size_t workspaceBytes, void* stream);
int slam_blur_clamp(const double* d_in, int32_t rows, int32_t cols, const double* d_taps, int32_t radius, double* d_tmp,
                    double* d_out, void* stream);

/* Heading prior thetaWeight (ScanMatcher_OGBased.py:105-108) for N particles:
 * tw[p][a][b] = coef * acos((xv*cos(phi_p) + yv*sin(phi_p)) / dist)^2, zeros where hasPhi[p] == 0.
 * coef = -1 / (2*turnSigma**2) evaluated by the host. */
int slam_motion_priors(int32_t N, int32_t nHalf, double coef, const double* d_phi, const int32_t* d_hasPhi,
                       double* d_tw, void* stream);

/* (visited, total) := (1, 2) for N lattices (OccupancyGrid.py:13-14). */
int slam_grid_init(const slam_geometry* geom, float* d_grid, int32_t N, void* stream);

/* OccupancyGrid.updateOccupancyGrid for N particles (OccupancyGrid.py:127-152).  Owner computes: every map cell of a
 * lattice is updated by exactly one thread (float reductions at the L2 -- the counts are small integers, so the adds are
 * exact and order-independent), reproducing numpy's once-per-statement fancy `+=` even when two lidar-local cells round
 * to the same map cell (poses half a cell off the lattice take a map-cell-owned path inside the same launch).  Particles
 * are grouped by their sector shift first, so the empty / hit classification of a local cell is evaluated once per
 * group.  numSpokes <= 8192.  d_pose [N][3]; d_status |= SLAM_ST_SCAN_OUTSIDE_MAP (cells outside the lattice are skipped). */
int slam_update_grid(const slam_geometry* geom, float* d_grid, int32_t N, const double* d_ranges,
                     const double* d_pose, int32_t* d_status, void* d_workspace, size_t workspaceBytes, void* stream);
/* Device scratch slam_update_grid needs for N particles (per call; may be shared by calls on one stream). */
size_t slam_update_workspace_bytes(int32_t N);
/* The same with a slot table: particle p lives in lattice d_slots[p] of d_grid (NULL = identity).  Copy-elided
 * resampling (slam_copy_lattices) permutes ownership of the lattices instead of moving them. */
int slam_update_grid_slots(const slam_geometry* geom, float* d_grid, const int32_t* d_slots, int32_t N,
                           const double* d_ranges, const double* d_pose, int32_t* d_status, void* d_workspace,
                           size_t workspaceBytes, void* stream);

/* Per-particle odometry proposal (FastSlam.py:77-106).  The raw-odometry part is identical for every
 * particle and is computed by the host: estTheta = (prevMatched.theta + rawTheta) - prevRawTheta (left to
 * right, :78); position = previous matched position.  mode 0: estMovingTheta is None for all particles;
 * mode 1: estMovingTheta = prevMatchedMovingTheta + rawTurn (rawTurn = rawMovingTheta - prevRawMovingTheta).
 *  d_prevMatched [N][3], d_prevHeading [N] + d_hasHeading [N]  ->  d_estPose [N][3], d_phi [N], d_hasPhi [N]
 *  In mode 1 a particle without a previous heading gets SLAM_ST_HEADING_MISSING (the reference raises
 *  TypeError on None + float there). */
int slam_propose_poses(int32_t N, const double* d_prevMatched, double rawTheta, double prevRawTheta, int32_t mode,
                       double rawTurn, const double* d_prevHeading, const int32_t* d_hasHeading, double* d_estPose,
                       double* d_phi, int32_t* d_hasPhi, int32_t* d_status, void* stream);

/* After matching (FastSlam.py:130-135): heading from the last trajectory point (getMovingTheta), then
 * prevMatched := matched, weight *= confidence.  d_prevMatched holds the last trajectory point on entry. */
int slam_finish_step(int32_t N, const double* d_matched, const double* d_conf, double* d_prevMatched,
                     double* d_prevHeading, int32_t* d_hasHeading, double* d_weights, void* stream);

/* normalizeWeights + weightUnbalanced (FastSlam.py:30-48) with the reference's sequential float64 order.
 * d_out[0] = variance, d_out[1] = 1.0 if the trigger fires else 0.0. */
int slam_normalize_weights(int32_t N, double* d_weights, double* d_out, void* stream);

/* The whole end-of-step trigger in one launch (FastSlam.py:30-48 + the status check): d_weightsOut[i] =
 * d_weightsIn[i] / sum (sequential float64 order; the two may alias), d_out[0] = variance, d_out[1] = trigger, and the
 * low 32 bits of d_out[2] = bitwise OR of the nStatus status words (untouched when nStatus == 0). */
int slam_step_trigger(int32_t N, const double* d_weightsIn, double* d_weightsOut, const int32_t* d_status,
                      int32_t nStatus, double* d_out, void* stream);

/* d_out[0] = bitwise OR of the N per-particle status words.  Status words are sticky: every kernel ORs its bits
 * in, nothing clears them but the caller (numpy would have raised at the first one: ScanMatcher_OGBased.py:125-138). */
int slam_status_reduce(int32_t N, const int32_t* d_status, int32_t* d_out, void* stream);

/* np.random.choice(arange(N), N, p=w) given the N uniforms it would draw (legacy RandomState):
 * cdf = cumsum(w) sequential; cdf /= cdf[-1]; idx = searchsorted(cdf, u, side='right').
 * d_cdfScratch [N] double. */
int slam_resample_indices(int32_t N, const double* d_weights, const double* d_uniforms, double* d_cdfScratch,
                          int32_t* d_idx, void* stream);

/* In-place lattice copies of a copy-elided resample (FastSlam.py:50-62 deep-copies every chosen particle; here only
 * the extra copies of multiply-chosen particles move): lattice d_dst[c] := lattice d_src[c], c < nCopies.  Sources
 * and destinations are disjoint sets of lattices (a source survives under its own slot). */
int slam_copy_lattices(const slam_geometry* geom, float* d_grid, int32_t nCopies, const int32_t* d_src,
                       const int32_t* d_dst, void* stream);

/* dst particle i := src particle idx[i] for the lattices and the per-particle state rows
 * (deepcopy loop, FastSlam.py:60-62); weights := 1/N.  src and dst must not alias. */
int slam_gather_particles(const slam_geometry* geom, int32_t N, const int32_t* d_idx, const float* d_gridSrc,
                          float* d_gridDst, const double* d_stateSrc, double* d_stateDst, int32_t stateCols,
                          double* d_weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLAM2D_B200_H */
