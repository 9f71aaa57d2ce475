/*
 * artist_b200 - C ABI of the B200-native heliostat ray-tracing hot path.
 *
 * The reference (ARTIST v2.0.0) is pure Python/PyTorch and has no FFI layer; its boundary for
 * this path is the Python class API.  This header is the boundary a maintainer
 * would bind instead of the eager-op bodies of the functions cited at each entry point.  Rules:
 *   - plain pointers and sizes only (no torch types); all tensors are contiguous fp32 / int32
 *     DEVICE buffers unless an entry point says "host";
 *   - the library never allocates or frees caller-visible memory: the caller (torch) owns
 *     every buffer, outputs are written in place;
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*); calls are
 *     re-entrant for distinct streams; no global mutable state;
 *   - return value: AB200_OK (0) or a negative AB200_E* code; ab200_error_string() explains.
 *
 * Layout vocabulary (as in the reference): N = active heliostat-samples, P = surface points
 * per heliostat (all facets), R = rays per point, bitmap = [U rows, E columns], T = target
 * areas (planar first, then cylindrical - artist/field/solar_tower.py:80-90).
 */
#ifndef ARTIST_B200_H
#define ARTIST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB200_OK 0
#define AB200_EINVAL (-1)   /* bad argument (null pointer, size <= 0, unsupported degree ...) */
#define AB200_ECUDA (-2)    /* a CUDA runtime call or launch failed */
#define AB200_ELIMIT (-3)   /* size beyond what the kernels support */

#define AB200_ABI_VERSION 4

/* trig source for the per-ray scatter rotation (artist/geometry/transforms.py:52-55) */
#define AB200_TRIG_SINCOSF 0 /* libdevice sincosf (default) */
#define AB200_TRIG_TABLE 1   /* strict parity: consume caller-computed cos/sin (args->trig) */
#define AB200_TRIG_POLY 2    /* small-angle polynomial (<= 1 ulp for |x| <= pi/4), sincosf otherwise; for |x| <= 0.01 (any physical
                              * sun shape) the cosine is torch's CPU cos incl. its rounding bias, so that no ray changes pixel
                              * against the reference (csrc/common.cuh; -DAB200_COS_CORRECTLY_ROUNDED: the correctly rounded one) */

/* flags */
#define AB200_FLAG_FP32_ACCUM 1 /* accumulate the bitmap with fp32 shared-memory atomics instead of the
                                   deterministic fixed-point histogram (order-dependent rounding) */
#define AB200_FLAG_ONE_CTA_PER_SAMPLE 2 /* run the one-CTA-per-sample kernels (normally chosen when 2 x n_local >= SM count)
                                           whatever the sample count: lets a small case exercise the kernels a full field
                                           runs (parity tests, bench.py's parity leg) */

/* Target-area SoA tensors: artist/field/tower_target_areas_planar.py:45-74 and
 * tower_target_areas_cylindrical.py:56-102. */
typedef struct ab200_targets {
    int32_t n_planar;
    int32_t n_cyl;
    const float* planar_centers; /* [n_planar,4] */
    const float* planar_normals; /* [n_planar,4] */
    const float* planar_dims;    /* [n_planar,2] (width e, height u) */
    const float* cyl_centers;    /* [n_cyl,4] */
    const float* cyl_normals;    /* [n_cyl,4] */
    const float* cyl_axes;       /* [n_cyl,4] */
    const float* cyl_radii;      /* [n_cyl] */
    const float* cyl_heights;    /* [n_cyl] */
    const float* cyl_opening;    /* [n_cyl] opening angles (rad) */
} ab200_targets;

/* Blocking (artist/raytracing/blocking.py): rectangle primitives of ALL heliostats of the scenario, packed by
 * ab200_blocking_pack, and per active sample the ordered candidate list built by ab200_blocking_candidates (it
 * replaces the per-batch LBVH filter :832-995).  n_blockers == 0 switches blocking off. */
typedef struct ab200_blockers {
    int32_t n_blockers;            /* H */
    int32_t max_candidates;        /* row length of cand_idx (<= 64) */
    const float* prims;            /* [H,16]: corner0(3) span_u(3) span_v(3) normal(3) |u|^2 |v|^2 u.v det_safe */
    const int32_t* cand_idx;       /* [N,max_candidates] candidate primitive rows per active sample */
    const int32_t* cand_count;     /* [N] */
    float softness;                /* 1000 */
    float alpha;                   /* 100 */
    float ray_origin_offset;       /* 0.05 */
    float epsilon;                 /* 1e-12 */
    float cull_angle;              /* bound on the scatter angle for the per-point candidate cull (rad); <= 0 disables it */
} ab200_blockers;

/*
 * ab200_trace_fwd - replaces the body of HeliostatRayTracer.trace_rays
 * (artist/raytracing/heliostat_ray_tracer.py:285-508): reflect (geometry.py:32-41), scatter
 * (transforms.py:52-83 + heliostat_ray_tracer.py:547-552), line-plane / line-cylinder
 * intersection (geometry.py:100-204 / 287-445), intensities (:482-487), bilinear splat
 * (:648-778) and the three factor outputs (:498-506), fused in one pass; no per-ray tensor
 * is materialised.  Rows of `flux`/factors that are not in `local_rows` are zero-filled
 * (the reference leaves them uninitialised - SURVEY.md 0.7).
 */
typedef struct ab200_trace_args {
    int32_t abi_version;     /* AB200_ABI_VERSION */
    int32_t n_samples;       /* N */
    int32_t n_points;        /* P */
    int32_t n_rays;          /* R */
    int32_t res_e;           /* E = bitmap_resolution[0] */
    int32_t res_u;           /* U = bitmap_resolution[1] */
    int32_t n_local;         /* rows traced by this rank */
    const int32_t* local_rows; /* [n_local] sample indices (sampler contract sampling.py:129-146); NULL = 0..N-1 */
    const float* points;     /* [N,P,4] active_surface_points (aligned) */
    const float* normals;    /* [N,P,4] active_surface_normals (aligned) */
    const float* incident;   /* [N,4] incident_ray_directions */
    const float* distortions;/* [N,R,P,2] (u,e) interleaved, exactly Sun.get_distortions' sample layout */
    const float* trig;       /* AB200_TRIG_TABLE only: [N,R,P,4] = cos u, sin u, cos e, sin e */
    const int32_t* target_idx; /* [N] target_area_indices */
    ab200_targets targets;
    ab200_blockers blockers;
    float ray_magnitude;         /* HeliostatRayTracer.ray_magnitude (:185-203) */
    float one_minus_extinction;  /* (float)(1 - ray_extinction_factor) */
    float reflectivity;          /* mirror_reflectivity */
    float scatter_sigma;         /* sqrt(sun covariance): sizes the shared-memory bitmap window; 0 = unknown */
    int32_t trig_mode;           /* AB200_TRIG_* */
    int32_t flags;               /* AB200_FLAG_* */
    float* flux;             /* out [N,U,E] */
    float* intercept;        /* out [N] */
    float* on_target;        /* out [N] */
    float* blocking;         /* out [N] */
    /* optional per-ray parity probes (NULL to skip): the values line_plane/cylinder_intersections return */
    float* dbg_be;           /* [N,R,P] */
    float* dbg_bu;           /* [N,R,P] */
    float* dbg_t;            /* [N,R,P] */
    float* dbg_lambert;      /* [N,R,P] */
    int64_t* stats;          /* optional diagnostics (NULL to skip), int64[20], accumulated: [0] threads that used the global
                                fallback path, [1] sum of window cells, [2] CTAs, [4..10] cycles of the forward CTA phases
                                (start-up, placement, clearing, ray loop, wait, flush, tap conversion), [12..16] of the
                                backward's (start-up, placement, staging, ray loop, reduction) - thread 0, summed over CTAs */
    const float* orientations; /* optional [N,4,4] (NULL = `points`/`normals` are already aligned): fuses
                                HeliostatGroupRigidBody.align_surfaces_with_* (heliostat_group_rigid_body.py:217-222,
                                265-270) into the trace - `points`/`normals` are then the UN-aligned active surface
                                rows and every CTA applies `row @ O^T` itself (same FMA chain as ab200_align_fwd), so
                                the aligned [N,P,4] tensors are never written to or re-read from HBM */
    int32_t* windows;        /* optional scratch [N,4] (NULL to skip): ab200_trace_fwd records the shared-memory bitmap window
                                (e0, u0, width, height) it placed for every sample of the one-CTA-per-sample mode,
                                ab200_trace_bwd of the same call re-uses it instead of sampling the surface and placing
                                the window again.  Correctness never depends on it (it only selects the fast path). */
    const float* distortions_planar; /* optional [2,N,R,P] (NULL to skip): the SAME samples as `distortions`, de-interleaved
                                (plane 0 = u, plane 1 = e).  The benchmark-shaped kernels (trace_v3.cuh) trace two adjacent
                                surface points per thread in packed fp32x2 registers and read each plane with one 8-byte
                                load that IS the register pair (the interleaved layout costs two register moves per ray);
                                ab200_sample_distortions writes both layouts in one pass.  Without it the general kernels run. */
    const int32_t* src_rows; /* optional [N] (NULL = identity): activation index map.  Sample h reads its surface from row
                                src_rows[h] of `points` / `normals`, which then hold the group's UN-replicated [Nh,P,4]
                                surfaces - replaces the repeat_interleave copies of HeliostatGroup.activate_heliostats
                                (artist/field/heliostat_group.py:256-315; 640 MB at config-3 size).  Everything else
                                (incident, target_idx, distortions, orientations, outputs, gradient rows) stays per sample;
                                ab200_replica_sum folds the per-sample gradient rows back onto the source rows. */
} ab200_trace_args;

int32_t ab200_trace_fwd(const ab200_trace_args* args, void* stream);

/* out[s,:] = sum of in[k,:] for k in [row_start[s], row_start[s+1]) - the backward of the activation index map (replicas of a
 * heliostat are contiguous samples); fixed summation order.  in [N,row_elems], row_start [n_src+1], out [n_src,row_elems]. */
int32_t ab200_replica_sum(const float* in, const int32_t* row_start, int32_t n_src, int64_t row_elems, float* out, void* stream);

/*
 * ab200_trace_bwd - explicit backward of ab200_trace_fwd (replaces the autograd graph the
 * reference retains, ~0.6 KB/ray): re-computes each ray, gathers the four bitmap-gradient taps
 * and accumulates d/d(points) and d/d(normals) in registers over the R rays of a point.
 * Gradients flow through be, bu and the intensity, not through trunc(), masks or counts
 * (SURVEY.md Appendix A).
 */
typedef struct ab200_trace_bwd_args {
    ab200_trace_args fwd;     /* same inputs as the forward call (outputs/dbg pointers ignored) */
    const float* grad_flux;   /* [N,U,E]; row n starts at grad_flux + n * grad_flux_stride */
    int64_t grad_flux_stride; /* in floats: U*E (or -1) for a dense tensor, 0 when every sample shares one [U,E] gradient
                                 (e.g. the loss is taken on the per-target sum) - no [N,U,E] expansion is materialised */
    float* grad_points;       /* out [N,P,4] (w component 0); grad_points and grad_normals may BOTH be NULL when only
                                 grad_orientations / grad_prims are wanted (motor-position optimisation): the kernel then
                                 skips the 64 B per point of gradient stores */
    float* grad_normals;      /* out [N,P,4] (w component 0) */
    float* grad_prims;        /* blocking only, may be NULL: [H,12] d/d(corner0, span_u, span_v, normal).  With
                                 grad_prims_scratch: OVERWRITTEN with sums taken in a fixed order (every CTA adds the few rays
                                 inside a sigmoid transition into shared-memory double accumulators, writes one row per
                                 candidate slot to the scratch, a second kernel sums the rows per primitive).  Without:
                                 accumulated with float atomics (caller zeroes it). */
    float* grad_orientations; /* fwd.orientations only, may be NULL: [N,4,4] dL/dO (caller zeroes it; row 3 stays 0).
                                 With fwd.orientations set, grad_points / grad_normals are w.r.t. the UN-aligned rows. */
    float* grad_prims_scratch;        /* device scratch for grad_prims, may be NULL: >= (N + 1024) * max_candidates * 12 floats */
    int64_t grad_prims_scratch_floats; /* its capacity in floats (too small: the atomics path runs) */
} ab200_trace_bwd_args;

int32_t ab200_trace_bwd(const ab200_trace_bwd_args* args, void* stream);

/*
 * ab200_bitmaps_per_target - HeliostatRayTracer.get_bitmaps_per_target
 * (heliostat_ray_tracer.py:593-608): out[t] = sum of the bitmaps of samples with target_idx == t,
 * summed in a fixed order (8 interleaved partial sums, then combined; deterministic).  `out` is [T,U,E].
 */
int32_t ab200_bitmaps_per_target(const float* bitmaps, const int32_t* target_idx, int32_t n_samples,
                                 int32_t n_targets, int32_t res_u, int32_t res_e, float* out, void* stream);

/*
 * ab200_nurbs_fwd / _bwd - NURBSSurfaces.calculate_surface_points_and_normals
 * (artist/nurbs/surfaces.py:475-689): span lookup (:198-207), basis + first derivative
 * (:325-415), tensor-product contraction (:592-613), normal = normalise(dS/du x dS/dv)
 * (:615-661), optional canting + facet translation (geometry/transforms.py:321-347,
 * surfaces.py:674-687).
 */
typedef struct ab200_nurbs_args {
    int32_t abi_version;
    int32_t n_surfaces;       /* N */
    int32_t n_facets;         /* F */
    int32_t n_eval;           /* evaluation points per facet */
    int32_t n_ctrl_u, n_ctrl_v;
    int32_t degree_u, degree_v;  /* 1..3 */
    const float* control_points; /* [N,F,cu,cv,3] */
    const float* eval_points;    /* (u,v) pairs; element (n,f,k) at eval_points + n*eval_stride_n + f*eval_stride_f + 2k */
    int64_t eval_stride_n;       /* in floats; 0 = shared by all surfaces */
    int64_t eval_stride_f;       /* in floats; 0 = shared by all facets */
    const float* knots_u;        /* [cu+degree_u+1] */
    const float* knots_v;        /* [cv+degree_v+1] */
    const float* canting;        /* [N,F,2,4] or NULL */
    const float* facet_translations; /* [N,F,4] or NULL (must be NULL iff canting is NULL) */
    int32_t grid_u, grid_v;      /* > 0: the evaluation points are a sorted cartesian grid u_i x v_j with v fastest
                                    (nurbs/utils.py:7-49) of grid_u x grid_v = n_eval points - enables the separable
                                    backward; 0 = arbitrary points */
    float* points;               /* out [N,F,n_eval,4] */
    float* normals;              /* out [N,F,n_eval,4] */
} ab200_nurbs_args;

int32_t ab200_nurbs_fwd(const ab200_nurbs_args* args, void* stream);

typedef struct ab200_nurbs_bwd_args {
    ab200_nurbs_args fwd;
    const float* grad_points;    /* [N,F,n_eval,4] */
    const float* grad_normals;   /* [N,F,n_eval,4] */
    float* grad_control_points;  /* out [N,F,cu,cv,3] (overwritten) */
} ab200_nurbs_bwd_args;

int32_t ab200_nurbs_bwd(const ab200_nurbs_bwd_args* args, void* stream);

/*
 * ab200_kinematics_* - RigidBody forward kinematics with linear / ideal actuators
 * (artist/field/kinematics_rigid_body.py:194-324,510-538; actuators_linear.py:79-291;
 * actuators_ideal.py:66-111) and the <=4-iteration alignment to incident ray directions
 * (:540-634 with the closed-form inverse kinematics :326-508).
 */
typedef struct ab200_kinematics_args {
    int32_t abi_version;
    int32_t n;                       /* active heliostat-samples */
    int32_t linear_actuators;        /* 1 = LinearActuators, 0 = IdealActuators */
    const float* positions;          /* [n,4] */
    const float* translation_dev;    /* [n,9] */
    const float* rotation_dev;       /* [n,4] */
    const float* actuator_non_opt;   /* [n,7,2] */
    const float* actuator_opt;       /* [n,2,2] (ignored for ideal actuators) */
    const float* orientation_offset; /* [4,4] RigidBody.initial_orientation_offsets (:186-190) */
} ab200_kinematics_args;

/* motor_positions [n,2] -> orientations [n,4,4] (motor_positions_to_orientations, :510-538) */
int32_t ab200_kinematics_fwd(const ab200_kinematics_args* args, const float* motor_positions,
                             float* orientations, void* stream);

/* backward of the above: grad_orientations [n,4,4] -> grads (any output pointer may be NULL) */
int32_t ab200_kinematics_bwd(const ab200_kinematics_args* args, const float* motor_positions,
                             const float* grad_orientations, float* grad_motor_positions /* [n,2] */,
                             float* grad_rotation_dev /* [n,4] */, float* grad_translation_dev /* [n,9] */,
                             float* grad_actuator_opt /* [n,2,2] */, float* grad_positions /* [n,4] */,
                             void* stream);

/* incident_ray_directions_to_orientations (:540-634): writes orientations [n,4,4] and the final
 * motor positions [n,2].  The reference stops iterating when ALL heliostats converged; the same
 * batch-wide rule is applied on the device (no host sync); max_iterations <= 6. */
int32_t ab200_kinematics_align_incident(const ab200_kinematics_args* args, const float* incident /* [n,4] */,
                                        const float* aim_points /* [n,4] */, int32_t max_iterations,
                                        float min_eps, float* orientations, float* motor_positions,
                                        float* scratch /* [4n + 8] floats of working storage */, void* stream);

/*
 * ab200_align_fwd / _bwd - HeliostatGroupRigidBody.align_surfaces_with_* tail
 * (artist/field/heliostat_group_rigid_body.py:217-222,265-270): points @ O^T, normals @ O^T for
 * row-vector [N,P,4] data.  `src_row` (NULL = identity) maps an active sample to its source
 * heliostat row, which folds HeliostatGroup.activate_heliostats' repeat_interleave
 * (heliostat_group.py:258-263) into the same pass.
 */
int32_t ab200_align_fwd(const float* points, const float* normals, const float* orientations,
                        const int32_t* src_row, int32_t n_samples, int32_t n_points,
                        float* out_points, float* out_normals, void* stream);

/* grads w.r.t. the un-aligned points/normals (per active sample, [N,P,4]) and w.r.t. the
 * orientations ([N,4,4]); any output pointer may be NULL. */
int32_t ab200_align_bwd(const float* points, const float* normals, const float* orientations,
                        const int32_t* src_row, int32_t n_samples, int32_t n_points,
                        const float* grad_out_points, const float* grad_out_normals,
                        float* grad_points, float* grad_normals, float* grad_orientations, void* stream);

/*
 * Stand-alone forms of the steps the reference exports as free functions (artist/raytracing/__init__.py:1-12) on the
 * MATERIALISED per-ray tensors of its API; trace_rays never uses them (the fused kernels keep per-ray state in
 * registers), they serve API parity and step-wise checks.  Forward only.
 *   ab200_reflect              geometry.reflect, geometry.py:11-41: out = i - 2 (i.n) n over all 4 components;
 *                              incident [N,4] (one direction per sample), normals / out [N,P,4]
 *   ab200_scatter_rays         HeliostatRayTracer.scatter_rays, heliostat_ray_tracer.py:510-561 + transforms.
 *                              rotate_distortions :7-83: distortions u, e [N,R,P], reflected directions [N,P,4] ->
 *                              scattered directions [N,R,P,4]
 *   ab200_line_intersections   geometry.line_plane_intersections :44-204 (cylindrical = 0) and
 *                              line_cylinder_intersections :207-445 (cylindrical = 1; target_idx then counts within the
 *                              cylindrical areas): ray_directions [N,R,P,4], ray_magnitudes [N,R,P], ray_origins [N,P,4],
 *                              target_idx [N] or NULL (= area 0) -> bitmap coordinates e (flipped for planes) and u,
 *                              intersection distances, Lambert-weighted intensities, each [N,R,P], zero where invalid
 *   ab200_bilinear_splatting   HeliostatRayTracer.bilinear_splatting, heliostat_ray_tracer.py:610-778: be, bu,
 *                              intensities [N,K] -> out [N,U,E] (zeroed here; fp32 atomics, row-flipped like the reference)
 */
int32_t ab200_reflect(const float* incident, const float* normals, int32_t n_samples, int32_t n_points, float* out,
                      void* stream);
int32_t ab200_scatter_rays(const float* distortions_u, const float* distortions_e, const float* reflected,
                           int32_t n_samples, int32_t n_rays, int32_t n_points, float* out, void* stream);
int32_t ab200_line_intersections(const float* ray_directions, const float* ray_magnitudes, const float* ray_origins,
                                 const ab200_targets* targets, const int32_t* target_idx, int32_t cylindrical,
                                 int32_t n_samples, int32_t n_rays, int32_t n_points, int32_t res_e, int32_t res_u,
                                 float* be, float* bu, float* distances, float* intensities, void* stream);
int32_t ab200_bilinear_splatting(const float* be, const float* bu, const float* intensities, int32_t n_samples,
                                 int64_t rays_per_sample, int32_t res_e, int32_t res_u, float* out, void* stream);

/*
 * Step after the ray tracer (flux post-processing, artist/flux/bitmap.py).
 *   ab200_flux_moments   get_center_of_mass :12-55: moments[b] = (sum, e-, u- centre of mass) of bitmaps [n,U,E]; pixel
 *                        units (normalised = 0) or the normalised [-1,1] coordinates the crop uses (normalised = 1)
 *   ab200_flux_crop_fwd  crop_flux_distributions_around_center :121-246: bilinear resampling (affine_grid + grid_sample,
 *                        align_corners, zero padding) of every bitmap around its centre of mass with
 *                        scale[b] = (crop_width / target width, crop_height / target height); writes moments [n,3]
 *                        (kept for the backward) and out [n,U,E]
 *   ab200_flux_crop_bwd  its backward (through the resampling AND through the centre of mass), a deterministic gather;
 *                        scratch = 2n floats
 */
int32_t ab200_flux_moments(const float* bitmaps, int32_t n_bitmaps, int32_t res_u, int32_t res_e, int32_t normalised,
                           float* moments, void* stream);
/* backward of the two centre coordinates: grad_centre [n,2] -> grad_bitmaps [n,U,E] */
int32_t ab200_flux_moments_bwd(const float* moments, const float* grad_centre, int32_t n_bitmaps, int32_t res_u, int32_t res_e,
                               int32_t normalised, float* grad_bitmaps, void* stream);
int32_t ab200_flux_crop_fwd(const float* bitmaps, const float* scale, int32_t n_bitmaps, int32_t res_u, int32_t res_e,
                            float* moments, float* out, void* stream);
int32_t ab200_flux_crop_bwd(const float* bitmaps, const float* scale, const float* moments, const float* grad_out,
                            int32_t n_bitmaps, int32_t res_u, int32_t res_e, float* scratch, float* grad_in, void* stream);

/*
 * Losses on flux bitmaps (artist/optim/loss.py), reduced over the whole bitmap (reduction_dimensions = (1, 2)):
 *   AB200_LOSS_PIXEL          PixelLoss :251-319:        loss[b] = sum (p - g)^2 / sum g
 *   AB200_LOSS_KL_DIVERGENCE  KLDivergenceLoss :322-410: both bitmaps L1-normalised (eps 1e-12), + 1e-12, log,
 *                                                        KLDivLoss(log_target=True), summed
 * prediction, ground_truth [n,U,E]; loss [n]; aux [n,4] (kept for the backward).  ab200_flux_loss_bwd writes
 * grad_prediction [n,U,E] = d loss[b] / d prediction * grad_loss[b] (the ground truth is a constant).
 */
#define AB200_LOSS_PIXEL 0
#define AB200_LOSS_KL_DIVERGENCE 1
int32_t ab200_flux_loss_fwd(const float* prediction, const float* ground_truth, int32_t n_bitmaps, int32_t res_u, int32_t res_e,
                            int32_t kind, float* loss, float* aux, void* stream);
int32_t ab200_flux_loss_bwd(const float* prediction, const float* ground_truth, const float* aux, const float* grad_loss,
                            int32_t n_bitmaps, int32_t res_u, int32_t res_e, int32_t kind, float* grad_prediction, void* stream);

/*
 * ab200_trace_host - end-to-end convenience entry with HOST buffers: uploads the per-call inputs
 * (incident directions, target indices, aligned points/normals if given on the host), traces,
 * and downloads the per-target bitmaps.  Device scratch is supplied by the caller.
 * Synchronises `stream` before it returns.  Exercised by tests/test_gpu_trace_parity.py (bit-identical to the
 * device-pointer path); bench.py's e2e leg goes through the class API with pinned torch tensors instead.
 */
typedef struct ab200_host_trace_args {
    ab200_trace_args dev;            /* device-side argument block; points/normals/incident/target_idx are
                                        DEVICE scratch buffers that this call fills from the host pointers below */
    const float* h_points;           /* host [N,P,4] or NULL (already resident in dev.points) */
    const float* h_normals;          /* host [N,P,4] or NULL */
    const float* h_incident;         /* host [N,4] */
    const int32_t* h_target_idx;     /* host [N] */
    float* d_target_bitmaps;         /* device scratch [T,U,E] */
    float* h_target_bitmaps;         /* host out [T,U,E] */
    float* h_factors;                /* host out [3,N] (intercept, on_target, blocking) or NULL */
} ab200_host_trace_args;

int32_t ab200_trace_host(const ab200_host_trace_args* args, void* stream);

/* pack the blocking primitives (corners [H,4,3], spans [H,2,3], normals [H,3] -> prims [H,16]) */
int32_t ab200_blocking_pack(const float* corners, const float* spans, const float* normals, int32_t n_prims, float epsilon,
                            float* prims, void* stream);
/* candidate lists: primitives whose bounding sphere touches the tapered capsule heliostat -> aim point
 * (radius from the heliostat's half diagonal to target_radius + spread_angle * distance); `overflow` (int, caller
 * zeroes) counts samples with more than max_candidates candidates. */
int32_t ab200_blocking_candidates(const float* prims, int32_t n_prims, const int32_t* sample_to_blocker,
                                  const float* aim_points /* [N,4] */, const float* target_radius /* [N] */,
                                  int32_t n_samples, float spread_angle, int32_t max_candidates, int32_t* cand_idx,
                                  int32_t* cand_count, int32_t* overflow, void* stream);

/*
 * Sun-shape distortion sampling (artist/scene/sun.py:199-234, Sun.get_distortions): out [n_pairs, 2] = (u, e) pairs in
 * the layout of `MultivariateNormal(mean, cov*I).sample((N, R, P))` (n_pairs = N*R*P), bit-identical to what torch's CUDA
 * generator produces for (seed, philox_offset): out[2i+c] = mean_c + fl(sigma_c * z[2i+c]), z = the `normal_()` Philox
 * stream (ATen/native/cuda/DistributionTemplates.h:64-89, curand_normal4).  sm_count / max_threads_per_sm decide torch's
 * counter layout; pass 0 to use the current device's.  philox_offset_after (host, may be NULL) receives the generator
 * offset torch would be left with.
 */
int32_t ab200_sample_distortions(float* out, int64_t n_pairs, uint64_t seed, uint64_t philox_offset, float sigma_u,
                                 float sigma_e, float mean_u, float mean_e, int32_t sm_count_override,
                                 int32_t max_threads_per_sm_override, uint64_t* philox_offset_after, float* out_planar,
                                 void* stream);
/* out_planar (optional, [2, n_pairs]): the same samples de-interleaved, see ab200_trace_args::distortions_planar */
/* de-interleave an existing [n_pairs,2] buffer into [2,n_pairs] */
int32_t ab200_deinterleave_distortions(const float* interleaved, int64_t n_pairs, float* out_planar, void* stream);

/* misc */
int32_t ab200_abi_version(void);
int64_t ab200_kernel_launch_count(void); /* kernels launched by this library so far (diagnostic) */
const char* ab200_error_string(int32_t code);
const char* ab200_last_error_detail(void); /* thread-local detail of the last failing call */
/* per-ray trig probe for parity tests: out_sin/out_cos [n] with the kernel's trig for `mode` */
/* parity probe: the kernels' exact constant-divisor quotient (q_fast) next to IEEE division (q_ieee), a[i] / b */
int32_t ab200_debug_const_div(const float* a, int32_t n, float b, float* q_fast, float* q_ieee, void* stream);
/* parity probe: the range-guarded division of the fast ray loops (q_fast) next to IEEE division, a[i] / b[i] */
int32_t ab200_debug_div_regular(const float* a, const float* b, int32_t n, float* q_fast, float* q_ieee, void* stream);
int32_t ab200_debug_trig(const float* angles, int32_t n, int32_t mode, float* out_sin, float* out_cos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARTIST_B200_H */
