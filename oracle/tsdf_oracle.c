/*
 * tsdf_oracle.c — CPU restatement of the Scoobadood/TSDF integrate + raycast hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (tsdf_b200/csrc) never links, imports or calls anything in oracle/.
 *
 * The reference has no CPU implementation at this commit; this file
 * restates the reference's CUDA kernels line by line in plain C, with every fp32
 * operation in the reference's order and NO fused multiply-add (build with
 * -ffp-contract=off).  All operations used are IEEE-754 correctly rounded on both
 * x86-64 and sm_100 (+ - * / sqrt floor round min max), so this oracle is bit-identical
 * to the reference kernels built with -fmad=false (asserted on the GPU box by
 * tests/test_ref_cuda.py against oracle/_ref).
 *
 * Parity pin: oracle/_ref (the reference's own .cu files compiled for sm_100) run on the
 * GPU box; TestData/t_100_2000_50.tsdf (clear(), truncation distance, file layout) via
 * tests/golden/.
 *
 * Reference citations are relative to /root/reference/src.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } f3;

/* Column-major 4x4 / 3x3, the memory layout of include/cuda_utilities.hpp:12-23
 * (Mat44{m11,m21,m31,m41,m12,...}) which is also Eigen's default storage.          */
#define M4(m, r, c) ((m)[((c) - 1) * 4 + ((r) - 1)])
#define M3(m, r, c) ((m)[((c) - 1) * 3 + ((r) - 1)])

/* GPU float -> int32 conversion (cvt.rzi.s32.f32): NaN -> 0, saturating.  The reference
 * relies on it implicitly (cuda_coordinate_transforms.cu:25-26, TSDF_utilities.cu:46-50);
 * in C the out-of-range conversion is undefined behaviour, so emulate it.            */
static inline int32_t gpu_f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* TSDFVolume::set_size, TSDF/TSDFVolume.cu:686-693:
 *   m_voxel_size = f3_div_elem(physical, dim3)      (cuda_utilities.hpp:78-81)
 *   m_truncation_distance = 1.1f * f3_norm(voxel)   (cuda_utilities.hpp:99-102)       */
void oracle_volume_params(uint32_t nx, uint32_t ny, uint32_t nz, const float phys[3],
                          float voxel[3], float *trunc) {
    voxel[0] = phys[0] / nx;
    voxel[1] = phys[1] / ny;
    voxel[2] = phys[2] / nz;
    *trunc = 1.1f * sqrtf(voxel[0] * voxel[0] + voxel[1] * voxel[1] + voxel[2] * voxel[2]);
}

/* TSDFVolume::clear, TSDF/TSDFVolume.cu:812-845: weights <- 0 (:822), distances <- trunc
 * (:829), colours untouched (cudaMemset arguments swapped, :835), deformation nodes <-
 * initialise_deformation (:768-794).  deform may be NULL (6 floats per voxel otherwise). */
void oracle_clear(float *dist, float *weight, float *deform, uint32_t nx, uint32_t ny,
                  uint32_t nz, const float voxel[3], const float grid_offset[3], float trunc) {
    size_t n = (size_t)nx * ny * nz;
    for (size_t i = 0; i < n; i++) { weight[i] = 0.0f; dist[i] = trunc; }
    if (!deform) return;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t vz = 0; vz < (int64_t)nz; vz++)
        for (int64_t vy = 0; vy < (int64_t)ny; vy++) {
            size_t idx = ((size_t)nx * ny) * vz + (size_t)nx * vy;
            for (int vx = 0; vx < (int)nx; vx++, idx++) {
                float *d = deform + 6 * idx;
                d[0] = ((vx + 0.5f) * voxel[0]) + grid_offset[0];      /* :783 */
                d[1] = (((int)vy + 0.5f) * voxel[1]) + grid_offset[1]; /* :784 */
                d[2] = (((int)vz + 0.5f) * voxel[2]) + grid_offset[2]; /* :785 */
                d[3] = d[4] = d[5] = 0.0f;                             /* :787-789 */
            }
        }
}

/* integrate_kernel, TSDF/TSDFVolume.cu:308-392, for z in [z_begin, z_end).
 * deform: 6 floats per voxel (translation xyz, rotation xyz) exactly as the reference
 * reads it (:343), or NULL to evaluate initialise_deformation's expression (:783-785)
 * with offset_at_clear in its place.  Returns the number of voxels rewritten.        */
uint64_t oracle_integrate(float *dist, float *weight, const float *deform, uint32_t nx,
                          uint32_t ny, uint32_t nz, const float voxel[3],
                          const float offset_at_clear[3], const float offset[3], float trunc,
                          const float inv_pose[16], const float k[9], const float kinv[9],
                          uint32_t width, uint32_t height, const uint16_t *depth,
                          uint32_t z_begin, uint32_t z_end) {
    uint64_t n_upd = 0;
    if (z_end > nz) z_end = nz;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : n_upd)
    for (int64_t vz = z_begin; vz < (int64_t)z_end; vz++)
        for (int64_t vy = 0; vy < (int64_t)ny; vy++) {
            size_t voxel_index = ((size_t)nx * ny) * vz + (size_t)nx * vy; /* :334 */
            for (int vx = 0; vx < (int)nx; vx++, voxel_index++) {          /* :337 */
                f3 tr;
                if (deform) {
                    tr.x = deform[6 * voxel_index + 0];
                    tr.y = deform[6 * voxel_index + 1];
                    tr.z = deform[6 * voxel_index + 2];
                } else {
                    tr.x = ((vx + 0.5f) * voxel[0]) + offset_at_clear[0];
                    tr.y = (((int)vy + 0.5f) * voxel[1]) + offset_at_clear[1];
                    tr.z = (((int)vz + 0.5f) * voxel[2]) + offset_at_clear[2];
                }
                /* f3_add(offset, translation) = translation + offset, cuda_utilities.hpp:46-48 */
                f3 c = { tr.x + offset[0], tr.y + offset[1], tr.z + offset[2] };

                /* world_to_pixel, Utilities/cuda_coordinate_transforms.cu:10-30 */
                f3 cam;
                cam.x = M4(inv_pose,1,1) * c.x + M4(inv_pose,1,2) * c.y + M4(inv_pose,1,3) * c.z + M4(inv_pose,1,4);
                cam.y = M4(inv_pose,2,1) * c.x + M4(inv_pose,2,2) * c.y + M4(inv_pose,2,3) * c.z + M4(inv_pose,2,4);
                cam.z = M4(inv_pose,3,1) * c.x + M4(inv_pose,3,2) * c.y + M4(inv_pose,3,3) * c.z + M4(inv_pose,3,4);
                f3 img;
                img.x = M3(k,1,1) * cam.x + M3(k,1,2) * cam.y + M3(k,1,3) * cam.z;
                img.y = M3(k,2,1) * cam.x + M3(k,2,2) * cam.y + M3(k,2,3) * cam.z;
                img.z = M3(k,3,1) * cam.x + M3(k,3,2) * cam.y + M3(k,3,3) * cam.z;
                int32_t px = gpu_f2i(roundf(img.x / img.z));               /* :25 */
                int32_t py = gpu_f2i(roundf(img.y / img.z));               /* :26 */

                /* TSDFVolume.cu:349 — (px < width) compares int with uint32_t */
                if (!(px >= 0 && (uint32_t)px < width && py >= 0 && (uint32_t)py < height)) continue;
                uint32_t pix_index = (uint32_t)py * width + (uint32_t)px;  /* :352 */
                uint16_t surface_depth = depth[pix_index];
                if (!(surface_depth > 0)) continue;                        /* :356 */

                /* pixel_to_camera, cuda_coordinate_transforms.cu:132-146 (only .z is consumed) */
                float ipc_z = M3(kinv,3,1) * px + M3(kinv,3,2) * py + M3(kinv,3,3);
                float scale = (float)surface_depth / ipc_z;
                float surf_z = ipc_z * scale;                              /* f3_mul_scalar: vec.z * scalar */

                /* world_to_camera, cuda_coordinate_transforms.cu:108-121 */
                float vc_z = (M4(inv_pose,3,1) * c.x) + (M4(inv_pose,3,2) * c.y) + (M4(inv_pose,3,3) * c.z) + M4(inv_pose,3,4);
                float w4   = (M4(inv_pose,4,1) * c.x) + (M4(inv_pose,4,2) * c.y) + (M4(inv_pose,4,3) * c.z) + M4(inv_pose,4,4);
                vc_z /= w4;

                float sdf = surf_z - vc_z;                                 /* :363 */
                if (!(sdf >= -trunc)) continue;                            /* :365 */
                float tsdf;
                if (sdf > 0) tsdf = fminf(sdf, trunc); else tsdf = sdf;    /* :368-372 */

                float prior_weight = weight[voxel_index];                  /* :375 */
                float current_weight = 1.0f;
                float new_weight = prior_weight + current_weight;          /* :377; clamp commented out :378 */
                float prior_distance = dist[voxel_index];
                float new_distance = ((prior_distance * prior_weight) + (tsdf * current_weight)) / new_weight; /* :381 */
                weight[voxel_index] = new_weight;                          /* :383 */
                dist[voxel_index] = new_distance;                          /* :384 */
                n_upd++;
            }
        }
    return n_upd;
}

/* tsdf_value_at, TSDF/TSDF_utilities.cu:29-37 (uint16_t parameters, clamped high) */
static inline float tsdf_value_at(int x, int y, int z, const float *v, uint32_t nx, uint32_t ny, uint32_t nz) {
    uint32_t ux = (uint16_t)x, uy = (uint16_t)y, uz = (uint16_t)z;
    if (ux > nx - 1) ux = nx - 1;
    if (uy > ny - 1) uy = ny - 1;
    if (uz > nz - 1) uz = nz - 1;
    uint32_t idx = nx * ny * uz + nx * uy + ux;   /* 32-bit unsigned arithmetic as in the reference */
    return v[(size_t)idx];
}

/* trilinearly_interpolate, RayCaster/GPURaycaster.cu:53-124 */
static float trilinear(f3 p, uint32_t nx, uint32_t ny, uint32_t nz, const float vs[3], const float *v) {
    f3 mx = { nx * vs[0], ny * vs[1], nz * vs[2] };                        /* :60-64 */
    f3 adj = p;
    if (p.x >= mx.x) adj.x = mx.x - (vs[0] / 10.0f);                       /* :66-68 */
    if (p.y >= mx.y) adj.y = mx.y - (vs[1] / 10.0f);
    if (p.z >= mx.z) adj.z = mx.z - (vs[2] / 10.0f);
    if (p.x < 0.0f) adj.x = 0.0f;                                          /* :69-71 */
    if (p.y < 0.0f) adj.y = 0.0f;
    if (p.z < 0.0f) adj.z = 0.0f;
    /* voxel_for_point, TSDF_utilities.cu:45-52 */
    int vx = gpu_f2i(floorf(adj.x / vs[0]));
    int vy = gpu_f2i(floorf(adj.y / vs[1]));
    int vz = gpu_f2i(floorf(adj.z / vs[2]));
    if (vx < 0 || vy < 0 || vz < 0 || (uint32_t)vx >= nx || (uint32_t)vy >= ny || (uint32_t)vz >= nz)
        return NAN;                                                        /* :77-80 (printf dropped) */
    /* centre_of_voxel_at with the default zero offset, TSDF_utilities.cu:10-17 */
    f3 ctr = { (vx + 0.5f) * vs[0] + 0.0f, (vy + 0.5f) * vs[1] + 0.0f, (vz + 0.5f) * vs[2] + 0.0f };
    int lx = (p.x < ctr.x) ? vx - 1 : vx;                                  /* :87-89 */
    int ly = (p.y < ctr.y) ? vy - 1 : vy;
    int lz = (p.z < ctr.z) ? vz - 1 : vz;
    if (lx < 0) lx = 0;                                                    /* :92-94 */
    if (ly < 0) ly = 0;
    if (lz < 0) lz = 0;
    f3 lc = { (lx + 0.5f) * vs[0] + 0.0f, (ly + 0.5f) * vs[1] + 0.0f, (lz + 0.5f) * vs[2] + 0.0f };
    float u = (p.x - lc.x) / vs[0];                                        /* :98-102 */
    float vv = (p.y - lc.y) / vs[1];
    float w = (p.z - lc.z) / vs[2];
    float c000 = tsdf_value_at(lx + 0, ly + 0, lz + 0, v, nx, ny, nz);     /* :105-112 */
    float c001 = tsdf_value_at(lx + 0, ly + 0, lz + 1, v, nx, ny, nz);
    float c010 = tsdf_value_at(lx + 0, ly + 1, lz + 0, v, nx, ny, nz);
    float c011 = tsdf_value_at(lx + 0, ly + 1, lz + 1, v, nx, ny, nz);
    float c100 = tsdf_value_at(lx + 1, ly + 0, lz + 0, v, nx, ny, nz);
    float c101 = tsdf_value_at(lx + 1, ly + 0, lz + 1, v, nx, ny, nz);
    float c110 = tsdf_value_at(lx + 1, ly + 1, lz + 0, v, nx, ny, nz);
    float c111 = tsdf_value_at(lx + 1, ly + 1, lz + 1, v, nx, ny, nz);
    float s = c000 * (1 - u) * (1 - vv) * (1 - w) +                        /* :114-121 */
              c001 * (1 - u) * (1 - vv) * w +
              c010 * (1 - u) * vv * (1 - w) +
              c011 * (1 - u) * vv * w +
              c100 * u * (1 - vv) * (1 - w) +
              c101 * u * (1 - vv) * w +
              c110 * u * vv * (1 - w) +
              c111 * u * vv * w;
    return s;
}

/* can_intersect_in_dimension, GPURaycaster.cu:138-181 */
static int can_intersect(float smin, float smax, float o, float d, float *near_t, float *far_t) {
    int ok = 1;
    if (d == 0) {
        if (o < smin || o > smax) ok = 0;
    } else {
        float d0 = (smin - o) / d;
        float d1 = (smax - o) / d;
        if (d0 > d1) { float t = d0; d0 = d1; d1 = t; }
        if (d0 > *near_t) *near_t = d0;
        if (d1 < *far_t) *far_t = d1;
        if (*near_t > *far_t) ok = 0;
        else if (*far_t < 0) ok = 0;
    }
    return ok;
}

/* compute_near_and_far_t, GPURaycaster.cu:197-251 */
static int near_far(f3 o, f3 d, f3 smin, f3 smax, float *near_t, float *far_t) {
    if (o.x >= smin.x && o.x <= smax.x && o.y >= smin.y && o.y <= smax.y && o.z >= smin.z && o.z <= smax.z) {
        *near_t = 0;
        float xt = NAN, yt = NAN, zt = NAN;
        if (d.x > 0) xt = (smax.x - o.x) / d.x; else if (d.x < 0) xt = (smin.x - o.x) / d.x;
        if (d.y > 0) yt = (smax.y - o.y) / d.y; else if (d.y < 0) yt = (smin.y - o.y) / d.y;
        if (d.z > 0) zt = (smax.z - o.z) / d.z; else if (d.z < 0) zt = (smin.z - o.z) / d.z;
        if (xt < yt) { if (xt < zt) *far_t = xt; else *far_t = zt; }
        else         { if (yt < zt) *far_t = yt; else *far_t = zt; }
        return 1;
    }
    *near_t = -INFINITY;
    *far_t = INFINITY;
    return can_intersect(smin.x, smax.x, o.x, d.x, near_t, far_t) &&
           can_intersect(smin.y, smax.y, o.y, d.y, near_t, far_t) &&
           can_intersect(smin.z, smax.z, o.z, d.z, near_t, far_t);
}

/* process_ray, GPURaycaster.cu:265-377, for every pixel.  vertices: 3 floats per pixel,
 * index y*w+x.  khit (optional): index k of the sample at which the ray terminated with
 * a hit, -1 for rays that end without one.  Only image rows y_begin, y_begin+y_step, ... are
 * processed (0,1 = all; used to time a bounded sample).  Returns the number of trilinear samples. */
uint64_t oracle_raycast(const float *dist, uint32_t nx, uint32_t ny, uint32_t nz, const float voxel[3],
                        const float space_min_[3], const float space_max_[3], float trunc,
                        const float origin_[3], const float rot[9], const float kinv[9],
                        uint32_t width, uint32_t height, float *vertices, int32_t *khit,
                        uint32_t y_begin, uint32_t y_step) {
    uint64_t n_samples = 0;
    f3 origin = { origin_[0], origin_[1], origin_[2] };
    f3 smin = { space_min_[0], space_min_[1], space_min_[2] };
    f3 smax = { space_max_[0], space_max_[1], space_max_[2] };
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_samples)
    for (int64_t imy = y_begin; imy < (int64_t)height; imy += (y_step ? y_step : 1))
        for (uint32_t imx = 0; imx < width; imx++) {
            size_t idx = (size_t)imy * width + imx;
            uint16_t pix_x = (uint16_t)imx, pix_y = (uint16_t)imy;
            /* compute_ray_direction_at_pixel, GPURaycaster.cu:24-44; f3_normalise takes its
             * argument by value (cuda_utilities.hpp:87-93) so the direction is NOT normalised */
            f3 rc = { pix_x * M3(kinv,1,1) + pix_y * M3(kinv,1,2) + M3(kinv,1,3),
                      pix_x * M3(kinv,2,1) + pix_y * M3(kinv,2,2) + M3(kinv,2,3),
                      pix_x * M3(kinv,3,1) + pix_y * M3(kinv,3,2) + M3(kinv,3,3) };
            f3 dir = { M3(rot,1,1) * rc.x + M3(rot,1,2) * rc.y + M3(rot,1,3) * rc.z,
                       M3(rot,2,1) * rc.x + M3(rot,2,2) * rc.y + M3(rot,2,3) * rc.z,
                       M3(rot,3,1) * rc.x + M3(rot,3,2) * rc.y + M3(rot,3,3) * rc.z };
            float near_t, far_t;
            int intersects = near_far(origin, dir, smin, smax, &near_t, &far_t);
            f3 ip = { NAN, NAN, NAN };
            int32_t kh = -1;
            if (intersects) {
                /* :306  f3_sub(f3_add(origin, f3_mul_scalar(near_t, direction)), space_min) */
                f3 start = { (dir.x * near_t + origin.x) - smin.x,
                             (dir.y * near_t + origin.y) - smin.y,
                             (dir.z * near_t + origin.z) - smin.z };
                int done = 0;
                float tsdf_outer = trunc;      /* :311; never updated: the loop's tsdf shadows it (:332) */
                float previous_tsdf = 0;
                float t = 0;
                float max_t = far_t - near_t;  /* :317 */
                int count = 0;
                float step_size = (float)((double)trunc * 0.05); /* :324 (double literal) */
                while (!done) {
                    f3 cp = { dir.x * t + start.x, dir.y * t + start.y, dir.z * t + start.z }; /* :326 */
                    previous_tsdf = tsdf_outer;                                                /* :329 */
                    float tsdf = trilinear(cp, nx, ny, nz, voxel, dist);                       /* :332 */
                    n_samples++;
                    if (tsdf <= 0) {
                        if (tsdf < 0) {
                            t = t - step_size;                                                 /* :338 */
                            t = t + (previous_tsdf / (previous_tsdf - tsdf)) * step_size;      /* :341 */
                        }
                        cp.x = dir.x * t + start.x; cp.y = dir.y * t + start.y; cp.z = dir.z * t + start.z; /* :345 */
                        ip.x = cp.x + smin.x; ip.y = cp.y + smin.y; ip.z = cp.z + smin.z;      /* :348 */
                        kh = count;
                        done = 1;
                    } else if (previous_tsdf < 0) {
                        done = 1;                                                              /* :354 (dead) */
                    } else {
                        t = t + step_size;                                                     /* :360 */
                        if (t >= max_t) done = 1;                                              /* :363 */
                    }
                    if (count++ > 4400) done = 1;                                              /* :369 (printf dropped) */
                }
            }
            vertices[3 * idx + 0] = ip.x;                                                      /* :376 */
            vertices[3 * idx + 1] = ip.y;
            vertices[3 * idx + 2] = ip.z;
            if (khit) khit[idx] = kh;
        }
    return n_samples;
}

/* ---- Z-sharded raycast (no reference counterpart: the reference is single-GPU) ----------------------------
 * Statement of the decomposition the multi-GPU path uses, so that it can be checked on the CPU: a rank that
 * owns the cells starting in [z_lo, z_hi) walks process_ray's loop but evaluates only its own samples (the
 * others count as "no hit here"); key = (k << 32 | bits(sample)) of its first hit, INT64_MAX if none.  The
 * minimum key over ranks is the first hit of the undivided loop, and oracle_resolve turns it into the vertex
 * with process_ray's formula (GPURaycaster.cu:336-348).                                                     */
static int cell_start_z(f3 p, uint32_t nz, const float vs[3]) {
    float mz = nz * vs[2];
    float adj = p.z;
    if (p.z >= mz) adj = mz - (vs[2] / 10.0f);
    if (p.z < 0.0f) adj = 0.0f;
    int vz = gpu_f2i(floorf(adj / vs[2]));
    float ctr = (vz + 0.5f) * vs[2] + 0.0f;
    int lz = (p.z < ctr) ? vz - 1 : vz;
    if (lz < 0) lz = 0;
    return lz;
}

static void ray_setup(const float origin_[3], const float rot[9], const float kinv[9], const float smin_[3],
                      const float smax_[3], uint32_t imx, uint32_t imy, f3 *dir, f3 *start, float *max_t, int *intersects) {
    f3 origin = { origin_[0], origin_[1], origin_[2] };
    f3 smin = { smin_[0], smin_[1], smin_[2] }, smax = { smax_[0], smax_[1], smax_[2] };
    uint16_t pix_x = (uint16_t)imx, pix_y = (uint16_t)imy;
    f3 rc = { pix_x * M3(kinv,1,1) + pix_y * M3(kinv,1,2) + M3(kinv,1,3),
              pix_x * M3(kinv,2,1) + pix_y * M3(kinv,2,2) + M3(kinv,2,3),
              pix_x * M3(kinv,3,1) + pix_y * M3(kinv,3,2) + M3(kinv,3,3) };
    dir->x = M3(rot,1,1) * rc.x + M3(rot,1,2) * rc.y + M3(rot,1,3) * rc.z;
    dir->y = M3(rot,2,1) * rc.x + M3(rot,2,2) * rc.y + M3(rot,2,3) * rc.z;
    dir->z = M3(rot,3,1) * rc.x + M3(rot,3,2) * rc.y + M3(rot,3,3) * rc.z;
    float near_t, far_t;
    *intersects = near_far(origin, *dir, smin, smax, &near_t, &far_t);
    start->x = (dir->x * near_t + origin.x) - smin.x;
    start->y = (dir->y * near_t + origin.y) - smin.y;
    start->z = (dir->z * near_t + origin.z) - smin.z;
    *max_t = far_t - near_t;
}

void oracle_raycast_slab(const float *dist, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z_lo, uint32_t z_hi,
                         const float voxel[3], const float space_min[3], const float space_max[3], float trunc,
                         const float origin[3], const float rot[9], const float kinv[9],
                         uint32_t width, uint32_t height, int64_t *keys) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t imy = 0; imy < (int64_t)height; imy++)
        for (uint32_t imx = 0; imx < width; imx++) {
            f3 dir, start; float max_t; int intersects;
            ray_setup(origin, rot, kinv, space_min, space_max, imx, (uint32_t)imy, &dir, &start, &max_t, &intersects);
            int64_t key = INT64_MAX;
            if (intersects) {
                float t = 0, step_size = (float)((double)trunc * 0.05);
                int count = 0, done = 0;
                while (!done) {
                    f3 cp = { dir.x * t + start.x, dir.y * t + start.y, dir.z * t + start.z };
                    int lz = cell_start_z(cp, nz, voxel);
                    float tsdf = trunc;                                   /* not mine: behaves like a positive sample */
                    if ((uint32_t)lz >= z_lo && (uint32_t)lz < z_hi) tsdf = trilinear(cp, nx, ny, nz, voxel, dist);
                    if (tsdf <= 0) {
                        uint32_t bits; memcpy(&bits, &tsdf, 4);
                        key = ((int64_t)count << 32) | (int64_t)bits;
                        done = 1;
                    } else {
                        t = t + step_size;
                        if (t >= max_t) done = 1;
                    }
                    if (count++ > 4400) done = 1;
                }
            }
            keys[(size_t)imy * width + imx] = key;
        }
}

void oracle_resolve(const int64_t *keys, const float space_min[3], const float space_max[3], float trunc,
                    const float origin[3], const float rot[9], const float kinv[9],
                    uint32_t width, uint32_t height, float *vertices, int32_t *khit) {
    float step_size = (float)((double)trunc * 0.05);
    for (uint32_t imy = 0; imy < height; imy++)
        for (uint32_t imx = 0; imx < width; imx++) {
            size_t idx = (size_t)imy * width + imx;
            f3 ip = { NAN, NAN, NAN };
            int32_t kh = -1;
            if (keys[idx] != INT64_MAX) {
                kh = (int32_t)(keys[idx] >> 32);
                uint32_t bits = (uint32_t)(keys[idx] & 0xffffffff);
                float s; memcpy(&s, &bits, 4);
                f3 dir, start; float max_t; int intersects;
                ray_setup(origin, rot, kinv, space_min, space_max, imx, imy, &dir, &start, &max_t, &intersects);
                float t = 0;
                for (int k = 0; k < kh; k++) t = t + step_size;          /* t_k by repeated addition (:360) */
                if (s < 0) { t = t - step_size; t = t + (trunc / (trunc - s)) * step_size; }
                ip.x = (dir.x * t + start.x) + space_min[0];
                ip.y = (dir.y * t + start.y) + space_min[1];
                ip.z = (dir.z * t + start.z) + space_min[2];
            }
            vertices[3 * idx + 0] = ip.x; vertices[3 * idx + 1] = ip.y; vertices[3 * idx + 2] = ip.z;
            if (khit) khit[idx] = kh;
        }
}

/* compute_normals kernel, GPURaycaster.cu:393-427 */
void oracle_normals(uint32_t width, uint32_t height, const float *V, float *N) {
#pragma omp parallel for schedule(static)
    for (int64_t imy = 0; imy < (int64_t)height; imy++)
        for (uint32_t imx = 0; imx < width; imx++) {
            size_t idx = (size_t)imy * width + imx;
            float *n = N + 3 * idx;
            if (imy == (int64_t)height - 1 || imx == width - 1) { n[0] = n[1] = n[2] = 0; continue; }
            const float *a = V + 3 * idx, *r = V + 3 * (idx + 1), *b = V + 3 * (idx + width);
            f3 v2 = { r[0] - a[0], r[1] - a[1], r[2] - a[2] };
            f3 v1 = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };
            float nx = v1.y * v2.z - v1.z * v2.y;
            float ny = v1.z * v2.x - v1.x * v2.z;
            float nz = v1.x * v2.y - v1.y * v2.x;
            float l = sqrtf(nx * nx + ny * ny + nz * nz);
            n[0] = nx / l; n[1] = ny / l; n[2] = nz / l;
        }
}

/* "Hit voxel index" of SURVEY.md §8(a): voxel_for_point(vertex - space_min) linearised
 * x + y*X + z*X*Y, -1 for misses.  Not a reference output; shared definition for tests. */
void oracle_hit_voxels(uint32_t n_pix, const float *vertices, const float space_min[3], const float voxel[3],
                       uint32_t nx, uint32_t ny, int64_t *out) {
    for (uint32_t i = 0; i < n_pix; i++) {
        const float *v = vertices + 3 * (size_t)i;
        if (v[0] != v[0]) { out[i] = -1; continue; }
        int64_t x = gpu_f2i(floorf((v[0] - space_min[0]) / voxel[0]));
        int64_t y = gpu_f2i(floorf((v[1] - space_min[1]) / voxel[1]));
        int64_t z = gpu_f2i(floorf((v[2] - space_min[2]) / voxel[2]));
        out[i] = x + y * (int64_t)nx + z * (int64_t)nx * ny;
    }
}

/* ---- marching cubes -------------------------------------------------------------------------
 * Restatement of extract_surface_ms (reference src/MarchingCubes/MarkAndSweepMC.cu): cube corner numbering
 * voxel_indices_for_cube_index :60-100, calculate_cube_type :110-124, per-cube vertex count :132-153, host
 * scan in ascending cube index :456-473, generate_vertices :218-304 with interpolate :44-58 and
 * centre_of_voxel_at (TSDF/TSDF_utilities.cu:10-17).  Triangle table: Bourke's, shared with the product
 * (tsdf_b200/csrc/mc_tables.h) and checked against the reference's copy by tests/test_mc_cpu.py.
 * Two passes: out == NULL counts, otherwise fills 3 floats per vertex.  Returns the vertex count.    */
#include "../tsdf_b200/csrc/mc_tables.h"

uint64_t oracle_mc_extract(const float *dist, uint32_t nx, uint32_t ny, uint32_t nz, const float voxel[3],
                           const float offset[3], float *out) {
    uint64_t n_out = 0;
    if (nx < 2 || ny < 2 || nz < 2) return 0;
    for (uint32_t z = 0; z + 1 < nz; z++)
        for (uint32_t y = 0; y + 1 < ny; y++)
            for (uint32_t x = 0; x + 1 < nx; x++) {
                /* corners 0..7: (x,z+1) (x+1,z+1) (x+1,z) (x,z) on plane y, then the same on y+1 */
                const uint32_t cx[8] = { x, x + 1, x + 1, x, x, x + 1, x + 1, x };
                const uint32_t cy[8] = { y, y, y, y, y + 1, y + 1, y + 1, y + 1 };
                const uint32_t cz[8] = { z + 1, z + 1, z, z, z + 1, z + 1, z, z };
                float w[8];
                unsigned type = 0;
                for (int c = 0; c < 8; c++) {
                    w[c] = dist[((size_t)nx * ny) * cz[c] + (size_t)nx * cy[c] + cx[c]];
                    if (w[c] < 0) type |= 1u << c;
                }
                const char *tri = kMcTriangles[type];
                if (!tri[0]) continue;
                if (!out) { n_out += strlen(tri); continue; }
                for (int i = 0; tri[i]; i++) {
                    const int e = tri[i] <= '9' ? tri[i] - '0' : tri[i] - 'a' + 10;
                    int a = kMcEdgeCorners[e][0], b = kMcEdgeCorners[e][1];
                    float w0 = w[a], w1 = w[b];
                    if (w0 > 0 && w1 < 0) { const float t = w0; w0 = w1; w1 = t; const int ti = a; a = b; b = ti; }
                    const float va[3] = { (cx[a] + 0.5f) * voxel[0] + offset[0], (cy[a] + 0.5f) * voxel[1] + offset[1], (cz[a] + 0.5f) * voxel[2] + offset[2] };
                    const float vb[3] = { (cx[b] + 0.5f) * voxel[0] + offset[0], (cy[b] + 0.5f) * voxel[1] + offset[1], (cz[b] + 0.5f) * voxel[2] + offset[2] };
                    const float ratio = -(w0) / (w1 - w0);
                    for (int k = 0; k < 3; k++) {
                        const float delta = vb[k] - va[k];
                        out[3 * n_out + k] = va[k] + delta * ratio;
                    }
                    n_out++;
                }
            }
    return n_out;
}

/* ---- bilateral filter -----------------------------------------------------------------------
 * Restatement of BilateralFilter::filter_bpp (reference src/BilateralFilter.cpp:53-121) with the look-up tables of the
 * constructor (:15-42) passed in: taps x-major then y (:81-82), kernel index advancing only for in-image taps (:105),
 * `double w = k*s` from a float product, float sum/total updated through double (:100-103), floorf(sum/total) (:109).
 * bits == 8 is the reference exactly (pinned against the compiled reference by tests/test_bilateral_cpu.py); bits == 16
 * is undefined behaviour in the reference (256-entry table indexed up to 65535, one output byte per pixel): here the
 * table has n_similarity entries (differences beyond it use the last one) and the output is 16-bit.              */
void oracle_bilateral(const void *in, void *out, int bits, int width, int height, const float *kernel, int kernel_size,
                      const float *similarity, int n_similarity) {
    const int radius = (kernel_size - 1) / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            const int centre = bits == 8 ? ((const uint8_t *)in)[(size_t)width * y + x] : ((const uint16_t *)in)[(size_t)width * y + x];
            float total = 0, sum = 0;
            int k = 0;
            for (int cx = x - radius; cx <= x + radius; cx++)
                for (int cy = y - radius; cy <= y + radius; cy++) {
                    if (cx < 0 || cx >= width || cy < 0 || cy >= height) continue;
                    const int v = bits == 8 ? ((const uint8_t *)in)[(size_t)width * cy + cx] : ((const uint16_t *)in)[(size_t)width * cy + cx];
                    int delta = abs(v - centre);
                    if (delta >= n_similarity) delta = n_similarity - 1;
                    const float kw = kernel[k] * similarity[delta];
                    const double w = kw;
                    sum += (w * v);
                    total += w;
                    k++;
                }
            const int r = (int)floorf(sum / total);
            if (bits == 8) ((uint8_t *)out)[(size_t)width * y + x] = (uint8_t)r;
            else ((uint16_t *)out)[(size_t)width * y + x] = (uint16_t)r;
        }
}
