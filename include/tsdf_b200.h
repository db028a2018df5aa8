/*
 * tsdf_b200.h — C-ABI of the B200-native TSDF integrate + raycast hot path.
 *
 * The reference (Scoobadood/TSDF) has no FFI of its own: its boundary is the C++ class
 * surface kinfu.cpp compiles against (src/include/TSDFVolume.hpp, Camera.hpp,
 * Raycaster.hpp, GPURaycaster.hpp).  This header is the seam *under* those classes:
 * level 1 mirrors the reference's kernel launches one to one (raw device pointers + POD
 * parameters, what TSDFVolume.cu / GPURaycaster.cu pass to their __global__ functions);
 * level 2 mirrors the class methods (opaque handles, HOST buffers, synchronous like the
 * reference) and is what the drop-in C++ classes in tsdf_b200/include forward to.
 *
 * Conventions: every function returns 0 on success, a cudaError_t value (>0) for CUDA
 * failures, or a negative TSDF_B200_E* code for argument errors.  Nothing here calls
 * exit().  Matrices are COLUMN-major float arrays — the storage of Eigen::Matrix4f /
 * Matrix3f::data() and of the reference's Mat44/Mat33 (src/include/cuda_utilities.hpp:12-23).
 * Volumes are x-fastest: index = x + y*nx + z*nx*ny (src/include/TSDFVolume.hpp:165-167).
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * All citations are relative to the reference tree's src/ directory.
 */
#ifndef TSDF_B200_H
#define TSDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSDF_B200_EINVAL (-1)   /* bad argument (null pointer, zero size, unsupported dimension) */
#define TSDF_B200_ENOMEM (-2)
#define TSDF_B200_EIO    (-3)
#define TSDF_B200_ESTATE (-4)

/* Longest ray: samples k = 0..4401 (RayCaster/GPURaycaster.cu:369 `count++ > 4400`). */
#define TSDF_B200_MAX_SAMPLES 4402
/* Floats in a ray-parameter table (tsdf_b200_ray_table). */
#define TSDF_B200_RAY_TABLE_LEN 4416
/* Edge of an occupancy brick in voxels (empty-space skipping grid). */
#define TSDF_B200_BRICK 8
/* Most GPUs one exchange call can address (tsdf_b200_bricks_push, tsdf_b200_raycast_tiles). */
#define TSDF_B200_MAX_PEERS 16
/* Bytes of an exported peer-memory handle (tsdf_b200_peer_alloc / tsdf_b200_peer_open). */
#define TSDF_B200_PEER_HANDLE_BYTES 64

const char *tsdf_b200_version(void);
/* Text for any return code of this library (cudaGetErrorString for positive codes). */
const char *tsdf_b200_strerror(int code);

/* ------------------------------------------------------------------------------------
 * Level 1 — kernel launches.  All pointers are DEVICE pointers unless marked host.
 * ---------------------------------------------------------------------------------- */

/* Volume constants as TSDFVolume::set_size derives them (TSDF/TSDFVolume.cu:686-693):
 * voxel = physical / size (element-wise), trunc = 1.1f * |voxel|.  Host-only arithmetic. */
int tsdf_b200_volume_params(uint32_t nx, uint32_t ny, uint32_t nz, const float physical[3],
                            float voxel_out[3], float *trunc_out);

/* Replaces set_memory_to_value x2 in TSDFVolume::clear (TSDF/TSDFVolume.cu:797-832):
 * weight <- 0, dist <- trunc.  d_occ (optional, tsdf_b200_occupancy_bytes() long) <- 0. */
int tsdf_b200_clear(float *d_dist, float *d_weight, uint32_t nx, uint32_t ny, uint32_t nz,
                    float trunc, uint8_t *d_occ, void *stream);

/* Replaces initialise_deformation (TSDF/TSDFVolume.cu:768-794): node = {((v+0.5)*voxel)+
 * grid_offset, 0}; d_deform holds 6 floats per voxel (TSDFVolume::DeformationNode).     */
int tsdf_b200_init_deformation(float *d_deform, uint32_t nx, uint32_t ny, uint32_t nz,
                               const float voxel[3], const float grid_offset[3], void *stream);

/* Staged depth frame for tsdf_b200_integrate: a max-pyramid of the frame (largest depth per 2^l x 2^l pixel tile).
 * The rigid-camera integrate kernel uses it to skip, per warp, slabs of voxels that lie entirely behind everything
 * they can project onto (or entirely outside the image) without projecting them one by one.  Purely an
 * acceleration structure: results are bit-identical with and without it.  d_staged holds
 * tsdf_b200_depth_staged_bytes(width, height) bytes.  Replaces nothing in the reference, which projects every
 * voxel of the volume for every frame (TSDF/TSDFVolume.cu:326-349).                                        */
size_t tsdf_b200_depth_staged_bytes(uint32_t width, uint32_t height);
int tsdf_b200_depth_stage(const uint16_t *d_depth, uint32_t width, uint32_t height, float *d_staged, void *stream);

/* Replaces integrate_kernel (TSDF/TSDFVolume.cu:308-392) for planes z in [z_begin, z_end) of
 * the arrays.  The arrays hold nz planes whose plane 0 is global plane z_base of the volume
 * (0 for a whole volume; a Z-slab of a sharded volume otherwise; a multiple of 8 with d_occ).
 * d_deform == NULL selects the analytic identity grid: translation = ((v+0.5)*voxel) +
 * offset_at_clear, bit-identical to reading the array clear() wrote, without the 24 B/voxel
 * read.  d_depth_staged (optional): the same frame after tsdf_b200_depth_stage (culling pyramid for the
 * rigid-camera kernel).  d_occ (optional): occupancy bricks are marked for every voxel whose new distance
 * leaves the "certainly positive" band.  d_n_updated (optional): incremented by the number
 * of voxels rewritten (device counter, unsigned long long).                             */
int tsdf_b200_integrate(float *d_dist, float *d_weight, const float *d_deform,
                        uint32_t nx, uint32_t ny, uint32_t nz, const float voxel[3],
                        const float offset_at_clear[3], const float offset[3], float trunc,
                        const float inv_pose[16], const float k[9], const float kinv[9],
                        uint32_t width, uint32_t height, const uint16_t *d_depth,
                        const float *d_depth_staged,
                        uint32_t z_begin, uint32_t z_end, uint32_t z_base, uint8_t *d_occ,
                        unsigned long long *d_n_updated, void *stream);

/* Test hook: tsdf_b200_integrate picks a specialised kernel when the camera is rigid with a
 * conventional K (same result bits, fewer instructions); on != 0 forces the general kernel.  */
void tsdf_b200_debug_force_generic_integrate(int on);
/* Test / measurement hook: how the rigid-camera kernel stages the dist / weight planes.  0 = per-thread cp.async, one pass
 * (default); 1 = TMA boxes (cp.async.bulk.tensor + mbarrier; environment TSDF_B200_TMA=1); 2 = box cull into a work list +
 * persistent kernel (TSDF_B200_LIST=1).  Same result bits in every variant.                                      */
void tsdf_b200_debug_integrate_variant(int variant);

/* Size in bytes of the occupancy buffer of a volume: three bytes per 8^3 brick — the brick flags that
 * integrate / occupancy_rebuild maintain, then the brick distance grid and a scratch copy that the
 * raycast derives from the flags on every call (the raycast entry points WRITE those two thirds).  */
size_t tsdf_b200_occupancy_bytes(uint32_t nx, uint32_t ny, uint32_t nz);

/* Recompute the occupancy grid from scratch (after set_distance_data / file load). */
int tsdf_b200_occupancy_rebuild(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                float trunc, uint8_t *d_occ, void *stream);

/* Ray parameter table t_k: t_0 = 0, t_{k+1} = t_k + step, step = (float)(trunc * 0.05)
 * (RayCaster/GPURaycaster.cu:316,324,360) — identical for every ray of a volume.
 * d_table holds TSDF_B200_RAY_TABLE_LEN floats.                                         */
int tsdf_b200_ray_table(float trunc, float *d_table, void *stream);

/* Replaces process_ray (RayCaster/GPURaycaster.cu:265-377).  d_vertices: 3 floats per
 * pixel, index y*width+x, NaN^3 for rays without a hit.  d_khit (optional): sample index
 * of the hit, -1 otherwise.  d_occ (optional): occupancy grid enabling exact empty-space
 * skipping.  d_n_samples (optional): incremented by trilinear samples actually evaluated. */
int tsdf_b200_raycast(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                      const float voxel[3], const float space_min[3], const float space_max[3],
                      float trunc, const float origin[3], const float rot[9], const float kinv[9],
                      uint32_t width, uint32_t height, const float *d_table,
                      const uint8_t *d_occ, float *d_vertices, int32_t *d_khit,
                      unsigned long long *d_n_samples, void *stream);

/* tsdf_b200_raycast with the division by the voxel size done as a 3-instruction reciprocal
 * sequence when fastdiv != 0.  Pass 1 only after tsdf_b200_selftest_division() returned zero
 * mismatches for voxel[0], voxel[1] and voxel[2]; results are then bit-identical.        */
int tsdf_b200_raycast_ex(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                         const float voxel[3], const float space_min[3], const float space_max[3],
                         float trunc, const float origin[3], const float rot[9], const float kinv[9],
                         uint32_t width, uint32_t height, const float *d_table,
                         const uint8_t *d_occ, float *d_vertices, int32_t *d_khit,
                         unsigned long long *d_n_samples, int fastdiv, void *stream);

/* tsdf_b200_raycast_ex that also writes the vertex map to `mirror` while it marches: meant for pinned host memory that
 * the device can address (the level-2 volume passes the caller's buffer when it is pinned), so that the transfer of the
 * vertex map overlaps the march instead of following it.  Each warp writes its 8x4-pixel tile as aligned 16-byte stores;
 * requires width % 8 == 0, height % 4 == 0 and a 16-byte aligned `mirror` (or mirror == NULL: plain raycast).        */
int tsdf_b200_raycast_mirrored(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                               const float voxel[3], const float space_min[3], const float space_max[3],
                               float trunc, const float origin[3], const float rot[9], const float kinv[9],
                               uint32_t width, uint32_t height, const float *d_table,
                               const uint8_t *d_occ, float *d_vertices, float *mirror,
                               unsigned long long *d_n_samples, int fastdiv, void *stream);

/* process_ray + compute_normals (RayCaster/GPURaycaster.cu:265-377, 393-427) in ONE launch sequence: vertex map and normal
 * map in device memory, and — when mirror_vertices / mirror_normals are given (pinned host memory the device can address;
 * same alignment rules as tsdf_b200_raycast_mirrored) — copies of both written while the march runs: a tile's vertices leave
 * when its last ray finishes, a tile's normals as soon as the tile, its right and its lower neighbour are complete
 * (compute_normals reads v(x+1, y) and v(x, y+1)).  d_tile_counters: tsdf_b200_raycast_tile_counters(width, height) 32-bit
 * words of device memory, ZERO before the first call (every launch leaves them zero again).  Without them, without d_occ, or
 * with an image that is not a whole number of 8x4 tiles the normals come from the separate kernel and the mirrors from a copy. */
size_t tsdf_b200_raycast_tile_counters(uint32_t width, uint32_t height);
int tsdf_b200_raycast_fused(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                            const float voxel[3], const float space_min[3], const float space_max[3],
                            float trunc, const float origin[3], const float rot[9], const float kinv[9],
                            uint32_t width, uint32_t height, const float *d_table,
                            const uint8_t *d_occ, float *d_vertices, float *d_normals,
                            float *mirror_vertices, float *mirror_normals, unsigned int *d_tile_counters,
                            unsigned long long *d_n_samples, int fastdiv, void *stream);

/* Z-sharded raycast, march phase.  d_dist_slab / d_occ_slab hold z_planes planes starting at global plane
 * z_base of an nx*ny*nz volume (owned planes plus the upper halo plane).  Every ray is marched, but only
 * samples whose interpolation cell starts in [z_lo, z_hi) are evaluated.  d_keys[pixel] receives
 * (k_hit << 32 | float_bits(sample)) or INT64_MAX: the minimum over ranks is the first hit along the ray
 * (the sample parameters t_k do not depend on the rank), so one all-reduce(min) merges the shards.     */
int tsdf_b200_raycast_slab(const float *d_dist_slab, uint32_t nx, uint32_t ny, uint32_t nz,
                           uint32_t z_base, uint32_t z_planes, uint32_t z_lo, uint32_t z_hi,
                           const float voxel[3], const float space_min[3], const float space_max[3],
                           float trunc, const float origin[3], const float rot[9], const float kinv[9],
                           uint32_t width, uint32_t height, const float *d_table,
                           const uint8_t *d_occ_slab, long long *d_keys,
                           unsigned long long *d_n_samples, int fastdiv, void *stream);

/* tsdf_b200_raycast_slab with the exchange fused into the march: instead of writing a key per pixel, every ray that hits
 * inside this slab min-merges its key into d_keys_min[pixel] with a 64-bit atomic — d_keys_min may be another GPU's key
 * map (peer memory over NVLink: cudaDeviceEnablePeerAccess in one process, tsdf_b200_peer_open across processes), so the
 * ranks of a sharded volume need no collective for the exchange, only a barrier before the owner of the map resolves it.
 * The map must hold INT64_MAX in every pixel before the first rank starts (tsdf_b200_raycast_resolve_reset leaves it so).  */
int tsdf_b200_raycast_slab_min(const float *d_dist_slab, uint32_t nx, uint32_t ny, uint32_t nz,
                               uint32_t z_base, uint32_t z_planes, uint32_t z_lo, uint32_t z_hi,
                               const float voxel[3], const float space_min[3], const float space_max[3],
                               float trunc, const float origin[3], const float rot[9], const float kinv[9],
                               uint32_t width, uint32_t height, const float *d_table,
                               const uint8_t *d_occ_slab, long long *d_keys_min,
                               unsigned long long *d_n_samples, int fastdiv, void *stream);

/* Z-sharded raycast, march phase, INTERLEAVED slabs: the volume's planes are cut into global slabs of slab_planes planes
 * (a multiple of 8) dealt to `world` ranks round robin (slab s belongs to rank s % world) — surfaces then spread over the
 * ranks instead of landing in one contiguous slab, and the march balances.  d_dist_local holds this rank's slabs back to
 * back in ascending order, each followed by ONE halo plane (slab_planes + 1 planes per slab; the halo of a slab that ends
 * the volume is unused).  d_occ_global is an occupancy grid of the WHOLE volume (tsdf_b200_occupancy_bytes(nx,ny,nz)) in
 * which only this rank's bricks were ever flagged: integrate each slab with d_occ = grid + brick offset of the slab.
 * Keys as in tsdf_b200_raycast_slab: min-reduce over ranks, then tsdf_b200_raycast_resolve.                      */
int tsdf_b200_raycast_interleaved(const float *d_dist_local, uint32_t nx, uint32_t ny, uint32_t nz,
                                  uint32_t slab_planes, uint32_t world, uint32_t rank,
                                  const float voxel[3], const float space_min[3], const float space_max[3],
                                  float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                  uint32_t width, uint32_t height, const float *d_table,
                                  const uint8_t *d_occ_global, long long *d_keys,
                                  unsigned long long *d_n_samples, int fastdiv, void *stream);

/* ---- Multi-GPU, image-sharded raycast over peer memory (no counterpart in the single-GPU reference; the per-sample
 * arithmetic is RayCaster/GPURaycaster.cu:265-377 as in tsdf_b200_raycast) ------------------------------------------------
 *
 * Every GPU holds, next to its own slabs, a full-size copy of the distance volume (a "replica") in which only surface bricks
 * are ever valid.  Per frame: integrate the owned slabs (interleaved layout of tsdf_b200_raycast_interleaved, occupancy
 * grid of the whole volume), max-reduce the brick flags over the ranks, tsdf_b200_bricks_push, barrier,
 * tsdf_b200_raycast_tiles, barrier, tsdf_b200_normals.
 *
 * tsdf_b200_bricks_push: copies every owned brick that has a flagged brick in its 27-neighbourhood (d_occ_global, merged
 * flags), and every owned brick on a low face of the volume (bx, by or bz == 0), from this rank's slabs into the n_dst
 * replicas d_dst[0..n_dst) (host array of device pointers; peers' replicas are written through NVLink).  That set covers
 * every voxel a ray can read: away from the low faces a sample is evaluated only inside a flagged brick b and reads voxels
 * of [8b-1, 8b+8]^3; voxel layer 0 of each axis is always evaluated (the reference extrapolates there) and reads layers 0
 * and 1.  d_n_bricks (optional) += bricks copied.                                                                      */
int tsdf_b200_bricks_push(const float *d_dist_local, uint32_t nx, uint32_t ny, uint32_t nz,
                          uint32_t slab_planes, uint32_t world, uint32_t rank,
                          const uint8_t *d_occ_global, uint32_t n_dst, float *const *d_dst,
                          unsigned long long *d_n_bricks, void *stream);

/* tsdf_b200_raycast_ex restricted to this rank's pixel tiles: of every `world` consecutive 8x4-pixel tiles rank `rank`
 * marches one (rotating from tile row to tile row), against d_dist = its replica and d_occ = the merged occupancy grid, and
 * stores each vertex into all n_out vertex maps d_vertices_out[0..n_out) (host array of device pointers, 3*width*height
 * floats each).  After every rank's call (and a barrier) each map holds the complete single-GPU result.             */
int tsdf_b200_raycast_tiles(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                            const float voxel[3], const float space_min[3], const float space_max[3],
                            float trunc, const float origin[3], const float rot[9], const float kinv[9],
                            uint32_t width, uint32_t height, const float *d_table,
                            const uint8_t *d_occ, uint32_t world, uint32_t rank,
                            uint32_t n_out, float *const *d_vertices_out,
                            unsigned long long *d_n_samples, int fastdiv, void *stream);

/* Peer memory for the two calls above: a cudaMalloc block plus its CUDA IPC handle (64 bytes, to be sent to the other
 * processes of the box); tsdf_b200_peer_open maps another process's block (peer access enabled on demand).          */
int tsdf_b200_peer_alloc(size_t bytes, void **d_ptr, unsigned char handle[TSDF_B200_PEER_HANDLE_BYTES]);
int tsdf_b200_peer_open(const unsigned char handle[TSDF_B200_PEER_HANDLE_BYTES], void **d_ptr);
int tsdf_b200_peer_close(void *d_ptr);
int tsdf_b200_peer_free(void *d_ptr);
int tsdf_b200_fill_f32(float *d_ptr, size_t count, float value, void *stream);

/* Z-sharded raycast, resolve phase: reduced keys -> vertices (NaN^3 for INT64_MAX) and optional k_hit,
 * with the hit formula of RayCaster/GPURaycaster.cu:336-348.                                        */
int tsdf_b200_raycast_resolve(const long long *d_keys, const float space_min[3], const float space_max[3],
                              float trunc, const float origin[3], const float rot[9], const float kinv[9],
                              uint32_t width, uint32_t height, const float *d_table,
                              float *d_vertices, int32_t *d_khit, void *stream);

/* tsdf_b200_raycast_resolve that also puts INT64_MAX back into every key it reads: the map is ready for the next frame's
 * tsdf_b200_raycast_slab_min without a separate fill.                                                              */
int tsdf_b200_raycast_resolve_reset(long long *d_keys, const float space_min[3], const float space_max[3],
                                    float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                    uint32_t width, uint32_t height, const float *d_table,
                                    float *d_vertices, int32_t *d_khit, void *stream);
/* count 64-bit words <- value (key maps: INT64_MAX). */
int tsdf_b200_fill_i64(long long *d_ptr, size_t count, long long value, void *stream);

/* Replaces the compute_normals kernel (RayCaster/GPURaycaster.cu:393-427). */
int tsdf_b200_normals(uint32_t width, uint32_t height, const float *d_vertices,
                      float *d_normals, void *stream);

/* Replaces extract_surface_ms: get_cube_contribution + host prefix sum + generate_vertices
 * (MarchingCubes/MarkAndSweepMC.cu:132-153, 456-473, 218-304, 390-497).  d_dist holds nz_planes planes starting at
 * global plane z_base; cubes whose base plane is in [cz_begin, cz_end) (local, clamped to nz_planes - 1) are
 * processed — a whole volume is (0, 0, nz - 1); a Z-shard passes its owned planes and relies on its halo plane.
 * *d_vertices_out receives a cudaMalloc'ed array of 3 floats per vertex (nullptr when empty; release it with
 * tsdf_b200_device_free), in the reference's order: ascending cube index (x fastest), triangle-table order within
 * a cube, three consecutive vertices per triangle.  Synchronises `stream`.                                   */
int tsdf_b200_mc_extract(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz_planes, uint32_t z_base,
                         uint32_t cz_begin, uint32_t cz_end, const float voxel[3], const float offset[3],
                         float **d_vertices_out, unsigned long long *n_vertices_out, void *stream);
void tsdf_b200_device_free(void *d_ptr);
/* cudaMemcpy device -> host for callers that do not link the CUDA runtime themselves. */
int tsdf_b200_copy_to_host(void *host, const void *device, size_t bytes);

/* Replaces BilateralFilter::filter_bpp (BilateralFilter.cpp:53-121), which is host code in the reference.  kernel:
 * kernel_size^2 spatial weights, similarity: n_similarity range weights — the look-up tables the reference's
 * constructor builds (:15-42); the caller computes them on the host so that the device evaluates no transcendental
 * and the result equals the host loop bit for bit.  d_out != d_in.  8-bit: the reference exactly; 16-bit: the
 * reference is undefined behaviour there (see csrc/bilateral.cu), the table may cover all 65536 differences.    */
int tsdf_b200_bilateral_u8(const uint8_t *d_in, uint8_t *d_out, uint32_t width, uint32_t height, const float *d_kernel,
                           uint32_t kernel_size, const float *d_similarity, uint32_t n_similarity, void *stream);
int tsdf_b200_bilateral_u16(const uint16_t *d_in, uint16_t *d_out, uint32_t width, uint32_t height, const float *d_kernel,
                            uint32_t kernel_size, const float *d_similarity, uint32_t n_similarity, void *stream);
/* In-place filtering of a HOST image (bits_per_pixel 8 or 16) with HOST tables: BilateralFilter::filter. */
int tsdf_b200_bilateral_host(void *host_image, int bits_per_pixel, uint32_t width, uint32_t height, const float *host_kernel,
                             uint32_t kernel_size, const float *host_similarity, uint32_t n_similarity);

/* Exhaustive check (every numerator bit pattern with 2^-100 <= |a| <= 2^100, and +-0) that the
 * 3-instruction reciprocal division used by the raycast kernel equals IEEE a/divisor.  *mismatches is a
 * host pointer.  The level-2 volume runs this once per voxel size and falls back to
 * IEEE division in the kernel when it is not zero.                                      */
int tsdf_b200_selftest_division(float divisor, unsigned long long *mismatches);

/* ------------------------------------------------------------------------------------
 * Level 2 — object API with HOST buffers (what kinfu.cpp reaches through the classes).
 * ---------------------------------------------------------------------------------- */
typedef struct tsdf_b200_volume tsdf_b200_volume;

/* TSDFVolume(UInt3, Float3) / set_size (TSDF/TSDFVolume.cu:430-437, 679-722). */
int tsdf_b200_volume_create(uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz,
                            tsdf_b200_volume **out);
/* The same volume sharded along Z over `ngpus` GPUs of the box, inside this process (csrc/multi.cu): GPU r owns a slab of
 * whole 8-voxel bricks plus one redundant halo plane; integrate needs no communication, raycast min-merges the per-slab
 * hits into GPU 0's key map with atomics over NVLink peer memory, marching cubes runs per slab.  Every entry point below
 * works on such a volume and returns the same bits as on one GPU; exceptions: the deformation grid is not distributed
 * (tsdf_b200_volume_deformation returns NULL, _set_deformation TSDF_B200_ESTATE), and distance_data() / weight_data()
 * gather the slabs into a full-size array on GPU 0 at every call.  tsdf_b200_volume_create itself creates a sharded volume
 * when the environment variable TSDF_NGPUS is > 1 — that is how the unchanged kinfu.cpp uses several GPUs.  Fewer GPUs or
 * brick layers than asked for: as many slabs as fit (one = the single-GPU volume).                                   */
int tsdf_b200_volume_create_sharded(uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz, int ngpus,
                                    tsdf_b200_volume **out);
/* GPUs the volume lives on (1 = not sharded). */
int tsdf_b200_volume_gpus(const tsdf_b200_volume *v);
/* TSDFVolume(const std::string&) load constructor (TSDF/TSDFVolume.cu:463-664). */
int tsdf_b200_volume_load(const char *path, tsdf_b200_volume **out);
void tsdf_b200_volume_destroy(tsdf_b200_volume *v);

int tsdf_b200_volume_get(const tsdf_b200_volume *v, uint32_t size[3], float physical[3],
                         float voxel[3], float offset[3], float *trunc, float *max_weight);
/* global_translation() / global_rotation() (include/TSDFVolume.hpp:216,223): zero unless the volume was loaded from a
 * .tsdf file whose header carries them (TSDF/TSDFVolume.cu:496-497). */
int tsdf_b200_volume_get_global(const tsdf_b200_volume *v, float translation[3], float rotation[3]);
/* TSDFVolume::offset(ox,oy,oz) (include/TSDFVolume.hpp:144-148). */
int tsdf_b200_volume_set_offset(tsdf_b200_volume *v, float ox, float oy, float oz);
/* TSDFVolume::clear (TSDF/TSDFVolume.cu:812-845). */
int tsdf_b200_volume_clear(tsdf_b200_volume *v);
/* distance_data()/weight_data(): raw device pointers.  The volume works on a blocking stream of its own, so kernels and
 * copies a caller issues on the legacy default stream are ordered against the volume's calls in both directions. */
const float *tsdf_b200_volume_distance_data(const tsdf_b200_volume *v);
const float *tsdf_b200_volume_weight_data(const tsdf_b200_volume *v);
/* deformation(): materialises the 24 B/voxel node array on first use. */
float *tsdf_b200_volume_deformation(tsdf_b200_volume *v);
/* set_distance_data / set_weight_data / set_deformation (TSDF/TSDFVolume.cu:729-755). */
int tsdf_b200_volume_set_distance_data(tsdf_b200_volume *v, const float *host);
int tsdf_b200_volume_set_weight_data(tsdf_b200_volume *v, const float *host);
int tsdf_b200_volume_set_deformation(tsdf_b200_volume *v, const float *host_nodes);
/* Device -> host copies for tests and tools. */
int tsdf_b200_volume_read(const tsdf_b200_volume *v, float *host_dist, float *host_weight);

/* Host memory for buffers that cross the bus every frame (depth maps, vertex / normal maps).  Blocks of 256 KiB and more come
 * from a pool of pinned (page-locked, device-addressable) memory and go back to it when freed — a caller that allocates its
 * result buffers per frame, as kinfu.cpp does with its Eigen matrices, pays for the pinning once; smaller blocks, and any block
 * when there is no CUDA device, come from malloc.  tsdf_b200_volume_raycast recognises pinned buffers by itself.            */
void *tsdf_b200_host_alloc(size_t bytes);
void tsdf_b200_host_free(void *p);

/* TSDFVolume::integrate (TSDF/TSDFVolume.cu:861-902): host depth map, camera matrices as
 * Camera::inverse_pose()/k()/kinv() .data().  Returns when the depth map has been read; the fusion
 * completes in stream order before any later call on the volume returns data (as if synchronous;
 * TSDF_B200_SYNC=1 waits for the kernels in this call).                                   */
int tsdf_b200_volume_integrate(tsdf_b200_volume *v, const uint16_t *host_depth, uint32_t width,
                               uint32_t height, const float inv_pose[16], const float k[9],
                               const float kinv[9]);
/* TSDFVolume::raycast -> GPURaycaster::raycast (TSDF/TSDFVolume.cu:1054-1058,
 * RayCaster/GPURaycaster.cu:519-547): pose = Camera::pose().data().  host_vertices /
 * host_normals receive 3*width*height floats each.  Synchronous.                        */
int tsdf_b200_volume_raycast(const tsdf_b200_volume *v, uint32_t width, uint32_t height,
                             const float pose[16], const float kinv[9], float *host_vertices,
                             float *host_normals);
/* TSDFVolume::save_to_file (TSDF/TSDFVolume.cu:911-1027), byte-compatible format. */
int tsdf_b200_volume_save(const tsdf_b200_volume *v, const char *path);

/* extract_surface_ms (MarchingCubes/MarkAndSweepMC.cu:390-497) of the volume: *d_vertices_out receives a cudaMalloc'ed array
 * of 3 floats per vertex on the volume's (first) GPU, in the reference's order (tsdf_b200_mc_extract); release it with
 * tsdf_b200_device_free.  A sharded volume extracts per slab and concatenates.                                    */
int tsdf_b200_volume_extract_mesh(const tsdf_b200_volume *v, float **d_vertices_out, unsigned long long *n_vertices_out);

/* Counters of the last integrate / raycast (voxels rewritten, samples evaluated). */
int tsdf_b200_volume_stats(const tsdf_b200_volume *v, unsigned long long *n_updated,
                           unsigned long long *n_samples);
/* 0 disables / 1 enables empty-space skipping in volume_raycast (default 1). */
int tsdf_b200_volume_set_skipping(tsdf_b200_volume *v, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* TSDF_B200_H */
