// volume_internal.h — the level-2 volume object shared by volume.cu (single GPU) and multi.cu (Z-slabs over the GPUs of a box).
#pragma once
#include "common.cuh"
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace tsdf { struct Multi; }

struct tsdf_b200_volume {
    uint32_t nx = 0, ny = 0, nz = 0;
    float phys[3] = {0, 0, 0}, vs[3] = {0, 0, 0};
    float off[3] = {0, 0, 0};          // m_offset
    float off_clear[3] = {0, 0, 0};    // m_offset at the time clear() wrote the deformation grid
    float trunc = 0, max_weight = 15.0f;
    float gtrans[3] = {0, 0, 0}, grot[3] = {0, 0, 0};
    float *d_dist = nullptr, *d_weight = nullptr;
    float *d_deform = nullptr;         // 6 floats / voxel, lazily materialised
    bool deform_identity = true;       // d_deform (if any) equals the grid clear() would write
    uint8_t *h_colour = nullptr;       // colours are never touched on the hot path; host copy only when loaded
    uint8_t *d_occ = nullptr;
    float *d_table = nullptr;
    uint16_t *d_depth = nullptr; size_t depth_cap = 0;
    float *d_staged = nullptr; size_t staged_cap = 0;   // staged depth frame (tsdf_b200_depth_stage)
    float *d_vn = nullptr; size_t pix_cap = 0;   // vertices then normals
    unsigned int *d_tiles = nullptr; size_t tile_cap = 0;   // per-tile counters of the fused raycast (zero between launches)
    unsigned long long *d_counters = nullptr;    // [0] voxels rewritten, [1] samples
    unsigned long long h_counters[2] = {0, 0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_depth = nullptr;    // the depth map of the last integrate has left the caller's buffer
    bool counters_stale = false;       // d_counters are ahead of h_counters (fetched on demand by volume_stats)
    int fastdiv = 0, skipping = 1, counting = 1;
    int device = 0;                    // the device the single-GPU arrays (and, when sharded, the merged results) live on
    tsdf::Multi *multi = nullptr;      // non-null: the volume is sharded along Z over several GPUs (multi.cu)
};

namespace tsdf {

// One Z-slab of a sharded volume: planes [z0, z1) owned, [z0, zs1) stored (one redundant halo plane, fused by the same
// integrate kernel, so that every trilinear cell / marching cube that starts in an owned plane is complete locally).
struct Shard {
    int dev = 0;
    uint32_t z0 = 0, z1 = 0, zs1 = 0;
    float *d_dist = nullptr, *d_weight = nullptr;
    uint8_t *d_occ = nullptr;
    float *d_table = nullptr;
    uint16_t *d_depth = nullptr; size_t depth_cap = 0;
    float *d_staged = nullptr; size_t staged_cap = 0;
    unsigned long long *d_counters = nullptr;
    unsigned long long h_counters[2] = {0, 0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_march = nullptr;
    int rc = 0;
};

// Worker threads, one per GPU: a call is handed to all of them and returns when every one is done.
struct Multi {
    std::vector<Shard> shards;
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::function<int(Shard &, int)> job;
    unsigned long long generation = 0;
    int pending = 0;
    bool quit = false;
    // shard 0 holds the merged results: two key maps (alternating frames), vertices + normals
    long long *d_keys[2] = {nullptr, nullptr}; size_t keys_cap = 0;
    unsigned frame = 0;
    float *d_full_dist = nullptr, *d_full_weight = nullptr;    // distance_data() / weight_data() gathered on demand
    int marched = 0;                   // host-side rendezvous of the raycast: how many shards have recorded ev_march
};

int multi_create(tsdf_b200_volume *v, int ngpus);
void multi_destroy(tsdf_b200_volume *v);
int multi_clear(tsdf_b200_volume *v);
int multi_integrate(tsdf_b200_volume *v, const uint16_t *host_depth, uint32_t width, uint32_t height, const float inv_pose[16],
                    const float k[9], const float kinv[9]);
int multi_raycast(tsdf_b200_volume *v, uint32_t width, uint32_t height, const float pose[16], const float kinv[9],
                  float *host_vertices, float *host_normals);
int multi_read(const tsdf_b200_volume *v, float *host_dist, float *host_weight);
int multi_write(tsdf_b200_volume *v, const float *host_dist, const float *host_weight);
int multi_gather_device(tsdf_b200_volume *v);     // fills d_full_dist / d_full_weight on shard 0's device
int multi_extract_mesh(tsdf_b200_volume *v, float **d_vertices_out, unsigned long long *n_vertices_out);

}  // namespace tsdf
