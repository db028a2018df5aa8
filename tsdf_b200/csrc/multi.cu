// multi.cu — the level-2 volume sharded along Z over the GPUs of one box, inside ONE process (TSDF_NGPUS / create_sharded).
//
// The reference is single-GPU (src/TSDF/TSDFVolume.cu, one device, default stream); this is the coordinator that lets an
// unchanged caller of the class surface — kinfu.cpp through TSDFVolume::integrate / ::raycast — use several B200s:
//   * GPU r owns the planes [z0, z1) of the volume (whole 8-voxel bricks) plus one redundant halo plane z1, fused by the
//     same integrate kernel: every voxel depends only on itself and the frame, so the copy is bit-identical to its owner's
//     and no voxel ever crosses NVLink;
//   * raycast: every GPU marches all rays through the samples whose interpolation cell starts in its slab
//     (tsdf_b200_raycast_slab_min) and min-merges the key (k_hit << 32 | sample bits) of each hit straight into GPU 0's key
//     map with 64-bit atomics over NVLink peer memory — the exchange is fused into the march, there is no collective;
//     GPU 0 waits for the other GPUs' march events, resolves the keys to vertices (the sample parameters t_k are
//     ray-independent, so the winning key reproduces the single-GPU vertex bit for bit), computes the normals and copies
//     both maps to the caller;
//   * marching cubes runs per slab; the parts are concatenated in slab order on GPU 0 (the reference's cube order).
// One worker thread per GPU issues that GPU's launches, so the host-side launch cost does not add up over the GPUs; a
// call returns when every GPU is done, like the reference's synchronous methods.
#include "volume_internal.h"
#include <atomic>
#include <new>

namespace tsdf {

namespace {

size_t plane_elems(const tsdf_b200_volume *v) { return (size_t)v->nx * v->ny; }

void worker_main(Multi *M, int index) {
    cudaSetDevice(M->shards[index].dev);
    unsigned long long seen = 0;
    for (;;) {
        std::function<int(Shard &, int)> job;
        {
            std::unique_lock<std::mutex> lk(M->m);
            M->cv_go.wait(lk, [&] { return M->quit || M->generation != seen; });
            if (M->quit) return;
            seen = M->generation;
            job = M->job;
        }
        const int rc = job(M->shards[index], index);
        {
            std::lock_guard<std::mutex> lk(M->m);
            M->shards[index].rc = rc;
            if (--M->pending == 0) M->cv_done.notify_all();
        }
    }
}

// Runs `job` for every shard on that shard's worker thread (its GPU is current there); first non-zero result wins.
int run_all(Multi *M, const std::function<int(Shard &, int)> &job) {
    {
        std::lock_guard<std::mutex> lk(M->m);
        M->job = job;
        M->pending = (int)M->shards.size();
        M->generation++;
    }
    M->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lk(M->m);
        M->cv_done.wait(lk, [&] { return M->pending == 0; });
    }
    for (const Shard &s : M->shards) if (s.rc) return s.rc;
    return 0;
}

void free_shard(Shard &s) {
    cudaSetDevice(s.dev);
    cudaFree(s.d_dist); cudaFree(s.d_weight); cudaFree(s.d_occ); cudaFree(s.d_table); cudaFree(s.d_depth);
    cudaFree(s.d_staged); cudaFree(s.d_counters);
    if (s.ev_march) cudaEventDestroy(s.ev_march);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Shard();
}

}  // namespace

int multi_create(tsdf_b200_volume *v, int ngpus) {
    int have = 0;
    TSDF_CUDA_TRY(cudaGetDeviceCount(&have));
    if (ngpus > have) ngpus = have;
    // whole bricks per slab, like tsdf_b200/sharded.py::shard_ranges
    const uint32_t bricks = (v->nz + TSDF_B200_BRICK - 1) / TSDF_B200_BRICK;
    if ((uint32_t)ngpus > bricks) ngpus = (int)bricks;
    if (ngpus < 2) return TSDF_B200_EINVAL;
    const uint32_t per = (bricks + (uint32_t)ngpus - 1) / (uint32_t)ngpus;
    while (ngpus > 1 && (uint32_t)(ngpus - 1) * per * TSDF_B200_BRICK >= v->nz) ngpus--;      // no empty slab at the end
    if (ngpus < 2) return TSDF_B200_EINVAL;
    int dev0 = 0;
    TSDF_CUDA_TRY(cudaGetDevice(&dev0));
    Multi *M = new (std::nothrow) Multi();
    if (!M) return TSDF_B200_ENOMEM;
    v->multi = M;
    v->device = 0;
    M->shards.resize((size_t)ngpus);
    int rc = 0;
    for (int r = 0; r < ngpus && !rc; r++) {
        Shard &s = M->shards[(size_t)r];
        s.dev = r;
        s.z0 = (uint32_t)r * per * TSDF_B200_BRICK;
        s.z1 = s.z0 + per * TSDF_B200_BRICK < v->nz ? s.z0 + per * TSDF_B200_BRICK : v->nz;
        s.zs1 = s.z1 < v->nz ? s.z1 + 1 : v->nz;
        const size_t n = plane_elems(v) * (s.zs1 - s.z0);
        cudaError_t e = cudaSetDevice(r);
        if (e == cudaSuccess && r > 0) {
            // the march of GPU r writes its hits into GPU 0's key map
            int ok = 0;
            e = cudaDeviceCanAccessPeer(&ok, r, 0);
            if (e == cudaSuccess && !ok) { rc = TSDF_B200_ESTATE; break; }
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(0, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { e = cudaSuccess; cudaGetLastError(); }
            }
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ev_march, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_dist, n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_weight, n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_occ, tsdf_b200_occupancy_bytes(v->nx, v->ny, s.zs1 - s.z0));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_table, TSDF_B200_RAY_TABLE_LEN * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&s.d_counters, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemsetAsync(s.d_counters, 0, 2 * sizeof(unsigned long long), s.stream);
        if (e != cudaSuccess) { rc = (int)e; break; }
        rc = tsdf_b200_ray_table(v->trunc, s.d_table, s.stream);
        if (!rc) rc = (int)cudaStreamSynchronize(s.stream);
    }
    cudaSetDevice(0);
    if (!rc) {
        // prove the reciprocal division for this volume's voxel sizes once (the GPUs are identical)
        v->fastdiv = 1;
        for (int a = 0; a < 3 && v->fastdiv && !rc; a++) {
            bool seen = false;
            for (int b = 0; b < a; b++) seen |= (v->vs[b] == v->vs[a]);
            if (seen) continue;
            unsigned long long bad = 1;
            rc = tsdf_b200_selftest_division(v->vs[a], &bad);
            if (bad) v->fastdiv = 0;
        }
    }
    if (rc) {
        for (Shard &s : M->shards) free_shard(s);
        delete M;
        v->multi = nullptr;
        cudaSetDevice(dev0);
        return rc;
    }
    v->stream = M->shards[0].stream;
    for (int r = 0; r < ngpus; r++) M->workers.emplace_back(worker_main, M, r);
    return 0;
}

void multi_destroy(tsdf_b200_volume *v) {
    Multi *M = v->multi;
    if (!M) return;
    {
        std::lock_guard<std::mutex> lk(M->m);
        M->quit = true;
    }
    M->cv_go.notify_all();
    for (std::thread &t : M->workers) t.join();
    for (Shard &s : M->shards) free_shard(s);
    cudaSetDevice(v->device);
    cudaFree(M->d_keys[0]); cudaFree(M->d_keys[1]); cudaFree(M->d_full_dist); cudaFree(M->d_full_weight);
    delete M;
    v->multi = nullptr;
    v->stream = nullptr;
}

int multi_clear(tsdf_b200_volume *v) {
    return run_all(v->multi, [v](Shard &s, int) -> int {
        int rc = tsdf_b200_clear(s.d_dist, s.d_weight, v->nx, v->ny, s.zs1 - s.z0, v->trunc, s.d_occ, s.stream);
        if (rc) return rc;
        return (int)cudaStreamSynchronize(s.stream);
    });
}

int multi_integrate(tsdf_b200_volume *v, const uint16_t *host_depth, uint32_t width, uint32_t height, const float inv_pose[16],
                    const float k[9], const float kinv[9]) {
    Multi *M = v->multi;
    const size_t npix = (size_t)width * height;
    const size_t staged_bytes = tsdf_b200_depth_staged_bytes(width, height);
    int rc = run_all(M, [=](Shard &s, int) -> int {
        if (npix > s.depth_cap) {
            cudaFree(s.d_depth); s.d_depth = nullptr; s.depth_cap = 0;
            TSDF_CUDA_TRY(cudaMalloc(&s.d_depth, npix * sizeof(uint16_t)));
            s.depth_cap = npix;
        }
        if (staged_bytes > s.staged_cap) {
            cudaFree(s.d_staged); s.d_staged = nullptr; s.staged_cap = 0;
            TSDF_CUDA_TRY(cudaMalloc(&s.d_staged, staged_bytes));
            s.staged_cap = staged_bytes;
        }
        // every GPU needs the whole frame (614 KB): its own copy from the caller's buffer
        TSDF_CUDA_TRY(cudaMemcpyAsync(s.d_depth, host_depth, npix * sizeof(uint16_t), cudaMemcpyHostToDevice, s.stream));
        int rc = tsdf_b200_depth_stage(s.d_depth, width, height, s.d_staged, s.stream);
        if (rc) return rc;
        if (v->counting) TSDF_CUDA_TRY(cudaMemsetAsync(s.d_counters, 0, sizeof(unsigned long long), s.stream));
        const uint32_t own = s.z1 - s.z0, stored = s.zs1 - s.z0;
        rc = tsdf_b200_integrate(s.d_dist, s.d_weight, nullptr, v->nx, v->ny, stored, v->vs, v->off_clear, v->off, v->trunc, inv_pose, k,
                                 kinv, width, height, s.d_depth, s.d_staged, 0, own, s.z0, s.d_occ, v->counting ? s.d_counters : nullptr,
                                 s.stream);
        if (rc) return rc;
        if (stored > own) {            // the redundant halo plane (not counted: its owner counts it)
            rc = tsdf_b200_integrate(s.d_dist, s.d_weight, nullptr, v->nx, v->ny, stored, v->vs, v->off_clear, v->off, v->trunc, inv_pose,
                                     k, kinv, width, height, s.d_depth, s.d_staged, own, stored, s.z0, s.d_occ, nullptr, s.stream);
            if (rc) return rc;
        }
        if (v->counting)
            TSDF_CUDA_TRY(cudaMemcpyAsync(&s.h_counters[0], s.d_counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        return (int)cudaStreamSynchronize(s.stream);
    });
    if (rc) return rc;
    v->h_counters[0] = 0;
    for (const Shard &s : M->shards) v->h_counters[0] += s.h_counters[0];
    return 0;
}

int multi_raycast(tsdf_b200_volume *v, uint32_t width, uint32_t height, const float pose[16], const float kinv[9],
                  float *host_vertices, float *host_normals) {
    Multi *M = v->multi;
    const size_t npix = (size_t)width * height;
    Shard &s0 = M->shards[0];
    TSDF_CUDA_TRY(cudaSetDevice(s0.dev));
    if (npix > M->keys_cap) {
        cudaFree(M->d_keys[0]); M->d_keys[0] = nullptr; M->keys_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&M->d_keys[0], npix * sizeof(long long)));
        M->keys_cap = npix;
        int rc = tsdf_b200_fill_i64(M->d_keys[0], npix, 0x7fffffffffffffffLL, s0.stream);
        if (rc) return rc;
        TSDF_CUDA_TRY(cudaStreamSynchronize(s0.stream));
    }
    if (npix > v->pix_cap) {
        cudaFree(v->d_vn); v->d_vn = nullptr; v->pix_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&v->d_vn, npix * 6 * sizeof(float)));
        v->pix_cap = npix;
    }
    float *d_vert = v->d_vn, *d_norm = v->d_vn + 3 * npix;
    long long *keys = M->d_keys[0];
    const float origin[3] = { pose[12], pose[13], pose[14] };
    const float rot[9] = { pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10] };
    float smin[3], smax[3];
    for (int i = 0; i < 3; i++) { smin[i] = v->off[i]; smax[i] = v->off[i] + v->phys[i]; }
    std::atomic<int> marched(0);
    const int n = (int)M->shards.size();
    int rc = run_all(M, [&, d_vert, d_norm, keys](Shard &s, int r) -> int {
        struct Arrive {             // the rendezvous below must see every shard, also one that failed to launch
            std::atomic<int> &c; bool done = false;
            void now() { if (!done) { done = true; c.fetch_add(1); } }
            ~Arrive() { now(); }
        } arrive{marched};
        cudaError_t e = cudaSuccess;
        if (v->counting) e = cudaMemsetAsync(s.d_counters + 1, 0, sizeof(unsigned long long), s.stream);
        int rc = (int)e;
        if (!rc) rc = tsdf_b200_raycast_slab_min(s.d_dist, v->nx, v->ny, v->nz, s.z0, s.zs1 - s.z0, s.z0, s.z1, v->vs, smin, smax, v->trunc,
                                                 origin, rot, kinv, width, height, s.d_table, v->skipping ? s.d_occ : nullptr, keys,
                                                 v->counting ? s.d_counters + 1 : nullptr, v->fastdiv, s.stream);
        if (!rc) rc = (int)cudaEventRecord(s.ev_march, s.stream);
        arrive.now();
        if (r == 0 && !rc) {
            // every other GPU's march event has been RECORDED (host side) before this GPU's stream is told to wait for it
            while (marched.load() < n) std::this_thread::yield();
            for (int q = 1; q < n && !rc; q++) rc = (int)cudaStreamWaitEvent(s.stream, M->shards[(size_t)q].ev_march, 0);
            if (!rc) rc = tsdf_b200_raycast_resolve_reset(keys, smin, smax, v->trunc, origin, rot, kinv, width, height, s.d_table, d_vert,
                                                          nullptr, s.stream);
            if (!rc) rc = tsdf_b200_normals(width, height, d_vert, d_norm, s.stream);
            if (!rc) rc = (int)cudaMemcpyAsync(host_vertices, d_vert, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, s.stream);
            if (!rc) rc = (int)cudaMemcpyAsync(host_normals, d_norm, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, s.stream);
        }
        if (!rc && v->counting)
            rc = (int)cudaMemcpyAsync(&s.h_counters[1], s.d_counters + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream);
        const int rs = (int)cudaStreamSynchronize(s.stream);
        return rc ? rc : rs;
    });
    if (rc) return rc;
    v->h_counters[1] = 0;
    for (const Shard &s : M->shards) v->h_counters[1] += s.h_counters[1];
    return 0;
}

int multi_read(const tsdf_b200_volume *v, float *host_dist, float *host_weight) {
    const size_t pe = plane_elems(v);
    return run_all(v->multi, [=](Shard &s, int) -> int {
        const size_t n = pe * (s.z1 - s.z0);
        if (host_dist) TSDF_CUDA_TRY(cudaMemcpyAsync(host_dist + pe * s.z0, s.d_dist, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        if (host_weight) TSDF_CUDA_TRY(cudaMemcpyAsync(host_weight + pe * s.z0, s.d_weight, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        return (int)cudaStreamSynchronize(s.stream);
    });
}

int multi_write(tsdf_b200_volume *v, const float *host_dist, const float *host_weight) {
    const size_t pe = plane_elems(v);
    return run_all(v->multi, [=](Shard &s, int) -> int {
        const size_t n = pe * (s.zs1 - s.z0);              // owned planes and the halo plane
        if (host_dist) {
            TSDF_CUDA_TRY(cudaMemcpyAsync(s.d_dist, host_dist + pe * s.z0, n * sizeof(float), cudaMemcpyHostToDevice, s.stream));
            int rc = tsdf_b200_occupancy_rebuild(s.d_dist, v->nx, v->ny, s.zs1 - s.z0, v->trunc, s.d_occ, s.stream);
            if (rc) return rc;
        }
        if (host_weight) TSDF_CUDA_TRY(cudaMemcpyAsync(s.d_weight, host_weight + pe * s.z0, n * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        return (int)cudaStreamSynchronize(s.stream);
    });
}

int multi_gather_device(tsdf_b200_volume *v) {
    Multi *M = v->multi;
    const size_t pe = plane_elems(v), n = pe * v->nz;
    TSDF_CUDA_TRY(cudaSetDevice(M->shards[0].dev));
    if (!M->d_full_dist) TSDF_CUDA_TRY(cudaMalloc(&M->d_full_dist, n * sizeof(float)));
    if (!M->d_full_weight) TSDF_CUDA_TRY(cudaMalloc(&M->d_full_weight, n * sizeof(float)));
    const int dev0 = M->shards[0].dev;
    float *fd = M->d_full_dist, *fw = M->d_full_weight;
    return run_all(M, [=](Shard &s, int) -> int {
        const size_t m = pe * (s.z1 - s.z0);
        TSDF_CUDA_TRY(cudaMemcpyPeerAsync(fd + pe * s.z0, dev0, s.d_dist, s.dev, m * sizeof(float), s.stream));
        TSDF_CUDA_TRY(cudaMemcpyPeerAsync(fw + pe * s.z0, dev0, s.d_weight, s.dev, m * sizeof(float), s.stream));
        return (int)cudaStreamSynchronize(s.stream);
    });
}

int multi_extract_mesh(tsdf_b200_volume *v, float **d_vertices_out, unsigned long long *n_vertices_out) {
    Multi *M = v->multi;
    const size_t ns = M->shards.size();
    std::vector<float *> parts(ns, nullptr);
    std::vector<unsigned long long> counts(ns, 0);
    int rc = run_all(M, [&](Shard &s, int r) -> int {
        // cubes based in the owned planes; the halo plane closes the cubes at the slab's upper face
        return tsdf_b200_mc_extract(s.d_dist, v->nx, v->ny, s.zs1 - s.z0, s.z0, 0, s.z1 - s.z0, v->vs, v->off, &parts[(size_t)r],
                                    &counts[(size_t)r], s.stream);
    });
    unsigned long long total = 0;
    for (unsigned long long c : counts) total += c;
    float *out = nullptr;
    const int dev0 = M->shards[0].dev;
    cudaSetDevice(dev0);
    if (!rc && total) {
        cudaError_t e = cudaMalloc(&out, total * 3 * sizeof(float));
        unsigned long long at = 0;
        for (size_t r = 0; r < ns && e == cudaSuccess; r++) {
            if (counts[r]) e = cudaMemcpyPeer(out + 3 * at, dev0, parts[r], M->shards[r].dev, counts[r] * 3 * sizeof(float));
            at += counts[r];
        }
        if (e != cudaSuccess) { cudaFree(out); out = nullptr; rc = (int)e; }
    }
    for (float *p : parts) if (p) cudaFree(p);
    if (rc) return rc;
    *d_vertices_out = out;
    *n_vertices_out = total;
    return 0;
}

}  // namespace tsdf
