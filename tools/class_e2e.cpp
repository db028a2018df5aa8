// class_e2e — frames per second through the drop-in C++ classes, driven the way the reference's kinfu.cpp drives them
// (src/Tools/kinfu.cpp:150-190): per frame a DepthImage, a camera pose, TSDFVolume::integrate, two Eigen result matrices
// declared INSIDE the loop, TSDFVolume::raycast.  Input: a file written by bench.py —
//   uint32 size, width, height, n_frames; float physical; float k[9] (row-major); then per frame float pose[16] (row-major)
//   and width*height uint16 depth.
// Output: one JSON object on stdout.  (Decoding PNGs is not part of the measurement: the DepthImages are built before the clock
// starts, one per frame, as the TUM loader would hand them over.)
#include "../tsdf_b200/include/Camera.hpp"
#include "../tsdf_b200/include/DepthImage.hpp"
#include "../tsdf_b200/include/TSDFVolume.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: class_e2e frames.bin warmup\n"); return 2; }
    std::FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const int warmup = std::atoi(argv[2]);
    uint32_t hdr[4];
    float physical, k[9];
    if (std::fread(hdr, 4, 4, f) != 4 || std::fread(&physical, 4, 1, f) != 1 || std::fread(k, 4, 9, f) != 9) return 2;
    const uint32_t size = hdr[0], width = hdr[1], height = hdr[2], n_frames = hdr[3];
    std::vector<Eigen::Matrix4f> poses(n_frames);
    std::vector<std::unique_ptr<DepthImage>> images;
    std::vector<uint16_t> px((size_t)width * height);
    for (uint32_t i = 0; i < n_frames; i++) {
        float p[16];
        if (std::fread(p, 4, 16, f) != 16 || std::fread(px.data(), 2, px.size(), f) != px.size()) return 2;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) poses[i](r, c) = p[4 * r + c];
        images.emplace_back(new DepthImage((uint16_t)width, (uint16_t)height, px.data()));
    }
    std::fclose(f);
    Eigen::Matrix3f K;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K(r, c) = k[3 * r + c];
    Camera camera(K);
    TSDFVolume volume(TSDFVolume::UInt3{size, size, size}, TSDFVolume::Float3{physical, physical, physical});

    double checksum = 0;
    size_t hits = 0;
    std::chrono::steady_clock::time_point t0;
    for (uint32_t i = 0; i < n_frames; i++) {
        if ((int)i == warmup) t0 = std::chrono::steady_clock::now();
        camera.set_pose(poses[i]);
        volume.integrate(images[i]->data(), width, height, camera);
        Eigen::Matrix<float, 3, Eigen::Dynamic> vertices;
        Eigen::Matrix<float, 3, Eigen::Dynamic> normals;
        volume.raycast((uint16_t)width, (uint16_t)height, camera, vertices, normals);
        const float z = vertices(2, (size_t)(height / 2) * width + width / 2);
        if (z == z) { checksum += z; hits++; }
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int timed = (int)n_frames - warmup;
    std::printf("{\"frames\": %d, \"ms_per_frame\": %.6f, \"frames_per_s\": %.3f, \"centre_hits\": %zu, \"centre_z_sum\": %.3f}\n",
                timed, 1e3 * s / timed, timed / s, hits, checksum);
    return 0;
}
