// GPURaycaster.hpp — the CUDA raycaster of the reference (src/include/GPURaycaster.hpp:19-41), forwarding to
// tsdf_b200_volume_raycast (sm_100a kernels in tsdf_b200/csrc/raycast.cu).
#ifndef GPURaycaster_hpp
#define GPURaycaster_hpp

#include <Eigen/Core>
#include "Raycaster.hpp"
#include "TSDFVolume.hpp"
#include "DepthImage.hpp"

class GPURaycaster : public Raycaster {
public:
    GPURaycaster(int width = 640, int height = 480) : Raycaster{width, height} {}

    virtual void raycast(const TSDFVolume &volume, const Camera &camera, Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                         Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const;

    // Camera-space z of every raycast vertex, rounded to u16 millimetres (GPURaycaster.cu:555-606; the reference's
    // debug PNG written to the author's desktop is not reproduced).  The caller owns the result.
    DepthImage *render_to_depth_image(const TSDFVolume &volume, const Camera &camera) const;
};
#endif /* GPURaycaster_hpp */
