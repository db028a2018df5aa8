// Raycaster.hpp — abstract raycaster, same surface as reference src/include/Raycaster.hpp:17-39.
#ifndef Raycaster_hpp
#define Raycaster_hpp

#include <Eigen/Core>
#include "TSDFVolume.hpp"
#include "Camera.hpp"

class Raycaster {
public:
    Raycaster(int width = 640, int height = 480) : m_width(static_cast<uint16_t>(width)), m_height(static_cast<uint16_t>(height)) {}
    virtual ~Raycaster() {}

    // vertices / normals are resized to 3 x (width*height); a pixel whose ray meets no surface gets a NaN vertex
    virtual void raycast(const TSDFVolume &volume, const Camera &camera, Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                         Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const = 0;

protected:
    uint16_t m_width;
    uint16_t m_height;
};
#endif /* Raycaster_hpp */
