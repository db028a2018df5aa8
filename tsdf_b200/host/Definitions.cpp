// Definitions.cpp — reference src/Utilities/Definitions.cpp.
#include "../include/Definitions.hpp"
#include <limits>

const Eigen::Vector3f BAD_VERTEX{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
