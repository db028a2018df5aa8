// Definitions.hpp — shared constants (reference src/include/Definitions.hpp).
#ifndef Definitions_hpp
#define Definitions_hpp
#include <Eigen/Core>

extern const Eigen::Vector3f BAD_VERTEX;        // (FLT_MAX, FLT_MAX, FLT_MAX): "no depth here"
#endif /* Definitions_hpp */
