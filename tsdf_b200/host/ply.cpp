// ply.cpp — reference src/Utilities/ply.cpp.
#include "../include/ply.hpp"

#include <fstream>
#include <iostream>

void write_to_ply(const std::string &file_name, const std::vector<float3> &vertices, const std::vector<int3> &triangles) {
    std::ofstream f{file_name};
    if (!f.is_open()) {
        std::cout << "Problem opening file for write " << file_name << std::endl;
        return;
    }
    f << "ply\nformat ascii 1.0\n";
    f << "element vertex " << vertices.size() << "\n";
    f << "property float x\nproperty float y\nproperty float z\n";
    f << "element face " << triangles.size() << "\n";
    f << "property list uchar int vertex_indices\nend_header\n";
    for (const float3 &v : vertices) f << v.x << " " << v.y << " " << v.z << "\n";
    for (const int3 &t : triangles) f << "3 " << t.x << " " << t.y << " " << t.z << "\n";
}
