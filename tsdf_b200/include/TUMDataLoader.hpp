// TUMDataLoader.hpp — TUM RGB-D sequence reader (reference src/include/TUMDataLoader.hpp:9-49):
// <dir>/ground_truth.txt lines "timestamp tx ty tz qx qy qz qw", frames in <dir>/depth/<timestamp>.png.
#ifndef TUM_DATA_LOADER_H
#define TUM_DATA_LOADER_H

#include "DepthImage.hpp"

#include <string>
#include <vector>
#include <Eigen/Dense>

class TUMDataLoader {
public:
    TUMDataLoader(const std::string &directory);
    ~TUMDataLoader();
    // Next frame (depth scaled to millimetres) and its pose (translation in millimetres); nullptr when exhausted
    // or when the frame's file is missing.  The caller owns the image.
    DepthImage *next(Eigen::Matrix4f &pose);

private:
    struct DATA_RECORD {
        std::string file_name;
        float data[7];
    };
    Eigen::Matrix4f to_pose(float vars[7]) const;
    void process_line(const std::string &line);
    void load_data_from(const std::string &gt_file_name);

    size_t m_current_idx;
    std::vector<struct DATA_RECORD> m_data_records;
    std::string m_directory_name;
};
#endif
