// TUMDataLoader.cpp — reference src/DataLoader/TUMDataLoader.cpp.
#include "../include/TUMDataLoader.hpp"
#include "../include/FileUtilities.hpp"

#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>

TUMDataLoader::TUMDataLoader(const std::string &directory) : m_current_idx(0) {
    bool is_directory = false;
    if (!file_exists(directory, is_directory) || !is_directory) throw std::invalid_argument("Directory not found " + directory);
    m_directory_name = directory;
    const std::string gt_file_name = directory + "/ground_truth.txt";
    if (!file_exists(gt_file_name, is_directory) || is_directory) throw std::invalid_argument("Ground truth file not found " + gt_file_name);
    load_data_from(gt_file_name);
}

TUMDataLoader::~TUMDataLoader() {}

Eigen::Matrix4f TUMDataLoader::to_pose(float vars[7]) const {
    // tx ty tz (metres) qx qy qz qw -> rotation from the quaternion, translation in millimetres (TUMDataLoader.cpp:47-76)
    const float x = vars[3], y = vars[4], z = vars[5], w = vars[6];
    Eigen::Matrix4f pose = Eigen::Matrix4f::Zero();
    pose(0, 0) = 1 - 2 * (y * y + z * z); pose(0, 1) = 2 * (x * y - w * z);     pose(0, 2) = 2 * (x * z + w * y);
    pose(1, 0) = 2 * (x * y + w * z);     pose(1, 1) = 1 - 2 * (x * x + z * z); pose(1, 2) = 2 * (y * z - w * x);
    pose(2, 0) = 2 * (x * z - w * y);     pose(2, 1) = 2 * (y * z + w * x);     pose(2, 2) = 1 - 2 * (x * x + y * y);
    pose(0, 3) = vars[0] * 1000.0f;
    pose(1, 3) = vars[1] * 1000.0f;
    pose(2, 3) = vars[2] * 1000.0f;
    pose(3, 3) = 1.0f;
    return pose;
}

DepthImage *TUMDataLoader::next(Eigen::Matrix4f &pose) {
    if (m_current_idx >= m_data_records.size()) return nullptr;
    DATA_RECORD &record = m_data_records[m_current_idx++];
    bool is_directory = false;
    if (!file_exists(record.file_name, is_directory) || is_directory) {
        std::cerr << "Couldn't find file " << record.file_name << std::endl;
        return nullptr;
    }
    DepthImage *image = new DepthImage(record.file_name);
    image->scale_depth(0.2f);          // TUM stores 5000 units per metre; the volume works in millimetres
    pose = to_pose(record.data);
    return image;
}

void TUMDataLoader::process_line(const std::string &line) {
    if (line.empty() || line[0] == '#') return;
    std::stringstream fields(line);
    DATA_RECORD record;
    std::string stamp;
    fields >> stamp;
    record.file_name = m_directory_name + "/depth/" + stamp + ".png";
    for (int i = 0; i < 7; i++) fields >> record.data[i];
    m_data_records.push_back(record);
}

void TUMDataLoader::load_data_from(const std::string &gt_file_name) {
    if (!process_file_by_lines(gt_file_name, [this](const std::string &line) { process_line(line); }))
        throw std::runtime_error("Failed to parse the ground truth file");
    m_current_idx = 0;
}
