// FileUtilities.hpp — the file helpers the loaders use (subset of reference src/include/FileUtilities.hpp).
#ifndef FILE_UTILITIES_H
#define FILE_UTILITIES_H
#include <functional>
#include <string>
#include <vector>

bool process_file_by_lines(const std::string &file_name, std::function<void(const std::string &)> processor);
bool file_exists(const std::string &file_name, bool &is_directory);
void files_in_directory(const std::string &directory, std::vector<std::string> &files, std::function<bool(const char *)> filter);
#endif
