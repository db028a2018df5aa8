// FileUtilities.cpp — subset of reference src/Utilities/FileUtilities.cpp used by the loaders.
#include "../include/FileUtilities.hpp"

#include <cstdio>
#include <dirent.h>
#include <fstream>
#include <sys/stat.h>

bool process_file_by_lines(const std::string &file_name, std::function<void(const std::string &)> processor) {
    std::ifstream f(file_name);
    if (!f.is_open()) perror(("error while opening file " + file_name).c_str());
    std::string line;
    while (std::getline(f, line)) processor(line);
    if (f.bad()) perror(("error while reading file " + file_name).c_str());
    return true;            // like the reference (FileUtilities.cpp:92-116): problems are reported, not returned
}

bool file_exists(const std::string &file_name, bool &is_directory) {
    struct stat info;
    is_directory = false;
    if (stat(file_name.c_str(), &info) != 0) return false;
    is_directory = S_ISDIR(info.st_mode);
    return true;
}

void files_in_directory(const std::string &directory, std::vector<std::string> &files, std::function<bool(const char *)> filter) {
    DIR *dir = opendir(directory.c_str());
    if (!dir) return;
    while (struct dirent *entry = readdir(dir))
        if (!filter || filter(entry->d_name)) files.push_back(entry->d_name);
    closedir(dir);
}
