// Build environment compatibility for compiling LIFE's own sources around liblife_b200 (life_b200/host/Makefile): LIFE uses a few
// Boost.Filesystem calls (exists, remove_all, create_directory, rename, remove, directory_iterator, path::stem / extension).
// Where Boost is not installed, every one of them exists with the same semantics in C++17 <filesystem>: alias the namespace.
// A site with Boost simply drops -Ilife_b200/host/compat from the flags.
#pragma once
#include <filesystem>
#include <fstream>
#include <array>
#include <vector>
#include <string>
#include <limits>
namespace boost { namespace filesystem = std::filesystem; }
