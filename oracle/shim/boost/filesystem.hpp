// Test-infrastructure shim (NOT product code): the compiled reference under oracle/_ref uses a
// handful of Boost.Filesystem calls (exists, remove_all, create_directory, rename, remove,
// directory_iterator, path::stem/extension).  Boost is not installed in this image; every one of
// those calls exists with identical semantics in C++17 <filesystem>, so alias the namespace.
#pragma once
#include <filesystem>
#include <fstream>
#include <array>
#include <vector>
#include <string>
#include <limits>
namespace boost { namespace filesystem = std::filesystem; }
