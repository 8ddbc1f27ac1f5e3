// Test-infrastructure shim (NOT product code), force-included ahead of every reference translation unit by
// oracle/Makefile.  Pulls in every standard header the reference uses *first* (their include guards then make the
// reference's own #includes no-ops) and only afterwards opens the reference's class internals, so that
// ref_harness.cpp can call GridClass::lbmKernel, ObjectsClass::ibmKernelInterp, ... directly.
// Changes access control only — no arithmetic, no layout.
#pragma once
#include <iostream>
#include <iomanip>
#include <sstream>
#include <fstream>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <numeric>
#include <functional>
#include <array>
#include <vector>
#include <string>
#include <limits>
#include <memory>
#include <map>
#include <filesystem>
#include <omp.h>
#include <unistd.h>
#define private public
