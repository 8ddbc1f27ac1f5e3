// Force-included ahead of LIFE's translation units by life_b200/host/Makefile (route (b) of INTEGRATION.md: no source change in the
// LIFE checkout).  life_host.cpp DEFINES member functions of LIFE's classes (GridClass::lbmKernel, ObjectsClass::ibmKernelInterp, ...)
// and reads their members; some of those are private.  A maintainer taking route (a) adds the definitions inside the classes' own
// sources and needs none of this.  The standard headers come first so that their contents are not affected; access control is the
// only thing that changes — no arithmetic, no layout.
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
