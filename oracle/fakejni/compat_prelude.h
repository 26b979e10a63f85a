// TEST INFRASTRUCTURE (oracle): force-included before every reference translation unit.  The reference relies on
// libc++/bionic transitive includes and on std::powf (absent from libstdc++ 13); nothing else is changed.
#pragma once
#include <cstdint>
#include <cfloat>
#include <cstring>
#include <mutex>
#include <cmath>
#include <limits>
#include <memory>
#include <algorithm>
#include <string>
#include <stdexcept>
#include <thread>
#include <cstdlib>
#include <math.h>
namespace std { using ::powf; using ::sqrtf; using ::roundf; using ::floorf; using ::ceilf; using ::fabsf; using ::expf; using ::logf; }
