// TEST INFRASTRUCTURE. Force-included (-include) in front of the UNMODIFIED /root/reference/src/terrain/terrain.cpp so that it
// compiles with g++: its line 468 throws std::exception("invalid offset"), a constructor only MSVC's STL has. Every standard
// header the translation unit uses is included first (their include guards then make the later #includes no-ops); after that
// the identifier `exception` is spelled `runtime_error`, which has the (const char*) constructor.
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <exception>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <cuda_runtime.h>
#include <thrust/random.h>
#define exception runtime_error
