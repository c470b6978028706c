// TEST INFRASTRUCTURE (oracle/): force-included (-include) in front of every reference translation unit.  The harness
// (oracle/srl_reference_harness.cpp) has to fill private members of lioOptimization / eskfEstimator (voxel_map, eskf_pro,
// R_imu_lidar, ...) that the reference initialises from ROS parameters: all standard headers come first, then `private`
// is opened up for the reference's own headers.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <tr1/unordered_map>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>
#include <math.h>
#include <Eigen/Core>
#include <tsl/robin_map.h>
#include "srl_shim_ext.h"
#define private public
#define protected public
