// Shim: boost::mt19937 -> std::mt19937 (engine/db/index/knn/nndescent_common.hpp:228).
#pragma once
#include <random>
namespace boost { using mt19937 = std::mt19937; }
