// Shim: only the namespace is needed (engine/db/index/knn/knn.hpp:5-6,24).
#pragma once
namespace boost {}
