// Shim (test infrastructure): shadows engine/db/index/spatial/geoindex.hpp (Boost.Geometry
// R-tree, out of scope).  API shape from engine/db/index/spatial/geoindex.hpp:20-40; no-ops.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>
namespace vectordb {
namespace engine {
namespace index {
class GeospatialIndex {
 public:
  struct point_t {
    double lat, lon;
    point_t() : lat(0), lon(0) {}
    point_t(double a, double b) : lat(a), lon(b) {}
  };
  typedef std::pair<point_t, int64_t> value_t;
  GeospatialIndex() {}
  ~GeospatialIndex() {}
  void insertPoint(double, double, int64_t) {}
  void deletePoint(double, double, int64_t) {}
  void searchWithinRadius(double, double, double, std::vector<value_t>&) const {}
  static double distance(const point_t&, const point_t&) { return 0.0; }
};
}  // namespace index
}  // namespace engine
}  // namespace vectordb
