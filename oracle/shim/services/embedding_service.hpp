// Shim (test infrastructure): shadows engine/services/embedding_service.hpp so the
// reference's table_segment_mvp.cpp compiles without oatpp.  Declares only the API
// surface at engine/services/embedding_service.hpp:78-114; every call is a no-op.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>
#include "db/vector.hpp"
#include "logger/logger.hpp"
#include "utils/json.hpp"
#include "utils/status.hpp"
namespace vectordb {
namespace engine {
struct EmbeddingModel {
  std::string model;
  size_t dim;
  bool dense;
  bool dimensionReduction;
};
class EmbeddingService {
 public:
  explicit EmbeddingService(const std::string&) {}
  Status getSupportedModels(std::vector<EmbeddingModel>&) { return Status::OK(); }
  Status denseEmbedDocuments(const std::string&, VariableLenAttrColumnContainer&, float*, size_t, size_t,
                             size_t, std::unordered_map<std::string, std::string>&, bool) {
    return Status::OK();
  }
  Status denseEmbedQuery(const std::string&, const std::string&, std::vector<engine::DenseVectorElement>&,
                         size_t, std::unordered_map<std::string, std::string>&, bool) {
    return Status::OK();
  }
};
}  // namespace engine
}  // namespace vectordb
