// yaml-lite: the block-style YAML subset used by configs/*.yaml, parsed into a Json DOM
// (the reference uses yaml-cpp: types.cpp:20-101, worker_service.cpp:25-108).
// Supported: nested block mappings by indentation, block sequences ("- x", "- k: v" items
// with continuation keys), flow sequences "[a, b]", flow mappings "{a: 1}", single/double
// quoted scalars, comments, typed plain scalars (int/float/bool/null), multi-document "---"
// (first document only).  Not supported: anchors, tags, block scalars (| >), multi-line flow.
#pragma once
#include <optional>
#include <string>
#include <string_view>

#include "common/json.h"

namespace bb {

std::optional<Json> parse_yaml(std::string_view text, std::string* err = nullptr);
std::optional<Json> load_yaml_file(const std::string& path, std::string* err = nullptr);

// "2147483648", "2GB", "32_GB", "10 MiB", "4k", "unlimited" -> bytes.  GB/MB/KB are binary
// multiples (as the reference's comments use them: "2147483648  # 2GB").
std::optional<uint64_t> parse_size(std::string_view s);

}  // namespace bb
