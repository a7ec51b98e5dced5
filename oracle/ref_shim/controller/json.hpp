// TEST INFRASTRUCTURE — shadows the reference's controller/json.hpp (nlohmann adapters): the
// JSON conversions are host configuration plumbing outside the hot path.
#pragma once
#define NLOHMANN_DEFINE_TYPE_INTRUSIVE(...)
#define NLOHMANN_JSON_SERIALIZE_ENUM(...)
#include <optional>
#include <string>
#include <iostream>
