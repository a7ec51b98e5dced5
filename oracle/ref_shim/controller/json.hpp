// TEST INFRASTRUCTURE — shadows the reference's controller/json.hpp (nlohmann adapters): the
// JSON conversions are host configuration plumbing outside the hot path.
#pragma once
#define NLOHMANN_DEFINE_TYPE_INTRUSIVE(...)
#define NLOHMANN_JSON_SERIALIZE_ENUM(...)
#include <cstddef>
#include <filesystem>
#include <optional>
#include <string>
#include <iostream>
// state.hpp / control.hpp define inline to_json / from_json over this type; never called here
struct json {
    using size_type = std::size_t;
    static json array() { return json(); }
    template <class T> void push_back(const T &) {}
    size_type size() const { return 0; }
    json operator[](size_type) const { return json(); }
    template <class T> T get() const { return T(); }
};
