#pragma once
// Tools::Error of src/HYMLS_Tools.cpp:191-194: throws (HYMLS::Exception there)
#include <stdexcept>
#include <string>
namespace HYMLS {
struct Tools {
  static void Error(const std::string& msg, const char* file, int line) {
    throw std::runtime_error(msg + " (" + file + ":" + std::to_string(line) + ")");
  }
};
}  // namespace HYMLS
