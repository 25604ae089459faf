#pragma once
#include <string>
namespace fqb {
void set_error(const std::string &e);
}
