#pragma once
#include "boost/shim_core.hpp"
#include <unordered_map>
namespace boost { using std::unordered_map; }
