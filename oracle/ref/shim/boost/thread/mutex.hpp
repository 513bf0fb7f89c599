#pragma once
#include "boost/shim_core.hpp"
#include <mutex>
namespace boost {
typedef std::mutex mutex;
template <class M> using unique_lock = std::unique_lock<M>;
template <class M> using lock_guard = std::lock_guard<M>;
}
