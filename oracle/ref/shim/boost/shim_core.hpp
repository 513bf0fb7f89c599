// No-op stand-ins for the parts of Boost that the LuxRays headers mention in declarations
// (serialization hooks, archives, foreach, lexical_cast, signbit).  Nothing here computes anything
// that reaches a Ray / RayHit / BVH node: the reference's arithmetic is compiled unchanged.
#pragma once
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <stdexcept>
#define BOOST_VERSION 107500
namespace boost {
namespace serialization {
class access {};
enum { object_serializable = 1, object_class_info = 2, track_never = 0, track_selectively = 1 };
template <class B, class D> inline B &base_object(D &d) { return d; }
template <class T> struct array_wrapper_ { T *p; size_t n; };
template <class T> inline array_wrapper_<T> make_array(T *p, size_t n) { array_wrapper_<T> a = { p, n }; return a; }
template <class A, class T> inline void split_member(A &, T &, const unsigned int) {}
}   // namespace serialization
namespace archive {
class binary_oarchive {
public:
	template <class S> binary_oarchive(S &, unsigned = 0) {}
	template <class T> binary_oarchive &operator&(const T &) { return *this; }
	template <class T> binary_oarchive &operator<<(const T &) { return *this; }
};
class binary_iarchive {
public:
	template <class S> binary_iarchive(S &, unsigned = 0) {}
	template <class T> binary_iarchive &operator&(T &) { return *this; }
	template <class T> binary_iarchive &operator>>(T &) { return *this; }
};
enum { no_header = 1 };
}   // namespace archive
namespace filesystem {
typedef std::ofstream ofstream;
typedef std::ifstream ifstream;
}
namespace iostreams {
struct filtering_ostream : public std::ostringstream { template <class T> void push(const T &) {} };
struct filtering_istream : public std::istringstream { template <class T> void push(const T &) {} };
struct gzip_compressor { gzip_compressor(int = 0) {} };
struct gzip_decompressor {};
}
namespace math {
template <class T> inline int signbit(T v) { return std::signbit(v) ? 1 : 0; }
}
template <class T, class S> inline T lexical_cast(const S &s) {
	std::stringstream ss;
	ss << s;
	T t;
	ss >> t;
	if (ss.fail())
		throw std::runtime_error("bad lexical_cast");
	return t;
}
class bad_lexical_cast : public std::runtime_error { public: bad_lexical_cast() : std::runtime_error("bad_lexical_cast") {} };
}   // namespace boost
#define BOOST_CLASS_IMPLEMENTATION(T, L)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_KEY2(T, N)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)
#define BOOST_CLASS_VERSION(T, V)
#define BOOST_CLASS_TRACKING(T, E)
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#define BOOST_SERIALIZATION_SPLIT_MEMBER() template <class A_> void serialize(A_ &, const unsigned int) {}
#define BOOST_FOREACH(decl, range) for (decl : range)
#include <cassert>
#include <cfloat>
#include <cstring>
#include <atomic>
#include <cstdint>
#define BOOST_ASSERT(x) assert(x)
#define BOOST_SERIALIZATION_BASE_OBJECT_NVP(T) 0
#define BOOST_SERIALIZATION_NVP(x) (x)
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
namespace boost { namespace interprocess { namespace ipcdetail {
inline uint32_t atomic_cas32(volatile uint32_t *mem, uint32_t with, uint32_t cmp) { return __sync_val_compare_and_swap(mem, cmp, with); }
inline uint32_t atomic_add32(volatile uint32_t *mem, uint32_t val) { return __sync_fetch_and_add(mem, val); }
inline uint32_t atomic_inc32(volatile uint32_t *mem) { return __sync_fetch_and_add(mem, 1u); }
inline uint32_t atomic_dec32(volatile uint32_t *mem) { return __sync_fetch_and_sub(mem, 1u); }
} } }
namespace boost {
using std::atomic;
using std::memory_order_acquire;
using std::memory_order_release;
using std::memory_order_relaxed;
}
