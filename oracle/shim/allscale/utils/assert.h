// TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for the assertion macros of the
// un-vendored AllScale API (github.com/allscale/allscale_api @ fd12c6e5, see
// /root/reference/cmake/allscale.cmake:42-54).  Written from the published macro surface:
//   assert_true(c) << "msg";  assert_eq(a,b) << ...;  assert_fail() << ...;  assert_decl(stmt)
// With NDEBUG the condition is not evaluated and the message stream is discarded.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace allscale { namespace utils { namespace shim_detail {

	// Collects the streamed message; aborts on destruction when armed.
	struct AssertSink {
		bool armed;
		std::ostringstream msg;
		AssertSink(bool armed, const char* what, const char* file, int line) : armed(armed) {
			if (armed) msg << "Assertion failed: " << what << " at " << file << ":" << line << " ";
		}
		AssertSink(AssertSink&& o) : armed(o.armed), msg(std::move(o.msg)) { o.armed = false; }
		~AssertSink() {
			if (armed) { std::cerr << msg.str() << std::endl; std::abort(); }
		}
		template<typename T> AssertSink& operator<<(const T& v) { if (armed) msg << v; return *this; }
		AssertSink& operator<<(std::ostream& (*f)(std::ostream&)) { if (armed) f(msg); return *this; }
	};

}}}

#ifdef NDEBUG
  #define ALLSCALE_SHIM_CHECK(cond, text) if (true) {} else ::allscale::utils::shim_detail::AssertSink(false, text, __FILE__, __LINE__)
  #define assert_decl(...)  ((void)0)
#else
  #define ALLSCALE_SHIM_CHECK(cond, text) if (cond) {} else ::allscale::utils::shim_detail::AssertSink(true, text, __FILE__, __LINE__)
  #define assert_decl(...)  __VA_ARGS__
#endif

#define assert_true(c)   ALLSCALE_SHIM_CHECK((bool)(c), #c)
#define assert_false(c)  ALLSCALE_SHIM_CHECK(!(bool)(c), "!(" #c ")")
#define assert_eq(a,b)   ALLSCALE_SHIM_CHECK((a) == (b), #a " == " #b)
#define assert_ne(a,b)   ALLSCALE_SHIM_CHECK((a) != (b), #a " != " #b)
#define assert_lt(a,b)   ALLSCALE_SHIM_CHECK((a) <  (b), #a " < "  #b)
#define assert_le(a,b)   ALLSCALE_SHIM_CHECK((a) <= (b), #a " <= " #b)
#define assert_gt(a,b)   ALLSCALE_SHIM_CHECK((a) >  (b), #a " > "  #b)
#define assert_ge(a,b)   ALLSCALE_SHIM_CHECK((a) >= (b), #a " >= " #b)
#define assert_fail()           ALLSCALE_SHIM_CHECK(false, "fail")
#define assert_not_implemented() ALLSCALE_SHIM_CHECK(false, "not implemented")
