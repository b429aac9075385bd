// TEST INFRASTRUCTURE ONLY (oracle/). A ~60-line stand-in for GoogleTest 1.8 (which the
// reference downloads, cmake/dependencies/googletest.cmake:17-25) so that the reference's own
// unit tests under /root/reference/code/app/test can be compiled UNMODIFIED against the
// AllScale shim to validate it.  Supports exactly the macros those tests use.
#pragma once
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
	struct Registry {
		struct Case { std::string name; std::function<void()> run; };
		std::vector<Case> cases; int failures = 0; bool currentFailed = false; bool fatal = false;
		static Registry& get() { static Registry r; return r; }
	};
	struct Registrar { Registrar(const std::string& n, std::function<void()> f) { Registry::get().cases.push_back({ n, f }); } };
	struct Message {
		bool failed, fatal; std::ostringstream text;
		Message(bool failed, bool fatal, const char* expr, const char* file, int line) : failed(failed), fatal(fatal) { if (failed) text << file << ":" << line << ": check failed: " << expr << " "; }
		Message(Message&& o) : failed(o.failed), fatal(o.fatal), text(std::move(o.text)) { o.failed = false; }
		~Message() { if (failed) { std::cerr << text.str() << std::endl; auto& r = Registry::get(); r.currentFailed = true; if (fatal) r.fatal = true; } }
		template<typename T> Message& operator<<(const T& v) { if (failed) text << v; return *this; }
		Message& operator<<(std::ostream& (*f)(std::ostream&)) { if (failed) f(text); return *this; }
	};
	inline void InitGoogleTest(int*, char**) {}
	class Test { public: virtual ~Test() {} };
}

#define TEST(suite, name) \
	static void suite##_##name##_body(); \
	static ::testing::Registrar suite##_##name##_reg(#suite "." #name, suite##_##name##_body); \
	static void suite##_##name##_body()

#define MINI_CHECK(ok, fatal, expr) ::testing::Message(!(ok), fatal, expr, __FILE__, __LINE__)
#define EXPECT_TRUE(c)  MINI_CHECK((bool)(c), false, #c)
#define EXPECT_FALSE(c) MINI_CHECK(!(bool)(c), false, "!" #c)
#define EXPECT_EQ(a,b)  MINI_CHECK((a) == (b), false, #a " == " #b)
#define EXPECT_NE(a,b)  MINI_CHECK((a) != (b), false, #a " != " #b)
#define EXPECT_LT(a,b)  MINI_CHECK((a) < (b), false, #a " < " #b)
#define EXPECT_NEAR(a,b,t) MINI_CHECK(std::fabs((double)(a) - (double)(b)) <= (double)(t), false, #a " ~ " #b) << " got " << (a) << " vs " << (b) << " "
#define ASSERT_TRUE(c)  if (c) {} else return (void)MINI_CHECK(false, true, #c)
#define ASSERT_FALSE(c) if (!(c)) {} else return (void)MINI_CHECK(false, true, "!" #c)
#define ASSERT_EQ(a,b)  if ((a) == (b)) {} else return (void)MINI_CHECK(false, true, #a " == " #b)
#define EXPECT_DEATH_IF_SUPPORTED(stmt, re) do { std::cerr << "[ SKIPPED death test ] " #stmt << std::endl; } while (0)

inline int RUN_ALL_TESTS() {
	auto& r = ::testing::Registry::get();
	for (auto& c : r.cases) {
		if (c.name.find("DISABLED_") != std::string::npos) { std::cout << "[ DISABLED ] " << c.name << std::endl; continue; }
		r.currentFailed = false; c.run();
		std::cout << (r.currentFailed ? "[  FAILED  ] " : "[       OK ] ") << c.name << std::endl;
		if (r.currentFailed) r.failures++;
	}
	std::cout << (r.cases.size()) << " cases, " << r.failures << " failed" << std::endl;
	return r.failures ? 1 : 0;
}
