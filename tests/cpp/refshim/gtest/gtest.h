// TEST INFRASTRUCTURE - a minimal stand-in for <gtest/gtest.h> (googletest is not installed in this image and
// cannot be fetched).  It provides exactly what the reference's libzen/*.test.cu use - TEST, TEST_F,
// ::testing::Test with SetUp / TearDown, EXPECT_{EQ,NE,NEAR,TRUE,FALSE,THROW}, ASSERT_*, and the gtest_main
// entry point - so that those sources compile UNCHANGED against zen_b200/include (oracle/Makefile, target reftests).
#ifndef ZEN_B200_GTEST_SHIM_H
#define ZEN_B200_GTEST_SHIM_H

#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Test {
public:
	virtual ~Test() {}
	virtual void SetUp() {}
	virtual void TearDown() {}
	virtual void TestBody() = 0;
};

namespace internal {
	struct Case {
		std::string suite, name;
		std::function<Test*()> make;
	};
	inline std::vector<Case>& registry()
	{
		static std::vector<Case> r;
		return r;
	}
	struct State {
		long failures = 0;       // failed expectations of the running test
		long printed = 0;
		bool fatal = false;
	};
	inline State& state()
	{
		static State s;
		return s;
	}
	struct Registrar {
		Registrar(const char* suite, const char* name, std::function<Test*()> make) { registry().push_back({suite, name, make}); }
	};
	// sink for `EXPECT_x(...) << "message"`
	struct Msg {
		bool on;
		std::ostringstream os;
		explicit Msg(bool o)
		    : on(o)
		{
		}
		Msg(const Msg& m)
		    : on(m.on)
		{
		}
		template <typename T>
		Msg& operator<<(const T& v)
		{
			if (on) os << v;
			return *this;
		}
		~Msg()
		{
			if (on && !os.str().empty() && state().printed <= 20) std::cerr << "    " << os.str() << "\n";
		}
	};
	inline Msg fail(const char* file, int line, const std::string& what, bool fatal)
	{
		State& s = state();
		s.failures++;
		if (fatal) s.fatal = true;
		if (s.printed++ < 20) std::cerr << file << ":" << line << ": Failure\n    " << what << "\n";
		return Msg(true);
	}
	template <typename A, typename B>
	std::string show2(const char* op, const char* ea, const char* eb, const A& a, const B& b)
	{
		std::ostringstream os;
		os << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: " << a << " vs " << b;
		return os.str();
	}
}  // namespace internal

inline void InitGoogleTest(int*, char**) {}

}  // namespace testing

inline int RUN_ALL_TESTS()
{
	using namespace testing::internal;
	long failed_tests = 0, ran = 0;
	std::vector<std::string> failed;
	for (auto& c : registry()) {
		std::printf("[ RUN      ] %s.%s\n", c.suite.c_str(), c.name.c_str());
		std::fflush(stdout);
		state() = State();
		bool threw = false;
		try {
			testing::Test* t = c.make();
			t->SetUp();
			if (!state().fatal) t->TestBody();
			t->TearDown();
			delete t;
		}
		catch (const std::exception& e) {
			threw = true;
			std::cerr << "    unexpected exception: " << e.what() << "\n";
		}
		catch (...) {
			threw = true;
			std::cerr << "    unexpected exception\n";
		}
		++ran;
		if (threw || state().failures) {
			++failed_tests;
			failed.push_back(c.suite + "." + c.name);
			std::printf("[  FAILED  ] %s.%s (%ld failed expectations)\n", c.suite.c_str(), c.name.c_str(), state().failures);
		}
		else
			std::printf("[       OK ] %s.%s\n", c.suite.c_str(), c.name.c_str());
	}
	std::printf("[==========] %ld tests ran.\n[  PASSED  ] %ld tests.\n", ran, ran - failed_tests);
	for (auto& f : failed)
		std::printf("[  FAILED  ] %s\n", f.c_str());
	return failed_tests ? 1 : 0;
}

#define ZEN_GT_CLASS(suite, name) suite##_##name##_Test
#define ZEN_GT_TEST(suite, name, parent)                                                                              \
	class ZEN_GT_CLASS(suite, name) : public parent {                                                                \
	public:                                                                                                           \
		void TestBody() override;                                                                                     \
	};                                                                                                                \
	static ::testing::internal::Registrar zen_gt_reg_##suite##_##name(#suite, #name,                                  \
	                                                                  []() -> ::testing::Test* { return new ZEN_GT_CLASS(suite, name)(); }); \
	void ZEN_GT_CLASS(suite, name)::TestBody()
#define TEST(suite, name) ZEN_GT_TEST(suite, name, ::testing::Test)
#define TEST_F(fixture, name) ZEN_GT_TEST(fixture, name, fixture)

#define ZEN_GT_CMP(op, opname, a, b, fatal)                                                                           \
	if (const auto& zen_gt_a = (a); true)                                                                             \
		if (const auto& zen_gt_b = (b); zen_gt_a op zen_gt_b)                                                         \
			;                                                                                                         \
		else if (fatal)                                                                                               \
			return (void)::testing::internal::fail(__FILE__, __LINE__, ::testing::internal::show2(opname, #a, #b, zen_gt_a, zen_gt_b), true); \
		else                                                                                                          \
			::testing::internal::fail(__FILE__, __LINE__, ::testing::internal::show2(opname, #a, #b, zen_gt_a, zen_gt_b), false)

#define EXPECT_EQ(a, b) ZEN_GT_CMP(==, "==", a, b, false)
#define EXPECT_NE(a, b) ZEN_GT_CMP(!=, "!=", a, b, false)
#define EXPECT_LT(a, b) ZEN_GT_CMP(<, "<", a, b, false)
#define EXPECT_LE(a, b) ZEN_GT_CMP(<=, "<=", a, b, false)
#define EXPECT_GT(a, b) ZEN_GT_CMP(>, ">", a, b, false)
#define EXPECT_GE(a, b) ZEN_GT_CMP(>=, ">=", a, b, false)
#define ASSERT_EQ(a, b) ZEN_GT_CMP(==, "==", a, b, true)
#define ASSERT_NE(a, b) ZEN_GT_CMP(!=, "!=", a, b, true)

#define EXPECT_TRUE(c)                                                                                                \
	if (c)                                                                                                            \
		;                                                                                                             \
	else                                                                                                              \
		::testing::internal::fail(__FILE__, __LINE__, std::string("Expected true: ") + #c, false)
#define EXPECT_FALSE(c)                                                                                               \
	if (!(c))                                                                                                         \
		;                                                                                                             \
	else                                                                                                              \
		::testing::internal::fail(__FILE__, __LINE__, std::string("Expected false: ") + #c, false)
#define ASSERT_TRUE(c)                                                                                                \
	if (c)                                                                                                            \
		;                                                                                                             \
	else                                                                                                              \
		return (void)::testing::internal::fail(__FILE__, __LINE__, std::string("Expected true: ") + #c, true)

#define EXPECT_NEAR(a, b, tol)                                                                                        \
	if (const double zen_gt_d = std::fabs((double)(a) - (double)(b)); zen_gt_d <= (double)(tol))                      \
		;                                                                                                             \
	else                                                                                                              \
		::testing::internal::fail(__FILE__, __LINE__,                                                                 \
		                          std::string("|") + #a + " - " + #b + "| = " + std::to_string(zen_gt_d) + " exceeds " + #tol, false)

#define EXPECT_THROW(stmt, ex)                                                                                        \
	if (bool zen_gt_caught = false; true) {                                                                           \
		try {                                                                                                         \
			stmt;                                                                                                     \
		}                                                                                                             \
		catch (const ex&) {                                                                                           \
			zen_gt_caught = true;                                                                                     \
		}                                                                                                             \
		catch (...) {                                                                                                 \
		}                                                                                                             \
		if (!zen_gt_caught) ::testing::internal::fail(__FILE__, __LINE__, std::string("Expected ") + #stmt + " to throw " + #ex, false); \
	}                                                                                                                 \
	else                                                                                                              \
		::testing::internal::Msg(false)

#ifndef ZEN_GT_NO_MAIN
// gtest_main
int main(int argc, char** argv)
{
	::testing::InitGoogleTest(&argc, argv);
	return RUN_ALL_TESTS();
}
#endif

#endif
