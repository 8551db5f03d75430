// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/TestFramework.h surface -- Suite, Runner, TEST, TEST_THROWS, Print --
// enough to compile the reference's OWN unit-test files (Media/Tests/Test*.cpp) unmodified and run them against the
// reference sources this oracle links (oracle/Makefile, target ref_suites).  A failed TEST is counted and printed, not fatal,
// as in ohNet; the totals are read by tests/test_reference_own_suites.py.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>
#include <OpenHome/OsWrapper.h>
#include <OpenHome/Private/Printer.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace OpenHome {
namespace TestFramework {

struct Totals
{
    static unsigned long& Passed() { static unsigned long n = 0; return n; }
    static unsigned long& Failed() { static unsigned long n = 0; return n; }
};

inline void Succeed(const TChar* /*aFile*/, TUint /*aLine*/) { Totals::Passed()++; }
inline void Fail(const TChar* aFile, TUint aLine, const TChar* aExp, const TChar* aMsg)
{
    Totals::Failed()++;
    std::printf("FAILURE: %s:%u: %s%s%s\n", aFile, aLine, aExp ? aExp : "", aMsg ? " -- " : "", aMsg ? aMsg : "");
}

inline TInt Print(const TChar* aFormat, ...)
{
    if (std::getenv("OHP_REF_SUITES_VERBOSE") == nullptr) return 0;
    va_list args;
    va_start(args, aFormat);
    const int n = std::vprintf(aFormat, args);
    va_end(args);
    return n;
}

class Suite
{
public:
    virtual ~Suite() {}
    virtual void Test() = 0;
    const TChar* Description() const { return iDesc; }
protected:
    Suite(const TChar* aDesc) : iDesc(aDesc) {}
private:
    const TChar* iDesc;
};

class Runner
{
public:
    Runner(const TChar* aDesc) : iDesc(aDesc) {}
    ~Runner() { for (auto* s : iSuites) delete s; }
    void Add(Suite* aSuite) { iSuites.push_back(aSuite); }
    void Run()
    {
        for (auto* s : iSuites) {
            const unsigned long p0 = Totals::Passed(), f0 = Totals::Failed();
            try {
                s->Test();
            }
            catch (Exception& e) {
                Fail(e.File(), e.Line(), "unexpected exception", e.Message());
            }
            std::printf("suite: %s: %lu passed, %lu failed\n", s->Description(), Totals::Passed() - p0, Totals::Failed() - f0);
            std::fflush(stdout);
        }
    }
private:
    const TChar* iDesc;
    std::vector<Suite*> iSuites;
};

} // namespace TestFramework
} // namespace OpenHome

#define TEST(aCondition)                                                                         \
    do {                                                                                         \
        if (aCondition) OpenHome::TestFramework::Succeed(__FILE__, __LINE__);                    \
        else OpenHome::TestFramework::Fail(__FILE__, __LINE__, #aCondition, nullptr);            \
    } while (0)

#define TEST_THROWS(aExp, aExceptionType)                                                        \
    do {                                                                                         \
        bool thrown_ = false;                                                                    \
        try { aExp; }                                                                            \
        catch (aExceptionType&) { thrown_ = true; }                                              \
        if (thrown_) OpenHome::TestFramework::Succeed(__FILE__, __LINE__);                       \
        else OpenHome::TestFramework::Fail(__FILE__, __LINE__, #aExp, "did not throw " #aExceptionType); \
    } while (0)
