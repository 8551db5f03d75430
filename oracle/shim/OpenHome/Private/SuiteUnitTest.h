// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/SuiteUnitTest.h surface -- a Suite of named tests, each run between
// Setup() and TearDown().
#pragma once
#include <OpenHome/Private/TestFramework.h>
#include <OpenHome/Functor.h>
#include <string>
#include <utility>
#include <vector>

namespace OpenHome {
namespace TestFramework {

class SuiteUnitTest : public Suite
{
protected:
    SuiteUnitTest(const TChar* aSuiteName) : Suite(aSuiteName) {}
    void AddTest(Functor aTest, const TChar* aName = "") { iTests.emplace_back(aTest, aName ? aName : ""); }
private:
    virtual void Setup() = 0;
    virtual void TearDown() = 0;
    void Test() override
    {
        for (auto& t : iTests) {
            Setup();
            try {
                t.first();
            }
            catch (Exception& e) {
                Fail(e.File(), e.Line(), t.second.c_str(), e.Message());
            }
            TearDown();
        }
    }
private:
    std::vector<std::pair<Functor, std::string>> iTests;
};

} // namespace TestFramework
} // namespace OpenHome
