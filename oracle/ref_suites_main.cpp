// ref_suites_main.cpp -- TEST INFRASTRUCTURE: runs the REFERENCE'S OWN unit-test suites (Media/Tests/TestMsg.cpp, TestRamper.cpp,
// TestMuter.cpp, TestVolumeRamper.cpp, TestStarvationRamper.cpp, TestFlywheelRamper.cpp, TestDecodedAudioAggregator.cpp, and of
// three more elements that ramp with the same Split + SetRamp idiom: TestSkipper.cpp, TestWaiter.cpp, TestVariableDelay.cpp;
// compiled unmodified from /root/reference) against the reference sources the oracle links behind its ohNet shim.  What
// passes here is what the reference's authors check about this path -- ramp algebra, the applicator, playables and splits,
// the element state machines, the flywheel's known answers -- holding on THIS build of it: the shim (oracle/shim) changes
// none of it.  Built by oracle/Makefile (target ref_suites) into oracle/_ref/ref_suites; run by
// tests/test_reference_own_suites.py.  None of the reference's text is in this file.
#include <OpenHome/Private/TestFramework.h>
#include <cstdio>
#include <cstring>

extern void TestMsg();
extern void TestRamper();
extern void TestMuter();
extern void TestVolumeRamper();
extern void TestStarvationRamper();
extern void TestFlywheelRamper(OpenHome::Environment&);
extern void TestDecodedAudioAggregator();
extern void TestSkipper();
extern void TestWaiter();
extern void TestVariableDelay();

int main(int argc, char** argv)
{
    OpenHome::Environment env;
    const char* which = argc > 1 ? argv[1] : "all";
    auto on = [&](const char* name) { return std::strcmp(which, "all") == 0 || std::strcmp(which, name) == 0; };
    if (on("msg")) TestMsg();
    if (on("ramper")) TestRamper();
    if (on("muter")) TestMuter();
    if (on("volume")) TestVolumeRamper();
    if (on("starvation")) TestStarvationRamper();
    if (on("flywheel")) TestFlywheelRamper(env);
    if (on("aggregator")) TestDecodedAudioAggregator();
    if (on("skipper")) TestSkipper();
    if (on("waiter")) TestWaiter();
    if (on("delay")) TestVariableDelay();
    std::printf("total: %lu passed, %lu failed\n", OpenHome::TestFramework::Totals::Passed(), OpenHome::TestFramework::Totals::Failed());
    return OpenHome::TestFramework::Totals::Failed() ? 1 : 0;
}
