// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/OsWrapper.h surface.  The reference sources the oracle links use none of
// it; Media/Tests/TestFlywheelRamper.cpp times its profiling test with Os::TimeInMs(env.OsCtx()).
#pragma once
#include <OpenHome/Types.h>
#include <chrono>

struct OsContext;

namespace OpenHome {

class Os
{
public:
    static TUint TimeInMs(OsContext* /*aContext*/)
    {
        return (TUint)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    static TUint64 TimeInUs(OsContext* /*aContext*/)
    {
        return (TUint64)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
};

// ohNet's Environment, as far as a test that only wants the time needs it
class Environment
{
public:
    OsContext* OsCtx() { return nullptr; }
};

namespace Net {
} // namespace Net

} // namespace OpenHome
