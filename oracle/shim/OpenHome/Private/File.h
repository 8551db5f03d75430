// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/File.h -- included by Media/Tests/TestFlywheelRamper.cpp, nothing of
// it used by the suites that file registers.
#pragma once
#include <OpenHome/Types.h>
