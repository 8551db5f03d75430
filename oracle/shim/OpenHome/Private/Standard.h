// ohNet shim (TEST INFRASTRUCTURE ONLY): OpenHome/Private/Standard.h surface.
#pragma once
#include <OpenHome/Types.h>
#include <OpenHome/Exception.h>

namespace OpenHome {

class INonCopyable
{
protected:
    INonCopyable() {}
private:
    INonCopyable(const INonCopyable&);
    void operator=(const INonCopyable&);
};

} // namespace OpenHome
